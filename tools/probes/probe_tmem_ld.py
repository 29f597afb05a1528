import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pytorch3dunet_b200
from tools.probes.probe_lib import ProbeLib
L = ProbeLib()
out = torch.zeros(4, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for N in (32, 96, 256):
    for mma_iters in (0, 4000):
        out.zero_()
        L.call("b200_probe_tmem_ld_contention", N, mma_iters, 200, out.data_ptr(), s)
        torch.cuda.synchronize()
        a, b = out[0].item(), out[1].item()
        print(f"N={N} mma stream {'ON ' if mma_iters else 'off'}: {a / 200:.1f} cycles per tcgen05.ld(32 cols x 4 warps)" + (f", {b / (mma_iters * 4):.1f} cycles per mma" if mma_iters else ""))
