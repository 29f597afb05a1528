import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pytorch3dunet_b200
from tools.probes.probe_lib import ProbeLib
L = ProbeLib()
s = torch.cuda.current_stream().cuda_stream
iters = 512
print("N issuers grid | cycles per MMA per issuer (max over CTAs)  -> aggregate cycles per MMA per SM")
for N in (32, 64, 128):
    for grid in (1, 148, 296):
        for ni in (1, 2, 4):
            if ni * N > 256:
                continue
            out = torch.zeros(grid * 4, dtype=torch.int64, device="cuda")
            L.call("b200_probe_umma_multi_issue", N, ni, iters, grid, out.data_ptr(), s)
            torch.cuda.synchronize()
            o = out.view(grid, 4)[:, :ni].double()
            per = o.max().item() / (iters * 4)
            ctas_per_sm = 2 if grid > 148 else 1
            print(f"{N:4d} {ni:3d} {grid:5d} | {per:8.1f} (mean {o.mean().item() / (iters * 4):.1f})  -> {per / ni / ctas_per_sm:8.1f}   (ideal {N / 2})")
