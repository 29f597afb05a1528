import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pytorch3dunet_b200
from tools.probes.probe_lib import ProbeLib
L = ProbeLib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
print("N n_acc rowbytes group_rows shift | issue_cyc/mma  total_cyc/mma   (ideal N/2)")
for N in (32, 64, 96, 128, 256):
    for (rb, gr, sh) in [(128, 8, 0), (128, 8, 1), (128, 10, 0), (128, 10, 3), (64, 8, 0), (64, 10, 0), (64, 10, 3), (32, 8, 0), (32, 10, 0), (32, 10, 3)]:
        iters = 256
        L.call("b200_probe_umma_issue", N, 1, iters, rb, gr, sh, out.data_ptr(), s)
        torch.cuda.synchronize()
        a, b = out.tolist()
        print(f"{N:4d}   1   {rb:4d}   {gr:3d}  {sh:3d}  |  {a / (iters * 4):8.1f}   {b / (iters * 4):8.1f}   {N / 2:6.1f}")
