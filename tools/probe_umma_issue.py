import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch3dunet_b200
from pytorch3dunet_b200._lib import lib
L = lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
print("N n_acc  issue_cyc/mma  total_cyc/mma   (ideal N/2)")
for N in (16, 32, 64, 128, 256):
    for n_acc in (1, 2, 4, 8, 16):
        if n_acc * N > 512:
            continue
        iters = 256
        L.call("b200_probe_umma_issue", N, n_acc, iters, out.data_ptr(), s)
        torch.cuda.synchronize()
        a, b = out.tolist()
        print(f"{N:4d} {n_acc:3d}   {a / (iters * 4):8.1f}   {b / (iters * 4):8.1f}   {N / 2:6.1f}")
