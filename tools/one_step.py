"""A few training steps of the headline workload (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pytorch3dunet_b200 as P
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
m = P.get_model(bench.CFG).cuda()
x = torch.rand(2, 1, 128, 128, 128, device="cuda")
t = (torch.rand_like(x) > 0.5).float()
for _ in range(steps):
    for p in m.parameters():
        p.grad = None
    o, l = m(x, return_logits=True)
    P.losses.bce_dice_loss(l, t).backward()
torch.cuda.synchronize()
print("done")
