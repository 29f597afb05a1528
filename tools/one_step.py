"""Two identical training steps of a bench workload (for `ncu` launch lists: the second step is the one summarised)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pytorch3dunet_b200 as P
wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
m = P.get_model(wl["cfg"]).cuda()
flat = P.optim.FlatParameters(m)
B, S = wl["batch"], wl["size"]
x = torch.rand(B, 1, S, S, S, device="cuda")
t = (torch.rand_like(x) > 0.5).float()
for _ in range(steps):
    o, l = m(x, return_logits=True)
    P.losses.bce_dice_loss(l, t, fused=True).backward()
torch.cuda.synchronize()
print("launches fwd,bwd:", P.last_launch_counts())
