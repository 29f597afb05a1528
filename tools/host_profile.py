"""Host-side cost of one training step: seconds inside each C-ABI entry point vs the whole step (finds hidden syncs)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch3dunet_b200 as P
from pytorch3dunet_b200 import engine as E
name, fm, B, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
m = P.get_model(dict(name=name, in_channels=1, out_channels=1, f_maps=fm, num_levels=4 if name == "UNet3D" else 5)).cuda()
x = torch.rand(B, 1, S, S, S, device="cuda")
t = (torch.rand_like(x) > 0.5).float()
def step():
    for p in m.parameters():
        p.grad = None
    o, l = m(x, return_logits=True)
    P.losses.bce_dice_loss(l, t).backward()
for _ in range(3):
    step()
torch.cuda.synchronize()
E.HOST_PROF = {}
h0 = time.perf_counter()
for _ in range(5):
    step()
h1 = time.perf_counter()
torch.cuda.synchronize()
h2 = time.perf_counter()
prof, E.HOST_PROF = E.HOST_PROF, None
print(f"{name}: host enqueue {1e3 * (h1 - h0) / 5:.2f} ms/step, with final sync {1e3 * (h2 - h0) / 5:.2f} ms/step")
tot = sum(v[1] for v in prof.values())
print(f"inside C-ABI calls: {1e3 * tot / 5:.2f} ms/step over {sum(v[0] for v in prof.values()) // 5} calls")
for k, (c, sec) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {k:40s} {c // 5:4d} calls/step  {1e6 * sec / c:8.1f} us/call  {1e3 * sec / 5:7.2f} ms/step")
