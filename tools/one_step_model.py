import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch3dunet_b200 as P
name, fm, B, S, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
torch.manual_seed(0)
m = P.get_model(dict(name=name, in_channels=1, out_channels=1, f_maps=fm, num_levels=4 if name == "UNet3D" else 5)).cuda()
x = torch.rand(B, 1, S, S, S, device="cuda")
t = (torch.rand_like(x) > 0.5).float()
for _ in range(steps):
    for p in m.parameters():
        p.grad = None
    o, l = m(x, return_logits=True)
    P.losses.bce_dice_loss(l, t).backward()
torch.cuda.synchronize()
