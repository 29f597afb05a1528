"""Quick timing of BASELINE cfg 3 shape (ResidualUNet3D f_maps=32, 5 levels, batch 4x1x96^3) and the SE variant at 64^3."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch3dunet_b200 as P
from pytorch3dunet_b200 import engine as E
for name, fm, B, S in [("ResidualUNet3D", 32, 4, 96), ("ResidualUNetSE3D", 64, 1, 96), ("UNet3D", 32, 2, 128), ("ResidualUNet3D", 32, 4, 96)]:
    torch.manual_seed(0)
    levels = 4 if name == "UNet3D" else 5
    m = P.get_model(dict(name=name, in_channels=1, out_channels=1, f_maps=fm, num_levels=levels)).cuda()
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
    E.TIMING = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
    h0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"host enqueue {1e3 * (h1 - h0) / 5:.2f} ms/step, cudaMallocs in timed region {torch.cuda.memory_stats().get('num_device_alloc', 0) - a0}")
    tm, E.TIMING = E.TIMING, None
    by = {}
    for tag, fl, a, b in tm:
        d = by.setdefault(tag, [0.0, 0.0]); d[0] += fl; d[1] += a.elapsed_time(b)
    ms = e0.elapsed_time(e1) / 5
    print(f"{name} f{fm} batch {B}x{S}^3: {ms:.2f} ms/step, {B / ms * 1e3:.1f} patches/s; conv kernels:",
          {k: f"{v[1] / 5:.2f} ms {v[0] / v[1] / 1e9:.0f} TF/s" for k, v in by.items()}, "launches", sum(P.last_launch_counts()))
