"""Steady-state check of the caching allocator: cudaMalloc calls / reserved bytes / step time over many steps."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch3dunet_b200 as P
name, fm, B, S, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
m = P.get_model(dict(name=name, in_channels=1, out_channels=1, f_maps=fm, num_levels=4 if name == "UNet3D" else 5)).cuda()
x = torch.rand(B, 1, S, S, S, device="cuda")
t = (torch.rand_like(x) > 0.5).float()
def step():
    for p in m.parameters():
        p.grad = None
    o, l = m(x, return_logits=True)
    P.losses.bce_dice_loss(l, t).backward()
print("alloc conf:", os.environ.get("PYTORCH_CUDA_ALLOC_CONF"))
for i in range(steps):
    st = torch.cuda.memory_stats()
    a0 = st.get("num_device_alloc", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record()
    step()
    e1.record()
    h1 = time.perf_counter()
    torch.cuda.synchronize()
    st = torch.cuda.memory_stats()
    print(f"step {i}: gpu {e0.elapsed_time(e1):7.2f} ms  host {1e3 * (h1 - h0):7.2f} ms  cudaMallocs {st.get('num_device_alloc', 0) - a0:3d}  "
          f"reserved {st['reserved_bytes.all.current'] / 2**30:6.2f} GiB  peak alloc {st['allocated_bytes.all.peak'] / 2**30:6.2f} GiB")
