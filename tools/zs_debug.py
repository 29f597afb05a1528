"""Where the z-stacked conv kernel's warps spend their cycles (debug build with wait counters):
    make -C pytorch3dunet_b200/csrc debug && B200UNET_LIB=pytorch3dunet_b200/libb200unet_dbg.so python tools/zs_debug.py
Prints, per variant, the kernel time and per-plane cycle counts averaged over the CTAs."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200UNET_LIB", os.path.join(ROOT, "pytorch3dunet_b200", "libb200unet_dbg.so"))
import torch  # noqa: E402
from tests import gpu_util as U  # noqa: E402
from pytorch3dunet_b200 import engine as E  # noqa: E402
from pytorch3dunet_b200._lib import lib  # noqa: E402

L = lib()
shapes = [(2, 128, 128, 128, 32, 32), (2, 128, 128, 128, 16, 32), (2, 64, 64, 64, 32, 64)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in sys.argv[1].split(","))]
for (N, D, H, W, Cin, Cout) in shapes:
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((N, D, H, W, Cin), device="cuda", generator=g).bfloat16()
    wf = (torch.randn((N, 27, Cout, Cin), device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn((N, 64, Cout), device="cuda", generator=g)
    for zs, flags, label in [("1", 0, "zs"), ("1", 1, "zs, no global stores"), ("1", 4, "zs, epilogue = tmem ld only"), ("1", 8, "zs, no MMAs issued"),
                             ("1", 12, "zs, no MMAs, ld only"), ("0", 0, "halo kernel")]:
        os.environ["B200UNET_ZS"] = zs
        os.environ["B200UNET_DBG_FLAGS"] = str(flags)
        dbg = torch.zeros((148 * 2, 16), dtype=torch.int64, device="cuda")
        L.query("b200_set_debug_buffer", dbg.data_ptr())
        for _ in range(2):
            U.run_conv3(E.IMPL_TCGEN05, x, wf, b, act=E.ACT_RELU, want_stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            U.run_conv3(E.IMPL_TCGEN05, x, wf, b, act=E.ACT_RELU, want_stats=True)
        e1.record()
        torch.cuda.synchronize()
        d = dbg.double().cpu()
        d = d[d[:, 4] > 0]
        planes = d[:, 7].clamp_min(1)
        per = lambda i: (d[:, i] / planes).mean().item()  # noqa: E731
        print(f"{(N, D, H, W, Cin, Cout)} {label:32s} {e0.elapsed_time(e1) / 5 * 1e3:8.1f} us | per plane/tile (cycles): producer wait {per(0):7.0f} "
              f"total {per(1):7.0f} | issuer wait a_full {per(2):7.0f} tmem_empty {per(3):7.0f} total {per(4):7.0f} | "
              f"epilogue wait tmem_full {per(5):7.0f} total {per(6):7.0f} ld {per(8):7.0f}   (planes/CTA {planes.mean().item():.0f})")
        L.query("b200_set_debug_buffer", None)
