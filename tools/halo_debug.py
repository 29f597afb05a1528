import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch3dunet_b200
from pytorch3dunet_b200 import engine as E
from pytorch3dunet_b200._lib import lib
from tests import gpu_util as U
L = lib()
for (Cin, Cout) in [(32, 32), (32, 96)]:
    N, D, H, W = 2, 128, 128, 128
    x = torch.randn((N, D, H, W, Cin), device="cuda").bfloat16()
    wf = (torch.randn((N, 27, Cout, Cin), device="cuda") * 0.05).bfloat16()
    b = torch.randn((N, 64, Cout), device="cuda") * 0.1
    dbg = torch.zeros((148, 16), dtype=torch.int64, device="cuda")
    for want in (True, False):
        L.query("b200_set_debug_buffer", dbg.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        U.run_conv3(E.IMPL_TCGEN05, x, wf, b, act=E.ACT_RELU, want_stats=want)
        torch.cuda.synchronize()
        e0.record()
        U.run_conv3(E.IMPL_TCGEN05, x, wf, b, act=E.ACT_RELU, want_stats=want)
        e1.record()
        torch.cuda.synchronize()
        L.query("b200_set_debug_buffer", None)
        d = dbg.double().mean(0).tolist()
        tiles = d[7]
        print(f"{Cin}->{Cout} stats={want}: {e0.elapsed_time(e1):.3f} ms; per CTA (cycles/tile): producer wait a_empty {d[0]/tiles:.0f} total {d[1]/tiles:.0f} | "
              f"mma wait a_full {d[2]/tiles:.0f} wait tmem_empty {d[3]/tiles:.0f} total {d[4]/tiles:.0f} | epi(grp0) wait tmem_full {d[5]/tiles:.0f} total {d[6]/tiles:.0f} "
              f"[per own tile: ld {2*d[8]/tiles:.0f} math+store {2*d[9]/tiles:.0f} stats {2*d[10]/tiles:.0f}] | tiles/CTA {tiles:.0f}")
