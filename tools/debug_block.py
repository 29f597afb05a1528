"""On-GPU debugging aid: run one golden block through the engine and compare every backward intermediate with torch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import pytorch3dunet_b200 as P
from pytorch3dunet_b200 import engine as E
from tests.helpers import load_golden

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
os.environ["B200UNET_CONV_IMPL"] = sys.argv[1] if len(sys.argv) > 1 else "direct"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


rec, sd, grads = load_golden("block_singleconv_gcr_16_32")
mod = P.SingleConv(16, 32, order="gcr", num_groups=8)
mod.load_state_dict(sd)
mod = mod.cuda()
x = rec["x"].cuda().requires_grad_(True)
E.DEBUG = {}
y = mod(x)
(y * rec["r"].cuda()).sum().backward()
torch.cuda.synchronize()
dbg = E.DEBUG[""]
print("y", rel(y, rec["y"]), "grad_x", rel(x.grad, rec["grad_x"]))
for k, p in mod.named_parameters():
    print(k, rel(p.grad, grads[k]))

# ---- torch restatement from the engine's own saved x / y -----------------------------------------
xb = dbg["x"].double()                      # [N,D,H,W,C] bf16 input as the engine saw it
N, D, H, W, C = xb.shape
xc = xb.permute(0, 4, 1, 2, 3)
Wt = sd["conv.weight"].double().cuda()
gam, bet = sd["groupnorm.weight"].double().cuda(), sd["groupnorm.bias"].double().cuda()
G = 8
cpg = C // G
xg = xc.reshape(N, G, -1)
mean, var = xg.mean(-1), xg.var(-1, unbiased=False)
rstd = 1 / torch.sqrt(var + 1e-5)
print("mean", rel(dbg["mean_rstd"][..., 0], mean), "rstd", rel(dbg["mean_rstd"][..., 1], rstd))
a = (gam.view(1, G, cpg) * rstd.view(N, G, 1)).view(N, C)
b = bet.view(1, C) - a * mean.repeat_interleave(cpg, 1)
print("a", rel(dbg["ab"][..., 0], a), "b", rel(dbg["ab"][..., 1], b))
yb = dbg["y"].double()
r = rec["r"].double().cuda().permute(0, 2, 3, 4, 1)
dz_ref = r * (yb > 0)
print("dz", rel(dbg["dz"], dz_ref), " nnz frac", (dbg["dz"] != 0).double().mean().item(), (dz_ref != 0).double().mean().item())
dz = dbg["dz"].double()
xp = F.pad(xc, (1, 1, 1, 1, 1, 1))
ones = F.pad(torch.ones((D, H, W), dtype=torch.float64, device="cuda"), (1, 1, 1, 1, 1, 1))
Gr = torch.zeros(N, 27, C, Wt.shape[0], dtype=torch.float64, device="cuda")
Tr = torch.zeros(N, 27, Wt.shape[0], dtype=torch.float64, device="cuda")
for td in range(3):
    for th in range(3):
        for tw in range(3):
            t = (td * 3 + th) * 3 + tw
            Gr[:, t] = torch.einsum("ncdhw,ndhwo->nco", xp[:, :, td:td + D, th:th + H, tw:tw + W], dz)
            Tr[:, t] = torch.einsum("dhw,ndhwo->no", ones[td:td + D, th:th + H, tw:tw + W], dz)
print("T", rel(dbg["T"], Tr), "G", rel(dbg["G"].double().sum(1), Gr))
dW = torch.einsum("nc,ntco->oct", a, Gr) + torch.einsum("nc,nto->oct", b, Tr)
print("dW(engine) vs torch-from-engine-dz", rel(dbg["dW"].reshape(dW.shape), dW), " vs golden", rel(dbg["dW"], grads["conv.weight"]))
print("dW torch-from-engine-dz vs golden", rel(dW.reshape(grads["conv.weight"].shape), grads["conv.weight"]))
Wr = Wt.reshape(Wt.shape[0], C, 27)
s1 = torch.einsum("oct,nto->nc", Wr, Tr)
s2 = torch.einsum("oct,ntco->nc", Wr, Gr)
print("sums2", rel(dbg["sums2"][..., 0], s1), rel(dbg["sums2"][..., 1], s2))
print("dbeta from s1 vs golden", rel(s1.sum(0), grads["groupnorm.bias"]))
# golden-side dz for comparison
yg = rec["y"].double().cuda().permute(0, 2, 3, 4, 1)
dz_g = r * (yg > 0)
print("engine dz vs golden dz", rel(dz, dz_g), "mask mismatch frac", ((yb > 0) != (yg > 0)).double().mean().item())
print("y(engine) vs golden", rel(yb, yg), " |y| small frac", (yg.abs() < 1e-2).double().mean().item())
