"""Helpers for the GPU parity tests: thin wrappers that call the C-ABI directly on torch tensors, and plain
PyTorch fp32 restatements of each kernel's contract (the checker side)."""
import torch
import torch.nn.functional as F

import pytorch3dunet_b200  # noqa: F401  (makes the package importable under its alias)
from pytorch3dunet_b200 import engine as E
from pytorch3dunet_b200._lib import lib


def stream():
    return torch.cuda.current_stream().cuda_stream


def p(t):
    return None if t is None else t.data_ptr()


def cls_map(D, H, W, device):
    """[D,H,W] int64 border class index (d_cls*16 + h_cls*4 + w_cls), as csrc/common.cuh axis_cls"""
    def ax(n):
        i = torch.arange(n, device=device)
        lo, hi = i == 0, i == n - 1
        return torch.where(lo & hi, 3, torch.where(lo, 0, torch.where(hi, 2, 1)))
    return ax(D)[:, None, None] * 16 + ax(H)[None, :, None] * 4 + ax(W)[None, None, :]


def act_ref(z, act, slope):
    if act == E.ACT_RELU:
        return F.relu(z)
    if act == E.ACT_LEAKY:
        return F.leaky_relu(z, slope)
    if act == E.ACT_ELU:
        return F.elu(z)
    return z


def conv3_contract_ref(x, wf, biascls, act=E.ACT_NONE, slope=0.0, residual=None):
    """fp32 torch restatement of b200_conv3_fwd.  x: [N,D,H,W,Cin] (bf16 or f32); wf: [n_w,27,Cout,Cin] bf16;
    biascls: [n_b,64,Cout] f32 or None.  Returns fp32 [N,D,H,W,Cout] (before bf16 rounding)."""
    N, D, H, W, Cin = x.shape
    n_w, _, Cout, _ = wf.shape
    out = []
    cm = cls_map(D, H, W, x.device)
    for n in range(N):
        w = wf[n if n_w > 1 else 0].float().reshape(3, 3, 3, Cout, Cin).permute(3, 4, 0, 1, 2).contiguous()
        xi = x[n].float().permute(3, 0, 1, 2).unsqueeze(0)
        y = F.conv3d(xi, w, None, padding=1)[0].permute(1, 2, 3, 0)
        if biascls is not None:
            b = biascls[n if biascls.shape[0] > 1 else 0]
            y = y + b[cm]
        if residual is not None:
            y = y + residual[n].float()
        out.append(act_ref(y, act, slope))
    return torch.stack(out)


def run_conv3(impl, x, wf, biascls=None, act=E.ACT_NONE, slope=0.0, residual=None, want_stats=False):
    L = lib()
    N, D, H, W, Cin = x.shape
    n_w, _, Cout, _ = wf.shape
    is_f32 = int(x.dtype == torch.float32)
    r = L.query("b200_conv3_resolve_impl", impl, N, D, H, W, Cin, Cout, is_f32)
    assert r > 0, "implementation unsupported for this shape"
    y = torch.empty((N, D, H, W, Cout), dtype=torch.bfloat16, device=x.device)
    partials = None
    if want_stats:
        P = L.query("b200_conv3_partials_count", r, N, D, H, W, Cin, Cout)
        partials = torch.full((N, P, Cout, 2), float("nan"), device=x.device)
    n_b = 0 if biascls is None else biascls.shape[0]
    L.call("b200_conv3_fwd", r, p(x), is_f32, p(wf), n_w, p(biascls), n_b, p(residual), act, float(slope),
           N, D, H, W, Cin, Cout, p(y), 1 if want_stats else 0, None, p(partials), stream())
    sums = None
    if want_stats:
        sums = partials.double().sum(dim=1)
    return y, sums


def run_wgrad(impl, x, dz):
    L = lib()
    N, D, H, W, Cin = x.shape
    Cout = dz.shape[-1]
    is_f32 = int(x.dtype == torch.float32)
    r = L.query("b200_conv3_wgrad_resolve_impl", impl, N, D, H, W, Cin, Cout, is_f32)
    assert r > 0, "implementation unsupported for this shape"
    S = L.query("b200_conv3_wgrad_splits", r, N, D, H, W, Cin, Cout, is_f32)
    G = torch.full((N, S, 27, Cin, Cout), float("nan"), device=x.device)
    L.call("b200_conv3_wgrad", r, p(x), is_f32, p(dz), N, D, H, W, Cin, Cout, p(G), stream())
    return G.double().sum(dim=1)  # [N,27,Cin,Cout]


def wgrad_contract_ref(x, dz):
    """G[n][tap][ci][co] = sum_v dz[n,v,co] * x[n,v+tap-1,ci] in fp64 via conv3d tricks (small sizes only)."""
    N, D, H, W, Cin = x.shape
    Cout = dz.shape[-1]
    xp = F.pad(x.double().permute(0, 4, 1, 2, 3), (1, 1, 1, 1, 1, 1))  # N,Cin,D+2,H+2,W+2
    dzz = dz.double()
    out = torch.zeros((N, 27, Cin, Cout), dtype=torch.float64, device=x.device)
    for td in range(3):
        for th in range(3):
            for tw in range(3):
                xs = xp[:, :, td:td + D, th:th + H, tw:tw + W]  # N,Cin,D,H,W
                out[:, (td * 3 + th) * 3 + tw] = torch.einsum("ncdhw,ndhwo->nco", xs, dzz)
    return out


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
