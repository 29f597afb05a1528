"""CPU: the algebra behind the "virtual concat" decoder convolution (csrc/upcat_conv.cu), restated in
fp64 torch and checked against conv3d(interpolate(b)) -- guards the tap tables the CUDA kernels are driven by:

  forward   y[2u+p] = sum_{j in {0,1}^3} Wp[p][j] . b[u + off(p,j)],   per axis p=0: off {-1,0} <- taps {-1},{0,+1};  p=1: off {0,+1} <- {-1,0},{+1}
  dgrad     db[u]   = sum_{e in {-1..2}^3} Wd[e]^T dz[2u+e],           per axis e=-1 <- {+1}; 0 <- {0,+1}; 1 <- {-1,0}; 2 <- {-1}
  wgrad     dW[t]   = sum_{r in {0,1}^3} Q[r - t],   Q[e] = sum_u dz[2u+e] (x) b[u]
"""
import itertools

import torch
import torch.nn.functional as F

OFF = {0: (-1, 0), 1: (0, 1)}                      # low-res offset of tap j for parity p
TAPS = {(0, 0): (-1,), (0, 1): (0, 1), (1, 0): (-1, 0), (1, 1): (1,)}   # full-res taps carried by (p, j)
ETAPS = {-1: (1,), 0: (0, 1), 1: (-1, 0), 2: (-1,)}                      # full-res taps carried by transpose offset e


def _setup(d=3, h=4, w=2, c1=3, co=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    b = torch.randn((1, c1, d, h, w), generator=g, dtype=torch.float64, requires_grad=True)
    W = torch.randn((co, c1, 3, 3, 3), generator=g, dtype=torch.float64, requires_grad=True)
    y = F.conv3d(F.interpolate(b, scale_factor=2, mode="nearest"), W, padding=1)
    return b, W, y


def _pad(t, lo, hi):
    return F.pad(t, (lo, hi, lo, hi, lo, hi))


def test_phase_decomposition_forward():
    b, W, y = _setup()
    d, h, w = b.shape[2:]
    bp = _pad(b.detach(), 1, 1)
    out = torch.zeros_like(y)
    for p in itertools.product((0, 1), repeat=3):
        acc = 0
        for j in itertools.product((0, 1), repeat=3):
            Wsum = sum(W.detach()[:, :, td + 1, th + 1, tw + 1] for td in TAPS[(p[0], j[0])] for th in TAPS[(p[1], j[1])]
                       for tw in TAPS[(p[2], j[2])])
            o = [OFF[p[a]][j[a]] + 1 for a in range(3)]
            acc = acc + torch.einsum("oc,ncdhw->nodhw", Wsum, bp[:, :, o[0]:o[0] + d, o[1]:o[1] + h, o[2]:o[2] + w])
        out[:, :, p[0]::2, p[1]::2, p[2]::2] = acc
    assert torch.allclose(out, y.detach(), atol=1e-12)


def test_transpose_offsets_dgrad_and_wgrad():
    b, W, y = _setup(seed=1)
    d, h, w = b.shape[2:]
    g = torch.Generator().manual_seed(2)
    dz = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dz)
    dzp = _pad(dz, 2, 2)   # dz[2u+e] for e in -1..2 needs 1 voxel below and 2 above; pad 2 keeps the slicing simple

    def lattice(e):        # dz[2u+e] over the low-res grid u
        s = [2 + e[a] for a in range(3)]
        return dzp[:, :, s[0]:s[0] + 2 * d:2, s[1]:s[1] + 2 * h:2, s[2]:s[2] + 2 * w:2]

    db = torch.zeros_like(b)
    Q = {}
    for e in itertools.product((-1, 0, 1, 2), repeat=3):
        Wd = sum(W.detach()[:, :, td + 1, th + 1, tw + 1] for td in ETAPS[e[0]] for th in ETAPS[e[1]] for tw in ETAPS[e[2]])
        L = lattice(e)
        db += torch.einsum("oc,nodhw->ncdhw", Wd, L)
        Q[e] = torch.einsum("nodhw,ncdhw->oc", L, b.detach())
    assert torch.allclose(db, b.grad, atol=1e-12)
    dW = torch.zeros_like(W)
    for t in itertools.product((-1, 0, 1), repeat=3):
        dW[:, :, t[0] + 1, t[1] + 1, t[2] + 1] = sum(Q[tuple(r[a] - t[a] for a in range(3))] for r in itertools.product((0, 1), repeat=3))
    assert torch.allclose(dW, W.grad, atol=1e-12)


def test_groupnorm_statistics_and_backward_scale_of_the_virtual_tensor():
    """sums of up2x(b) are 8x the sums of b, and d/db of sum_children f = A*sum(dxhat) + 8B*b + 8C (engine: coefficients (1,8,8))."""
    g = torch.Generator().manual_seed(3)
    b = torch.randn((2, 3, 2, 3, 2), generator=g, dtype=torch.float64)
    up = F.interpolate(b, scale_factor=2, mode="nearest")
    assert torch.allclose(up.sum((2, 3, 4)), 8 * b.sum((2, 3, 4)))
    assert torch.allclose((up * up).sum((2, 3, 4)), 8 * (b * b).sum((2, 3, 4)))
    A, B, C = 0.7, -0.3, 0.2
    dxhat = torch.randn(up.shape, generator=g, dtype=torch.float64)
    full = A * dxhat + B * up + C                      # GroupNorm backward, element-wise on the virtual tensor
    pooled = 8 * F.avg_pool3d(full, 2)                 # gradient reaching b = sum over its 8 copies
    assert torch.allclose(pooled, A * 8 * F.avg_pool3d(dxhat, 2) + 8 * B * b + 8 * C)
