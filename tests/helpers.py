"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    sd = {k[3:]: v for k, v in rec.items() if k.startswith("sd/")}
    grads = {k[5:]: v for k, v in rec.items() if k.startswith("grad/")}
    return rec, sd, grads


def rel_l2(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    den = b.norm().item()
    if den == 0.0:
        return a.norm().item()
    return (a - b).norm().item() / den


def assert_close_l2(a, b, rtol, atol=0.0, msg=""):
    """|a-b|_2 <= rtol*|b|_2 + atol*sqrt(numel): relative L2 with an absolute floor for ~zero tensors."""
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    assert a.shape == b.shape, f"{msg}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).norm().item()
    bound = rtol * b.norm().item() + atol * (b.numel() ** 0.5)
    assert err <= bound, f"{msg}: |a-b|={err:.3e} > {bound:.3e} (|b|={b.norm().item():.3e}, rel={err / max(b.norm().item(), 1e-30):.3e})"
