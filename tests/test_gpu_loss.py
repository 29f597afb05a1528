"""GPU: the fused BCEDiceLoss kernels (csrc/loss_ops.cu) against the oracle's fp64 formulation
(reference losses.py:187-201) -- value and gradient w.r.t. the logits."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,alpha", [((2, 1, 16, 16, 16), 1.0), ((1, 3, 5, 7, 9), 0.5), ((2, 2, 33, 20, 17), 2.0), ((2, 1, 64, 64, 64), 1.0)])
def test_fused_bce_dice_matches_oracle(shape, alpha):
    import pytorch3dunet_b200 as P
    from oracle import unet3d_oracle as O
    g = torch.Generator(device="cuda").manual_seed(11)
    logits = (torch.randn(shape, device="cuda", generator=g) * 2.0).requires_grad_(True)
    target = (torch.rand(shape, device="cuda", generator=g) > 0.6).float()
    loss = P.losses.bce_dice_loss(logits, target, alpha=alpha, fused=True)
    loss.backward()
    ref_x = logits.detach().double().requires_grad_(True)
    ref = O.bce_dice_loss(ref_x, target.double(), alpha=alpha)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 2e-6 * max(1.0, abs(ref.item()))
    err = (logits.grad.double() - ref_x.grad).norm() / ref_x.grad.norm()
    assert err < 1e-5, err
    # upstream gradient other than 1, and the eager path gives the same numbers
    logits2 = logits.detach().clone().requires_grad_(True)
    (3.0 * P.losses.bce_dice_loss(logits2, target, alpha=alpha, fused=True)).backward()
    assert torch.allclose(logits2.grad, 3.0 * logits.grad, rtol=1e-6, atol=0)
    logits3 = logits.detach().clone().requires_grad_(True)
    eager = P.losses.bce_dice_loss(logits3, target, alpha=alpha, fused=False)
    eager.backward()
    assert abs(eager.item() - loss.item()) < 1e-5
    assert (logits3.grad - logits.grad).norm() / logits.grad.norm() < 1e-4


def test_fused_loss_rejects_cpu_tensors():
    import pytorch3dunet_b200 as P
    with pytest.raises(RuntimeError):
        P.losses.bce_dice_loss(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4), fused=True)
