"""Loss helpers used by bench.py / smoke() to close the forward+backward loop the way the reference trainer does
(trainer.py:364-365: loss_criterion(logits, target)).  The reference's own `losses.py` keeps working unchanged on the
engine's outputs; these are small functional equivalents of DiceLoss (losses.py:130-145) and BCEDiceLoss (:187-201)
for places where the reference package is not importable (the GPU box).  Losses are outside the hot path (SURVEY.md
section 2, row 4): a few passes over the C_out-channel logits.
"""
import os

import torch
import torch.nn.functional as F

# BCEDiceLoss as two passes over the logits in the engine's library (csrc/loss_ops.cu) instead of
# ~20 ATen kernels.  Selected by bce_dice_loss(..., fused=True) or B200UNET_FUSED_LOSS=1.
FUSED_LOSS = os.environ.get("B200UNET_FUSED_LOSS", "0") == "1"


def _per_channel(t):
    return t.transpose(0, 1).reshape(t.size(1), -1)


def dice_loss(logits, target, eps=1e-6):
    p = _per_channel(torch.sigmoid(logits))
    t = _per_channel(target).float()
    inter = (p * t).sum(-1)
    den = (p * p).sum(-1) + (t * t).sum(-1)
    return 1.0 - (2.0 * inter / den.clamp(min=eps)).mean()


class _FusedBCEDice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, alpha, eps):
        from ._lib import lib
        L = lib()
        x, t = logits.detach().contiguous(), target.detach().contiguous().float()
        n, c = x.shape[0], x.shape[1]
        v = x[0, 0].numel()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            P = L.query("b200_bce_dice_partials_count", n, c, v)
            partials = torch.empty((n * c, P, 4), dtype=torch.float32, device=x.device)
            loss = torch.empty((1,), dtype=torch.float32, device=x.device)
            coef = torch.empty((1 + 2 * c,), dtype=torch.float32, device=x.device)
            L.call("b200_bce_dice_fwd", x.data_ptr(), t.data_ptr(), n, c, v, float(alpha), float(eps), partials.data_ptr(),
                   loss.data_ptr(), coef.data_ptr(), stream)
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                L.call("b200_bce_dice_bwd", x.data_ptr(), t.data_ptr(), coef.data_ptr(), n, c, v, dx.data_ptr(), stream)
                ctx.save_for_backward(dx)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None, None, None


def bce_dice_loss(logits, target, alpha=1.0, fused=None, eps=1e-6):
    """BCEDiceLoss (reference losses.py:187-201).  fused: run the engine's two-pass kernels (fp32 CUDA logits only)."""
    if fused is None:
        fused = FUSED_LOSS
    if fused:
        if not (logits.is_cuda and logits.dtype == torch.float32 and logits.shape == target.shape and logits.dim() >= 3):
            raise RuntimeError("fused bce_dice_loss needs float32 CUDA logits and a target of the same shape")
        return _FusedBCEDice.apply(logits, target, alpha, eps)
    return F.binary_cross_entropy_with_logits(logits, target) + alpha * dice_loss(logits, target, eps)
