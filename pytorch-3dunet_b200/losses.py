"""Loss helpers used by bench.py / smoke() to close the forward+backward loop the way the reference trainer does
(trainer.py:364-365: loss_criterion(logits, target)).  The reference's own `losses.py` keeps working unchanged on the
engine's outputs; these are small functional equivalents of DiceLoss (losses.py:130-145) and BCEDiceLoss (:187-201)
for places where the reference package is not importable (the GPU box).  Losses are outside the hot path (SURVEY.md
section 2, row 4): a few passes over the C_out-channel logits.
"""
import torch
import torch.nn.functional as F


def _per_channel(t):
    return t.transpose(0, 1).reshape(t.size(1), -1)


def dice_loss(logits, target, eps=1e-6):
    p = _per_channel(torch.sigmoid(logits))
    t = _per_channel(target).float()
    inter = (p * t).sum(-1)
    den = (p * p).sum(-1) + (t * t).sum(-1)
    return 1.0 - (2.0 * inter / den.clamp(min=eps)).mean()


def bce_dice_loss(logits, target, alpha=1.0):
    return F.binary_cross_entropy_with_logits(logits, target) + alpha * dice_loss(logits, target)
