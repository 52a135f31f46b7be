"""B200-native drop-in for `losses.losses` (/root/reference/losses/losses.py).

Same function names and argument meaning. Each called loss maps onto the fused sm_100a loss kernels
(csrc/loss.cu) through pacingpseudo_b200.functional; inputs must be CUDA tensors (there is no CPU path).
The two functions the reference defines but never calls (bidirectional_kl_loss,
multi_label_soft_margin_loss) keep their torch bodies, as SURVEY.md section 2 row 4 records.
"""
import torch
import torch.nn.functional as F

from pacingpseudo_b200 import functional as PF


def _as_4d(input, target):
    if input.dim() == 2:  # (N, C) logits, (N,) targets — e.g. the memory-bank logits
        return input[:, :, None, None], target[:, None, None]
    return input, target


def entropy_minimization_loss(input, valid_mask=None):
    """losses.py:9-24. input: logits (N, C, H, W); valid_mask: (N, 1, H, W) or None."""
    return PF.scribble_losses(input, None, -1, mask=valid_mask, do_ent=True)['loss_ent']


def cross_entropy_loss(input, target):
    """losses.py:26-33: F.cross_entropy with its default ignore_index (-100)."""
    input, target = _as_4d(input, target)
    return PF.scribble_losses(input, target, -100)['loss_pce']


def partial_cross_entropy_loss(input, target, ignore_index):
    """losses.py:35-43: mean over the non-ignored pixels of -log softmax(input)[target]."""
    input, target = _as_4d(input, target)
    return PF.scribble_losses(input, target, ignore_index)['loss_pce']


def soft_label_cross_entropy_loss(input, target, valid_mask=None):
    """losses.py:45-62. input: logits; target: probabilities."""
    return PF.pair_loss(input, target, valid_mask, 'ce_loss')


def l1_loss(input, target, valid_mask=None):
    """losses.py:64-79. input / target: probabilities."""
    return PF.pair_loss(input, target, valid_mask, 'l1_loss')


def l2_loss(input, target, valid_mask=None):
    """losses.py:81-96. input / target: probabilities."""
    return PF.pair_loss(input, target, valid_mask, 'l2_loss')


def kl_loss(input, target, valid_mask=None):
    """losses.py:98-116: KL(softmax(target) || softmax(input)) on two logits tensors."""
    return PF.scribble_losses(target, None, -1, zs=input, mask=valid_mask, cr_variant='kl_loss')['loss_cr']


def bidirectional_kl_loss(input, target, valid_mask=None):
    """losses.py:118-145 (never called by the reference; torch body kept)."""
    input_ll = F.log_softmax(input, dim=1)
    target_ll = F.log_softmax(target, dim=1)
    p_loss = F.kl_div(input_ll, target_ll, log_target=True, reduction='none')
    q_loss = F.kl_div(target_ll, input_ll, log_target=True, reduction='none')
    if valid_mask is not None:
        p_loss = (p_loss * valid_mask).sum() / max(valid_mask.sum(), 1e-8)
        q_loss = (q_loss * valid_mask).sum() / max(valid_mask.sum(), 1e-8)
    else:
        p_loss = p_loss.mean()
        q_loss = q_loss.mean()
    return (p_loss + q_loss) / 2


def dice_loss_fn(input, target):
    """losses.py:147-162. input: logits (N, C, H, W); target: one-hot encodings (N, C, H, W)."""
    return PF.DiceFunction.apply(input, target)


def multi_label_soft_margin_loss(input, target):
    """losses.py:164-171 (never called by the reference; torch body kept)."""
    return F.multilabel_soft_margin_loss(input, target)
