"""Validation Dice on the GPU (reference: utils/metrics.py:7-34 `compute_dice`, called per
sample on host copies of the softmax values by train_chaos.py:386-390 and inference.py:161-176).

`compute_dice_batch(scores, label)` scores the whole batch in one pass of a hand-written kernel and returns an
(N, C) tensor (NaN where a class is absent from both prediction and label, which the caller skips exactly as it
skips the reference's np.nan). `compute_dice(input, target)` keeps the reference's per-sample signature and return
type (a list of C floats) for C x H x W inputs given as numpy arrays or tensors. It is not part of the drop-in module
tree on purpose: a `utils` package there would shadow the reference's own `utils` (schedules, meters); the caller
switches with `from pacingpseudo_b200.metrics import compute_dice`. No CPU path: inputs are moved to the GPU.
"""
import numpy as np
import torch

from .lib import current_stream, get_lib, ptr


def compute_dice_batch(scores, label):
    """scores, label: (N, C, H, W) tensors (label one-hot) -> (N, C) fp32 CUDA tensor of per-sample class Dice."""
    if scores.shape != label.shape or scores.dim() != 4:
        raise ValueError("compute_dice_batch: scores %s and label %s must be equal-shaped (N, C, H, W)" % (
            tuple(scores.shape), tuple(label.shape)))
    if not scores.is_cuda:
        raise RuntimeError("compute_dice_batch: scores must be a CUDA tensor (there is no CPU path)")
    dev = scores.device
    scores = scores.detach().contiguous().float()
    label = label.detach().to(dev).contiguous().float()
    N, C, H, W = scores.shape
    lib = get_lib()
    lib.ensure_init(dev.index if dev.index is not None else 0)
    with torch.cuda.device(dev):
        dice = torch.empty((N, C), dtype=torch.float32, device=dev)
        scratch = torch.empty(3 * N * C * 8 + N * C * 4, dtype=torch.uint8, device=dev)
        lib.call("pp_dice_metric", ptr(scores), ptr(label), ptr(dice), ptr(scratch), N, C, H * W, current_stream(dev))
    return dice


def compute_dice(input, target):
    """Reference signature (utils/metrics.py:7): C x H x W softmax values and one-hot target of one sample -> list of
    C Dice values (np.nan for a class absent from both)."""
    assert tuple(input.shape) == tuple(target.shape)
    dev = torch.device("cuda", torch.cuda.current_device())
    s = torch.as_tensor(np.ascontiguousarray(input) if isinstance(input, np.ndarray) else input).to(dev)
    t = torch.as_tensor(np.ascontiguousarray(target) if isinstance(target, np.ndarray) else target).to(dev)
    out = compute_dice_batch(s[None].float(), t[None].float())[0].cpu().numpy().astype(np.float64)
    return [float(v) if not np.isnan(v) else np.nan for v in out]
