"""Host-side scalar schedules of the training loop (the reference keeps them in utils/utils.py; its scripts use their
own copies unchanged — these are for drivers that do not import the reference, e.g. bench.py)."""
import math


def loss_weight_ramp_up(epoch, base_value, max_epoch=80, scale=5.0):
    """Weight of the entropy / consistency terms at `epoch` (utils/utils.py:53-65 `gaussian_ramp_up`, called with
    scale=8 by train_chaos.py:279,287): base * exp(-scale * (1 - epoch / max_epoch)) until max_epoch, then base."""
    if epoch < max_epoch:
        return base_value * math.exp(-scale * (1.0 - epoch / max_epoch))
    return base_value
