"""Adam over one flat fp32 parameter buffer (train_chaos.py:219 `torch.optim.Adam(lr, weight_decay)` semantics:
L2 weight decay folded into the gradient, bias-corrected moments) as a single hand-written kernel launch.

Parameters and their gradients are re-homed as views of two flat buffers, so (a) the optimizer step is one
launch instead of 165 tensors x several ops, and (b) the data-parallel gradient exchange is an all-reduce of
contiguous bucket slices of one buffer (pacingpseudo_b200/dp.py). `param_groups[0]['lr']` is honoured, so the
reference's `poly_lr_decay(optimizer, ...)` (utils/utils.py:37-51) works unchanged.
"""
import torch

from .lib import current_stream, get_lib, ptr


class FlatAdam:
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam: parameters must live on a CUDA device (no CPU path)")
        # every view starts on a 16-byte boundary (vector loads in the kernels); the padding elements stay zero
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.flat_param = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        for p, off in zip(self.params, self.offsets):
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
            p.grad = self.flat_grad[off:off + n].view_as(p)
            p._pp_direct_grad = True  # UNetFunction.backward accumulates straight into this view
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.param_groups = [dict(params=self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)]
        self.step_count = 0
        self.grad_scale = 1.0  # e.g. 1 / world_size after a summing all-reduce

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        for p, off in zip(self.params, self.offsets):  # re-attach if a caller dropped the views
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + p.numel()].view_as(p)

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        self.step_count += 1
        dev = self.flat_param.device
        with torch.cuda.device(dev):
            get_lib().call("pp_adam_step", ptr(self.flat_param), ptr(self.flat_grad), ptr(self.exp_avg),
                           ptr(self.exp_avg_sq), self.numel, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                           float(g["eps"]), float(g["weight_decay"]), self.step_count, float(self.grad_scale),
                           current_stream(dev))
