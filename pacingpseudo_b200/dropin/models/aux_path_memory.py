"""B200-native drop-in for `models.aux_path_memory` (/root/reference/models/aux_path_memory.py).

Same kwargs, parameter names and memory-bank semantics (including the reference's quirks: only sample 0
updates the bank, the bank row is L2-normalised before the EMA in cosine mode, an all-zero row is
initialised with the plain mean). Execution is hand-written CUDA through pacingpseudo_b200.functional.
"""
import torch
import torch.nn as nn

from pacingpseudo_b200 import functional as PF
from pacingpseudo_b200.functional import AuxPathFunction, MemoryLossFunction


class AuxPath(nn.Module):
    """Auxiliary classification path + memory bank (aux_path_memory.py:10-66)."""

    def __init__(self, **kwargs):
        super(AuxPath, self).__init__()
        self.num_classes = kwargs['num_classes']
        self.feat_stage = kwargs['feat_stage']
        self.feat_ch = kwargs['feat_ch']
        self.hid_ch = kwargs['hid_ch']
        self.aux_drop_prob = kwargs['aux_drop_prob']
        # parameter holders, reference construction order (aux_path_memory.py:22-33)
        self.layer_bottleneck = nn.Sequential(
            nn.Dropout2d(self.aux_drop_prob),
            nn.Conv2d(sum(self.feat_ch), self.hid_ch, 3, 1, 1),
            nn.BatchNorm2d(self.hid_ch),
            nn.LeakyReLU(1e-2),
        )
        self.fc_cls = nn.Sequential(
            nn.Dropout2d(self.aux_drop_prob),
            nn.Conv2d(self.hid_ch, self.num_classes, 1, bias=False),
        )
        self.do_memory = kwargs['do_memory']
        self.max_step = kwargs['max_step']
        self.momentum = kwargs['update_momentum']
        self.ensemble_mode = kwargs['ensemble_mode']
        self.memory_bank = nn.Parameter(
            torch.zeros((self.num_classes, self.hid_ch, 1, 1), dtype=torch.float32), requires_grad=False)
        self._memory_target = None
        self.bank_sync = None  # data-parallel hook: called with the bank right after memory_update
        if len(self.feat_stage) not in (1, 2):
            raise NotImplementedError("pacingpseudo_b200 AuxPath: feat_stage must name one or two end points")

    @property
    def memory_target(self):
        dev = self.memory_bank.device
        if self._memory_target is None or self._memory_target.device != dev:
            self._memory_target = torch.arange(self.num_classes, dtype=torch.long, device=dev)
        return self._memory_target

    # ---- native path ---------------------------------------------------------------------------
    def run_native(self, feats, scribble, step, code):
        """feats: native NHWC tensors in feat_stage order. -> (logits_aux_low NCHW fp32 (N, C, h, w), aux_features NHWC).
        The logits are NOT up-sampled here (aux_path_memory.py:52): PF.upsample_planes(logits, scribble.shape[-2:])
        gives the reference's tensor, and the fused scribble loss interpolates at the labelled pixels itself."""
        conv, bn = self.layer_bottleneck[1], self.layer_bottleneck[2]
        fa = feats[0]
        fb = feats[1] if len(feats) > 1 else None
        w_conv = conv.weight
        true_in = list(self.feat_ch)[:len(feats)]
        if [f.shape[-1] for f in feats] != true_in:   # zero-padded stages (max_ch = 728): matching zero weight columns
            from models.unet import pad_in_channels
            w_conv = pad_in_channels(w_conv, [(t, f.shape[-1]) for t, f in zip(true_in, feats)]).contiguous()
        drop = None
        self._bank_drop = None
        if self.training and self.aux_drop_prob > 0:   # the two nn.Dropout2d layers (aux_path_memory.py:23,31)
            n_in = fa.shape[-1] + (fb.shape[-1] if fb is not None else 0)
            drop = (self.drop_factors(fa.shape[0], n_in, fa.device), self.drop_factors(fa.shape[0], self.hid_ch, fa.device))
            if self.do_memory:   # fc_cls (with its Dropout2d) is also applied to the bank (aux_path_memory.py:60)
                self._bank_drop = self.drop_factors(self.num_classes, self.hid_ch, fa.device)
        logits, aux_features = AuxPathFunction.apply(
            code, (bn.running_mean, bn.running_var, bn.num_batches_tracked), self.training,
            drop, fa, fb, w_conv, conv.bias, bn.weight, bn.bias, self.fc_cls[1].weight)
        if self.do_memory:
            self.memory_update(aux_features, scribble, step, _code=code)
            if self.bank_sync is not None:
                self.bank_sync(self.memory_bank.data)
        return logits, aux_features

    def drop_factors(self, n, channels, device):
        """Dropout2d as per-(sample, channel) factors: 0 with probability p, else 1/(1-p). Drawn from torch's CUDA
        generator (the reference draws from the same distribution; the streams differ, as they do between torch
        versions). Tests replace this method to pin the masks."""
        p = float(self.aux_drop_prob)
        if p >= 1.0:
            return torch.zeros((n, channels), dtype=torch.float32, device=device)
        keep = torch.rand((n, channels), device=device) >= p
        return keep.float() / (1.0 - p)

    def memory_loss(self):
        bank = self.memory_bank.data.view(self.num_classes, self.hid_ch)
        return MemoryLossFunction.apply(bank, self.fc_cls[1].weight, getattr(self, "_bank_drop", None))

    def forward(self, end_points, scribble, step):
        native = getattr(end_points, 'native', None)
        if native is None or any(s not in native for s in self.feat_stage):
            raise RuntimeError("pacingpseudo_b200 AuxPath needs the end points of the pacingpseudo_b200 UNet "
                               "(elab_end_points=True)")
        code = PF.BF16 if native[self.feat_stage[0]].dtype == torch.bfloat16 else PF.F32
        logits_low, _ = self.run_native([native[s] for s in self.feat_stage], scribble, step, code)
        out = {'logits_aux_cls': PF.upsample_planes(logits_low, scribble.shape[-2:]),   # aux_path_memory.py:52
               'aux_targets': (scribble if scribble.dim() == 3 else PF.onehot_argmax(scribble)).long()}
        if self.do_memory:
            # logits_memory is produced for API parity; its loss/gradient go through memory_loss()
            out['logits_memory'] = PF.bank_logits(self.memory_bank.data, self.fc_cls[1].weight)
            out['memory_target'] = self.memory_target
        return out

    @torch.no_grad()
    def memory_update(self, aux_features, scribble, step, _code=None, layout=None):
        """aux_path_memory.py:68-116. aux_features: (N, hid, h, w) in the reference's NCHW layout (the public
        signature; layout=None or 'nchw') or the native NHWC tensor (layout='nhwc', what run_native passes with
        _code). The layout is never guessed from the shape: a 64-wide NCHW map with hid_ch = 64 is ambiguous."""
        if layout is None:
            layout = 'nhwc' if _code is not None else 'nchw'
        if layout not in ('nchw', 'nhwc'):
            raise ValueError("memory_update: layout must be 'nchw' or 'nhwc'")
        if aux_features.dim() != 4 or aux_features.shape[1 if layout == 'nchw' else 3] != self.hid_ch:
            raise RuntimeError("memory_update: expected %s features with %d channels, got shape %s" % (
                layout.upper(), self.hid_ch, tuple(aux_features.shape)))
        if layout == 'nchw':
            aux_features = aux_features.permute(0, 2, 3, 1).contiguous()
        code = _code if _code is not None else (PF.BF16 if aux_features.dtype == torch.bfloat16 else PF.F32)
        m = _ramp_up_mo(step, self.max_step, self.momentum)
        PF.memory_update(code, aux_features.contiguous(), scribble, self.memory_bank.data, self.ensemble_mode, m)


def _ramp_up_mo(step, max_step, base_mo=0.9, gamma=0.9):
    """Momentum schedule of the bank EMA (aux_path_memory.py:118-120)."""
    return (1 - step / max_step) ** gamma * base_mo
