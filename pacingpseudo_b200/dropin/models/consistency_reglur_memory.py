"""B200-native drop-in for `models.consistency_reglur_memory`
(/root/reference/models/consistency_reglur_memory.py:13-102).

Same constructor, `forward(names_to_data, mode, step) -> dict` and returned keys. Differences are in
execution only: the weak and the strong image go through the backbone as ONE batched pass with
per-branch BatchNorm statistics (identical maths to the reference's two passes, running statistics are
updated weak-then-strong), and all per-pixel losses are one fused kernel.
"""
import torch
import torch.nn as nn

from pacingpseudo_b200 import functional as PF
from models.unet import UNet
from losses.losses import *  # noqa: F401,F403  (the reference re-exports the loss functions here)
from .aux_path_memory import *  # noqa: F401,F403
from .aux_path_memory import AuxPath


class _Pending(object):
    __slots__ = ("make",)

    def __init__(self, make):
        self.make = make


class NetOutputs(dict):
    """The reference's output dict (consistency_reglur_memory.py:84-100) with values that are produced on first
    access. `logits_aux_cls` is the only one: the training step never reads it (train_chaos.py:355 looks at it for
    the TensorBoard images only; the loss kernels interpolate the low-resolution aux logits at the labelled pixels
    themselves), so its F.interpolate (aux_path_memory.py:52) runs when somebody asks for the key. Every read path
    (`[]`, get, items, values, pop, iteration-based copies such as dict(out) / other.update(out)) resolves it."""

    def set_lazy(self, key, make):
        dict.__setitem__(self, key, _Pending(make))

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if isinstance(v, _Pending):
            v = v.make()
            dict.__setitem__(self, key, v)
        return v

    def get(self, key, default=None):
        return self[key] if key in self else default

    def __iter__(self):   # (an overridden __iter__ also keeps dict(out) / d.update(out) off CPython's raw-value fast path)
        return iter(list(dict.keys(self)))

    def keys(self):
        return list(dict.keys(self))

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def pop(self, key, *default):
        if key in self:
            v = self[key]
            dict.__delitem__(self, key)
            return v
        if default:
            return default[0]
        raise KeyError(key)

    def copy(self):
        return dict(self.items())


class ConsistencyRegulr(nn.Module):
    def __init__(self, kwargs_unet, kwargs_aux_path=None, args_parser=None):
        super(ConsistencyRegulr, self).__init__()
        self.kwargs_unet = kwargs_unet
        self.kwargs_aux_path = kwargs_aux_path
        self.args = args_parser
        self.backbone = UNet(**kwargs_unet)
        self.aux_path = AuxPath(**kwargs_aux_path)

    def forward(self, names_to_data, mode=None, step=None):
        assert mode in ['train', 'val', None]
        args = self.args
        net_outputs = NetOutputs()
        image = names_to_data['image']
        n = image.shape[0]
        train = mode == 'train'
        do_cr = train and args.do_decoder_consistency
        do_aux = train and args.do_aux_path
        do_ent = train and args.do_loss_ent

        # one backbone pass; with consistency on, weak and strong are batched (2 statistics groups)
        x = torch.cat((image, names_to_data['image_strong']), 0) if do_cr else image
        feat_names = tuple(self.aux_path.feat_stage) if do_aux else ()
        logits_all, feats = self.backbone.run_native(x, groups=2 if do_cr else 1, feat_names=feat_names)
        logits_weak = logits_all[:n]
        self.backbone.end_points._set_native(feats)
        self.backbone.end_points.update({'segmentation/logits': logits_all[n:] if do_cr else logits_weak})

        # fp32 one-hot (N, C+1, H, W) as ToTorchTensor builds it, or the compact uint8 class-index map (N, H, W)
        scribble = names_to_data['scribble']
        scb_target = scribble.to(torch.uint8) if scribble.dim() == 3 else PF.onehot_argmax(scribble)
        valid_mask = names_to_data.get('valid_mask')

        logits_aux = None   # low resolution (N, C, h, w): the loss kernels interpolate at the labelled pixels
        if do_aux:
            # the reference hands the aux path the LAST end_points written, i.e. the strong branch's
            # features whenever the consistency branch ran (unet.py:23 instance-owned dict)
            sel = [(f[n:] if do_cr else f) for f in (feats[s] for s in feat_names)]
            logits_aux, _ = self.aux_path.run_native(sel, scribble, step, self.backbone.engine.code)

        variant = None
        if do_cr:
            variant = args.loss_cr_variants
            if variant not in ('ce_loss', 'l1_loss', 'l2_loss', 'kl_loss'):
                raise ValueError('The loss is not implemented.')
        losses = PF.scribble_losses(
            logits_all, scb_target, args.ignored_index, za=logits_aux, mask=valid_mask if (do_ent or do_cr) else None,
            do_ent=do_ent, cr_variant=variant, detach_weak=bool(getattr(args, 'detach_weak_cr', False)),
            siamese=do_cr)

        net_outputs.update({'segmentation/logits': logits_weak, 'loss_pce': losses['loss_pce']})
        if do_ent:
            net_outputs.update({'loss_ent': losses['loss_ent']})
        if do_cr:
            net_outputs.update({'loss_cr': losses['loss_cr'], 'segmentation/logits_strong': logits_all[n:]})
        if do_aux:
            out_hw = tuple(scribble.shape[-2:])
            net_outputs.set_lazy('logits_aux_cls', lambda: PF.upsample_planes(logits_aux, out_hw))   # aux_path_memory.py:52
            net_outputs.update({'loss_aux_cls': losses['loss_aux']})
            if args.do_memory:
                net_outputs.update({'loss_memory': self.aux_path.memory_loss()})
        return net_outputs
