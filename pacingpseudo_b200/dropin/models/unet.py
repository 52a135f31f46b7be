"""B200-native drop-in for the reference `models.unet` (/root/reference/models/unet.py).

Same constructor, same `forward(x) -> end_points dict`, same state-dict keys/shapes (fp32 OIHW master
weights, BatchNorm buffers) — but `forward` is ONE autograd Function that runs the whole network as
hand-written sm_100a kernels on NHWC bf16 activations (or fp32 with PP_PRECISION=fp32). The nn.Conv2d /
nn.BatchNorm2d objects below only HOLD parameters, in the reference's construction order so that a given
torch seed yields the reference's initial weights; they are never called.
"""
import os

import torch
import torch.nn as nn

from pacingpseudo_b200.functional import UNetEngine, UNetFunction

_STAGES = ["encoder/stage%d" % i for i in range(1, 7)] + ["decoder/stage%d" % i for i in range(5, 0, -1)]


def default_precision():
    return os.environ.get("PP_PRECISION", "bf16")


class EndPoints(dict):
    """The instance-owned dict UNet.forward returns (unet.py:23,78-98). Feature maps live on the device as
    NHWC tensors in the activation dtype (`.native`); the reference-format NCHW fp32 view of a stage is
    materialised on first access."""

    def __init__(self):
        super().__init__()
        self.native = {}
        self.true_channels = {}   # stages whose native tensor carries zero padding channels (max_ch = 728)

    def _set_native(self, feats):
        self.native = dict(feats)
        for k in feats:
            dict.pop(self, k, None)

    def _materialise(self, key):
        if not dict.__contains__(self, key) and key in self.native:
            t = self.native[key]
            c = self.true_channels.get(key, t.shape[-1])
            dict.__setitem__(self, key, t[..., :c].permute(0, 3, 1, 2).float())

    def __getitem__(self, key):
        self._materialise(key)
        return dict.__getitem__(self, key)

    def get(self, key, default=None):
        self._materialise(key)
        return dict.get(self, key, default)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self.native

    def _all(self):
        for k in list(self.native):
            self._materialise(k)

    def keys(self):
        self._all()
        return dict.keys(self)

    def items(self):
        self._all()
        return dict.items(self)

    def values(self):
        self._all()
        return dict.values(self)

    def __iter__(self):
        self._all()
        return dict.__iter__(self)

    def __len__(self):
        self._all()
        return dict.__len__(self)


class ConvLayer(nn.Module):
    """Parameter holder for Conv2d(bias) -> BatchNorm2d -> LeakyReLU(0.01) (unet.py:178-193)."""

    def __init__(self, in_ch, out_ch, kernel_size=3, stride=1, padding=1, dilation=1,
                 norm_op=nn.BatchNorm2d, nonlin_op=nn.LeakyReLU, negative_slop=1e-2):
        super().__init__()
        self.conv = nn.Conv2d(in_ch, out_ch, kernel_size, stride, padding, dilation)
        self.norm_op = norm_op(out_ch)
        self.nonlin_op = nonlin_op(negative_slop)


class DoubleConv(nn.Module):
    def __init__(self, in_ch, out_ch, ks1=3, stride1=1, padding1=1, dilation1=1,
                 ks2=3, stride2=1, padding2=1, dilation2=1):
        super().__init__()
        self.conv_layer1 = ConvLayer(in_ch, out_ch, ks1, stride1, padding1, dilation1)
        self.conv_layer2 = ConvLayer(out_ch, out_ch, ks2, stride2, padding2, dilation2)


class EncBlock(nn.Module):
    """unet.py:100-127: max-pool + double conv, or (is_stride_conv) a stride-2 first conv, or dilated convs."""

    def __init__(self, in_ch, out_ch, do_subsamp=True, is_stride_conv=False, dilation=1):
        super().__init__()
        self.pooling = nn.MaxPool2d(2, 2) if (do_subsamp and not is_stride_conv) else None
        stride1 = 2 if (do_subsamp and is_stride_conv) else 1
        self.conv_block = DoubleConv(in_ch, out_ch, stride1=stride1, padding1=dilation, dilation1=dilation,
                                     padding2=dilation, dilation2=dilation)


class DecBlock(nn.Module):
    """unet.py:129-152: bilinear up-sampling, or (is_trans_conv) ConvTranspose2d(lower_ch, skip_ch, ks, stride,
    bias=False) that also adjusts the channels, then the double conv over cat((up, skip), 1)."""

    def __init__(self, lower_ch, skip_ch, out_ch, trans_ks=2, trans_stride=2, is_trans_conv=False):
        super().__init__()
        if is_trans_conv:
            self.up_samp = nn.ConvTranspose2d(lower_ch, skip_ch, trans_ks, trans_stride, bias=False)
            self.conv_block = DoubleConv(2 * skip_ch, out_ch)
        else:
            self.up_samp = nn.Upsample(scale_factor=trans_stride, mode='bilinear', align_corners=True)
            self.conv_block = DoubleConv(lower_ch + skip_ch, skip_ch)


_KERNEL_WIDTHS = (32, 64, 128, 256, 512, 1024)   # channel counts the conv / BatchNorm kernels are built for


def kernel_width(c):
    """Smallest supported channel count >= c. A stage of another width (`--max_ch 728`, train_chaos.py:71) runs with
    zero weights, zero gamma / beta / bias on the padding channels: they stay exactly 0 through conv, BatchNorm and
    LeakyReLU and contribute nothing downstream, so the network computes what the reference computes."""
    for w in _KERNEL_WIDTHS:
        if c <= w:
            return w
    raise ValueError("pacingpseudo_b200 UNet: %d channels exceed the widest supported stage (1024)" % c)


def pad_in_channels(w, segments):
    """w: (Cout, sum of true widths, kh, kw); segments: [(true, internal), ...] in concat order -> zero columns are
    inserted after each segment (differentiable: autograd slices the gradient back)."""
    if all(t == i for t, i in segments):
        return w
    parts, off = [], 0
    for t, i in segments:
        parts.append(w[:, off:off + t])
        if i > t:
            parts.append(w.new_zeros((w.shape[0], i - t) + tuple(w.shape[2:])))
        off += t
    return torch.cat(parts, 1)


def pad_vector(v, n, value=0.0):
    return v if v.shape[0] == n else torch.cat((v, v.new_full((n - v.shape[0],), value)))


class UNet(nn.Module):
    def __init__(self, input_ch=1, init_ch=32, max_ch=512, num_classes=4, output_stride=32,
                 is_stride_conv=False, is_trans_conv=False, elab_end_points=False, precision=None):
        super().__init__()
        self.elab_end_points = elab_end_points
        self.end_points = EndPoints()
        assert is_trans_conv == is_stride_conv, \
            "Only combo of stride_conv and trans_conv or maxpool and upsample is allowed."
        assert output_stride in [8, 16, 32]
        sc, tc = is_stride_conv, is_trans_conv
        ch_ls = [min(max_ch, 2 ** k * init_ch) for k in range(6)]
        self.enc_block1 = EncBlock(input_ch, ch_ls[0], do_subsamp=False, is_stride_conv=sc)
        self.enc_block2 = EncBlock(ch_ls[0], ch_ls[1], do_subsamp=True, is_stride_conv=sc)
        self.enc_block3 = EncBlock(ch_ls[1], ch_ls[2], do_subsamp=True, is_stride_conv=sc)
        self.enc_block4 = EncBlock(ch_ls[2], ch_ls[3], do_subsamp=True, is_stride_conv=sc)
        if output_stride == 32:
            self.enc_block5 = EncBlock(ch_ls[3], ch_ls[4], do_subsamp=True, is_stride_conv=sc)
            self.enc_block6 = EncBlock(ch_ls[4], ch_ls[5], do_subsamp=True, is_stride_conv=sc)
            self.dec_block5 = DecBlock(ch_ls[5], ch_ls[4], ch_ls[4], 2, 2, is_trans_conv=tc)
            self.dec_block4 = DecBlock(ch_ls[4], ch_ls[3], ch_ls[3], 2, 2, is_trans_conv=tc)
        elif output_stride == 16:
            self.enc_block5 = EncBlock(ch_ls[3], ch_ls[4], do_subsamp=True, is_stride_conv=sc)
            self.enc_block6 = EncBlock(ch_ls[4], ch_ls[5], do_subsamp=False, is_stride_conv=sc, dilation=2)
            self.dec_block5 = DecBlock(ch_ls[5], ch_ls[4], ch_ls[4], 1, 1, is_trans_conv=tc)
            self.dec_block4 = DecBlock(ch_ls[4], ch_ls[3], ch_ls[3], 2, 2, is_trans_conv=tc)
        else:
            self.enc_block5 = EncBlock(ch_ls[3], ch_ls[4], do_subsamp=False, is_stride_conv=sc, dilation=2)
            self.enc_block6 = EncBlock(ch_ls[4], ch_ls[5], do_subsamp=False, is_stride_conv=sc, dilation=4)
            self.dec_block5 = DecBlock(ch_ls[5], ch_ls[4], ch_ls[4], 1, 1, is_trans_conv=tc)
            self.dec_block4 = DecBlock(ch_ls[4], ch_ls[3], ch_ls[3], 1, 1, is_trans_conv=tc)
        self.dec_block3 = DecBlock(ch_ls[3], ch_ls[2], ch_ls[2], is_trans_conv=tc)
        self.dec_block2 = DecBlock(ch_ls[2], ch_ls[1], ch_ls[1], is_trans_conv=tc)
        self.dec_block1 = DecBlock(ch_ls[1], ch_ls[0], ch_ls[0], is_trans_conv=tc)
        self.final_conv = nn.Conv2d(ch_ls[0], num_classes, 1, 1)

        # kernel-side widths: stages that are not 32 * 2^k wide are zero-padded up to the next such width
        self._ch_true = ch_ls
        self._ch_int = [kernel_width(c) for c in ch_ls]
        self._padded = self._ch_int != ch_ls
        if self._padded and is_stride_conv:
            raise NotImplementedError("pacingpseudo_b200 UNet: max_ch=%d with is_stride_conv/is_trans_conv" % max_ch)
        for i, name in enumerate(_STAGES):
            k = i if i < 6 else 10 - i           # encoder stage i+1 -> ch[i]; decoder stage s -> ch[s-1]
            if self._ch_int[k] != ch_ls[k]:
                self.end_points.true_channels[name] = ch_ls[k]
        self._cfg = (input_ch, init_ch, max(self._ch_int), num_classes, output_stride)
        self._strided = bool(is_stride_conv)
        self._precision = precision or default_precision()
        self._engine = None

    # ---- native execution ----------------------------------------------------------------------
    @property
    def engine(self):
        if self._engine is None:
            self._engine = UNetEngine(*self._cfg, self._precision, strided=self._strided)
        return self._engine

    def _layer_modules(self):
        mods = []
        for name, _cin, _cout, _dil in self.engine.layers:
            m = self
            for part in name.split('.'):
                m = getattr(m, part)
            mods.append(m)
        return mods

    def run_native(self, x, groups=1, feat_names=()):
        """-> (logits NCHW fp32, {name: native NHWC feature}). `groups` = BatchNorm statistics groups."""
        learnable, buffers, copy_back = [], [], []
        for (name, _cin, cout_i, _dil), m in zip(self.engine.layers, self._layer_modules()):
            if isinstance(m, nn.ConvTranspose2d):   # weight only (bias=False, no BatchNorm): unet.py:141
                learnable += [m.weight, None, None, None]
                buffers += [None, None, None]
                continue
            w, b, g, bt = m.conv.weight, m.conv.bias, m.norm_op.weight, m.norm_op.bias
            rm, rv = m.norm_op.running_mean, m.norm_op.running_var
            if self._padded:
                w = pad_in_channels(w, self._in_segments(name))
                if w.shape[0] != cout_i:
                    n_t = w.shape[0]
                    w = torch.cat((w, w.new_zeros((cout_i - n_t,) + tuple(w.shape[1:]))), 0)
                    b, g, bt = pad_vector(b, cout_i), pad_vector(g, cout_i), pad_vector(bt, cout_i)
                    rm_p, rv_p = pad_vector(rm.detach(), cout_i), pad_vector(rv.detach(), cout_i, 1.0)
                    copy_back.append((rm, rm_p, rv, rv_p, n_t))
                    rm, rv = rm_p, rv_p
                w = w.contiguous()
            learnable += [w, b, g, bt]
            buffers += [rm, rv, m.norm_op.num_batches_tracked]
        learnable += [self.final_conv.weight, self.final_conv.bias]
        feat_names = tuple(feat_names)
        outs = UNetFunction.apply(self.engine, buffers, groups, self.training, feat_names, x, *learnable)
        if self.training:
            with torch.no_grad():
                for rm, rm_p, rv, rv_p, n_t in copy_back:   # the kernels updated the padded running statistics
                    rm.copy_(rm_p[:n_t])
                    rv.copy_(rv_p[:n_t])
        return outs[0], dict(zip(feat_names, outs[1:]))

    def _in_segments(self, name):
        """[(true, internal)] input-channel segments of a conv layer, in the reference's concat order (up, skip)."""
        ct, ci = self._ch_true, self._ch_int
        block, _, layer = name.partition('.conv_block.')
        k = int(block[-1])
        if block.startswith('enc'):
            if layer == 'conv_layer2':
                return [(ct[k - 1], ci[k - 1])]
            return [(self._cfg[0], self._cfg[0])] if k == 1 else [(ct[k - 2], ci[k - 2])]
        if layer == 'conv_layer2':
            return [(ct[k - 1], ci[k - 1])]
        return [(ct[k], ci[k]), (ct[k - 1], ci[k - 1])]   # DecBlock(lower = ch[k], skip = ch[k-1]) (unet.py:36-58)

    def forward(self, x):
        names = _STAGES if self.elab_end_points else ()
        logits, feats = self.run_native(x, 1, names)
        self.end_points._set_native(feats)
        self.end_points.update({"segmentation/logits": logits})
        return self.end_points
