"""CPU oracle for the PacingPseudo training-step hot path — TEST INFRASTRUCTURE ONLY.

A from-scratch, functional restatement (plain torch CPU tensor arithmetic on a state dict, fp32 or
fp64) of the reference algorithm. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product path (pacingpseudo_b200/) never does.

Pinning: the reference ships no tests or golden vectors, so the pin is the reference
itself run in the build container: oracle/gen_golden.py imports /root/reference, runs it on seeded
synthetic inputs and commits the outputs to tests/golden/; tests/test_oracle.py checks this restatement
against those vectors (and, where /root/reference is present, against the live reference).

Every function cites the reference lines it restates (paths relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F

EPS_BN = 1e-5
MOM_BN = 0.1
SLOPE = 1e-2


# ------------------------------------------------------------------------------------------------
# network structure (models/unet.py:10-60)
# ------------------------------------------------------------------------------------------------
def unet_structure(init_ch=32, max_ch=512, output_stride=8):
    """-> (channels per stage, encoder [(pool?, dilation)] x6, decoder scales for stages 5..1)."""
    ch = [min(max_ch, 2 ** k * init_ch) for k in range(6)]
    if output_stride == 32:
        enc = [(False, 1), (True, 1), (True, 1), (True, 1), (True, 1), (True, 1)]
        scales = [2, 2, 2, 2, 2]
    elif output_stride == 16:
        enc = [(False, 1), (True, 1), (True, 1), (True, 1), (True, 1), (False, 2)]
        scales = [1, 2, 2, 2, 2]
    else:
        enc = [(False, 1), (True, 1), (True, 1), (True, 1), (False, 2), (False, 4)]
        scales = [1, 1, 2, 2, 2]
    return ch, enc, scales


def unet_param_shapes(input_ch=1, init_ch=32, max_ch=512, num_classes=5, output_stride=8, strided=False):
    """Ordered {state-dict key: shape} of models.unet.UNet (165-entry ConsistencyRegulr minus aux).
    strided: the is_stride_conv + is_trans_conv variant adds dec_blockK.up_samp.weight (unet.py:141)."""
    ch, _enc, _sc = unet_structure(init_ch, max_ch, output_stride)
    shapes = {}

    def conv_layer(prefix, cin, cout):
        shapes[prefix + '.conv.weight'] = (cout, cin, 3, 3)
        shapes[prefix + '.conv.bias'] = (cout,)
        shapes[prefix + '.norm_op.weight'] = (cout,)
        shapes[prefix + '.norm_op.bias'] = (cout,)
        shapes[prefix + '.norm_op.running_mean'] = (cout,)
        shapes[prefix + '.norm_op.running_var'] = (cout,)
        shapes[prefix + '.norm_op.num_batches_tracked'] = ()

    def block(name, cin, cout):
        conv_layer(name + '.conv_block.conv_layer1', cin, cout)
        conv_layer(name + '.conv_block.conv_layer2', cout, cout)

    cin = input_ch
    for k in range(6):
        block('enc_block%d' % (k + 1), cin, ch[k])
        cin = ch[k]
    for stage in (5, 4, 3, 2, 1):  # DecBlock(lower, skip, .): DoubleConv(lower + skip, skip), unet.py:145
        lower = ch[stage]
        skip = ch[stage - 1]
        if strided:   # ConvTranspose2d(lower, skip, ks, stride, bias=False) then DoubleConv(2 * skip, skip)
            ks = _sc[5 - stage]
            shapes['dec_block%d.up_samp.weight' % stage] = (lower, skip, ks, ks)
            lower = skip
        block('dec_block%d' % stage, lower + skip, skip)
    shapes['final_conv.weight'] = (num_classes, ch[0], 1, 1)
    shapes['final_conv.bias'] = (num_classes,)
    return shapes


def aux_param_shapes(num_classes=5, feat_ch=(512, 512), hid_ch=64):
    """aux_path_memory.py:22-43."""
    return {
        'layer_bottleneck.1.weight': (hid_ch, sum(feat_ch), 3, 3),
        'layer_bottleneck.1.bias': (hid_ch,),
        'layer_bottleneck.2.weight': (hid_ch,),
        'layer_bottleneck.2.bias': (hid_ch,),
        'layer_bottleneck.2.running_mean': (hid_ch,),
        'layer_bottleneck.2.running_var': (hid_ch,),
        'layer_bottleneck.2.num_batches_tracked': (),
        'fc_cls.1.weight': (num_classes, hid_ch, 1, 1),
        'memory_bank': (num_classes, hid_ch, 1, 1),
    }


def synth_state_dict(shapes, seed, dtype=torch.float32):
    """Deterministic, reference-independent parameter values (so fixtures need not ship 80 MB of weights)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith('num_batches_tracked'):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith('running_var'):
            sd[k] = (0.5 + torch.rand(shp, generator=g)).to(dtype)
        elif k.endswith('running_mean'):
            sd[k] = (0.1 * torch.randn(shp, generator=g)).to(dtype)
        elif k.endswith('norm_op.weight') or k.endswith('layer_bottleneck.2.weight'):
            sd[k] = (1.0 + 0.1 * torch.randn(shp, generator=g)).to(dtype)
        elif k.endswith('.bias'):
            sd[k] = (0.05 * torch.randn(shp, generator=g)).to(dtype)
        elif k == 'memory_bank':
            sd[k] = torch.zeros(shp, dtype=dtype)
        elif k.endswith('up_samp.weight'):   # (Cin, Cout, ks, ks)
            sd[k] = (torch.randn(shp, generator=g) * math.sqrt(1.0 / shp[0])).to(dtype)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = (torch.randn(shp, generator=g) * math.sqrt(2.0 / fan_in)).to(dtype)
    return sd


# ------------------------------------------------------------------------------------------------
# bf16 emulation: the CUDA path stores activations, activation gradients and 3x3-conv weights in bf16
# (fp32 accumulation, fp32 BatchNorm statistics, fp32 heads / first conv / losses). With quant=True the
# oracle rounds at exactly those storage points (straight-through forward, rounded gradient backward), so
# a bf16 run can be checked TIGHTLY against "the reference algorithm under the same storage rounding";
# the distance to the un-rounded reference is reported separately against the north-star tolerances.
# ------------------------------------------------------------------------------------------------
class _BF16Store(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().to(g.dtype)


def _q(x, quant):
    return _BF16Store.apply(x) if quant else x


def _qw(w, quant):
    # straight-through: forward value bf16(w), gradient exactly 1 w.r.t. w. (The dtype casts are differentiable in
    # autograd, so the rounding residue must be detached as a whole — `w.bfloat16().to() + (w - w.detach())` would
    # count the gradient twice; round 1 shipped that and its emulated weight gradients were 2x too large.)
    return w + (w.bfloat16().to(w.dtype) - w).detach() if quant else w


# ------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------
def conv_bn_lrelu(x, sd, prefix, dilation, training, quant=False, stride=1):
    """ConvLayer (unet.py:178-193): Conv2d(3x3, stride, pad=dil, bias) -> BatchNorm2d -> LeakyReLU(0.01).
    Running statistics in `sd` are updated in place when training (momentum 0.1, unbiased variance)."""
    w = sd[prefix + '.conv.weight']
    y = F.conv2d(x, _qw(w, quant and w.shape[1] > 1), sd[prefix + '.conv.bias'], stride, dilation, dilation)
    y = _q(y, quant)
    g, b = sd[prefix + '.norm_op.weight'], sd[prefix + '.norm_op.bias']
    rm, rv = sd[prefix + '.norm_op.running_mean'], sd[prefix + '.norm_op.running_var']
    if training:
        mean = y.mean(dim=(0, 2, 3))
        var = y.var(dim=(0, 2, 3), unbiased=False)
        n = y.numel() // y.shape[1]
        with torch.no_grad():
            rm.mul_(1 - MOM_BN).add_(MOM_BN * mean.detach().to(rm.dtype))
            rv.mul_(1 - MOM_BN).add_(MOM_BN * (var.detach() * n / max(n - 1, 1)).to(rv.dtype))
            sd[prefix + '.norm_op.num_batches_tracked'] += 1
    else:
        mean, var = rm.to(y.dtype), rv.to(y.dtype)
    yhat = (y - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + EPS_BN)
    z = yhat * g[None, :, None, None] + b[None, :, None, None]
    return _q(torch.where(z > 0, z, z * SLOPE), quant)


def double_conv(x, sd, prefix, dilation, training, quant=False, stride1=1):
    x = conv_bn_lrelu(x, sd, prefix + '.conv_block.conv_layer1', dilation, training, quant, stride1)
    return conv_bn_lrelu(x, sd, prefix + '.conv_block.conv_layer2', dilation, training, quant)


def upsample_bilinear_ac(x, size):
    """nn.Upsample(bilinear, align_corners=True) (unet.py:144) restated with explicit gather weights:
    src = dst * (in - 1) / (out - 1); lerp between floor(src) and min(floor(src) + 1, in - 1)."""
    n, c, h, w = x.shape
    H, W = size

    def axis(inp, out):
        scale = (inp - 1) / (out - 1) if out > 1 else 0.0
        src = torch.arange(out, dtype=torch.float32) * torch.tensor(scale, dtype=torch.float32)
        i0 = src.floor().long().clamp(max=inp - 1)
        i1 = (i0 + 1).clamp(max=inp - 1)
        w1 = (src - i0.to(torch.float32)).to(x.dtype)
        return i0, i1, 1 - w1, w1

    y0, y1, wy0, wy1 = axis(h, H)
    x0, x1, wx0, wx1 = axis(w, W)
    top = x[:, :, y0][:, :, :, x0] * wx0 + x[:, :, y0][:, :, :, x1] * wx1
    bot = x[:, :, y1][:, :, :, x0] * wx0 + x[:, :, y1][:, :, :, x1] * wx1
    return top * wy0[:, None] + bot * wy1[:, None]


def conv_transpose_ks(x, w):
    """ConvTranspose2d(kernel = stride = S, bias=False) (unet.py:141) restated without F.conv_transpose2d:
    out[n, co, S*y + a, S*x + b] = sum_ci x[n, ci, y, x] * w[ci, co, a, b]."""
    n, _ci, h, wd = x.shape
    co, S = w.shape[1], w.shape[2]
    y = torch.einsum('nihw,ioab->nohawb', x, w)
    return y.reshape(n, co, h * S, wd * S)


def unet_forward(sd, x, training, init_ch=32, max_ch=512, output_stride=8, prefix='', quant=False, strided=False):
    """UNet.forward (unet.py:62-98) -> dict of the 12 end points. strided: stride-2 first conv instead of the
    max-pool (unet.py:113-116) and ConvTranspose2d instead of the bilinear up-sampling (unet.py:141)."""
    _ch, enc_cfg, scales = unet_structure(init_ch, max_ch, output_stride)
    ep = {}
    enc = []
    cur = x
    for k, (pool, dil) in enumerate(enc_cfg):
        if pool and not strided:
            cur = F.max_pool2d(cur, 2, 2)  # unet.py:109
        cur = double_conv(cur, sd, '%senc_block%d' % (prefix, k + 1), dil, training, quant,
                          stride1=2 if (pool and strided) else 1)
        enc.append(cur)
        ep['encoder/stage%d' % (k + 1)] = cur
    for i, stage in enumerate((5, 4, 3, 2, 1)):
        skip = enc[stage - 1]
        s = scales[i]
        if strided:
            wt = sd['%sdec_block%d.up_samp.weight' % (prefix, stage)]
            up = _q(conv_transpose_ks(cur, _qw(wt, quant)), quant)
        else:
            up = _q(upsample_bilinear_ac(cur, (cur.shape[2] * s, cur.shape[3] * s)), quant) if s > 1 else cur
        cur = double_conv(torch.cat((up, skip), 1), sd, '%sdec_block%d' % (prefix, stage), 1, training, quant)  # :151
        ep['decoder/stage%d' % stage] = cur
    ep['segmentation/logits'] = F.conv2d(cur, sd[prefix + 'final_conv.weight'], sd[prefix + 'final_conv.bias'])
    return ep


# ------------------------------------------------------------------------------------------------
# losses (losses/losses.py)
# ------------------------------------------------------------------------------------------------
def _masked_mean(loss, valid_mask):
    """losses.py:19-23 / 57-61: sum(loss * mask) / max(sum(mask), 1e-8), else the element mean."""
    if valid_mask is not None:
        return (loss * valid_mask).sum() / valid_mask.sum().clamp(min=1e-8)
    return loss.mean()


def partial_cross_entropy(logits, target, ignore_index):
    """losses.py:35-43: mean over pixels with target != ignore_index of -log softmax(z)[target]."""
    logp = torch.log_softmax(logits, 1)
    keep = target != ignore_index
    t = target.clamp(0, logits.shape[1] - 1)
    nll = -logp.gather(1, t[:, None]).squeeze(1)
    return (nll * keep).sum() / keep.sum()


def entropy_minimization(logits, valid_mask=None):
    """losses.py:9-24."""
    return _masked_mean(-torch.softmax(logits, 1) * torch.log_softmax(logits, 1), valid_mask)


def soft_label_cross_entropy(logits, target_prob, valid_mask=None):
    """losses.py:45-62."""
    return _masked_mean(-target_prob * torch.log_softmax(logits, 1), valid_mask)


def l1(p, q, valid_mask=None):
    """losses.py:64-79."""
    return _masked_mean((p - q).abs().sum(1, keepdim=True), valid_mask)


def l2(p, q, valid_mask=None):
    """losses.py:81-96."""
    return _masked_mean(((p - q) ** 2).sum(1, keepdim=True), valid_mask)


def kl(logits_in, logits_tgt, valid_mask=None):
    """losses.py:98-116: exp(t) * (t - i) elementwise on log-probabilities."""
    i, t = torch.log_softmax(logits_in, 1), torch.log_softmax(logits_tgt, 1)
    return _masked_mean(t.exp() * (t - i), valid_mask)


def dice(logits, onehot):
    """losses.py:147-162."""
    p = torch.softmax(logits, 1).flatten(2)
    t = onehot.flatten(2).to(p.dtype)
    return -(2 * (p * t).sum(2) / (p.sum(2) + t.sum(2) + 1e-5)).mean()


# ------------------------------------------------------------------------------------------------
# strong colour augmentation (datasets/augmentations.py:98-166 as chained by chaos_aug_configs.py:63-86)
# ------------------------------------------------------------------------------------------------
def strong_color_augment_np(image, params):
    """Brightness -> Contrast -> GammaAugmentation(retain_stats=True, invert_data=False) on ONE float32 slice (H, W)
    with explicit draws: params = [apply_brightness, b, apply_contrast, a, apply_gamma, gamma, ...]."""
    import numpy as np
    eps = np.float32(1e-8)
    x = np.asarray(image, dtype=np.float32)
    ab, b, ac, a, ag, gamma = (np.float32(v) for v in params[:6])
    if ab:
        x = x + b                                                  # augmentations.py:110
    if ac:
        mean_, max_, min_ = np.mean(x), np.max(x), np.min(x)       # :125-127
        x = np.clip((x - mean_) * a + mean_, min_, max_)           # :128
    if ag:
        mean_, std_, max_, min_ = np.mean(x), np.std(x), np.max(x), np.min(x)   # :146-149
        x = np.power((x - min_) / (max_ - min_ + eps), gamma)      # :156
        x = (x - np.mean(x)) / (np.std(x) + eps)                   # :159
        x = x * std_ + mean_                                       # :160
    return x.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# aux path + memory bank (models/aux_path_memory.py)
# ------------------------------------------------------------------------------------------------
def compute_dice_np(scores, target):
    """utils/metrics.py:7-34 compute_dice, restated: scores, target (C, H, W) numpy -> list of C Dice values. pred =
    argmax over classes (first maximum); per class 2*sum(pred_c*t_c) / (sum(pred_c) + sum(t_c) + 1e-5); np.nan when
    the class is absent from both the prediction and the target."""
    import numpy as np
    C = scores.shape[0]
    pred = np.argmax(scores, axis=0)
    out = []
    for c in range(C):
        pc = (pred == c).astype(np.float64).reshape(-1)
        tc = np.asarray(target[c], dtype=np.float64).reshape(-1)
        if not pc.any() and not tc.any():
            out.append(np.nan)
        else:
            out.append(2.0 * float((pc * tc).sum()) / (float(pc.sum()) + float(tc.sum()) + 1e-5))
    return out


def synth_drop_factors(seed, N, cin, hid, C, p):
    """Deterministic Dropout2d factors (0 or 1/(1-p)) for the three dropout calls of one aux-path forward, in call
    order: input features [N, cin], bottleneck output [N, hid], memory bank [C, hid] (aux_path_memory.py:23,31,60)."""
    g = torch.Generator().manual_seed(seed)
    return tuple((torch.rand(shape, generator=g) >= p).float() / (1.0 - p) for shape in ((N, cin), (N, hid), (C, hid)))


def ramp_up_mo(step, max_step, base_mo=0.9, gamma=0.9):
    """aux_path_memory.py:118-120."""
    return (1 - step / max_step) ** gamma * base_mo


def aux_forward(sd, feats, out_hw, training, prefix='aux_path.', quant=False, drop=None):
    """aux_path_memory.py:49-52. drop = None (Dropout2d(p=0) is the identity) or (s_in [N, Cin], s_hid [N, hid]):
    the two nn.Dropout2d layers (aux_path_memory.py:23,31) as explicit per-(sample, channel) factors, 0 or 1/(1-p)."""
    x = torch.cat(feats, 1)
    if drop is not None:
        x = _q(x * drop[0].to(x.dtype)[:, :, None, None], quant)
    y = F.conv2d(x, _qw(sd[prefix + 'layer_bottleneck.1.weight'], quant), sd[prefix + 'layer_bottleneck.1.bias'], 1, 1)
    y = _q(y, quant)
    g, b = sd[prefix + 'layer_bottleneck.2.weight'], sd[prefix + 'layer_bottleneck.2.bias']
    rm, rv = sd[prefix + 'layer_bottleneck.2.running_mean'], sd[prefix + 'layer_bottleneck.2.running_var']
    if training:
        mean, var = y.mean(dim=(0, 2, 3)), y.var(dim=(0, 2, 3), unbiased=False)
        n = y.numel() // y.shape[1]
        with torch.no_grad():
            rm.mul_(1 - MOM_BN).add_(MOM_BN * mean.detach().to(rm.dtype))
            rv.mul_(1 - MOM_BN).add_(MOM_BN * (var.detach() * n / max(n - 1, 1)).to(rv.dtype))
            sd[prefix + 'layer_bottleneck.2.num_batches_tracked'] += 1
    else:
        mean, var = rm.to(y.dtype), rv.to(y.dtype)
    z = (y - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + EPS_BN)
    z = z * g[None, :, None, None] + b[None, :, None, None]
    aux_features = _q(torch.where(z > 0, z, z * SLOPE), quant)
    fc_in = aux_features if drop is None else _q(aux_features * drop[1].to(x.dtype)[:, :, None, None], quant)
    low = F.conv2d(fc_in, sd[prefix + 'fc_cls.1.weight'])
    return upsample_bilinear_ac(low, out_hw), aux_features


@torch.no_grad()
def memory_update(bank, aux_features, scribble, step, max_step, momentum=0.9, mode='cosine_similarity'):
    """aux_path_memory.py:68-116 in closed form. Visits sample 0 only (the reference returns inside the
    per-sample loop). bank: (C, hid, 1, 1), updated in place."""
    C, hid = bank.shape[0], bank.shape[1]
    H, W = scribble.shape[-2:]
    emb = upsample_bilinear_ac(aux_features[:1], (H, W))[0].permute(1, 2, 0).reshape(H * W, hid)
    m = ramp_up_mo(step, max_step, momentum)
    for c in range(C):
        sel = scribble[0, c].reshape(-1) == 1
        if not bool(sel.any()):
            continue
        e = emb[sel]
        row = bank[c, :, 0, 0]
        if bool((row == 0).all()):
            bank[c, :, 0, 0] = e.mean(0)
            continue
        if mode == 'mean':
            upd = e.mean(0)
            base = row
        else:
            e = e / (e.pow(2).sum(1, keepdim=True).sqrt() + 1e-8)
            base = row / (row.pow(2).sum().sqrt() + 1e-8)
            wgt = 1 - (e * base[None]).sum(1, keepdim=True)
            wgt = wgt / (wgt.sum() + 1e-8)
            upd = (e * wgt).sum(0)
        bank[c, :, 0, 0] = (1 - m) * base + m * upd


# ------------------------------------------------------------------------------------------------
# the pacingpseudo step (models/consistency_reglur_memory.py:24-102)
# ------------------------------------------------------------------------------------------------
class StepConfig:
    def __init__(self, num_classes=5, ignored_index=5, do_loss_ent=True, do_decoder_consistency=True,
                 detach_weak_cr=False, loss_cr_variants='ce_loss', do_aux_path=True, do_memory=True,
                 feat_stage=('encoder/stage6', 'encoder/stage5'), max_step=400, update_momentum=0.9,
                 ensemble_mode='cosine_similarity', init_ch=32, max_ch=512, output_stride=8, quant=False, strided=False):
        self.__dict__.update(locals())
        del self.__dict__['self']


def consistency_forward(sd, batch, cfg, mode='train', step=0, training=True, drop=None):
    """ConsistencyRegulr.forward. `sd` keys carry the `backbone.` / `aux_path.` prefixes of the reference
    state dict. `training` is the module's BatchNorm mode (train_chaos.py never re-enters .train())."""
    net = cfg
    out = {}
    kw = dict(init_ch=net.init_ch, max_ch=net.max_ch, output_stride=net.output_stride, prefix='backbone.',
              quant=net.quant, strided=net.strided)
    ep = unet_forward(sd, batch['image'], training, **kw)
    zw = ep['segmentation/logits']
    target = batch['scribble'].argmax(1)
    out['segmentation/logits'] = zw
    out['loss_pce'] = partial_cross_entropy(zw, target, net.ignored_index)
    mask = batch.get('valid_mask')
    if mode == 'train' and net.do_loss_ent:
        out['loss_ent'] = entropy_minimization(zw, mask)
    if mode == 'train' and net.do_decoder_consistency:
        ep = unet_forward(sd, batch['image_strong'], training, **kw)  # the end-point dict is overwritten (T1)
        zs = ep['segmentation/logits']
        pw = torch.softmax(zw, 1)
        if net.detach_weak_cr:
            pw = pw.detach()
        v = net.loss_cr_variants
        if v == 'ce_loss':
            out['loss_cr'] = soft_label_cross_entropy(zs, pw, mask)
        elif v == 'l1_loss':
            out['loss_cr'] = l1(torch.softmax(zs, 1), pw, mask)
        elif v == 'l2_loss':
            out['loss_cr'] = l2(torch.softmax(zs, 1), pw, mask)
        elif v == 'kl_loss':
            out['loss_cr'] = kl(zs, zw, mask)
        else:
            raise ValueError('The loss is not implemented.')
        out['segmentation/logits_strong'] = zs
    if mode == 'train' and net.do_aux_path:
        feats = [ep[s] for s in net.feat_stage]
        # drop = (s_in, s_hid, s_bank): explicit Dropout2d factors of the aux path (aux_drop_prob > 0, train mode)
        za, aux_features = aux_forward(sd, feats, batch['scribble'].shape[-2:], training, quant=net.quant,
                                       drop=None if drop is None else drop[:2])
        out['logits_aux_cls'] = za
        out['loss_aux_cls'] = partial_cross_entropy(za, target, net.ignored_index)
        if net.do_memory:
            memory_update(sd['aux_path.memory_bank'], aux_features.detach(), batch['scribble'], step, net.max_step,
                          net.update_momentum, net.ensemble_mode)
            bank = sd['aux_path.memory_bank']
            if drop is not None:   # fc_cls = Dropout2d + 1x1 conv is applied to the bank as well (aux_path_memory.py:60)
                bank = bank * drop[2].to(bank.dtype)[:, :, None, None]
            lm = F.conv2d(bank, sd['aux_path.fc_cls.1.weight'])[:, :, 0, 0]
            out['loss_memory'] = partial_cross_entropy(lm, torch.arange(lm.shape[0]), -100)
    return out


def gaussian_ramp_up(t, base_value, max_t=80, scale=5.0):
    """utils/utils.py:53-65."""
    return base_value * math.exp(-scale * (1 - t / max_t)) if t < max_t else base_value


def total_loss(out, epoch, loss_ent_weight=1.0, loss_cr_weight=1.0, loss_aux_weight=0.01, loss_memory_weight=1.0,
               ramp_up_scale=8.0):
    """Loss weighting of train_chaos.py:273-310 (non-mutating restatement)."""
    loss = out['loss_pce']
    if 'loss_ent' in out:
        loss = loss + out['loss_ent'] * gaussian_ramp_up(epoch, loss_ent_weight, scale=ramp_up_scale)
    if 'loss_cr' in out:
        loss = loss + out['loss_cr'] * gaussian_ramp_up(epoch, loss_cr_weight, scale=ramp_up_scale)
    if 'loss_aux_cls' in out:
        loss = loss + out['loss_aux_cls'] * loss_aux_weight
    if 'loss_memory' in out:
        loss = loss + out['loss_memory'] * loss_memory_weight
    return loss


def adam_step(params, grads, state, lr, weight_decay, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam with L2 weight decay folded into the gradient (train_chaos.py:219)."""
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    for k, p in params.items():
        g = grads[k] + weight_decay * p
        m, v = state.setdefault(k, (torch.zeros_like(p), torch.zeros_like(p)))
        m.mul_(beta1).add_((1 - beta1) * g)
        v.mul_(beta2).add_((1 - beta2) * g * g)
        p.sub_((lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps))
