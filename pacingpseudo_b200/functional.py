"""torch.autograd.Function wrappers around the C ABI (include/pacingpseudo_b200.h).

PyTorch is used here for device memory (caching allocator), streams and the autograd graph at the
granularity of whole fused regions: one Function for the UNet, one for the aux path, one for the
fused scribble loss, one each for Dice and the bank loss. Every Function calls hand-written CUDA
through ctypes; none of them computes with torch ops.
"""
import ctypes

import torch
from torch.autograd.function import once_differentiable

from .lib import BF16, CR_VARIANTS, F32, current_stream, get_lib, ptr, require_cuda

_PRECISIONS = {"bf16": BF16, "fp32": F32}


def dtype_code(precision):
    if precision not in _PRECISIONS:
        raise ValueError("precision must be 'bf16' or 'fp32', got %r" % (precision,))
    return _PRECISIONS[precision]


def act_dtype(code):
    return torch.bfloat16 if code == BF16 else torch.float32


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _view_bytes(ws, offset, shape, dtype):
    n = 1
    for s in shape:
        n *= s
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    return ws[offset:offset + nbytes].view(dtype).view(*shape)


# --------------------------------------------------------------------------------------------
# UNet
# --------------------------------------------------------------------------------------------
class UNetEngine:
    """C-side executor handle + layer table for one UNet configuration (models/unet.py:10-60)."""

    def __init__(self, input_ch, init_ch, max_ch, num_classes, output_stride, precision, strided=False):
        self.lib = get_lib()
        self.code = dtype_code(precision)
        self.precision = precision
        self.num_classes = num_classes
        self.input_ch = input_ch
        h = ctypes.c_void_p()
        self.lib.call("pp_unet_create_ex", input_ch, init_ch, max_ch, num_classes, output_stride, self.code,
                      int(bool(strided)), ctypes.byref(h))
        self.handle = h
        self.nconv = self.lib.cdll.pp_unet_num_convs(self.handle)
        self.layers = []
        self.kinds = []   # per layer (kind, scale): 0 conv+BN, 1 stride-2 conv+BN, 2 ConvTranspose2d (weight only)
        for i in range(self.nconv):
            cin, cout, dil = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            name = ctypes.c_char_p()
            self.lib.call("pp_unet_conv_info", self.handle, i, ctypes.byref(cin), ctypes.byref(cout),
                          ctypes.byref(dil), ctypes.byref(name))
            self.layers.append((name.value.decode(), cin.value, cout.value, dil.value))
            kind, scale = ctypes.c_int(), ctypes.c_int()
            self.lib.call("pp_unet_conv_kind", self.handle, i, ctypes.byref(kind), ctypes.byref(scale))
            self.kinds.append((kind.value, scale.value))

    def __del__(self):
        try:
            self.lib.cdll.pp_unet_destroy(self.handle)
        except Exception:
            pass

    def workspace_bytes(self, N, H, W, G):
        n = self.lib.cdll.pp_unet_workspace_bytes(self.handle, N, H, W, G)
        if n < 0:
            raise RuntimeError("pp_unet_workspace_bytes: %s" % self.lib.last_error())
        return n

    def activation(self, name, N, H, W, G):
        act, off = ctypes.c_int(), ctypes.c_longlong()
        C, h, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.lib.call("pp_unet_activation", self.handle, name.encode(), N, H, W, G, ctypes.byref(act),
                      ctypes.byref(off), ctypes.byref(C), ctypes.byref(h), ctypes.byref(w))
        return act.value, off.value, C.value, h.value, w.value


class UNetFunction(torch.autograd.Function):
    """(x, *learnable) -> (logits NCHW fp32, *native end points NHWC).

    `learnable` = per conv layer [weight, bias, gamma, beta] in engine order, then head [weight, bias].
    `buffers`   = per conv layer [running_mean, running_var, num_batches_tracked] (updated in place).
    A ConvTranspose2d layer (engine.kinds[i][0] == 2) has a weight only: its other slots are None.
    """

    @staticmethod
    def forward(ctx, engine, buffers, groups, training, feat_names, x, *learnable):
        require_cuda(x, "UNet input")
        lib = engine.lib
        x = x.contiguous().float()
        N, cin, H, W = x.shape
        if cin != engine.input_ch:
            raise RuntimeError("pacingpseudo_b200 UNet built for %d input channel(s), got %d" % (engine.input_ch, cin))
        dev = x.device
        nconv = engine.nconv
        params = []
        for i in range(nconv):
            w, b, g, bt = learnable[4 * i:4 * i + 4]
            rm, rv, nbt = buffers[3 * i:3 * i + 3]
            params += [w, b, g, bt, rm, rv, nbt]
        params += [learnable[4 * nconv], learnable[4 * nconv + 1]]
        for t in params:
            if t is not None and (not t.is_cuda or not t.is_contiguous()):
                raise RuntimeError("UNet parameters/buffers must be contiguous CUDA tensors")
        with torch.cuda.device(dev):
            ws = torch.empty(engine.workspace_bytes(N, H, W, groups), dtype=torch.uint8, device=dev)
            logits = torch.empty((N, engine.num_classes, H, W), dtype=torch.float32, device=dev)
            lib.call("pp_unet_forward", engine.handle, ptr(x), _ptr_array(params), ptr(ws), N, H, W, groups,
                     int(training), ptr(logits), current_stream(dev))
        feats, act_ids = [], []
        for name in feat_names:
            act, off, C, h, w = engine.activation(name, N, H, W, groups)
            feats.append(_view_bytes(ws, off, (N, h, w, C), act_dtype(engine.code)))
            act_ids.append(act)
        ctx.engine, ctx.groups, ctx.training, ctx.act_ids = engine, groups, int(training), act_ids
        ctx.shape = (N, H, W)
        ctx.set_materialize_grads(False)  # unused end points must arrive as None, not as zero tensors
        ctx.buffers = buffers
        ctx.params = learnable  # the Parameter objects themselves (direct gradient accumulation, see backward)
        ctx.save_for_backward(x, ws, *learnable)
        return (logits, *feats)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_logits, *g_feats):
        engine = ctx.engine
        lib = engine.lib
        x, ws = ctx.saved_tensors[:2]
        learnable = ctx.saved_tensors[2:]
        N, H, W = ctx.shape
        dev = x.device
        nconv = engine.nconv
        params = []
        for i in range(nconv):
            params += list(learnable[4 * i:4 * i + 4]) + list(ctx.buffers[3 * i:3 * i + 3])
        params += [learnable[4 * nconv], learnable[4 * nconv + 1]]
        with torch.cuda.device(dev):
            if g_logits is None:
                g_logits = torch.zeros((N, engine.num_classes, H, W), dtype=torch.float32, device=dev)
            g_logits = g_logits.contiguous().float()
            # The kernels ACCUMULATE (+=) into the gradient buffers. Parameters whose .grad has been pre-attached by
            # pacingpseudo_b200.optim.FlatAdam (views of one flat buffer, marked _pp_direct_grad) are written in place
            # and reported to autograd as None: no per-parameter `grad += g` kernels, no temporary gradient copy.
            direct = [t is not None and getattr(t, "_pp_direct_grad", False) and t.grad is not None
                      and t.grad.is_contiguous() and t.grad.dtype == torch.float32 for t in ctx.params]
            sizes = [0 if (d or t is None) else t.numel() for d, t in zip(direct, ctx.params)]
            flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            grads, ret = [], []
            for d, t, g in zip(direct, ctx.params, flat.split(sizes)):
                if t is None:
                    grads.append(None)
                    ret.append(None)
                elif d:
                    grads.append(t.grad)
                    ret.append(None)
                else:
                    grads.append(g.view_as(t))
                    ret.append(grads[-1])
            ids, dfeat = [], []
            for act, g in zip(ctx.act_ids, g_feats):
                if g is not None:
                    ids.append(act)
                    dfeat.append(g.contiguous().to(act_dtype(engine.code)))
            id_arr = (ctypes.c_int * max(1, len(ids)))(*ids)
            lib.call("pp_unet_backward", engine.handle, ptr(x), _ptr_array(params), ptr(ws), N, H, W, ctx.groups,
                     ctx.training, ptr(g_logits), len(ids), id_arr, _ptr_array(dfeat) if dfeat else None,
                     _ptr_array(grads), current_stream(dev))
        return (None, None, None, None, None, None, *ret)


# --------------------------------------------------------------------------------------------
# Aux path (models/aux_path_memory.py:46-66): cat -> conv3x3 -> BN -> LeakyReLU -> 1x1 -> bilinear x8
# --------------------------------------------------------------------------------------------
class UpsamplePlanesFunction(torch.autograd.Function):
    """F.interpolate(x, size, mode='bilinear', align_corners=True) of fp32 NCHW planes (aux_path_memory.py:52)."""

    @staticmethod
    def forward(ctx, x, out_hw):
        require_cuda(x, "planes")
        x = x.contiguous().float()
        N, C, h, w = x.shape
        H, W = out_hw
        y = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            get_lib().call("pp_upsample_planes_fwd", ptr(x), ptr(y), N * C, h, w, H, W, current_stream(x.device))
        ctx.dims = (N, C, h, w, H, W)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        N, C, h, w, H, W = ctx.dims
        g = g.contiguous().float()
        gx = torch.empty((N, C, h, w), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            get_lib().call("pp_upsample_planes_bwd", ptr(g), ptr(gx), N * C, h, w, H, W, current_stream(g.device))
        return gx, None


def upsample_planes(x, out_hw):
    return UpsamplePlanesFunction.apply(x, tuple(int(v) for v in out_hw))


class AuxPathFunction(torch.autograd.Function):
    """(feat_a, feat_b, conv_w, conv_b, gamma, beta, fc_w) -> (logits_aux_low NCHW fp32 [N, C, h, w], aux_features NHWC).

    The logits are returned at the resolution fc_cls produces them (aux_path_memory.py:51). The reference's
    F.interpolate to the label size (aux_path_memory.py:52) is either folded into the fused scribble loss, which
    interpolates at the labelled pixels only (pp_scribble_loss_lowaux_*), or done by upsample_planes() when a caller
    asks for the full-resolution tensor.

    drop = None, or the two nn.Dropout2d layers of aux_path_memory.py:23,31 as per-(sample, channel) factors
    (s_in [N, Ca + Cb], s_hid [N, hid]; entries 0 or 1/(1-p)): s_in scales the concatenated input features,
    s_hid the bottleneck output in front of the 1x1 classifier (the returned aux_features stay un-dropped)."""

    @staticmethod
    def forward(ctx, code, buffers, training, drop, feat_a, feat_b, conv_w, conv_b, gamma, beta, fc_w):
        require_cuda(feat_a, "aux features")
        lib = get_lib()
        dev = feat_a.device
        adt = act_dtype(code)
        feat_a = feat_a.contiguous()
        feat_b = feat_b.contiguous() if feat_b is not None else None
        N, h, w, Ca = feat_a.shape
        Cb = feat_b.shape[3] if feat_b is not None else 0
        hid = conv_w.shape[0]
        C = fc_w.shape[0]
        rm, rv, nbt = buffers
        es = 2 if code == BF16 else 4
        with torch.cuda.device(dev):
            st = current_stream(dev)
            wf = torch.empty(9 * hid * (Ca + Cb) * es, dtype=torch.uint8, device=dev)
            wd = torch.empty_like(wf)
            lib.call("pp_pack_weights", code, ptr(conv_w), ptr(wf), ptr(wd), hid, Ca + Cb, st)
            s_in = s_hid = None
            if drop is not None:
                s_in, s_hid = (t.contiguous().float() for t in drop)
                fa_d = torch.empty_like(feat_a)
                lib.call("pp_channel_scale", code, ptr(feat_a), ptr(s_in), ptr(fa_d), N, h * w, Ca, Ca + Cb, st)
                feat_a = fa_d
                if feat_b is not None:
                    fb_d = torch.empty_like(feat_b)
                    lib.call("pp_channel_scale", code, ptr(feat_b), ctypes.c_void_p(s_in.data_ptr() + 4 * Ca), ptr(fb_d),
                             N, h * w, Cb, Ca + Cb, st)
                    feat_b = fb_d
            yraw = torch.empty((N, h, w, hid), dtype=adt, device=dev)
            lib.call("pp_conv3x3", code, ptr(feat_a), Ca, ptr(feat_b), Cb, ptr(wf), ptr(conv_b), ptr(yraw), hid, 0,
                     None, 0, 0, N, h, w, 1, st)
            sums = torch.zeros(2 * hid, dtype=torch.float64, device=dev)
            coef = torch.empty(4 * hid, dtype=torch.float32, device=dev)
            Pg = N * h * w
            if training:
                lib.call("pp_bn_stats", code, ptr(yraw), ptr(sums), 1, Pg, hid, st)
            lib.call("pp_bn_finalize", ptr(sums), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), ptr(nbt), ptr(coef), 1, Pg,
                     hid, int(training), 1e-5, 0.1, st)
            act = torch.empty_like(yraw)
            lib.call("pp_bn_apply", code, ptr(yraw), ptr(coef), ptr(act), 1, Pg, hid, 0.01, st)
            low = torch.empty((N, C, h, w), dtype=torch.float32, device=dev)
            act_in = act
            if s_hid is not None:
                act_in = torch.empty_like(act)
                lib.call("pp_channel_scale", code, ptr(act), ptr(s_hid), ptr(act_in), N, h * w, hid, hid, st)
            lib.call("pp_head_fwd", code, ptr(act_in), ptr(fc_w), None, ptr(low), Pg, h * w, hid, C, st)
        ctx.code, ctx.training, ctx.dims = code, int(training), (N, h, w, Ca, Cb, hid, C)
        ctx.save_for_backward(feat_a, feat_b, wd, yraw, coef, act_in, fc_w, s_in, s_hid)
        ctx.mark_non_differentiable(act)
        return low, act

    @staticmethod
    @once_differentiable
    def backward(ctx, g_low, _g_act):
        lib = get_lib()
        feat_a, feat_b, wd, yraw, coef, act, fc_w, s_in, s_hid = ctx.saved_tensors   # feat_*/act: after dropout
        code = ctx.code
        N, h, w, Ca, Cb, hid, C = ctx.dims
        dev = feat_a.device
        adt = act_dtype(code)
        Pg = N * h * w
        with torch.cuda.device(dev):
            st = current_stream(dev)
            g_low = g_low.contiguous().float()
            d_act = torch.empty((N, h, w, hid), dtype=adt, device=dev)
            d_fc = torch.zeros_like(fc_w)
            lib.call("pp_head_bwd", code, ptr(g_low), ptr(act), ptr(fc_w), ptr(d_act), ptr(d_fc), None, Pg, h * w, hid,
                     C, st)
            if s_hid is not None:
                lib.call("pp_channel_scale", code, ptr(d_act), ptr(s_hid), ptr(d_act), N, h * w, hid, hid, st)
            bsums = torch.empty(2 * hid, dtype=torch.float64, device=dev)
            bcoef = torch.empty(2 * hid, dtype=torch.float32, device=dev)
            d_gamma = torch.zeros(hid, dtype=torch.float32, device=dev)
            d_beta = torch.zeros_like(d_gamma)
            d_bias = torch.zeros_like(d_gamma)
            dy = torch.empty_like(d_act)
            lib.call("pp_bn_bwd", code, ptr(d_act), ptr(yraw), ptr(coef), ptr(bsums), ptr(bcoef), ptr(d_gamma),
                     ptr(d_beta), ptr(d_bias), ptr(dy), 1, Pg, hid, ctx.training, 0.01, st)
            dwp = torch.zeros(9 * hid * (Ca + Cb), dtype=torch.float32, device=dev)
            lib.call("pp_conv3x3_wgrad", code, ptr(dy), hid, ptr(feat_a), Ca, ptr(feat_b), Cb, ptr(dwp), N, h, w, 1, st)
            d_w = torch.empty((hid, Ca + Cb, 3, 3), dtype=torch.float32, device=dev)
            lib.call("pp_unpack_wgrad", ptr(dwp), ptr(d_w), hid, Ca + Cb, 0, st)
            g_a = torch.empty_like(feat_a)
            g_b = torch.empty_like(feat_b) if feat_b is not None else None
            lib.call("pp_conv3x3", code, ptr(dy), hid, None, 0, ptr(wd), None, ptr(g_a), Ca, 0, ptr(g_b), Cb, 0, N, h, w,
                     1, st)
            if s_in is not None:
                lib.call("pp_channel_scale", code, ptr(g_a), ptr(s_in), ptr(g_a), N, h * w, Ca, Ca + Cb, st)
                if g_b is not None:
                    lib.call("pp_channel_scale", code, ptr(g_b), ctypes.c_void_p(s_in.data_ptr() + 4 * Ca), ptr(g_b), N,
                             h * w, Cb, Ca + Cb, st)
        return None, None, None, None, g_a, g_b, d_w, d_bias, d_gamma, d_beta, d_fc


# --------------------------------------------------------------------------------------------
# Losses
# --------------------------------------------------------------------------------------------
def onehot_argmax(x):
    """torch.argmax(x, 1) of an fp32 NCHW one-hot tensor as a uint8 index map (N, H, W)."""
    require_cuda(x, "one-hot tensor")
    x = x.contiguous().float()
    N, K, H, W = x.shape
    out = torch.empty((N, H, W), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        get_lib().call("pp_onehot_argmax", ptr(x), ptr(out), N, K, H * W, current_stream(x.device))
    return out


class ScribbleLossFunction(torch.autograd.Function):
    """One fused pass for partial CE + entropy + consistency + aux partial CE.

    forward(cfg, zw, zs, za, target_u8, mask) -> (loss_pce, loss_ent, loss_cr, loss_aux), each an
    independent 0-dim fp32 tensor (the caller mutates them in place, train_chaos.py:274-309).
    cfg = (ignore_index, do_ent, cr_variant, detach_weak). If `zs` is None and cr_variant != none the
    strong logits are the second half of `zw` (batched siamese tensor). `za` is either the full-resolution aux
    logits (N, C, H, W) or, when its spatial size differs from the label map's, the low-resolution tensor
    (N, C, h, w) of aux_path_memory.py:51: the kernels then interpolate it at the labelled pixels themselves
    (bilinear, align_corners=True, aux_path_memory.py:52) and the gradient comes back at (N, C, h, w).
    """

    @staticmethod
    def forward(ctx, cfg, zw, zs, za, target, mask):
        ignore_index, do_ent, cr_variant, detach_weak, siamese = cfg
        require_cuda(zw, "logits")
        lib = get_lib()
        dev = zw.device
        zw = zw.contiguous().float()
        zs = zs.contiguous().float() if zs is not None else None
        za = za.contiguous().float() if za is not None else None
        mask = mask.contiguous().float() if mask is not None else None
        Nall, C, H, W = zw.shape
        bad_any = None
        if target is not None:
            target = target.contiguous()
            if target.dtype != torch.uint8:
                # int64 targets of the reference API (F.cross_entropy semantics): the ignore label may be any integer
                # (torch's default is -100). It is mapped to the sentinel 255 BEFORE the narrowing cast, so a negative
                # or > 255 ignore label can never alias a class id. Labels outside [0, C) that are not the ignore
                # label are a device-side assert in torch; here they poison the loss with NaN (no host sync on the
                # training path, but never a silently skipped pixel).
                if C >= 255:
                    raise RuntimeError("scribble loss: at most 254 classes (255 is the ignore sentinel)")
                ign = target == ignore_index
                bad_any = (((target < 0) | (target >= C)) & ~ign).any()
                target = torch.where(ign, torch.full_like(target, 255), target).clamp(0, 255).to(torch.uint8)
                ignore_index = 255
            elif not 0 <= ignore_index <= 255:
                ignore_index = 255   # a uint8 map cannot hold this label: nothing is ignored unless it says 255
        N = Nall // 2 if siamese else Nall
        zw_p = zw.data_ptr()
        zs_p = (zw_p + N * C * H * W * 4) if siamese else (zs.data_ptr() if zs is not None else None)
        if target is not None and target.shape[0] != N:
            raise RuntimeError("scribble loss: target batch %d != logits batch %d" % (target.shape[0], N))
        aux_low = za is not None and tuple(za.shape[-2:]) != (H, W)
        if za is not None and (za.dim() != 4 or za.shape[0] != N or za.shape[1] != C):
            raise RuntimeError("scribble loss: aux logits %s do not match %d samples of %d classes" % (
                tuple(za.shape), N, C))
        with torch.cuda.device(dev):
            acc = torch.empty(8, dtype=torch.float64, device=dev)
            outs = [torch.zeros((), dtype=torch.float32, device=dev) for _ in range(4)]
            if aux_low:
                lib.call("pp_scribble_loss_lowaux_fwd", ctypes.c_void_p(zw_p), ctypes.c_void_p(zs_p) if zs_p else None,
                         ptr(za), za.shape[2], za.shape[3], ptr(target), ptr(mask), ptr(acc), ptr(outs[0]), ptr(outs[1]),
                         ptr(outs[2]), ptr(outs[3]), N, C, H, W, ignore_index, int(do_ent), cr_variant,
                         current_stream(dev))
            else:
                lib.call("pp_scribble_loss_fwd", ctypes.c_void_p(zw_p), ctypes.c_void_p(zs_p) if zs_p else None, ptr(za),
                         ptr(target), ptr(mask), ptr(acc), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), N, C,
                         H * W, ignore_index, int(do_ent), cr_variant, current_stream(dev))
            if bad_any is not None:
                outs[0].masked_fill_(bad_any, float("nan"))
        cfg = (ignore_index, do_ent, cr_variant, detach_weak, siamese)   # the (possibly remapped) ignore label
        ctx.cfg, ctx.dims = cfg, (N, C, H, W)
        ctx.has = (zs is not None, za is not None, mask is not None)
        ctx.aux_low = aux_low
        ctx.save_for_backward(zw, zs, za, target, mask, acc)
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, g_pce, g_ent, g_cr, g_aux):
        ignore_index, do_ent, cr_variant, detach_weak, siamese = ctx.cfg
        lib = get_lib()
        zw, zs, za, target, mask, acc = ctx.saved_tensors
        N, C, H, W = ctx.dims
        dev = zw.device
        with torch.cuda.device(dev):
            def scal(g):
                return None if g is None else g.contiguous().float()
            g_pce, g_ent, g_cr, g_aux = scal(g_pce), scal(g_ent), scal(g_cr), scal(g_aux)
            dzw = torch.empty_like(zw)
            dzs = torch.empty_like(zs) if zs is not None else None
            dza = torch.empty_like(za) if za is not None else None
            zw_p = zw.data_ptr()
            half = N * C * H * W * 4
            zs_p = (zw_p + half) if siamese else (zs.data_ptr() if zs is not None else None)
            dzs_p = (dzw.data_ptr() + half) if siamese else (dzs.data_ptr() if dzs is not None else None)
            if ctx.aux_low:   # accumulated in a fixed-point scratch with integer atomics (order-independent), then -> dza
                scratch = torch.empty(za.numel(), dtype=torch.int64, device=dev)
                lib.call("pp_scribble_loss_lowaux_bwd", ctypes.c_void_p(zw_p), ctypes.c_void_p(zs_p) if zs_p else None,
                         ptr(za), za.shape[2], za.shape[3], ptr(target), ptr(mask), ptr(acc), ptr(g_pce), ptr(g_ent),
                         ptr(g_cr), ptr(g_aux), ctypes.c_void_p(dzw.data_ptr()), ctypes.c_void_p(dzs_p) if dzs_p else None,
                         ptr(dza), ptr(scratch), N, C, H, W, ignore_index, int(do_ent), cr_variant, int(detach_weak),
                         current_stream(dev))
            else:
                lib.call("pp_scribble_loss_bwd", ctypes.c_void_p(zw_p), ctypes.c_void_p(zs_p) if zs_p else None, ptr(za),
                         ptr(target), ptr(mask), ptr(acc), ptr(g_pce), ptr(g_ent), ptr(g_cr), ptr(g_aux),
                         ctypes.c_void_p(dzw.data_ptr()), ctypes.c_void_p(dzs_p) if dzs_p else None, ptr(dza), N, C,
                         H * W, ignore_index, int(do_ent), cr_variant, int(detach_weak), current_stream(dev))
        return None, dzw, dzs, dza, None, None


def scribble_losses(zw, target, ignore_index, zs=None, za=None, mask=None, do_ent=False, cr_variant=None,
                    detach_weak=False, siamese=False):
    """Fused losses; returns dict with loss_pce / loss_ent / loss_cr / loss_aux (None when not requested)."""
    var = CR_VARIANTS[cr_variant] if not isinstance(cr_variant, int) else cr_variant
    cfg = (int(ignore_index), bool(do_ent), var, bool(detach_weak), bool(siamese))
    pce, ent, cr, aux = ScribbleLossFunction.apply(cfg, zw, zs, za, target, mask)
    return {"loss_pce": pce, "loss_ent": ent if do_ent else None, "loss_cr": cr if var else None,
            "loss_aux": aux if za is not None else None}


class PairLossFunction(torch.autograd.Function):
    """soft-label CE (a = logits, b = probabilities) / L1 / L2 (a, b = probabilities), masked mean."""

    @staticmethod
    def forward(ctx, variant, a, b, mask):
        require_cuda(a, "loss input")
        a, b = a.contiguous().float(), b.contiguous().float()
        mask = mask.contiguous().float() if mask is not None else None
        N, C, H, W = a.shape
        dev = a.device
        with torch.cuda.device(dev):
            pacc = torch.empty(8, dtype=torch.float64, device=dev)
            loss = torch.zeros((), dtype=torch.float32, device=dev)
            get_lib().call("pp_pair_loss_fwd", ptr(a), ptr(b), ptr(mask), ptr(pacc), ptr(loss), N, C, H * W, variant,
                           current_stream(dev))
        ctx.variant = variant
        ctx.save_for_backward(a, b, mask, pacc)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        a, b, mask, pacc = ctx.saved_tensors
        N, C, H, W = a.shape
        dev = a.device
        with torch.cuda.device(dev):
            da = torch.empty_like(a) if ctx.needs_input_grad[1] else None
            db = torch.empty_like(b) if ctx.needs_input_grad[2] else None
            get_lib().call("pp_pair_loss_bwd", ptr(a), ptr(b), ptr(mask), ptr(pacc), ptr(g.contiguous().float()),
                           ptr(da), ptr(db), N, C, H * W, ctx.variant, current_stream(dev))
        return None, da, db, None


def pair_loss(a, b, mask, variant):
    return PairLossFunction.apply(CR_VARIANTS[variant], a, b, mask)


class DiceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, label):
        require_cuda(z, "logits")
        lib = get_lib()
        z = z.contiguous().float()
        label = label.contiguous().float()
        N, C, H, W = z.shape
        dev = z.device
        with torch.cuda.device(dev):
            sums = torch.empty(3 * N * C, dtype=torch.float64, device=dev)
            coef = torch.empty(2 * N * C, dtype=torch.float32, device=dev)
            loss = torch.zeros((), dtype=torch.float32, device=dev)
            lib.call("pp_dice_fwd", ptr(z), ptr(label), ptr(sums), ptr(coef), ptr(loss), N, C, H * W,
                     current_stream(dev))
        ctx.save_for_backward(z, label, coef)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        z, label, coef = ctx.saved_tensors
        N, C, H, W = z.shape
        dev = z.device
        with torch.cuda.device(dev):
            dz = torch.empty_like(z)
            get_lib().call("pp_dice_bwd", ptr(z), ptr(label), ptr(coef), ptr(g.contiguous().float()), ptr(dz), N, C,
                           H * W, 0, current_stream(dev))
        return dz, None


class MemoryLossFunction(torch.autograd.Function):
    """cross_entropy(fc_cls(memory_bank), arange(C)); gradient flows to fc_cls.weight only.
    s_bank: optional [C, hid] Dropout2d factors (fc_cls[0] also acts on the bank, aux_path_memory.py:31,60)."""

    @staticmethod
    def forward(ctx, bank, fc_w, s_bank=None):
        require_cuda(fc_w, "fc_cls weight")
        C, hid = bank.shape[0], bank.shape[1]
        dev = fc_w.device
        bank_c = bank.detach().contiguous().float().clone()  # the bank is mutated in place by later steps
        with torch.cuda.device(dev):
            if s_bank is not None:
                get_lib().call("pp_channel_scale", F32, ptr(bank_c), ptr(s_bank.contiguous().float()), ptr(bank_c), C, 1,
                               hid, hid, current_stream(dev))
            loss = torch.zeros((), dtype=torch.float32, device=dev)
            probs = torch.empty(C * C, dtype=torch.float32, device=dev)
            get_lib().call("pp_memory_loss_fwd", ptr(bank_c), ptr(fc_w.contiguous()), ptr(loss), ptr(probs), C, hid,
                           current_stream(dev))
        ctx.save_for_backward(bank_c, probs, fc_w)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        bank_c, probs, fc_w = ctx.saved_tensors
        C, hid = bank_c.shape[0], bank_c.shape[1]
        dev = fc_w.device
        with torch.cuda.device(dev):
            d = torch.zeros_like(fc_w)
            get_lib().call("pp_memory_loss_bwd", ptr(bank_c), ptr(probs), ptr(g.contiguous().float()), ptr(d), C, hid,
                           current_stream(dev))
        return None, d, None


def bank_logits(bank, fc_w):
    """fc_cls applied to the memory bank (aux_path_memory.py:60): (C, hid, 1, 1) x (K, hid, 1, 1) -> (C, K, 1, 1),
    through the 1x1 head kernel (the bank rows are C "pixels" of hid channels). No gradient: the bank loss and its
    gradient go through MemoryLossFunction."""
    require_cuda(fc_w, "fc_cls weight")
    C, hid = bank.shape[0], bank.shape[1]
    K = fc_w.shape[0]
    dev = fc_w.device
    with torch.cuda.device(dev):
        out = torch.empty((C, K, 1, 1), dtype=torch.float32, device=dev)
        get_lib().call("pp_head_fwd", F32, ptr(bank.detach().contiguous().float().view(C, hid)),
                       ptr(fc_w.detach().contiguous().float()), None, ptr(out), C, 1, hid, K, current_stream(dev))
    return out


def memory_update(code, aux_features, scribble, bank, mode, m):
    """In-place bank update from sample 0 (aux_path_memory.py:68-116). aux_features: native NHWC.
    scribble: fp32 one-hot (N, C+1, H, W) (reference format) or a uint8 class-index map (N, H, W)."""
    lib = get_lib()
    N, h, w, hid = aux_features.shape
    H, W = scribble.shape[-2:]
    C = bank.shape[0]
    index_map = scribble.dim() == 3
    scribble = scribble.contiguous().to(torch.uint8) if index_map else scribble.contiguous().float()
    with torch.cuda.device(bank.device):
        scratch = torch.empty(lib.cdll.pp_memory_update_scratch_floats(C, hid), dtype=torch.float32, device=bank.device)
        lib.call("pp_memory_update_idx" if index_map else "pp_memory_update", code, ptr(aux_features), ptr(scribble),
                 ptr(bank), ptr(scratch), C, h, w, H, W, hid, 1 if mode == "cosine_similarity" else 0, float(m),
                 float(1.0 - m), current_stream(bank.device))
