"""GPU (-m gpu): parity of the hand-written sm_100a path, called through the C ABI, against the CPU oracle
(oracle/pp_oracle.py), the reference's golden vectors (tests/golden/) and size-independent properties at
BASELINE.json's full sizes. Tolerances are the north-star's: bf16 logits/losses 2e-2, gradients 5e-2,
argmax agreement >= 99.9 %; fp32 mode logits/losses 1e-4."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import harness as Hn
from oracle import pp_oracle as O
from oracle.gen_golden import CASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pp():
    from pacingpseudo_b200 import lib as pplib
    from pacingpseudo_b200 import functional as PF
    L = pplib.get_lib()
    L.ensure_init(0)
    torch.cuda.set_device(0)
    return L, PF, pplib


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _nhwc(t, dtype):
    return t.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


# ------------------------------------------------------------------------------------------------
# operator level
# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # N, H, W, C0, C1, Cout, dil
    (2, 32, 32, 64, 0, 64, 1), (2, 32, 32, 32, 0, 32, 1), (2, 16, 16, 256, 0, 512, 2), (2, 8, 8, 512, 512, 512, 1),
    (3, 16, 16, 64, 32, 32, 1), (1, 64, 64, 128, 64, 64, 1), (2, 8, 8, 512, 0, 512, 4), (1, 28, 28, 64, 0, 128, 1),
    (2, 8, 8, 512, 512, 64, 1),
    # wide rows: smem-resident (halo) forward / dgrad and the row variant of the narrow weight gradient
    (1, 8, 128, 32, 0, 32, 1), (1, 6, 192, 64, 32, 32, 1), (2, 4, 64, 64, 0, 64, 1), (1, 5, 128, 32, 0, 64, 1),
    (1, 7, 112, 64, 32, 32, 1), (1, 5, 224, 32, 0, 32, 1),   # ragged last 64-pixel block (ACDC / LVSC crops)
]


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv3x3_forward_dgrad_wgrad_vs_oracle(pp, case, precision):
    """nn.Conv2d forward / input gradient / weight gradient (unet.py:188) against torch CPU autograd."""
    L, PF, pplib = pp
    N, H, W, C0, C1, Co, dil = case
    code = PF.dtype_code(precision)
    adt = PF.act_dtype(code)
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(N, C0 + C1, H, W, generator=g)
    w = torch.randn(Co, C0 + C1, 3, 3, generator=g) / (3 * (C0 + C1) ** 0.5)
    b = torch.randn(Co, generator=g)
    gy = torch.randn(N, Co, H, W, generator=g)
    if precision == "bf16":  # the oracle sees the same bf16-rounded operands the tensor cores see
        x, w, gy = x.bfloat16().float(), w.bfloat16().float(), gy.bfloat16().float()
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    y_ref = F.conv2d(xr, wr, b, 1, dil, dil)
    y_ref.backward(gy)

    es = 2 if precision == "bf16" else 4
    wf = torch.empty(9 * Co * (C0 + C1) * es, dtype=torch.uint8, device="cuda")
    wd = torch.empty_like(wf)
    w_d, b_d = w.cuda(), b.cuda()
    L.call("pp_pack_weights", code, _p(w_d), _p(wf), _p(wd), Co, C0 + C1, _st())
    x0 = _nhwc(x[:, :C0], adt)
    x1 = _nhwc(x[:, C0:], adt) if C1 else None
    y = torch.empty(N, H, W, Co, dtype=adt, device="cuda")
    L.call("pp_conv3x3", code, _p(x0), C0, _p(x1), C1, _p(wf), _p(b_d), _p(y), Co, 0, None, 0, 0, N, H, W, dil, _st())
    tol = 1e-2 if precision == "bf16" else 1e-5   # bf16: output rounding only (inputs are shared)
    assert _rel(y.float().permute(0, 3, 1, 2), y_ref.detach()) < tol

    dy = _nhwc(gy, adt)
    g0 = torch.randn(N, H, W, C0, device="cuda").to(adt)   # accumulate into source 0, overwrite source 1
    g0_init = g0.float().clone()
    g1 = torch.empty(N, H, W, C1, dtype=adt, device="cuda") if C1 else None
    L.call("pp_conv3x3", code, _p(dy), Co, None, 0, _p(wd), None, _p(g0), C0, 1, _p(g1), C1, 0, N, H, W, dil, _st())
    gx = xr.grad
    assert _rel(g0.float().permute(0, 3, 1, 2) - g0_init.permute(0, 3, 1, 2), gx[:, :C0]) < (3e-2 if precision == "bf16" else 1e-5)
    if C1:
        assert _rel(g1.float().permute(0, 3, 1, 2), gx[:, C0:]) < tol

    dwp = torch.zeros(9 * Co * (C0 + C1), dtype=torch.float32, device="cuda")
    L.call("pp_conv3x3_wgrad", code, _p(dy), Co, _p(x0), C0, _p(x1), C1, _p(dwp), N, H, W, dil, _st())
    gw = torch.empty(Co, C0 + C1, 3, 3, device="cuda")
    L.call("pp_unpack_wgrad", _p(dwp), _p(gw), Co, C0 + C1, 0, _st())
    assert _rel(gw, wr.grad) < 1e-4

    if precision == "bf16":  # tcgen05 kernel vs its CUDA-core twin on identical inputs
        y2 = torch.empty_like(y)
        L.call("pp_conv3x3_reference", code, _p(x0), C0, _p(x1), C1, _p(wf), _p(b_d), _p(y2), Co, 0, None, 0, 0, N,
               H, W, dil, _st())
        assert _rel(y.float(), y2.float()) < 2e-3


# The layer shapes bench.py actually runs (24 slices = 12 weak + 12 strong, or one 12-slice branch): full tile counts,
# so the MT = 2 / 4 multi-tile CTAs, the 256-row deterministic split-K weight gradient at full K, the halo kernels'
# long row rings and the row variant of the narrow weight gradient are all checked at their real grid sizes
# (operator tests only covered M <= 4096 pixels).
FULL_TILE_CASES = [
    # N, H, W, C0, C1, Cout, dil
    (24, 32, 32, 512, 0, 512, 4),     # enc_block6: M = 24*1024, 512 -> 512, dilation 4
    (24, 32, 32, 512, 512, 512, 1),   # dec_block5 conv1: concat 1024 -> 512
    (12, 32, 32, 512, 256, 256, 1),   # dec_block4 conv1 (one branch): 768 -> 256 (BLOCK_N 192 dgrad)
    (12, 64, 64, 256, 128, 128, 1),   # dec_block3 conv1
    (12, 128, 128, 128, 64, 64, 1),   # dec_block2 conv1 (BLOCK_N 64, MT 2)
    (24, 256, 256, 32, 0, 32, 1),     # enc_block1 conv2 / dec_block1 conv2: M = 24*65536, halo kernel + row wgrad
    (12, 256, 256, 64, 32, 32, 1),    # dec_block1 conv1: concat 96 -> 32 at full resolution
    (12, 28, 28, 512, 0, 512, 2),     # ACDC / LVSC 224^2 bottleneck: ragged 28 x 28 tiles, dilation 2
]


@pytest.mark.parametrize("case", FULL_TILE_CASES)
def test_conv3x3_full_tile_counts_vs_torch(pp, case):
    """tcgen05 forward / dgrad / wgrad at the bench's real launch shapes against torch CPU autograd on the same
    bf16-rounded operands. Tolerances: forward / dgrad 3e-3 rel-L2 (one bf16 rounding of the fp32 accumulator is
    ~1.2e-3), weight gradient 5e-4 (fp32 output; only the summation order differs)."""
    L, PF, pplib = pp
    N, H, W, C0, C1, Co, dil = case
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, C0 + C1, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Co, C0 + C1, 3, 3, generator=g) / (3 * (C0 + C1) ** 0.5)).bfloat16().float()
    b = torch.randn(Co, generator=g)
    gy = torch.randn(N, Co, H, W, generator=g).bfloat16().float()
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    y_ref = F.conv2d(xr, wr, b, 1, dil, dil)
    y_ref.backward(gy)
    code, adt = PF.BF16, torch.bfloat16
    wf = torch.empty(9 * Co * (C0 + C1) * 2, dtype=torch.uint8, device="cuda")
    wd = torch.empty_like(wf)
    w_d, b_d = w.cuda(), b.cuda()
    L.call("pp_pack_weights", code, _p(w_d), _p(wf), _p(wd), Co, C0 + C1, _st())
    x0 = _nhwc(x[:, :C0], adt)
    x1 = _nhwc(x[:, C0:], adt) if C1 else None
    y = torch.empty(N, H, W, Co, dtype=adt, device="cuda")
    L.call("pp_conv3x3", code, _p(x0), C0, _p(x1), C1, _p(wf), _p(b_d), _p(y), Co, 0, None, 0, 0, N, H, W, dil, _st())
    e_fwd = _rel(y.float().permute(0, 3, 1, 2), y_ref.detach())
    dy = _nhwc(gy, adt)
    g0 = torch.empty(N, H, W, C0, dtype=adt, device="cuda")
    g1 = torch.empty(N, H, W, C1, dtype=adt, device="cuda") if C1 else None
    L.call("pp_conv3x3", code, _p(dy), Co, None, 0, _p(wd), None, _p(g0), C0, 0, _p(g1), C1, 0, N, H, W, dil, _st())
    gx = xr.grad
    e_d0 = _rel(g0.float().permute(0, 3, 1, 2), gx[:, :C0])
    e_d1 = _rel(g1.float().permute(0, 3, 1, 2), gx[:, C0:]) if C1 else 0.0
    # the training path: OIHW gradient accumulated in place, deterministic split-K scratch for the wide sources
    gw = torch.zeros(Co, C0 + C1, 3, 3, device="cuda")
    dwp = torch.empty(9 * Co * (C0 + C1), dtype=torch.float32, device="cuda")
    wss = torch.empty(2 * 9 * Co * (C0 + C1), dtype=torch.float32, device="cuda")
    L.call("pp_conv3x3_wgrad_oihw", _p(dy), Co, _p(x0), C0, _p(x1), C1, _p(dwp), _p(gw), _p(wss), wss.numel(), N, H, W,
           dil, _st())
    e_w = _rel(gw, wr.grad)
    # the forward variant the training step launches: BatchNorm batch statistics in the epilogue, two groups
    reps = L.cdll.pp_stat_replicas()
    G = 2
    stats = torch.zeros(reps * G * Co * 2, dtype=torch.float64, device="cuda")
    y2 = torch.empty_like(y)
    L.call("pp_conv3x3_bn_stats", _p(x0), C0, _p(x1), C1, _p(wf), _p(b_d), _p(y2), Co, _p(stats), G, N, H, W, dil, _st())
    # (the two calls are tuned separately and may run different kernels: same values up to the fp32 summation order,
    # i.e. a bf16 ulp on a few elements; the statistics must match the output of THEIR launch)
    assert _rel(y2.float(), y.float()) < 1e-3
    st = stats.view(reps, G, Co, 2).sum(0).cpu()
    yq = y2.float().cpu().view(G, -1, Co).double()
    e_s = max(_rel(st[:, :, 0], yq.sum(1)), _rel(st[:, :, 1], (yq * yq).sum(1)))
    print("full-tile case %s: fwd %.2e dgrad %.2e / %.2e wgrad %.2e stats %.2e" % (case, e_fwd, e_d0, e_d1, e_w, e_s))
    assert e_fwd < 3e-3 and e_d0 < 3e-3 and e_d1 < 3e-3 and e_w < 5e-4 and e_s < 1e-5, (case, e_fwd, e_d0, e_d1, e_w, e_s)


@pytest.mark.parametrize("case", [(2, 32, 32, 64, 0, 128, 1), (2, 16, 16, 256, 256, 256, 2), (1, 8, 256, 64, 32, 32, 1),
                                  (2, 16, 128, 32, 0, 64, 1), (3, 28, 28, 128, 0, 512, 4)])
def test_conv_bn_eval_fused_forward_backward_vs_torch(pp, case):
    """Eval-mode BatchNorm + LeakyReLU folded into the conv epilogue (unet.py:188-190 with running statistics), and its
    one-pass backward from the saved activation, against torch CPU autograd of conv2d -> batch_norm(eval) ->
    leaky_relu on the same bf16-rounded operands (generic and halo kernels, concat sources, dilation)."""
    L, PF, pplib = pp
    N, H, W, C0, C1, Co, dil = case
    g = torch.Generator().manual_seed(sum(case) + 5)
    x = torch.randn(N, C0 + C1, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Co, C0 + C1, 3, 3, generator=g) / (3 * (C0 + C1) ** 0.5)).bfloat16().float()
    b = 0.1 * torch.randn(Co, generator=g)
    gamma = 1.0 + 0.3 * torch.randn(Co, generator=g)
    beta = 0.2 * torch.randn(Co, generator=g)
    rm = 0.1 * torch.randn(Co, generator=g)
    rv = 0.5 + torch.rand(Co, generator=g)
    ga = torch.randn(N, Co, H, W, generator=g).bfloat16().float()
    gam_r, bet_r, b_r = gamma.clone().requires_grad_(), beta.clone().requires_grad_(), b.clone().requires_grad_()
    y = F.conv2d(x, w, b_r, 1, dil, dil)
    a_ref = F.leaky_relu(F.batch_norm(y, rm, rv, gam_r, bet_r, False, 0.1, 1e-5), 0.01)
    y.retain_grad()
    a_ref.backward(ga)
    code, adt = PF.BF16, torch.bfloat16
    wf = torch.empty(9 * Co * (C0 + C1) * 2, dtype=torch.uint8, device="cuda")
    wd = torch.empty_like(wf)
    w_d = w.cuda()
    L.call("pp_pack_weights", code, _p(w_d), _p(wf), _p(wd), Co, C0 + C1, _st())
    coef = torch.empty(4 * Co, device="cuda")
    g_d, be_d, rm_d, rv_d, b_d = gamma.cuda(), beta.cuda(), rm.cuda(), rv.cuda(), b.cuda()   # keep the buffers alive
    L.call("pp_bn_eval_coef", _p(g_d), _p(be_d), _p(rm_d), _p(rv_d), _p(b_d), _p(coef), Co, 1e-5, _st())
    x0 = _nhwc(x[:, :C0], adt)
    x1 = _nhwc(x[:, C0:], adt) if C1 else None
    a = torch.empty(N, H, W, Co, dtype=adt, device="cuda")
    L.call("pp_conv3x3_bn_eval", _p(x0), C0, _p(x1), C1, _p(wf), _p(coef), _p(a), Co, 0.01, N, H, W, dil, _st())
    assert _rel(a.float().permute(0, 3, 1, 2), a_ref.detach()) < 3e-3
    # backward from the activation the oracle produced (rounded to bf16), so both sides see the same LeakyReLU mask
    a_in = _nhwc(a_ref.detach(), adt)
    da = _nhwc(ga, adt)
    sums = torch.zeros(2 * Co + 2, dtype=torch.float64, device="cuda")
    dgam, dbet, dbias = (torch.zeros(Co, device="cuda") for _ in range(3))
    dy = torch.empty_like(da)
    L.call("pp_bn_bwd_eval", code, _p(da), _p(a_in), _p(coef), _p(sums), _p(dgam), _p(dbet), _p(dbias), _p(dy),
           N * H * W, Co, 0.01, _st())
    assert _rel(dy.float().permute(0, 3, 1, 2), y.grad) < 3e-3
    assert _rel(dbet, bet_r.grad) < 1e-4 and _rel(dbias, b_r.grad) < 1e-4
    assert _rel(dgam, gam_r.grad) < 5e-3    # xhat is recovered from the bf16 activation


@pytest.mark.parametrize("case", [(2, 16, 16, 64, 0, 128, 1), (1, 16, 128, 32, 0, 32, 1), (1, 8, 256, 64, 32, 32, 1)])
def test_conv3x3_bias_pointer_alignment(pp, case):
    """The C ABI accepts any 4-byte aligned bias pointer (parameters re-homed as views of a flat buffer need not be
    16-byte aligned): generic and smem-resident (halo) kernels, aligned vs. misaligned pointer, bit-identical."""
    L, PF, _ = pp
    N, H, W, C0, C1, Co, dil = case
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(N, H, W, C0, generator=g).bfloat16().cuda()
    x1 = torch.randn(N, H, W, C1, generator=g).bfloat16().cuda() if C1 else None
    w = (torch.randn(Co, C0 + C1, 3, 3, generator=g) / (3 * (C0 + C1) ** 0.5)).cuda()
    b = torch.randn(Co, generator=g)
    wf = torch.empty(9 * Co * (C0 + C1) * 2, dtype=torch.uint8, device="cuda")
    wd = torch.empty_like(wf)
    L.call("pp_pack_weights", PF.BF16, _p(w), _p(wf), _p(wd), Co, C0 + C1, _st())
    outs = []
    for shift in (0, 1, 2, 3):
        buf = torch.zeros(Co + 8, device="cuda")
        assert buf.data_ptr() % 16 == 0
        bias = buf[shift:shift + Co]
        bias.copy_(b)
        y = torch.empty(N, H, W, Co, dtype=torch.bfloat16, device="cuda")
        L.call("pp_conv3x3", PF.BF16, _p(x0), C0, _p(x1), C1, _p(wf), _p(bias), _p(y), Co, 0, None, 0, 0, N, H, W, dil,
               _st())
        torch.cuda.synchronize()
        outs.append(y)
    for y in outs[1:]:
        assert torch.equal(outs[0], y)


@pytest.mark.parametrize("case", [(4, 32, 32, 64, 0, 128, 2), (2, 128, 128, 32, 0, 32, 1), (2, 64, 64, 128, 64, 64, 2),
                                  (4, 8, 8, 64, 0, 64, 2)])
def test_conv3x3_fused_batchnorm_statistics(pp, case):
    """The conv epilogue's per-group sum / sum-of-squares equal a separate pass over the stored (rounded) output,
    and two runs are bit-identical (the in-CTA reduction order is fixed)."""
    L, PF, _ = pp
    N, H, W, C0, C1, Co, G = case
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(N, H, W, C0, generator=g).bfloat16().cuda()
    x1 = torch.randn(N, H, W, C1, generator=g).bfloat16().cuda() if C1 else None
    w = (torch.randn(Co, C0 + C1, 3, 3, generator=g) / (3 * (C0 + C1) ** 0.5)).cuda()
    b = torch.randn(Co, generator=g).cuda()
    wf = torch.empty(9 * Co * (C0 + C1) * 2, dtype=torch.uint8, device="cuda")
    wd = torch.empty_like(wf)
    L.call("pp_pack_weights", PF.BF16, _p(w), _p(wf), _p(wd), Co, C0 + C1, _st())
    R = L.cdll.pp_stat_replicas()
    stats = torch.zeros(R, G, Co, 2, dtype=torch.float64, device="cuda")
    y = torch.empty(N, H, W, Co, dtype=torch.bfloat16, device="cuda")
    L.call("pp_conv3x3_bn_stats", _p(x0), C0, _p(x1), C1, _p(wf), _p(b), _p(y), Co, _p(stats), G, N, H, W, 1, _st())
    y2 = torch.empty_like(y)
    L.call("pp_conv3x3", PF.BF16, _p(x0), C0, _p(x1), C1, _p(wf), _p(b), _p(y2), Co, 0, None, 0, 0, N, H, W, 1, _st())
    assert torch.equal(y, y2)
    yg = y.double().view(G, -1, Co)
    ref = torch.stack([yg.sum(1), (yg * yg).sum(1)], dim=-1)
    assert _rel(stats.sum(0), ref) < 1e-6


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_first_conv_head_pool_upsample_bn_vs_oracle(pp, precision):
    L, PF, _ = pp
    code = PF.dtype_code(precision)
    adt = PF.act_dtype(code)
    tol = 1e-5 if precision == "fp32" else 1e-2
    g = torch.Generator().manual_seed(11)
    N, H, W = 3, 24, 16
    # --- first conv (Cin = 1) forward + weight grad
    x = torch.randn(N, 1, H, W, generator=g)
    w = torch.randn(32, 1, 3, 3, generator=g).requires_grad_()
    b = torch.randn(32, generator=g)
    y_ref = F.conv2d(x, w, b, 1, 1)
    y = torch.empty(N, H, W, 32, dtype=adt, device="cuda")
    x_d, w_d, b_d = x.cuda(), w.detach().cuda(), b.cuda()   # keep every device buffer alive across the async launch
    L.call("pp_first_conv_fwd", code, _p(x_d), _p(w_d), _p(b_d), _p(y), N, H, W, 32, _st())
    assert _rel(y.float().permute(0, 3, 1, 2), y_ref.detach()) < tol
    gy = torch.randn(N, 32, H, W, generator=g)
    if precision == "bf16":
        gy = gy.bfloat16().float()
    y_ref.backward(gy)
    dw = torch.zeros(32, 1, 3, 3, device="cuda")
    gy_d = _nhwc(gy, adt)
    L.call("pp_first_conv_wgrad", code, _p(gy_d), _p(x_d), _p(dw), N, H, W, 32, _st())
    assert _rel(dw, w.grad) < 1e-4
    # --- 1x1 head forward / backward (Cin 32 with bias, Cin 64 without)
    for cin, use_bias, C in ((32, True, 5), (64, False, 4)):
        a = torch.randn(N, cin, H, W, generator=g)
        if precision == "bf16":
            a = a.bfloat16().float()
        ar = a.clone().requires_grad_()
        hw_ = torch.randn(C, cin, 1, 1, generator=g).requires_grad_()
        hb = torch.randn(C, generator=g).requires_grad_() if use_bias else None
        z_ref = F.conv2d(ar, hw_, hb)
        dz = torch.randn(N, C, H, W, generator=g)
        z_ref.backward(dz)
        a_d = _nhwc(a, adt)
        z = torch.empty(N, C, H, W, device="cuda")
        hw_d, hb_d, dz_d = hw_.detach().cuda(), (hb.detach().cuda() if use_bias else None), dz.cuda()
        L.call("pp_head_fwd", code, _p(a_d), _p(hw_d), _p(hb_d), _p(z), N * H * W, H * W, cin, C, _st())
        assert _rel(z, z_ref.detach()) < 1e-5
        da = torch.empty(N, H, W, cin, dtype=adt, device="cuda")
        dwh = torch.zeros(C, cin, device="cuda")
        dbh = torch.zeros(C, device="cuda")
        L.call("pp_head_bwd", code, _p(dz_d), _p(a_d), _p(hw_d), _p(da), _p(dwh),
               _p(dbh) if use_bias else None, N * H * W, H * W, cin, C, _st())
        assert _rel(da.float().permute(0, 3, 1, 2), ar.grad) < tol
        assert _rel(dwh, hw_.grad.view(C, cin)) < 1e-4
        if use_bias:
            assert _rel(dbh, hb.grad) < 1e-4
    # --- max pool forward/backward (accumulating) incl. ties
    C = 64
    a = torch.randn(N, C, H, W, generator=g).round()  # rounding makes ties frequent
    ar = a.clone().requires_grad_()
    p_ref = F.max_pool2d(ar, 2, 2)
    gp = torch.randn(N, C, H // 2, W // 2, generator=g)
    if precision == "bf16":
        gp = gp.bfloat16().float()
    p_ref.backward(gp)
    a_d = _nhwc(a, adt)
    p = torch.empty(N, H // 2, W // 2, C, dtype=adt, device="cuda")
    L.call("pp_maxpool_fwd", code, _p(a_d), _p(p), N, H, W, C, _st())
    assert torch.equal(p.float().permute(0, 3, 1, 2).cpu(), p_ref.detach())
    gx = torch.ones(N, H, W, C, dtype=adt, device="cuda")
    gp_d = _nhwc(gp, adt)
    L.call("pp_maxpool_bwd", code, _p(a_d), _p(gp_d), _p(gx), N, H, W, C, 1, _st())
    assert _rel(gx.float().permute(0, 3, 1, 2) - 1, ar.grad) < tol
    # --- bilinear x2 (NHWC) and x8 (planes), align_corners=True, forward + backward
    a = torch.randn(N, C, 6, 4, generator=g)
    if precision == "bf16":
        a = a.bfloat16().float()
    ar = a.clone().requires_grad_()
    u_ref = F.interpolate(ar, scale_factor=2, mode="bilinear", align_corners=True)
    gu = torch.randn(N, C, 12, 8, generator=g)
    if precision == "bf16":
        gu = gu.bfloat16().float()
    u_ref.backward(gu)
    u = torch.empty(N, 12, 8, C, dtype=adt, device="cuda")
    a_d, gu_d = _nhwc(a, adt), _nhwc(gu, adt)
    L.call("pp_upsample_nhwc_fwd", code, _p(a_d), _p(u), N, 6, 4, 12, 8, C, _st())
    assert _rel(u.float().permute(0, 3, 1, 2), u_ref.detach()) < tol
    ga = torch.empty(N, 6, 4, C, dtype=adt, device="cuda")
    L.call("pp_upsample_nhwc_bwd", code, _p(gu_d), _p(ga), N, 6, 4, 12, 8, C, 0, _st())
    assert _rel(ga.float().permute(0, 3, 1, 2), ar.grad) < tol
    lo = torch.randn(N, 5, 4, 7, generator=g).requires_grad_()
    f_ref = F.interpolate(lo, size=(32, 56), mode="bilinear", align_corners=True)
    gf = torch.randn(N, 5, 32, 56, generator=g)
    f_ref.backward(gf)
    f = torch.empty(N, 5, 32, 56, device="cuda")
    lo_d, gf_d = lo.detach().cuda(), gf.cuda()
    L.call("pp_upsample_planes_fwd", _p(lo_d), _p(f), N * 5, 4, 7, 32, 56, _st())
    assert _rel(f, f_ref.detach()) < 1e-5
    gl = torch.empty(N, 5, 4, 7, device="cuda")
    L.call("pp_upsample_planes_bwd", _p(gf_d), _p(gl), N * 5, 4, 7, 32, 56, _st())
    assert _rel(gl, lo.grad) < 1e-5
    # --- BatchNorm (2 statistics groups) + LeakyReLU forward/backward, train and eval
    for training in (True, False):
        G, Ng, C = 2, 2, 64
        y = torch.randn(G * Ng, C, 8, 8, generator=g) * 2 + 0.5
        if precision == "bf16":
            y = y.bfloat16().float()
        gam, bet = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
        rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
        da_ = torch.randn(G * Ng, C, 8, 8, generator=g)
        if precision == "bf16":
            da_ = da_.bfloat16().float()
        yr, gr, br = y.clone().requires_grad_(), gam.clone().requires_grad_(), bet.clone().requires_grad_()
        rm_ref, rv_ref = rm.clone(), rv.clone()
        outs = []
        for gi in range(G):  # two reference passes, weak then strong
            o = F.batch_norm(yr[gi * Ng:(gi + 1) * Ng], rm_ref, rv_ref, gr, br, training, 0.1, 1e-5)
            outs.append(F.leaky_relu(o, 0.01))
        a_ref = torch.cat(outs)
        a_ref.backward(da_)
        Pg = Ng * 64
        y_d = _nhwc(y, adt)
        sums = torch.zeros(2 * G * C, dtype=torch.float64, device="cuda")
        coef = torch.empty(4 * G * C, device="cuda")
        rm_d, rv_d, gam_d, bet_d, da_d = rm.cuda(), rv.cuda(), gam.cuda(), bet.cuda(), _nhwc(da_, adt)
        nbt = torch.zeros((), dtype=torch.long, device="cuda")
        if training:
            L.call("pp_bn_stats", code, _p(y_d), _p(sums), G, Pg, C, _st())
        L.call("pp_bn_finalize", _p(sums), _p(gam_d), _p(bet_d), _p(rm_d), _p(rv_d), _p(nbt), _p(coef), G, Pg,
               C, int(training), 1e-5, 0.1, _st())
        a_d = torch.empty_like(y_d)
        L.call("pp_bn_apply", code, _p(y_d), _p(coef), _p(a_d), G, Pg, C, 0.01, _st())
        assert _rel(a_d.float().permute(0, 3, 1, 2), a_ref.detach()) < tol
        assert _rel(rm_d, rm_ref) < 1e-5 and _rel(rv_d, rv_ref) < 1e-5
        assert int(nbt) == (G if training else 0)
        bs = torch.empty(2 * G * C, dtype=torch.float64, device="cuda")
        bc = torch.empty(2 * G * C, device="cuda")
        dg, dbt, dbs = (torch.zeros(C, device="cuda") for _ in range(3))
        dy = torch.empty_like(y_d)
        L.call("pp_bn_bwd", code, _p(da_d), _p(y_d), _p(coef), _p(bs), _p(bc), _p(dg), _p(dbt), _p(dbs),
               _p(dy), G, Pg, C, int(training), 0.01, _st())
        assert _rel(dy.float().permute(0, 3, 1, 2), yr.grad) < (2e-2 if precision == "bf16" else 1e-4)
        assert _rel(dg, gr.grad) < 1e-4 and _rel(dbt, br.grad) < 1e-4
        if not training:
            assert _rel(dbs, yr.grad.sum(dim=(0, 2, 3))) < 1e-4


def test_loss_functions_vs_reference_golden(pp):
    """Every called function of losses/losses.py: value and gradient against the reference's own outputs."""
    L, PF, _ = pp
    sys.path.insert(0, os.path.join(Hn.ROOT, "pacingpseudo_b200", "dropin"))
    from losses import losses as DL
    g = Hn.load_golden("loss_functions")
    C = g["za"].shape[1]
    mask = torch.tensor(g["mask"]).cuda()
    target = torch.tensor(g["target"]).cuda()
    onehot = torch.tensor(g["onehot"]).cuda()

    def check(name, fn):
        za = torch.tensor(g["za"]).cuda().requires_grad_()
        zb = torch.tensor(g["zb"]).cuda().requires_grad_()
        v = fn(za, zb)
        (v * 1.0).backward()
        assert abs(v.item() - float(g[name])) <= 2e-5 * max(1.0, abs(float(g[name]))), (name, v.item(), float(g[name]))
        for t, key in ((za, name + "/dza"), (zb, name + "/dzb")):
            ref = g[key]
            if ref.size == 0:
                assert t.grad is None or float(t.grad.abs().max()) == 0, key
            else:
                np.testing.assert_allclose(t.grad.cpu().numpy(), ref, rtol=2e-3, atol=2e-7, err_msg=key)

    check("pce", lambda a, b: DL.partial_cross_entropy_loss(a, target, C))
    check("ce", lambda a, b: DL.cross_entropy_loss(a, target.clamp(max=C - 1)))
    for tag, m in (("mask", mask), ("nomask", None)):
        check("ent_" + tag, lambda a, b: DL.entropy_minimization_loss(a, m))
        check("softce_" + tag, lambda a, b: DL.soft_label_cross_entropy_loss(b, torch.softmax(a, 1), m))
        check("l1_" + tag, lambda a, b: DL.l1_loss(torch.softmax(b, 1), torch.softmax(a, 1), m))
        check("l2_" + tag, lambda a, b: DL.l2_loss(torch.softmax(b, 1), torch.softmax(a, 1), m))
        check("kl_" + tag, lambda a, b: DL.kl_loss(b, a, m))
    check("dice", lambda a, b: DL.dice_loss_fn(a, onehot))
    # all pixels ignored -> NaN, like F.cross_entropy
    z = torch.randn(1, 3, 4, 4, device="cuda")
    assert torch.isnan(DL.partial_cross_entropy_loss(z, torch.full((1, 4, 4), 3, device="cuda"), 3))
    # the caller mutates the returned losses in place and then backpropagates
    z = torch.randn(2, 3, 4, 4, device="cuda", requires_grad=True)
    t = torch.randint(0, 3, (2, 4, 4), device="cuda")
    out = PF.scribble_losses(z, t, 3, do_ent=True)
    loss = out["loss_pce"]
    loss += out["loss_ent"] * 0.5
    ent = out["loss_ent"]
    ent *= 2.0
    loss.backward()
    ref = z.detach().cpu().requires_grad_()
    (O.partial_cross_entropy(ref, t.cpu(), 3) + 0.5 * O.entropy_minimization(ref)).backward()
    np.testing.assert_allclose(z.grad.cpu().numpy(), ref.grad.numpy(), rtol=1e-3, atol=1e-7)


def _oracle_fused_losses(zw, zs, za, target, mask, C, variant, detach_weak):
    """The four terms of consistency_reglur_memory.py:32-81 from the oracle's loss functions (fp64, CPU)."""
    pce = O.partial_cross_entropy(zw, target, C)
    ent = O.entropy_minimization(zw, mask)
    pw = torch.softmax(zw.detach() if detach_weak else zw, 1)
    if variant == "ce_loss":
        cr = O.soft_label_cross_entropy(zs, pw, mask)
    elif variant == "l1_loss":
        cr = O.l1(torch.softmax(zs, 1), pw, mask)
    elif variant == "l2_loss":
        cr = O.l2(torch.softmax(zs, 1), pw, mask)
    else:
        cr = O.kl(zs, zw, mask)
    aux = O.partial_cross_entropy(za, target, C)
    return pce, ent, cr, aux


@pytest.mark.parametrize("variant", ["ce_loss", "l1_loss", "l2_loss", "kl_loss"])
@pytest.mark.parametrize("C,shape,use_mask,detach", [(5, (3, 16, 24), True, False), (2, (2, 12, 20), True, True),
                                                      (3, (2, 8, 16), False, False), (4, (2, 14, 16), True, False),
                                                      (7, (1, 8, 12), True, False), (5, (2, 5, 7), True, False)])
def test_fused_scribble_loss_all_terms_vs_oracle(pp, variant, C, shape, use_mask, detach):
    """All four terms in one call, with different upstream weights per term: the compile-time-C kernels
    (C in 2..5, H*W % 4 == 0), the generic kernels (C = 7, ragged 5x7, or PP_LOSS_GENERIC=1) and the fp64 oracle."""
    L, PF, _ = pp
    N, H, W = shape
    g = torch.Generator().manual_seed(100 * C + H)
    zw0, zs0, za0 = (3.0 * torch.randn(N, C, H, W, generator=g, dtype=torch.float64) for _ in range(3))
    target = torch.randint(0, C + 1, (N, H, W), generator=g)
    target[torch.rand(N, H, W, generator=g) < 0.6] = C
    target[0, 0, :4] = torch.tensor([0, C, 1, C])
    mask = (torch.rand(N, 1, H, W, generator=g) < 0.7).double() if use_mask else None
    wts = (1.0, 0.37, 0.81, 0.01)

    ref_in = [t.clone().requires_grad_() for t in (zw0, zs0, za0)]
    ref_terms = _oracle_fused_losses(*ref_in, target, mask, C, variant, detach)
    sum(w * t for w, t in zip(wts, ref_terms)).backward()

    results = {}
    for forced in ("0", "1"):
        os.environ["PP_LOSS_GENERIC"] = forced
        try:
            dev_in = [t.float().cuda().requires_grad_() for t in (zw0, zs0, za0)]
            out = PF.scribble_losses(dev_in[0], target.cuda(), C, zs=dev_in[1], za=dev_in[2],
                                     mask=None if mask is None else mask.float().cuda(), do_ent=True,
                                     cr_variant=variant, detach_weak=detach)
            terms = [out["loss_pce"], out["loss_ent"], out["loss_cr"], out["loss_aux"]]
            sum(w * t for w, t in zip(wts, terms)).backward()
            torch.cuda.synchronize()
        finally:
            os.environ.pop("PP_LOSS_GENERIC", None)
        for name, t, r in zip(("pce", "ent", "cr", "aux"), terms, ref_terms):
            assert abs(t.item() - r.item()) <= 1e-4 * max(1.0, abs(r.item())), (forced, name, t.item(), r.item())
        for name, t, r in zip(("dzw", "dzs", "dza"), dev_in, ref_in):
            assert _rel(t.grad, r.grad) < 1e-4, (forced, name, _rel(t.grad, r.grad))
        results[forced] = [t.grad.clone() for t in dev_in]
    for a, b in zip(results["0"], results["1"]):
        assert _rel(a, b) < 1e-5


@pytest.mark.parametrize("C,shape,low", [(5, (3, 32, 48), (4, 6)), (2, (2, 24, 24), (3, 3)), (4, (2, 28, 28), (7, 7)),
                                         (7, (2, 16, 24), (2, 3)), (5, (2, 9, 7), (3, 2)), (5, (1, 16, 16), (16, 16))])
def test_fused_scribble_loss_low_resolution_aux_logits(pp, C, shape, low):
    """The aux logits handed over at fc_cls's resolution (aux_path_memory.py:51): the loss kernels interpolate them at
    the labelled pixels (pp_scribble_loss_lowaux_*). Checked against F.interpolate(bilinear, align_corners=True) +
    the fp64 oracle losses (value of every term, gradient w.r.t. the LOW-resolution tensor), against the library's own
    full-resolution path (PF.upsample_planes + pp_scribble_loss_*), for the lean and the generic kernels."""
    L, PF, _ = pp
    N, H, W = shape
    g = torch.Generator().manual_seed(7 * C + H + low[0])
    zw0, zs0 = (3.0 * torch.randn(N, C, H, W, generator=g, dtype=torch.float64) for _ in range(2))
    za0 = 3.0 * torch.randn(N, C, low[0], low[1], generator=g, dtype=torch.float64)
    target = torch.randint(0, C + 1, (N, H, W), generator=g)
    target[torch.rand(N, H, W, generator=g) < 0.8] = C
    target[0, 0, :4] = torch.tensor([0, C, 1, C])
    target[-1, -1, -1] = C - 1                               # the last pixel reads the last taps
    mask = (torch.rand(N, 1, H, W, generator=g) < 0.7).double()
    wts = (1.0, 0.37, 0.81, 0.01)

    ref_in = [t.clone().requires_grad_() for t in (zw0, zs0, za0)]
    za_full = F.interpolate(ref_in[2], size=(H, W), mode="bilinear", align_corners=True)
    ref_terms = _oracle_fused_losses(ref_in[0], ref_in[1], za_full, target, mask, C, "ce_loss", False)
    sum(w * t for w, t in zip(wts, ref_terms)).backward()

    grads = {}
    for how in ("low", "low-generic", "full"):
        os.environ["PP_LOSS_GENERIC"] = "1" if how == "low-generic" else "0"
        try:
            dev_in = [t.float().cuda().requires_grad_() for t in (zw0, zs0, za0)]
            za = dev_in[2] if how != "full" else PF.upsample_planes(dev_in[2], (H, W))
            out = PF.scribble_losses(dev_in[0], target.cuda(), C, zs=dev_in[1], za=za, mask=mask.float().cuda(),
                                     do_ent=True, cr_variant="ce_loss")
            terms = [out["loss_pce"], out["loss_ent"], out["loss_cr"], out["loss_aux"]]
            sum(w * t for w, t in zip(wts, terms)).backward()
            torch.cuda.synchronize()
        finally:
            os.environ.pop("PP_LOSS_GENERIC", None)
        for name, t, r in zip(("pce", "ent", "cr", "aux"), terms, ref_terms):
            assert abs(t.item() - r.item()) <= 1e-4 * max(1.0, abs(r.item())), (how, name, t.item(), r.item())
        assert tuple(dev_in[2].grad.shape) == (N, C) + tuple(low)
        for name, t, r in zip(("dzw", "dzs", "dza_low"), dev_in, ref_in):
            assert _rel(t.grad, r.grad) < 1e-4, (how, name, _rel(t.grad, r.grad))
        grads[how] = [t.grad.clone() for t in dev_in]
    for how in ("low-generic", "full"):
        for a, b in zip(grads["low"], grads[how]):
            assert _rel(a, b) < 1e-5, how


@pytest.mark.parametrize("mode", ["cosine_similarity", "mean"])
def test_memory_update_vs_oracle(pp, mode):
    """aux_path_memory.py:68-116 incl. sample-0-only, first-touch mean, absent classes, in-place normalisation."""
    L, PF, _ = pp
    g = torch.Generator().manual_seed(3)
    C, hid, N, h, w, H, W = 5, 64, 3, 8, 8, 64, 64
    feats = torch.randn(N, hid, h, w, generator=g)
    scrib = torch.zeros(N, C + 1, H, W)
    scrib[0, 0, 5, 3:40] = 1
    scrib[0, 2, 20:50, 7] = 1
    scrib[0, 3, 63, 63] = 1          # a single pixel on the border
    scrib[1, 1, 9, 9] = 1            # class 1 only in sample 1: must never be touched
    bank_ref = torch.zeros(C, hid, 1, 1)
    bank = torch.zeros(C, hid, device="cuda")
    feats_d = feats.permute(0, 2, 3, 1).contiguous().cuda()
    for step in (0, 50, 399):
        O.memory_update(bank_ref, feats, scrib, step, 400, 0.9, mode)
        PF.memory_update(PF.F32, feats_d, scrib.cuda(), bank, mode, O.ramp_up_mo(step, 400, 0.9))
        assert _rel(bank, bank_ref.view(C, hid)) < 1e-5
        assert float(bank[1].abs().sum()) == 0 and float(bank[4].abs().sum()) == 0
        feats = feats + 0.3 * torch.randn(N, hid, h, w, generator=g)
        feats_d = feats.permute(0, 2, 3, 1).contiguous().cuda()


# ------------------------------------------------------------------------------------------------
# module level: the reference's golden vectors
# ------------------------------------------------------------------------------------------------
GOLDEN_CASES = list(CASES)


def test_cuda_graph_replay_of_whole_passes_is_identical_to_eager(pp):
    """pp_unet_forward / pp_unet_backward are replayed as CUDA graphs once the same call (pointers, shapes) recurs: the
    eager first pass, the captured second one and the replays must agree (forward: bit-identical logits in eval mode;
    backward into FlatAdam's flat gradient buffer: identical up to the fp32 atomics of the narrow weight gradients)."""
    L, PF, _ = pp
    from pacingpseudo_b200.optim import FlatAdam
    from pacingpseudo_b200.synth import make_batch
    case = dict(kind="baseline", C=5, os=8, training=False)
    model = Hn.build_cuda_model(case, "bf16").eval()
    opt = FlatAdam(model.parameters(), lr=0.0)
    batch = make_batch(4, 5, 128, 128, seed=31)
    x = batch["image"].cuda()
    target = batch["scribble"].argmax(1).to(torch.uint8).cuda()
    from losses import losses as DL
    r0 = L.cdll.pp_graph_replays()
    # results go into storage allocated up front: the caching allocator then hands the SAME blocks (workspace, logits,
    # loss gradient) to every iteration, which is what a steady-state training loop looks like and what replay keys on
    logits = torch.empty(6, 4, 5, 128, 128, device="cuda")
    grads = torch.empty(6, opt.flat_grad.numel(), device="cuda")
    for it in range(6):
        opt.zero_grad()
        z = model(x)["segmentation/logits"]
        loss = DL.partial_cross_entropy_loss(z, target, 5)
        loss.backward()
        logits[it].copy_(z.detach())
        grads[it].copy_(opt.flat_grad)
        del z, loss
    torch.cuda.synchronize()
    if os.environ.get("PP_GRAPHS", "1") != "0":
        assert L.cdll.pp_graph_replays() - r0 >= 4, "the repeated passes were not replayed as graphs"
    for z in logits[1:]:
        assert torch.equal(z, logits[0])
    for g in grads[1:]:
        assert torch.isfinite(g).all() and _rel(g, grads[0]) < 1e-4


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_dropin_modules_fp32_mode_vs_reference_golden(pp, name):
    rec = Hn.run_case_cuda(name, "fp32")
    report = []
    fails = Hn.compare(rec, Hn.load_golden(name), Hn.TOL["fp32"], CASES[name].get("steps", 1), report)
    assert not fails, "\n".join(fails + report)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_dropin_modules_bf16_vs_reference_golden(pp, name):
    """bf16 path on the (tiny) golden cases against the reference's fp32 vectors: the five losses within the
    north-star 2e-2, logits / gradient-free tensor norms within 2e-2, bank within 5e-2, running statistics 2e-2.
    Element-wise logits and gradients are judged at full size in tests/test_gpu_fullsize.py (bf16 storage noise is
    amplified ~40x by this network at the seeded initial state)."""
    rec = Hn.run_case_cuda(name, "bf16")
    report = []
    fails = Hn.compare(rec, Hn.load_golden(name), Hn.TOL["bf16_small"], CASES[name].get("steps", 1), report)
    print("\n".join(report))
    assert not fails, "\n".join(fails + report)


# ------------------------------------------------------------------------------------------------
# full size (BASELINE.json configs): oracle on one case + size-independent properties
# ------------------------------------------------------------------------------------------------
def _full_model(C=5, precision="bf16", seed=1):
    case = dict(kind="pacing", C=C, os=8, training=True, cr="ce_loss", mode="cosine_similarity")
    torch.manual_seed(seed)
    return Hn.build_cuda_model(case, precision), case


def test_full_size_step_properties(pp):
    """N=12, 256x256, C=5 (config 2). Properties that hold at any size:
    (a) strong == weak image and eval-mode BN  =>  logits_strong == logits_weak and loss_cr == loss_ent;
    (b) gradients are linear in the loss weight (x4: exact in bf16); (c) every loss/grad is finite; (d) weak-only val pass matches
    the weak half of the batched train pass."""
    from pacingpseudo_b200.synth import make_batch
    model, case = _full_model()
    batch = {k: v.cuda() for k, v in make_batch(12, 5, 256, 256, seed=5).items() if k != "label"}
    model.eval()
    same = dict(batch, image_strong=batch["image"])
    out = model(same, mode="train", step=3)
    assert torch.equal(out["segmentation/logits"], out["segmentation/logits_strong"])
    assert abs(out["loss_cr"].item() - out["loss_ent"].item()) < 1e-5 * max(1.0, abs(out["loss_ent"].item()))
    with torch.no_grad():
        val = model(batch, mode="val")
    assert torch.equal(val["segmentation/logits"], out["segmentation/logits"])
    assert set(val) == {"segmentation/logits", "loss_pce"}

    model.train()
    grads = []
    for wgt in (1.0, 4.0):   # a power of two: bf16 / fp32 scale exactly, so backward must be linear to rounding-order noise
        model.zero_grad(set_to_none=True)
        model.aux_path.memory_bank.data.zero_()
        sd0 = {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}
        out = model(batch, mode="train", step=3)
        (wgt * O.total_loss(out, epoch=40)).backward()
        model.load_state_dict(sd0, strict=False)  # same BN running stats for both runs
        grads.append([p.grad.clone() for p in model.parameters() if p.grad is not None])
        for k in ("loss_pce", "loss_ent", "loss_cr", "loss_aux_cls", "loss_memory"):
            assert torch.isfinite(out[k]), k
    for a, b in zip(*grads):
        assert torch.isfinite(a).all()
        assert _rel(4.0 * a, b) < 1e-3


# The full-size comparisons against the oracle (configs 1-5, seeded and trained state, fp32 mode and bf16) live in
# tests/test_gpu_fullsize.py; their measured distances are committed as profiles/r02_parity_fullsize.txt.


@pytest.mark.parametrize("max_ch,os_,strided,size", [(1024, 8, False, 64), (1024, 32, True, 64), (512, 16, True, 128),
                                                      (512, 8, True, 112), (728, 8, False, 64), (728, 16, False, 64)])
def test_unet_variants_fp32_mode_vs_oracle(pp, max_ch, os_, strided, size):
    """train_chaos.py:71 allows max_ch 1024; unet.py:113-116,141 the stride-2 / ConvTranspose2d variant, here at larger
    and ragged (112 = 7 * 16) sizes than the golden cases: logits, pCE and every gradient against the fp32 CPU oracle
    in the library's fp32 mode, plus a bf16 run against the north-star bf16 tolerances on the loss."""
    from pacingpseudo_b200.synth import make_batch
    from pacingpseudo_b200.dropin import DROPIN_PATH
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    from models.unet import UNet
    from losses import losses as DL
    C, N = 4, 2
    sd = O.synth_state_dict(O.unet_param_shapes(1, 32, max_ch, C, os_, strided=strided), seed=11)
    batch = make_batch(N, C, size, size, seed=21)
    target = batch["scribble"].argmax(1)
    names = [k for k in sd if sd[k].is_floating_point() and "running" not in k]
    s_ = {k: v.clone() for k, v in sd.items()}
    for k in names:
        s_[k].requires_grad_(True)
    z_ref = O.unet_forward(s_, batch["image"], True, max_ch=max_ch, output_stride=os_, strided=strided)["segmentation/logits"]
    l_ref = O.partial_cross_entropy(z_ref, target, C)
    l_ref.backward()
    for precision in ("fp32", "bf16"):
        model = UNet(1, 32, max_ch, C, os_, strided, strided, True, precision=precision)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().train()
        z = model(batch["image"].cuda())["segmentation/logits"]
        loss = DL.partial_cross_entropy_loss(z, target.cuda(), C)
        loss.backward()
        tol = Hn.TOL[precision]
        assert abs(loss.item() - l_ref.item()) <= tol["loss"] * abs(l_ref.item()), (precision, loss.item(), l_ref.item())
        if precision == "fp32":
            assert _rel(z.detach(), z_ref.detach()) < tol["logits"]
            # conv biases in front of a batch-statistics BatchNorm have a mathematically zero gradient (noise / noise)
            worst = max((_rel(p.grad, s_[k].grad), k) for k, p in model.named_parameters()
                        if not k.endswith(".conv.bias"))
            assert worst[0] < tol["grad"], worst
        for k, p in model.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
        if max_ch == 728:   # zero-padded stage: the end point is handed out at its true width, running stats come back
            assert model.end_points["encoder/stage6"].shape[1] == 728
            rm = model.enc_block6.conv_block.conv_layer2.norm_op.running_mean
            assert rm.shape[0] == 728
            # bf16 activations move a batch mean by ~3e-3 relative; the fp32 mode stays within 1e-3
            assert _rel(rm, s_["enc_block6.conv_block.conv_layer2.norm_op.running_mean"]) < (1e-3 if precision == "fp32" else 2e-2)


def test_pacing_step_max_ch_728_vs_oracle(pp):
    """Full pacingpseudo step with `--max_ch 728` (train_chaos.py:71; aux path fed by the 728-wide stage 6): the five
    losses, the bank and the aux-conv gradient against the fp32 oracle, fp32 mode."""
    import argparse
    from pacingpseudo_b200.synth import make_batch
    from pacingpseudo_b200.dropin import DROPIN_PATH
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    from models.consistency_reglur_memory import ConsistencyRegulr
    C = 3
    shapes = {"backbone." + k: v for k, v in O.unet_param_shapes(1, 32, 728, C, 8).items()}
    shapes.update({"aux_path." + k: v for k, v in O.aux_param_shapes(C, (728, 512), 64).items()})
    sd = O.synth_state_dict(shapes, seed=5)
    batch = make_batch(2, C, 64, 64, seed=31)
    batch.pop("label")
    ns = argparse.Namespace(ignored_index=C, do_loss_ent=True, do_decoder_consistency=True, detach_weak_cr=False,
                            loss_cr_variants="ce_loss", do_aux_path=True, do_memory=True)
    model = ConsistencyRegulr(
        kwargs_unet=dict(input_ch=1, init_ch=32, max_ch=728, num_classes=C, output_stride=8, is_stride_conv=False,
                         is_trans_conv=False, elab_end_points=True, precision="fp32"),
        kwargs_aux_path=dict(num_classes=C, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[728, 512],
                             hid_ch=64, aux_drop_prob=0., do_memory=True, max_step=400, update_momentum=0.9,
                             ensemble_mode='cosine_similarity'),
        args_parser=ns)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    out = model({k: v.cuda() for k, v in batch.items()}, mode="train", step=5)
    O.total_loss(out, epoch=40).backward()

    s_ = {k: v.clone() for k, v in sd.items()}
    learn = [k for k in s_ if s_[k].is_floating_point() and "running" not in k and not k.endswith("memory_bank")]
    for k in learn:
        s_[k].requires_grad_(True)
    cfg = O.StepConfig(num_classes=C, ignored_index=C, max_ch=728)
    ref = O.consistency_forward(s_, batch, cfg, mode="train", step=5, training=True)
    O.total_loss(ref, epoch=40).backward()
    for k in ("loss_pce", "loss_ent", "loss_cr", "loss_aux_cls", "loss_memory"):
        assert abs(out[k].item() - ref[k].item()) <= 1e-4 * max(1.0, abs(ref[k].item())), (k, out[k].item(), ref[k].item())
    assert _rel(model.aux_path.memory_bank.detach(), s_["aux_path.memory_bank"]) < 1e-4
    for k in ("aux_path.layer_bottleneck.1.weight", "backbone.enc_block6.conv_block.conv_layer2.conv.weight",
              "backbone.dec_block5.conv_block.conv_layer1.conv.weight", "backbone.enc_block1.conv_block.conv_layer1.conv.weight"):
        g = dict(model.named_parameters())[k].grad
        assert g.shape == s_[k].grad.shape and _rel(g, s_[k].grad) < 2e-2, (k, _rel(g, s_[k].grad))


def test_unet_multi_channel_input_vs_oracle(pp):
    """--input_ch 3 (train_chaos.py:65): NCHW fp32 input with three channels through the generic first-conv kernels,
    two BatchNorm statistics groups (the two parts of the batch run on two streams) against the fp32 oracle."""
    from pacingpseudo_b200.dropin import DROPIN_PATH
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    from models.unet import UNet
    from losses import losses as DL
    C, N, Cin = 4, 4, 3
    sd = O.synth_state_dict(O.unet_param_shapes(Cin, 32, 512, C, 16), seed=13)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(N, Cin, 48, 64, generator=g)
    target = torch.randint(0, C + 1, (N, 48, 64), generator=g)
    names = [k for k in sd if sd[k].is_floating_point() and "running" not in k]
    s_ = {k: v.clone() for k, v in sd.items()}
    for k in names:
        s_[k].requires_grad_(True)
    # two statistics groups == two independent forward passes over the halves (running stats: first half, then second)
    z_ref = torch.cat([O.unet_forward(s_, x[i:i + 2], True, output_stride=16)["segmentation/logits"] for i in (0, 2)], 0)
    l_ref = O.partial_cross_entropy(z_ref, target, C)
    l_ref.backward()
    for precision in ("fp32", "bf16"):
        model = UNet(Cin, 32, 512, C, 16, False, False, True, precision=precision)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().train()
        z, _ = model.run_native(x.cuda(), groups=2)
        loss = DL.partial_cross_entropy_loss(z, target.cuda(), C)
        loss.backward()
        tol = Hn.TOL[precision]
        assert abs(loss.item() - l_ref.item()) <= tol["loss"] * abs(l_ref.item()), (precision, loss.item(), l_ref.item())
        if precision == "fp32":
            assert _rel(z.detach(), z_ref.detach()) < tol["logits"]
            for k in ("enc_block1.conv_block.conv_layer1.conv.weight", "enc_block1.conv_block.conv_layer2.conv.weight",
                      "final_conv.weight"):
                gk = dict(model.named_parameters())[k].grad
                assert _rel(gk, s_[k].grad) < tol["grad"], (k, _rel(gk, s_[k].grad))
    with pytest.raises(RuntimeError, match="input channel"):
        model(torch.zeros(1, 1, 16, 16, device="cuda"))


def test_space_depth_and_channel_scale_operators(pp):
    """pp_space_to_depth / pp_depth_to_space (exact permutations, += variant) and pp_channel_scale against torch."""
    L, _, pplib = pp
    g = torch.Generator().manual_seed(3)
    for dtype, code in ((torch.bfloat16, pplib.BF16), (torch.float32, pplib.F32)):
        N, Hs, Ws, C = 2, 5, 7, 24
        x = torch.randn(N, 2 * Hs, 2 * Ws, C, generator=g).to(dtype).cuda()
        y = torch.empty(N, Hs, Ws, 4 * C, dtype=dtype, device="cuda")
        L.call("pp_space_to_depth", code, _p(x), _p(y), N, Hs, Ws, C, _st())
        ref = x.view(N, Hs, 2, Ws, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(N, Hs, Ws, 4 * C)
        assert torch.equal(y, ref)
        back = torch.empty_like(x)
        L.call("pp_depth_to_space", code, _p(y), _p(back), N, Hs, Ws, C, 0, _st())
        assert torch.equal(back, x)
        L.call("pp_depth_to_space", code, _p(y), _p(back), N, Hs, Ws, C, 1, _st())
        assert torch.equal(back.float(), (x.float() * 2).to(dtype).float())
        sc = (torch.rand(N, C + 8, generator=g) > 0.5).float().cuda() * 2.0
        out = torch.empty_like(x)
        L.call("pp_channel_scale", code, _p(x), ctypes.c_void_p(sc.data_ptr() + 4 * 8), _p(out), N, 4 * Hs * Ws, C, C + 8, _st())
        assert torch.equal(out.float(), (x.float() * sc[:, None, None, 8:]).to(dtype).float())


def test_strong_color_augment_kernel_vs_reference_golden(pp):
    """pp_strong_color_augment against the reference's own augmentation chain (tests/golden/strong_augment.npz), the
    numpy oracle at BASELINE's full size, and ragged sizes (H*W not a multiple of the block)."""
    from oracle.gen_golden import strong_augment_inputs
    from pacingpseudo_b200.data import sample_strong_params, strong_color_augment
    gold = Hn.load_golden("strong_augment")
    imgs, params = strong_augment_inputs()
    out = strong_color_augment(torch.tensor(imgs)[:, None].cuda(), torch.tensor(params)).cpu().numpy()[:, 0]
    np.testing.assert_allclose(out, gold["out"], rtol=2e-5, atol=2e-5)
    assert np.array_equal(out[7], imgs[7])
    from pacingpseudo_b200.synth import make_batch
    for n, size in ((12, 256), (3, 37)):
        x = make_batch(n, 5, size, size, seed=5)["image"]
        p = sample_strong_params(n, 1.0, torch.Generator().manual_seed(size))
        got = strong_color_augment(x.cuda(), p).cpu().numpy()
        for i in range(n):
            ref = O.strong_color_augment_np(x[i, 0].numpy(), p[i].numpy())
            np.testing.assert_allclose(got[i, 0], ref, rtol=5e-5, atol=5e-5, err_msg="size %d image %d" % (size, i))


def test_compact_index_map_scribble_is_equivalent(pp):
    """the drop-in accepts the scribble as a uint8 class-index map; the step is bit-identical to the
    one-hot path (same losses, same gradients, same bank), with 24x fewer scribble bytes crossing PCIe."""
    from pacingpseudo_b200.data import compact_batch
    from pacingpseudo_b200.synth import make_batch
    case = dict(kind="pacing", C=5, os=8, training=True, cr="ce_loss", mode="cosine_similarity")
    batch = make_batch(3, 5, 64, 64, seed=4)
    results = []
    for compact in (False, True):
        model = Hn.build_cuda_model(case, "bf16")
        b = dict(compact_batch(batch, 5), image_strong=batch["image_strong"]) if compact else \
            {k: v for k, v in batch.items() if k != "label"}
        assert (b["scribble"].dim() == 3) == compact
        out = model({k: v.cuda() for k, v in b.items()}, mode="train", step=3)
        loss = O.total_loss(out, epoch=40)
        loss.backward()
        results.append(([out[k].item() for k in ("loss_pce", "loss_ent", "loss_cr", "loss_aux_cls", "loss_memory")],
                        {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
                        model.aux_path.memory_bank.detach().clone()))
    assert results[0][0] == results[1][0]
    assert torch.equal(results[0][2], results[1][2])
    # weight gradients use fp32 atomics in a few kernels: compare tightly rather than bitwise
    for k in results[0][1]:
        assert _rel(results[1][1][k], results[0][1][k]) < 1e-5 or k.endswith(".conv.bias"), k


def test_data_parallel_equivalence_emulated(pp):
    """DP(G ranks) == mean over ranks of single-GPU gradients on each rank's local batch. Emulated on one GPU
    by looping the ranks sequentially (the all-reduce itself is exercised by tests/test_dp_gloo.py on CPU)."""
    from pacingpseudo_b200.synth import make_batch
    case = dict(kind="pacing", C=4, os=8, training=True, cr="ce_loss", mode="cosine_similarity")
    model = Hn.build_cuda_model(case, "fp32")
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    per_rank = []
    for rank in range(2):
        model.load_state_dict(sd0)
        model.zero_grad(set_to_none=True)
        b = {k: v.cuda() for k, v in make_batch(2, 4, 64, 64, seed=1234 + 1000 * rank).items() if k != "label"}
        O.total_loss(model(b, mode="train", step=0), epoch=10).backward()
        per_rank.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    from pacingpseudo_b200.dp import average_gradients_emulated
    avg = average_gradients_emulated(per_rank)
    for k in avg:
        assert torch.allclose(avg[k], 0.5 * (per_rank[0][k] + per_rank[1][k]), rtol=1e-6, atol=1e-9)


def test_flat_adam_and_direct_gradient_accumulation(pp):
    """pacingpseudo_b200.optim.FlatAdam == torch.optim.Adam(lr, weight_decay) (train_chaos.py:219), and gradients
    accumulated directly into its flat buffer equal the gradients autograd would have delivered."""
    from pacingpseudo_b200.optim import FlatAdam
    from pacingpseudo_b200.synth import make_batch
    case = dict(kind="pacing", C=4, os=8, training=True, cr="ce_loss", mode="cosine_similarity")
    ma = Hn.build_cuda_model(case, "fp32")
    mb = Hn.build_cuda_model(case, "fp32")
    oa = torch.optim.Adam(ma.parameters(), lr=1e-3, weight_decay=3e-4)
    ob = FlatAdam(mb.parameters(), lr=1e-3, weight_decay=3e-4)
    assert all(p.grad is not None and p._pp_direct_grad for p in mb.parameters() if p.requires_grad)
    for step in range(3):
        b = {k: v.cuda() for k, v in make_batch(2, 4, 64, 64, seed=40 + step).items() if k != "label"}
        for m, o in ((ma, oa), (mb, ob)):
            loss = O.total_loss(m(b, mode="train", step=step), epoch=30)
            o.zero_grad()
            loss.backward()
        if step == 0:
            for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
                if pa.grad is not None:
                    assert pb.grad.data_ptr() >= ob.flat_grad.data_ptr()
                    assert _rel(pb.grad, pa.grad) < 1e-4 or float(pa.grad.abs().max()) < 1e-7, k
        oa.step()
        ob.step()
        if step == 0:
            # one Adam step from identical states: identical updates (lr * m_hat / (sqrt(v_hat) + eps)), except where the
            # gradient is numerically zero (conv bias under batch-statistics BN: the sign of rounding noise decides)
            for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
                if pa.requires_grad and not k.endswith("conv.bias") and not k.endswith("layer_bottleneck.1.bias"):
                    d = (pa.detach() - pb.detach()).abs()
                    assert float((d > 2e-6).float().mean()) < 2e-3, (k, float(d.max()))
    # later steps: the two runs drift apart chaotically (tiny case), but every element moved by <= ~lr per step
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert float((pa.detach() - pb.detach()).abs().max()) <= 2.5 * 3 * 1e-3, k


def test_device_prefetcher_and_loss_reader(pp):
    """Host->device staging (two persistent buffers, copy stream one batch ahead) delivers every batch intact even
    when the consumer is slow, and LossReader returns each step's scalars exactly one push later."""
    from pacingpseudo_b200.data import DevicePrefetcher, LossReader
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    host = [{"image": torch.randn(4, 1, 64, 64, generator=g).pin_memory(),
             "scribble": torch.randn(4, 6, 64, 64, generator=g).pin_memory(), "tag": i} for i in range(7)]
    reader = LossReader(dev)
    seen, lagged = 0, []
    big = torch.randn(2048, 2048, device=dev)
    pf = DevicePrefetcher(iter(host), dev)
    for i, b in enumerate(pf):
        assert b["tag"] == i
        s = (b["image"].double().sum() + b["scribble"].double().sum()).float()
        for _ in range(3):
            big = big @ big * 1e-3     # keep the consumer stream busy so that copies really run ahead
        lagged.append(reader.push([s, s * 2]))
        assert torch.equal(b["image"].cpu(), host[i]["image"]) and torch.equal(b["scribble"].cpu(), host[i]["scribble"])
        seen += 1
    lagged.append(reader.flush())
    assert seen == 7 and lagged[0] is None
    for i in range(7):
        ref = float(host[i]["image"].double().sum() + host[i]["scribble"].double().sum())
        assert abs(lagged[i + 1][0] - ref) < 1e-3 * max(1.0, abs(ref)) and abs(lagged[i + 1][1] - 2 * ref) < 2e-3 * max(1.0, abs(ref))
    assert pf.h2d_bytes == sum(v.numel() * 4 for b in host for v in b.values() if torch.is_tensor(v))
    ptrs = {k: v.data_ptr() for s in pf.slots for k, v in s.items()}
    for i, b in enumerate(pf.reset(iter(host[:3]))):     # next epoch: same staging buffers, data still intact
        assert torch.equal(b["scribble"].cpu(), host[i]["scribble"])
    assert ptrs == {k: v.data_ptr() for s in pf.slots for k, v in s.items()} and i == 2


@pytest.mark.parametrize("case", [(4, 8, 8, 512, 512, 512, 1), (3, 16, 16, 512, 256, 256, 1), (2, 32, 32, 128, 0, 256, 2),
                                  (2, 32, 32, 128, 64, 64, 1), (5, 8, 8, 512, 0, 512, 4), (1, 28, 28, 256, 0, 128, 1),
                                  # row variant of the narrow kernel with more K blocks per CTA than pipeline stages
                                  # (regression: the unused MN chunk of the last stage read past the allocation)
                                  (3, 256, 256, 64, 0, 32, 1), (4, 128, 128, 32, 0, 64, 1), (2, 256, 256, 64, 32, 32, 1)])
def test_conv3x3_wgrad_oihw_split_k_deterministic(pp, case):
    """Weight gradient accumulated into the OIHW gradient through the split-K scratch path (256-row tiles, fixed-order
    reduction): matches torch CPU autograd, accumulates (+=), and two runs are bit-identical; the atomic path agrees."""
    L, PF, _ = pp
    N, H, W, C0, C1, Co, dil = case
    g = torch.Generator().manual_seed(sum(case))
    x = torch.randn(N, C0 + C1, H, W, generator=g).bfloat16().float()
    gy = torch.randn(N, Co, H, W, generator=g).bfloat16().float()
    w = torch.zeros(Co, C0 + C1, 3, 3, requires_grad=True)
    F.conv2d(x, w, None, 1, dil, dil).backward(gy)
    x0 = _nhwc(x[:, :C0], torch.bfloat16)
    x1 = _nhwc(x[:, C0:], torch.bfloat16) if C1 else None
    dy = _nhwc(gy, torch.bfloat16)
    n = 9 * Co * (C0 + C1)
    dwp = torch.empty(n, device="cuda")
    outs = []
    for ws_floats in (2 * n, 2 * n, 5 * n, 0):
        ws = torch.full((max(ws_floats, 1),), float("nan"), device="cuda")   # stale scratch must not leak into the result
        gw = torch.ones(Co, C0 + C1, 3, 3, device="cuda")
        L.call("pp_conv3x3_wgrad_oihw", _p(dy), Co, _p(x0), C0, _p(x1), C1, _p(dwp), _p(gw),
               _p(ws) if ws_floats else None, ws_floats, N, H, W, dil, _st())
        torch.cuda.synchronize()
        outs.append(gw - 1.0)
    for o in outs:
        assert _rel(o, w.grad) < 1e-4
    narrow = any(c in (32, 64) and Co in (32, 64) for c in (C0, C1))   # 32/64-channel sources accumulate atomically
    if not narrow:
        assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("case", [(2, 32, 32, 2, 128), (1, 28, 28, 2, 256), (2, 64, 64, 2, 64), (3, 6, 4, 2, 32),
                                  (1, 4, 5, 8, 32), (2, 16, 16, 1, 64), (1, 8, 8, 3, 64)])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_upsample_nhwc_forward_backward_shapes(pp, case, precision):
    """nn.Upsample(bilinear, align_corners=True) (unet.py:144) forward and backward at realistic sizes: the strip
    backward kernel (scale 2, incl. odd widths), its accumulate (+=) mode, and the per-pixel gather kernel that other
    scale factors (1, 3, 8) fall back to."""
    L, PF, _ = pp
    N, h, w, sc, C = case
    code = PF.dtype_code(precision)
    adt = PF.act_dtype(code)
    tol = 1e-2 if precision == "bf16" else 1e-5
    g = torch.Generator().manual_seed(h * 131 + w + sc)
    a = torch.randn(N, C, h, w, generator=g)
    gu = torch.randn(N, C, h * sc, w * sc, generator=g)
    base = torch.randn(N, C, h, w, generator=g)
    if precision == "bf16":
        a, gu, base = a.bfloat16().float(), gu.bfloat16().float(), base.bfloat16().float()
    ar = a.clone().requires_grad_()
    u_ref = F.interpolate(ar, size=(h * sc, w * sc), mode="bilinear", align_corners=True)
    u_ref.backward(gu)
    u = torch.empty(N, h * sc, w * sc, C, dtype=adt, device="cuda")
    L.call("pp_upsample_nhwc_fwd", code, _p(_nhwc(a, adt)), _p(u), N, h, w, h * sc, w * sc, C, _st())
    assert _rel(u.float().permute(0, 3, 1, 2), u_ref.detach()) < tol
    gu_d = _nhwc(gu, adt)
    ga = torch.empty(N, h, w, C, dtype=adt, device="cuda")
    L.call("pp_upsample_nhwc_bwd", code, _p(gu_d), _p(ga), N, h, w, h * sc, w * sc, C, 0, _st())
    assert _rel(ga.float().permute(0, 3, 1, 2), ar.grad) < tol
    gacc = _nhwc(base, adt)
    L.call("pp_upsample_nhwc_bwd", code, _p(gu_d), _p(gacc), N, h, w, h * sc, w * sc, C, 1, _st())
    assert _rel(gacc.float().permute(0, 3, 1, 2), ar.grad + base) < 2 * tol


def test_dice_metric_vs_reference_golden_and_oracle(pp):
    """pp_dice_metric (whole batch, one pass) against the reference's compute_dice vectors (tests/golden/dice_metric.npz)
    and the oracle at a full-size batch; the per-sample wrapper keeps the reference's return type."""
    from oracle.gen_golden import dice_metric_inputs
    from pacingpseudo_b200.metrics import compute_dice, compute_dice_batch
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dice_metric.npz"))["dice"]
    scores, onehot = dice_metric_inputs()
    got = compute_dice_batch(torch.from_numpy(scores).cuda(), torch.from_numpy(onehot).cuda()).cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(gold))
    assert np.allclose(got, gold, rtol=1e-6, atol=1e-7, equal_nan=True)
    one = compute_dice(scores[0], onehot[0])
    assert isinstance(one, list) and len(one) == 5 and np.isnan(one[4])
    assert np.allclose(one, gold[0], rtol=1e-6, equal_nan=True)
    # full size (12 x 5 x 256 x 256), logits instead of softmax values (same arg-max)
    g = torch.Generator().manual_seed(3)
    z = torch.randn(12, 5, 256, 256, generator=g)
    lab = F.one_hot(torch.randint(0, 4, (12, 256, 256), generator=g), 5).permute(0, 3, 1, 2).float()   # class 4 never labelled
    z[:, 4] -= 100.0                                                                                     # ... nor predicted
    ref = np.array([O.compute_dice_np(z[n].numpy(), lab[n].numpy()) for n in range(12)])
    got = compute_dice_batch(z.cuda(), lab.cuda()).cpu().numpy()
    assert np.isnan(got[:, 4]).all() and np.isnan(ref[:, 4]).all()
    assert np.allclose(got, ref, rtol=1e-6, atol=1e-7, equal_nan=True)
