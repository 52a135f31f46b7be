"""CPU: the C-ABI shared library loads and exports every symbol include/pacingpseudo_b200.h declares;
host-side logic that needs no GPU (header parser, plan construction, workspace sizing, error convention)."""
import ctypes
import os
import subprocess

import pytest

from pacingpseudo_b200 import lib as pplib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(pplib.LIB_PATH):
        from pacingpseudo_b200.build import build
        build()
    return pplib.get_lib()


def test_header_declares_expected_surface():
    protos = pplib.parse_header()
    for name in ("pp_init", "pp_last_error", "pp_unet_forward", "pp_unet_backward", "pp_conv3x3", "pp_conv3x3_wgrad",
                 "pp_scribble_loss_fwd", "pp_scribble_loss_bwd", "pp_memory_update", "pp_dice_fwd", "pp_adam_step"):
        assert name in protos, name
    assert len(protos) >= 44


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", pplib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in pplib.parse_header() if n not in exported]
    assert not missing, missing


def test_library_is_sm100a_tcgen05():
    """The shipped binary contains Blackwell tensor-core / TMA / TMEM instructions (SASS mnemonics)."""
    sass = subprocess.run(["cuobjdump", "-sass", pplib.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    for mnem in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnem in sass, mnem


def test_unet_plan_matches_reference_layer_table(lib):
    h = ctypes.c_void_p()
    lib.call("pp_unet_create", 1, 32, 512, 5, 8, pplib.BF16, ctypes.byref(h))
    n = lib.cdll.pp_unet_num_convs(h)
    assert n == 22  # 22 3x3 convs in the backbone
    rows = []
    for i in range(n):
        cin, cout, dil, name = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_char_p()
        lib.call("pp_unet_conv_info", h, i, ctypes.byref(cin), ctypes.byref(cout), ctypes.byref(dil), ctypes.byref(name))
        rows.append((name.value.decode(), cin.value, cout.value, dil.value))
    assert rows[0] == ("enc_block1.conv_block.conv_layer1", 1, 32, 1)
    assert rows[8] == ("enc_block5.conv_block.conv_layer1", 256, 512, 2)
    assert rows[10] == ("enc_block6.conv_block.conv_layer1", 512, 512, 4)
    assert rows[12] == ("dec_block5.conv_block.conv_layer1", 1024, 512, 1)
    assert rows[14] == ("dec_block4.conv_block.conv_layer1", 768, 256, 1)
    assert rows[20] == ("dec_block1.conv_block.conv_layer1", 96, 32, 1)
    # the oracle's independent layer table agrees
    from oracle import pp_oracle as O
    shapes = O.unet_param_shapes(1, 32, 512, 5, 8)
    for name, cin, cout, _ in rows:
        assert shapes[name + ".conv.weight"] == (cout, cin, 3, 3)
    ws = lib.cdll.pp_unet_workspace_bytes(h, 24, 256, 256, 2)
    assert 2 * 2**30 < ws < 8 * 2**30
    assert lib.cdll.pp_unet_workspace_bytes(h, 3, 256, 256, 2) == -1  # batch not divisible into 2 groups
    assert "groups" in lib.last_error()
    act, off, C, hh, ww = ctypes.c_int(), ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib.call("pp_unet_activation", h, b"encoder/stage6", 24, 256, 256, 2, ctypes.byref(act), ctypes.byref(off),
             ctypes.byref(C), ctypes.byref(hh), ctypes.byref(ww))
    assert (C.value, hh.value, ww.value) == (512, 32, 32)
    lib.cdll.pp_unet_destroy(h)


def test_error_convention(lib):
    h = ctypes.c_void_p()
    with pytest.raises(RuntimeError, match="input_ch"):
        lib.call("pp_unet_create", 17, 32, 512, 5, 8, pplib.BF16, ctypes.byref(h))
    with pytest.raises(RuntimeError, match="output_stride"):
        lib.call("pp_unet_create", 1, 32, 512, 5, 4, pplib.BF16, ctypes.byref(h))


def test_dropin_modules_mirror_reference_state_dict():
    """Same keys, shapes and dtypes as the reference modules (checkpoints load both ways)."""
    import sys
    import torch
    from pacingpseudo_b200.dropin import DROPIN_PATH
    from oracle import pp_oracle as O
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    from models.unet import UNet
    from models.consistency_reglur_memory import ConsistencyRegulr
    import argparse
    for os_ in (8, 16, 32):
        m = UNet(1, 32, 512, 5, os_, False, False, True)
        sd = m.state_dict()
        exp = O.unet_param_shapes(1, 32, 512, 5, os_)
        assert list(sd) == list(exp)
        assert all(tuple(sd[k].shape) == tuple(exp[k]) for k in exp)
    cr = ConsistencyRegulr(
        kwargs_unet=dict(input_ch=1, init_ch=32, max_ch=512, num_classes=5, output_stride=8, is_stride_conv=False,
                         is_trans_conv=False, elab_end_points=True),
        kwargs_aux_path=dict(num_classes=5, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[512, 512],
                             hid_ch=64, aux_drop_prob=0., do_memory=True, max_step=400, update_momentum=0.9,
                             ensemble_mode='cosine_similarity'),
        args_parser=argparse.Namespace())
    sd = cr.state_dict()
    assert len(sd) == 165
    assert sum(p.numel() for p in cr.parameters()) == 20245381
    assert sum(p.numel() for p in cr.parameters() if p.requires_grad) == 20245061
    # the is_stride_conv + is_trans_conv variant (unet.py:113-116,141): stride-2 holders, ConvTranspose2d weights
    for os_ in (8, 16, 32):
        m = UNet(1, 32, 512, 4, os_, True, True, True)
        sd = m.state_dict()
        exp = O.unet_param_shapes(1, 32, 512, 4, os_, strided=True)
        assert list(sd) == list(exp)
        assert all(tuple(sd[k].shape) == tuple(exp[k]) for k in exp)
        assert m.enc_block2.conv_block.conv_layer1.conv.stride == (2, 2) and m.enc_block2.pooling is None
    with pytest.raises(AssertionError):
        UNet(is_stride_conv=True, is_trans_conv=False)   # unet.py:25
    # max_ch = 728 (train_chaos.py:71): reference-shaped parameters, kernel-side stage width padded to 1024
    from models.unet import pad_in_channels
    m = UNet(1, 32, 728, 5, 8, False, False, True)
    exp = O.unet_param_shapes(1, 32, 728, 5, 8)
    assert list(m.state_dict()) == list(exp) and all(tuple(m.state_dict()[k].shape) == tuple(exp[k]) for k in exp)
    assert m._ch_int == [32, 64, 128, 256, 512, 1024] and m._cfg[2] == 1024
    assert m.end_points.true_channels == {"encoder/stage6": 728}
    assert m._in_segments("dec_block5.conv_block.conv_layer1") == [(728, 1024), (512, 512)]
    w = torch.arange(10.0).view(2, 5, 1, 1).requires_grad_()
    p = pad_in_channels(w, [(2, 4), (3, 3)])
    assert p.view(2, -1).tolist() == [[0, 1, 0, 0, 2, 3, 4], [5, 6, 0, 0, 7, 8, 9]]
    p.sum().backward()
    assert torch.equal(w.grad, torch.ones_like(w))
    with pytest.raises(RuntimeError, match="CUDA"):
        UNet(1, 32, 512, 5, 8)(torch.zeros(1, 1, 16, 16))  # no CPU fallback


def test_unet_plan_layer_table_matches_the_modules():
    """The C-side executor's layer table (pp_unet_conv_info / pp_unet_conv_kind: host-only, no GPU needed) against the
    drop-in modules for every output stride and both UNet variants: each layer names a module whose parameter has the
    reported (Cout, Cin) and kind; workspace sizes and end-point shapes are consistent."""
    import sys
    import torch.nn as nn
    from pacingpseudo_b200.dropin import DROPIN_PATH
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    from models.unet import UNet
    for strided in (False, True):
        for os_ in (8, 16, 32):
            m = UNet(1, 32, 512, 4, os_, strided, strided, True, precision="bf16")
            eng = m.engine
            mods = m._layer_modules()
            assert len(mods) == eng.nconv == (22 + (5 if strided else 0))
            n_s2 = 0
            for (name, cin, cout, dil), (kind, scale), mod in zip(eng.layers, eng.kinds, mods):
                if kind == 2:
                    assert isinstance(mod, nn.ConvTranspose2d) and name.endswith(".up_samp")
                    assert tuple(mod.weight.shape) == (cin, cout, scale, scale) and mod.bias is None
                    continue
                assert tuple(mod.conv.weight.shape) == (cout, cin, 3, 3), name
                assert mod.conv.dilation == (dil, dil) and mod.conv.stride == ((2, 2) if kind == 1 else (1, 1)), name
                n_s2 += kind == 1
            assert n_s2 == (0 if not strided else {8: 3, 16: 4, 32: 5}[os_])
            # every trainable parameter of the module tree is covered exactly once (plus the 1x1 head)
            covered = sum(p.numel() for mod in mods for p in mod.parameters()) + sum(p.numel() for p in m.final_conv.parameters())
            assert covered == sum(p.numel() for p in m.parameters())
            assert eng.workspace_bytes(4, 64, 64, 2) > eng.workspace_bytes(2, 64, 64, 1) > 0
            _, _, C6, h6, w6 = eng.activation("encoder/stage6", 2, 64, 64, 1)
            assert (C6, h6, w6) == (512, 64 // os_, 64 // os_)
            _, _, C1, h1, w1 = eng.activation("decoder/stage1", 2, 64, 64, 1)
            assert (C1, h1, w1) == (32, 64, 64)


def test_checkpoint_round_trip_in_reference_format(tmp_path):
    """a ConsistencyRegulr checkpoint written by the drop-in loads (a) back into the drop-in, (b) into a
    plain UNet through the `backbone.` prefix stripping of inference.py:138-146, and (c) into the REFERENCE modules and
    back when the reference is present (build container) - same keys, shapes, dtypes and values."""
    import argparse
    import sys
    from collections import OrderedDict
    import torch
    from pacingpseudo_b200.dropin import DROPIN_PATH
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    from models.unet import UNet
    from models.consistency_reglur_memory import ConsistencyRegulr

    def make(cls_cr, strided=False):
        return cls_cr(
            kwargs_unet=dict(input_ch=1, init_ch=32, max_ch=512, num_classes=3, output_stride=16, is_stride_conv=strided,
                             is_trans_conv=strided, elab_end_points=True),
            kwargs_aux_path=dict(num_classes=3, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[512, 512],
                                 hid_ch=64, aux_drop_prob=0., do_memory=True, max_step=400, update_momentum=0.9,
                                 ensemble_mode='cosine_similarity'),
            args_parser=argparse.Namespace())

    for strided in (False, True):
        torch.manual_seed(3)
        cr = make(ConsistencyRegulr, strided)
        with torch.no_grad():
            cr.aux_path.memory_bank.normal_()
        path = str(tmp_path / ("ckpt_%d.pth" % strided))
        torch.save(cr.state_dict(), path)                     # train_chaos.py:420 torch.save(model.state_dict(), ...)
        state = torch.load(path)
        cr2 = make(ConsistencyRegulr, strided)
        cr2.load_state_dict(state)                            # strict
        assert all(torch.equal(a, b) for a, b in zip(cr.state_dict().values(), cr2.state_dict().values()))
        unet = UNet(1, 32, 512, 3, 16, strided, strided, True)
        with pytest.raises(RuntimeError):
            unet.load_state_dict(state)
        stripped = OrderedDict((k.partition('.')[-1], v) for k, v in state.items() if 'backbone' in k)   # inference.py:141-145
        unet.load_state_dict(stripped)
        assert all(torch.equal(stripped[k], v) for k, v in unet.state_dict().items())

    if not os.path.isdir("/root/reference"):
        return
    torch.Tensor.cuda, saved_cuda = (lambda self, *a, **k: self), torch.Tensor.cuda   # AuxPath.__init__ calls .cuda()
    mods = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split('.')[0] in ("models", "losses")}
    # the reference's `models` directory has no __init__.py: a regular package of that name anywhere on sys.path (the
    # drop-in) would win over it, so the drop-in path is taken off sys.path while the reference is imported
    sys.path.remove(DROPIN_PATH)
    sys.path.insert(0, "/root/reference")
    try:
        from models.consistency_reglur_memory import ConsistencyRegulr as RefCR
        assert "/root/reference" in sys.modules["models.consistency_reglur_memory"].__file__
        for strided in (False, True):
            ref = make(RefCR, strided)
            ref.load_state_dict(torch.load(str(tmp_path / ("ckpt_%d.pth" % strided))))       # drop-in -> reference
            back = make(ConsistencyRegulr, strided)
            back.load_state_dict(ref.state_dict())                                             # reference -> drop-in
            sd_ref = ref.state_dict()
            assert list(sd_ref) == list(back.state_dict())
            assert all(torch.equal(sd_ref[k], v) and sd_ref[k].dtype == v.dtype for k, v in back.state_dict().items())
    finally:
        sys.path.remove("/root/reference")
        sys.path.insert(0, DROPIN_PATH)
        torch.Tensor.cuda = saved_cuda
        for k in [k for k in sys.modules if k.split('.')[0] in ("models", "losses")]:
            del sys.modules[k]
        sys.modules.update(mods)


def test_graft_build_compiles_library_and_standalone_checker():
    """__graft_entry__.build() (the driver's 'does it build' check): the shared library and the standalone
    tcgen05 cross-check binary (tests/cuda/test_conv_tc.cu, linked against the kernel objects) must both build, so a
    changed kernel-launcher signature cannot silently break the binary."""
    import __graft_entry__ as g
    path = g.build()
    assert os.path.exists(path)
    assert os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "test_conv_tc"))


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pacingpseudo_b200/ may import it, and bench.py may do so only
    inside the CPU-baseline / reference-arm function."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "pacingpseudo_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
    bench = open(os.path.join(root, "bench.py")).read()
    ours = bench[bench.index("def run_ours("):bench.index("def main(")]
    assert not re.search(r"^\s*(from|import)\s+oracle\b", ours, re.M)


def test_loss_weight_ramp_up_matches_the_oracle_schedule():
    from oracle.pp_oracle import gaussian_ramp_up
    from pacingpseudo_b200.schedules import loss_weight_ramp_up
    for t in (0, 1, 40, 79, 80, 200):
        assert loss_weight_ramp_up(t, 1.0, scale=8.0) == gaussian_ramp_up(t, 1.0, scale=8.0)


def test_net_outputs_resolves_lazy_values_on_every_read_path():
    """The drop-in ConsistencyRegulr returns the reference's dict; `logits_aux_cls` is produced on first access
    (dropin/models/consistency_reglur_memory.py: NetOutputs). Every way a caller can read a dict must see the value."""
    import sys
    from pacingpseudo_b200.dropin import DROPIN_PATH
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    from models.consistency_reglur_memory import NetOutputs

    def fresh():
        made = []
        o = NetOutputs()
        o["loss_pce"] = 1.0
        o.set_lazy("logits_aux_cls", lambda: (made.append(1), "tensor")[1])
        o.update({"loss_aux_cls": 2.0})
        return o, made

    o, made = fresh()
    assert isinstance(o, dict) and "logits_aux_cls" in o and len(o) == 3 and not made     # nothing produced yet
    assert list(o) == list(o.keys()) == ["loss_pce", "logits_aux_cls", "loss_aux_cls"] and not made
    assert o["logits_aux_cls"] == "tensor" and o["logits_aux_cls"] == "tensor" and made == [1]   # produced once
    for read in (lambda d: d.get("logits_aux_cls"), lambda d: dict(d)["logits_aux_cls"], lambda d: {**d}["logits_aux_cls"],
                 lambda d: dict(d.items())["logits_aux_cls"], lambda d: d.values()[1], lambda d: d.copy()["logits_aux_cls"],
                 lambda d: d.pop("logits_aux_cls"), lambda d: (lambda t: (t.update(d), t)[1])({})["logits_aux_cls"]):
        o, made = fresh()
        assert read(o) == "tensor" and made == [1]
    o, _ = fresh()
    assert o.get("missing", 5) == 5 and o.pop("missing", 6) == 6
    with pytest.raises(KeyError):
        o.pop("missing")
