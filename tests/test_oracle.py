"""CPU: the oracle restatement (oracle/pp_oracle.py) against the golden vectors the unmodified reference
produced (oracle/gen_golden.py), and against the live reference when /root/reference is present."""
import os
import sys

import numpy as np
import pytest
import torch

import harness as Hn
from oracle import pp_oracle as O
from oracle.gen_golden import CASES

FAST_CASES = ["pacing_train_bn", "pacing_eval_bn", "pacing_acdc_kl_mean", "pacing_l1_detach", "pacing_l2_nomask",
              "baseline_pce", "upperbound_ce_dice", "unet_os16", "unet_os32",
              "unet_strided_os32", "unet_strided_os16_eval", "pacing_strided_os8", "pacing_dropout"]


@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_matches_reference_golden(name):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    rec = Hn.run_case_oracle(name)
    gold = Hn.load_golden(name)
    report = []
    fails = Hn.compare(rec, gold, Hn.TOL["oracle"], CASES[name].get("steps", 1), report)
    assert not fails, "\n".join(fails + report)


def test_oracle_loss_functions_values_and_grads():
    g = Hn.load_golden("loss_functions")
    C = g["za"].shape[1]
    mask = torch.tensor(g["mask"])
    target = torch.tensor(g["target"])
    onehot = torch.tensor(g["onehot"])

    def check(name, fn):
        za = torch.tensor(g["za"], requires_grad=True)
        zb = torch.tensor(g["zb"], requires_grad=True)
        v = fn(za, zb)
        v.backward()
        assert abs(v.item() - float(g[name])) <= 2e-6 * max(1, abs(float(g[name]))), name
        for t, key in ((za, name + "/dza"), (zb, name + "/dzb")):
            ref = g[key]
            if ref.size == 0:
                assert t.grad is None or float(t.grad.abs().max()) == 0, key
            else:
                np.testing.assert_allclose(t.grad.numpy(), ref, rtol=2e-4, atol=2e-8, err_msg=key)

    check("pce", lambda a, b: O.partial_cross_entropy(a, target, C))
    check("ce", lambda a, b: O.partial_cross_entropy(a, target.clamp(max=C - 1), -100))
    for tag, m in (("mask", mask), ("nomask", None)):
        check("ent_" + tag, lambda a, b: O.entropy_minimization(a, m))
        check("softce_" + tag, lambda a, b: O.soft_label_cross_entropy(b, torch.softmax(a, 1), m))
        check("l1_" + tag, lambda a, b: O.l1(torch.softmax(b, 1), torch.softmax(a, 1), m))
        check("l2_" + tag, lambda a, b: O.l2(torch.softmax(b, 1), torch.softmax(a, 1), m))
        check("kl_" + tag, lambda a, b: O.kl(b, a, m))
    check("dice", lambda a, b: O.dice(a, onehot))


def test_oracle_pce_all_ignored_is_nan():
    """F.cross_entropy over an empty selection is NaN."""
    z = torch.randn(1, 3, 4, 4)
    assert torch.isnan(O.partial_cross_entropy(z, torch.full((1, 4, 4), 3), 3))


def test_oracle_memory_update_rules():
    """only sample 0 is visited; first touch = plain mean; cosine mode normalises the stored row."""
    torch.manual_seed(0)
    C, hid = 3, 64
    bank = torch.zeros(C, hid, 1, 1)
    feats = torch.randn(2, hid, 4, 4)
    scrib = torch.zeros(2, C + 1, 8, 8)
    scrib[0, 0, 1, 1:5] = 1
    scrib[1, 1, 2, 2] = 1            # class 1 only in sample 1 -> must stay untouched
    O.memory_update(bank, feats, scrib, step=0, max_step=400)
    assert float(bank[1].abs().sum()) == 0 and float(bank[2].abs().sum()) == 0
    emb = O.upsample_bilinear_ac(feats[:1], (8, 8))[0][:, 1, 1:5].mean(1)
    np.testing.assert_allclose(bank[0, :, 0, 0].numpy(), emb.numpy(), rtol=1e-5, atol=1e-6)
    before = bank[0, :, 0, 0].clone()
    O.memory_update(bank, feats, scrib, step=10, max_step=400)
    m = O.ramp_up_mo(10, 400)
    e = O.upsample_bilinear_ac(feats[:1], (8, 8))[0][:, 1, 1:5].t()
    e = e / (e.norm(dim=1, keepdim=True) + 1e-8)
    r = before / (before.norm() + 1e-8)
    w = 1 - (e * r).sum(1, keepdim=True)
    exp = (1 - m) * r + m * (e * (w / (w.sum() + 1e-8))).sum(0)
    np.testing.assert_allclose(bank[0, :, 0, 0].numpy(), exp.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only exists in the build container")
def test_oracle_matches_live_reference_upsample_and_unet():
    import torch.nn.functional as F
    x = torch.randn(2, 3, 5, 7)
    for size in ((10, 14), (40, 56), (5, 7)):
        np.testing.assert_allclose(O.upsample_bilinear_ac(x, size).numpy(),
                                   F.interpolate(x, size=size, mode="bilinear", align_corners=True).numpy(),
                                   rtol=1e-5, atol=1e-6)
    sys.path.insert(0, "/root/reference")
    try:
        import importlib
        ref_unet = importlib.import_module("models.unet")
        if "pacingpseudo_b200" in (getattr(ref_unet, "__file__", "") or ""):
            pytest.skip("drop-in shadows the reference in this process")
        sd = O.synth_state_dict(O.unet_param_shapes(1, 32, 512, 3, 16), seed=3)
        m = ref_unet.UNet(1, 32, 512, 3, 16, False, False, True)
        m.load_state_dict(sd)
        m.eval()
        xin = torch.randn(1, 1, 32, 48)
        with torch.no_grad():
            ref = m(xin)
            got = O.unet_forward(sd, xin, False, output_stride=16)
        for k in ref:
            np.testing.assert_allclose(got[k].numpy(), ref[k].numpy(), rtol=1e-4, atol=1e-5, err_msg=k)
    finally:
        sys.path.remove("/root/reference")
        for mod in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "losses" or k.startswith("losses.")]:
            del sys.modules[mod]


def test_dice_metric_oracle_matches_reference_golden():
    """oracle.compute_dice_np (restating utils/metrics.py:7-34) against the vectors the reference's own compute_dice
    produced (tests/golden/dice_metric.npz, oracle/gen_golden.py): incl. nan for a class absent from prediction and
    label, 0 for a class absent from the label only, and first-maximum tie breaking."""
    import numpy as np
    from oracle import pp_oracle as O
    from oracle.gen_golden import dice_metric_inputs
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dice_metric.npz"))["dice"]
    scores, onehot = dice_metric_inputs()
    mine = np.array([O.compute_dice_np(scores[n], onehot[n]) for n in range(scores.shape[0])])
    assert np.isnan(gold[0, 4]) and gold[1, 2] == 0.0 and gold[2, 1] == 0.0
    assert np.array_equal(np.isnan(mine), np.isnan(gold))
    assert np.allclose(mine, gold, rtol=1e-12, atol=0, equal_nan=True)


def test_strong_color_augment_oracle_matches_reference_golden():
    """oracle.strong_color_augment_np (restating datasets/augmentations.py:98-166) against the outputs of the
    reference's own Brightness / Contrast / GammaAugmentation chain with pinned draws (oracle/gen_golden.py)."""
    from oracle.gen_golden import strong_augment_inputs
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "strong_augment.npz"))
    imgs, params = strong_augment_inputs()
    np.testing.assert_array_equal(params, gold["params"])
    for i in range(len(imgs)):
        got = O.strong_color_augment_np(imgs[i], params[i])
        np.testing.assert_allclose(got, gold["out"][i], rtol=1e-6, atol=1e-6, err_msg="slice %d" % i)
    np.testing.assert_array_equal(O.strong_color_augment_np(imgs[7], params[7]), imgs[7])   # nothing applied


def test_sample_strong_params_ranges():
    """Draw ranges / probabilities of TransformsColor.get_strong_transforms (chaos_aug_configs.py:63-86)."""
    from pacingpseudo_b200.data import sample_strong_params
    for strength in (1.0, 0.25):
        p = sample_strong_params(4000, strength, torch.Generator().manual_seed(0))
        lo, hi = max(0.0, 1 - 0.8 * strength), 1 + 0.8 * strength
        assert p.shape == (4000, 8) and p.dtype == torch.float32
        for col in (0, 2, 4):
            assert set(p[:, col].unique().tolist()) <= {0.0, 1.0} and abs(p[:, col].mean().item() - 0.8) < 0.03
        assert p[:, 1].abs().max() <= 0.8 * strength and p[:, 1].min() < 0 < p[:, 1].max()
        assert lo <= p[:, 3].min() and p[:, 3].max() <= hi
        assert lo <= p[:, 5].min() and p[:, 5].max() <= hi
        assert abs((p[:, 5] < 1).float().mean().item() - 0.5) < 0.04   # half of the gammas below 1 (augmentations.py:150)


def test_zero_padded_stage_width_is_an_identity():
    """max_ch = 728 runs on the kernels as a 1024-wide stage with zero weights / gamma / beta / bias on the padding
    channels (dropin/models/unet.py). On the CPU oracle: the padded 1024-wide network built with the drop-in's own
    padding helpers reproduces the 728-wide logits, end points and running statistics, train and eval BatchNorm."""
    from pacingpseudo_b200.dropin import DROPIN_PATH
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    from models.unet import UNet, pad_in_channels, pad_vector
    C = 3
    sd = O.synth_state_dict(O.unet_param_shapes(1, 32, 728, C, 16), seed=21)
    m = UNet(1, 32, 728, C, 16, False, False, True)
    padded = {}
    for (name, _cin, cout_i, _dil) in m.engine.layers:          # host-only executor table: no GPU needed
        w = pad_in_channels(sd[name + ".conv.weight"], m._in_segments(name))
        n_t = w.shape[0]
        padded[name + ".conv.weight"] = torch.cat((w, w.new_zeros((cout_i - n_t,) + tuple(w.shape[1:]))), 0)
        for key, fill in ((".conv.bias", 0.0), (".norm_op.weight", 0.0), (".norm_op.bias", 0.0),
                          (".norm_op.running_mean", 0.0), (".norm_op.running_var", 1.0)):
            padded[name + key] = pad_vector(sd[name + key], cout_i, fill)
        padded[name + ".norm_op.num_batches_tracked"] = sd[name + ".norm_op.num_batches_tracked"].clone()
    padded["final_conv.weight"], padded["final_conv.bias"] = sd["final_conv.weight"], sd["final_conv.bias"]
    assert {k: tuple(v.shape) for k, v in padded.items()} == \
        {k: tuple(v) for k, v in O.unet_param_shapes(1, 32, 1024, C, 16).items()}
    x = torch.randn(2, 1, 32, 48, generator=torch.Generator().manual_seed(4))
    for training in (True, False):
        a = {k: v.clone() for k, v in sd.items()}
        b = {k: v.clone() for k, v in padded.items()}
        ref = O.unet_forward(a, x, training, max_ch=728, output_stride=16)
        got = O.unet_forward(b, x, training, max_ch=1024, output_stride=16)
        # same mathematics, different fp32 summation order inside the CPU convolutions (other channel counts)
        np.testing.assert_allclose(got["segmentation/logits"].numpy(), ref["segmentation/logits"].numpy(), rtol=1e-4, atol=3e-5)
        np.testing.assert_allclose(got["encoder/stage6"][:, :728].numpy(), ref["encoder/stage6"].numpy(), rtol=1e-4, atol=3e-5)
        assert float(got["encoder/stage6"][:, 728:].abs().max()) == 0.0
        k = "enc_block6.conv_block.conv_layer2.norm_op.running_var"
        np.testing.assert_allclose(b[k][:728].numpy(), a[k].numpy(), rtol=1e-4)


def test_strided_variant_identities_on_the_cpu():
    """the two identities that put the stride-2 conv and ConvTranspose2d(k = s) on the stride-1 conv
    kernels, restated in torch with the index maps of csrc/ops.cu (embed_s2_weight / embed_ct_weight /
    space_depth_kernel) and checked against F.conv2d(stride=2) / F.conv_transpose2d, incl. odd block counts."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(9)

    def space_to_depth(x):      # NCHW: y[n, (sy*2+sx)*C + c, Y, X] = x[n, c, 2Y+sy, 2X+sx]
        n, c, h, w = x.shape
        return x.view(n, c, h // 2, 2, w // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, h // 2, w // 2)

    def depth_to_space(y, S):   # inverse, channel (a*S+b)*C + c
        n, cc, h, w = y.shape
        c = cc // (S * S)
        return y.view(n, S, S, c, h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, c, h * S, w * S)

    def k_of(t, sub):           # ops.cu s2_k_of: embedded tap, sub-row -> original tap (or None: zero)
        return (0 if sub == 1 else None) if t == 0 else ((1 + sub) if t == 1 else None)

    cout, c = 5, 3
    w = torch.randn(cout, c, 3, 3, generator=g, dtype=torch.float64)
    we = torch.zeros(cout, 4 * c, 3, 3, dtype=torch.float64)
    for q in range(4):
        for ty in range(3):
            for tx in range(3):
                ky, kx = k_of(ty, q >> 1), k_of(tx, q & 1)
                if ky is not None and kx is not None:
                    we[:, q * c:(q + 1) * c, ty, tx] = w[:, :, ky, kx]
    assert int((we != 0).any(dim=(0, 1)).sum()) == 4            # only the taps (dy, dx) in {-1, 0}^2 carry weights
    for h, wd in ((8, 8), (6, 10), (2, 2)):
        x = torch.randn(2, c, h, wd, generator=g, dtype=torch.float64)
        ref = F.conv2d(x, w, stride=2, padding=1)
        got = F.conv2d(space_to_depth(x), we, padding=1)
        assert torch.allclose(got, ref, atol=1e-12), (h, wd)

    cin, cout = 4, 3
    for S in (1, 2):
        wt = torch.randn(cin, cout, S, S, generator=g, dtype=torch.float64)
        we = torch.zeros(S * S * cout, cin, 3, 3, dtype=torch.float64)
        for a in range(S):
            for b in range(S):
                q = a * S + b
                we[q * cout:(q + 1) * cout, :, 1, 1] = wt[:, :, a, b].t()
        x = torch.randn(2, cin, 5, 7, generator=g, dtype=torch.float64)
        ref = F.conv_transpose2d(x, wt, stride=S)
        got = depth_to_space(F.conv2d(x, we, padding=1), S)
        assert torch.allclose(got, ref, atol=1e-12), S
        assert torch.allclose(O.conv_transpose_ks(x, wt), ref, atol=1e-12)
