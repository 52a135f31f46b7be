"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the reference does not travel to the GPU box):
    python oracle/gen_golden.py
Inputs are NOT stored: they are regenerated from seeds by pacingpseudo_b200.synth.make_batch and
oracle.pp_oracle.synth_state_dict (both committed), so a fixture is a few KB of outputs: the five loss
terms, the total loss, logits samples + checksums, per-parameter gradient norms + samples, BatchNorm
running-stat checksums, and the memory bank after each step. Records torch version in the file.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PP_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import pp_oracle as O  # noqa: E402
from pacingpseudo_b200.synth import make_batch  # noqa: E402

CASES = {
    # name: dict(kind, N, C, H, W, output_stride, training, flags...)
    "pacing_train_bn": dict(kind="pacing", N=2, C=5, H=64, W=64, os=8, training=True, cr="ce_loss", mode="cosine_similarity", steps=3),
    "pacing_eval_bn": dict(kind="pacing", N=2, C=5, H=64, W=64, os=8, training=False, cr="ce_loss", mode="cosine_similarity", steps=2),
    "pacing_acdc_kl_mean": dict(kind="pacing", N=3, C=4, H=56, W=56, os=8, training=True, cr="kl_loss", mode="mean", steps=2),
    "pacing_l1_detach": dict(kind="pacing", N=3, C=2, H=64, W=64, os=8, training=True, cr="l1_loss", mode="cosine_similarity", steps=1, detach=True),
    "pacing_l2_nomask": dict(kind="pacing", N=3, C=5, H=64, W=64, os=8, training=True, cr="l2_loss", mode="cosine_similarity", steps=1, nomask=True),
    "baseline_pce": dict(kind="baseline", N=2, C=5, H=64, W=64, os=8, training=True),
    "upperbound_ce_dice": dict(kind="upper", N=2, C=5, H=64, W=64, os=8, training=True),
    "unet_os16": dict(kind="baseline", N=2, C=4, H=64, W=64, os=16, training=True),
    "unet_os32": dict(kind="baseline", N=2, C=4, H=64, W=64, os=32, training=True),
    # is_stride_conv + is_trans_conv (unet.py:113-116,141): stride-2 first convs, ConvTranspose2d up-sampling
    "unet_strided_os32": dict(kind="baseline", N=2, C=4, H=64, W=64, os=32, training=True, strided=True),
    "unet_strided_os16_eval": dict(kind="upper", N=2, C=5, H=64, W=64, os=16, training=False, strided=True),
    # aux_drop_prob = 0.5 with PINNED Dropout2d masks (F.dropout2d replaced by O.synth_drop_factors in call order)
    "pacing_dropout": dict(kind="pacing", N=3, C=5, H=64, W=64, os=8, training=True, cr="ce_loss",
                           mode="cosine_similarity", steps=2, drop_p=0.5),
    "pacing_strided_os8": dict(kind="pacing", N=2, C=5, H=64, W=64, os=8, training=True, cr="ce_loss",
                               mode="cosine_similarity", steps=2, strided=True),
}


def sample_idx(numel, k=16):
    g = torch.Generator().manual_seed(numel % 9973 + 17)
    return torch.randint(0, numel, (min(k, numel),), generator=g)


def summarize(t):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), t.norm().item()]), t[sample_idx(t.numel())].numpy()


def build_state(case):
    shapes = {}
    for k, s in O.unet_param_shapes(1, 32, 512, case["C"], case["os"], strided=bool(case.get("strided"))).items():
        shapes["backbone." + k if case["kind"] == "pacing" else k] = s
    if case["kind"] == "pacing":
        for k, s in O.aux_param_shapes(case["C"], (512, 512), 64).items():
            shapes["aux_path." + k] = s
    return O.synth_state_dict(shapes, seed=7)


def ref_args(case):
    return argparse.Namespace(
        ignored_index=case["C"], do_loss_ent=True, do_decoder_consistency=True, detach_weak_cr=bool(case.get("detach")),
        loss_cr_variants=case.get("cr", "ce_loss"), do_aux_path=True, do_memory=True)


def run_reference(name, case):
    from models.unet import UNet
    from models.consistency_reglur_memory import ConsistencyRegulr
    from losses import losses as RL
    sd = build_state(case)
    rec = {"torch_version": np.array(torch.__version__)}
    C = case["C"]
    if case["kind"] == "pacing":
        model = ConsistencyRegulr(
            kwargs_unet=dict(input_ch=1, init_ch=32, max_ch=512, num_classes=C, output_stride=case["os"],
                             is_stride_conv=bool(case.get("strided")), is_trans_conv=bool(case.get("strided")),
                             elab_end_points=True),
            kwargs_aux_path=dict(num_classes=C, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[512, 512],
                                 hid_ch=64, aux_drop_prob=float(case.get("drop_p", 0.)), do_memory=True, max_step=400,
                                 update_momentum=0.9, ensemble_mode=case["mode"]),
            args_parser=ref_args(case))
    else:
        model = UNet(1, 32, 512, C, case["os"], bool(case.get("strided")), bool(case.get("strided")), True)
    model.load_state_dict(sd, strict=True)
    model.train(case["training"])
    for step in range(case.get("steps", 1)):
        batch = make_batch(case["N"], C, case["H"], case["W"], seed=100 + step,
                           absent_class_in_sample0=(1 if step == 1 else None))
        if case.get("nomask"):
            batch.pop("valid_mask")
        model.zero_grad()
        if case["kind"] == "pacing":
            if case.get("drop_p"):   # pin the masks: the reference's nn.Dropout2d calls F.dropout2d, in this order
                queue = list(O.synth_drop_factors(500 + step, case["N"], 1024, 64, C, case["drop_p"]))

                def pinned_dropout2d(x, p=0.5, training=True, inplace=False, _q=queue):
                    f = _q.pop(0)
                    assert training and tuple(f.shape) == tuple(x.shape[:2]), (f.shape, x.shape)
                    return x * f[:, :, None, None]
                orig_dropout2d = torch.nn.functional.dropout2d
                torch.nn.functional.dropout2d = pinned_dropout2d
            out = model({k: v for k, v in batch.items() if k != "label"}, mode="train", step=step * 37)
            if case.get("drop_p"):
                torch.nn.functional.dropout2d = orig_dropout2d
                assert not queue, "the reference made fewer dropout calls than expected"
            loss = O.total_loss(out, epoch=40)
            for k in ("loss_pce", "loss_ent", "loss_cr", "loss_aux_cls", "loss_memory"):
                rec["s%d/%s" % (step, k)] = np.array(out[k].item())
            for k in ("segmentation/logits", "segmentation/logits_strong", "logits_aux_cls"):
                rec["s%d/%s/sum" % (step, k)], rec["s%d/%s/samples" % (step, k)] = summarize(out[k])
            rec["s%d/argmax_weak" % step] = out["segmentation/logits"].argmax(1).to(torch.uint8).numpy()
        else:
            logits = model(batch["image"])["segmentation/logits"]
            if case["kind"] == "baseline":
                loss = RL.partial_cross_entropy_loss(logits, batch["scribble"].argmax(1), C)
                rec["s%d/loss_pce" % step] = np.array(loss.item())
            else:
                target = batch["label"].argmax(1)
                lce = RL.partial_cross_entropy_loss(logits, target, C)
                ldice = RL.dice_loss_fn(logits, batch["label"])
                rec["s%d/loss_ce" % step] = np.array(lce.item())
                rec["s%d/loss_dice" % step] = np.array(ldice.item())
                loss = lce + ldice
            rec["s%d/segmentation/logits/sum" % step], rec["s%d/segmentation/logits/samples" % step] = summarize(logits)
            rec["s%d/argmax_weak" % step] = logits.argmax(1).to(torch.uint8).numpy()
        rec["s%d/total" % step] = np.array(loss.item())
        loss.backward()
        names, norms, samples = [], [], []
        for k, p in model.named_parameters():
            if p.grad is None:
                continue
            names.append(k)
            s, smp = summarize(p.grad)
            norms.append(s)
            samples.append(np.pad(smp, (0, 16 - len(smp))))
        rec["s%d/grad_names" % step] = np.array(names)
        rec["s%d/grad_sums" % step] = np.stack(norms)
        rec["s%d/grad_samples" % step] = np.stack(samples)
        if case["kind"] == "pacing":
            rec["s%d/memory_bank" % step] = model.aux_path.memory_bank.detach().double().numpy().reshape(C, 64)
    buf = {k: v for k, v in model.state_dict().items() if "running" in k}
    rec["running_names"] = np.array(list(buf))
    rec["running_sums"] = np.stack([summarize(v)[0] for v in buf.values()])
    return rec


def loss_function_vectors():
    """Direct known-answer vectors for every called function of losses/losses.py, values and gradients."""
    from losses import losses as RL
    g = torch.Generator().manual_seed(5)
    N, C, H, W = 3, 5, 9, 7
    rec = {}
    za = (2 * torch.randn(N, C, H, W, generator=g)).requires_grad_()
    zb = (2 * torch.randn(N, C, H, W, generator=g)).requires_grad_()
    mask = (torch.rand(N, 1, H, W, generator=g) > 0.3).float()
    target = torch.randint(0, C + 1, (N, H, W), generator=g)
    onehot = torch.nn.functional.one_hot(torch.randint(0, C, (N, H, W), generator=g), C).permute(0, 3, 1, 2).float()
    rec["za"], rec["zb"], rec["mask"], rec["target"], rec["onehot"] = za.detach().numpy(), zb.detach().numpy(), mask.numpy(), target.numpy(), onehot.numpy()

    def emit(name, fn):
        for t in (za, zb):
            t.grad = None
        v = fn()
        v.backward()
        rec[name] = np.array(v.item())
        rec[name + "/dza"] = za.grad.numpy().copy() if za.grad is not None else np.zeros(0)
        rec[name + "/dzb"] = zb.grad.numpy().copy() if zb.grad is not None else np.zeros(0)

    emit("pce", lambda: RL.partial_cross_entropy_loss(za, target, C))
    emit("ce", lambda: RL.cross_entropy_loss(za, target.clamp(max=C - 1)))
    for tag, m in (("mask", mask), ("nomask", None)):
        emit("ent_" + tag, lambda: RL.entropy_minimization_loss(za, m))
        emit("softce_" + tag, lambda: RL.soft_label_cross_entropy_loss(zb, torch.softmax(za, 1), m))
        emit("l1_" + tag, lambda: RL.l1_loss(torch.softmax(zb, 1), torch.softmax(za, 1), m))
        emit("l2_" + tag, lambda: RL.l2_loss(torch.softmax(zb, 1), torch.softmax(za, 1), m))
        emit("kl_" + tag, lambda: RL.kl_loss(zb, za, m))
    emit("dice", lambda: RL.dice_loss_fn(za, onehot))
    return rec


def dice_metric_inputs():
    """Seeded inputs of the Dice-metric vectors (shared with the tests): softmax scores and one-hot labels of 4 samples,
    with a class absent from prediction and label (-> nan), one absent from the label only, and exact score ties."""
    g = torch.Generator().manual_seed(11)
    N, C, H, W = 4, 5, 24, 20
    z = 2 * torch.randn(N, C, H, W, generator=g)
    z[0, 4] = -50.0                       # class 4 never predicted in sample 0 ...
    z[1, 2] = -50.0
    z[3, :, :4] = 0.25                    # exact ties: the first maximum wins
    lab = torch.randint(0, C, (N, H, W), generator=g)
    lab[0][lab[0] == 4] = 0               # ... and absent from its label: nan
    lab[2][lab[2] == 1] = 3               # absent from the label only: 0
    onehot = torch.nn.functional.one_hot(lab, C).permute(0, 3, 1, 2).float()
    return torch.softmax(z, 1).numpy(), onehot.numpy()


def dice_metric_vectors():
    """Known answers of utils/metrics.py compute_dice (the reference's own function) on the seeded inputs."""
    from utils.metrics import compute_dice
    scores, onehot = dice_metric_inputs()
    return {"dice": np.array([compute_dice(scores[n], onehot[n]) for n in range(scores.shape[0])], dtype=np.float64)}


def strong_augment_inputs():
    """Seeded slices (base-transformed: mean/std normalised, zero outside a valid rectangle) and per-slice draws for
    the strong colour chain: every apply / skip combination of the three transforms, gamma below and above 1."""
    from pacingpseudo_b200.synth import make_batch
    imgs = make_batch(8, 5, 48, 40, seed=77)["image"][:, 0].numpy().astype(np.float32)
    params = np.array([
        # apply_b, b, apply_c, a, apply_g, gamma, 0, 0
        [1, 0.55, 1, 1.62, 1, 0.43, 0, 0],
        [1, -0.7, 1, 0.31, 1, 1.71, 0, 0],
        [0, 0.0, 1, 1.2, 1, 1.3, 0, 0],
        [1, 0.2, 0, 1.0, 1, 0.8, 0, 0],
        [1, -0.3, 1, 0.75, 0, 1.0, 0, 0],
        [0, 0.0, 0, 1.0, 1, 0.25, 0, 0],
        [1, 0.79, 0, 1.0, 0, 1.0, 0, 0],
        [0, 0.0, 0, 1.0, 0, 1.0, 0, 0],
    ], dtype=np.float32)
    return imgs, params


def strong_augment_vectors():
    """The reference's own Brightness / Contrast / GammaAugmentation classes (TransformsColor(strength=1) chain) run on
    the seeded slices with np.random.uniform replaced by the pinned draws, in the order the classes consume them."""
    import types
    for name in ("skimage", "skimage.transform"):      # imported by datasets/augmentations.py:8, unused by these classes
        sys.modules.setdefault(name, types.ModuleType(name))
    # the HuggingFace `datasets` package in site-packages shadows the reference's (namespace) `datasets` directory, so
    # its two files are loaded by path under that name: augmentations.py, then chaos_aug_configs.py (which does
    # `from datasets.augmentations import *`); the chain is TransformsColor(strength=1).strong_transforms itself
    import importlib.util
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "datasets" or k.startswith("datasets.")}
    try:
        pkg = types.ModuleType("datasets")
        pkg.__path__ = [os.path.join(REF, "datasets")]
        sys.modules["datasets"] = pkg
        for name, rel in (("datasets.augmentations", "augmentations.py"),
                          ("datasets.chaos_aug_configs", os.path.join("chaos", "chaos_aug_configs.py"))):
            spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "datasets", rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
        chain = sys.modules["datasets.chaos_aug_configs"].TransformsColor(1.0).strong_transforms
    finally:
        for k in [k for k in sys.modules if k == "datasets" or k.startswith("datasets.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    imgs, params = strong_augment_inputs()
    outs = []
    orig = np.random.uniform
    try:
        for img, p in zip(imgs, params):
            ab, b, ac, a, ag, gamma = p[:6]
            draws = [0.0 if ab else 0.99] + ([float(b)] if ab else []) + [0.0 if ac else 0.99] + ([float(a)] if ac else [])
            if ag:   # apply draw, the branch draw (< 0.5 -> gamma from [lo, 1)), the gamma draw
                draws += [0.0, 0.0 if gamma < 1 else 0.99, float(gamma)]
            else:
                draws += [0.99]
            np.random.uniform = lambda *a_, _q=draws, **k_: _q.pop(0)
            data = {"image": img.copy()}
            for t in chain:
                data = t(data)
            assert not draws, "the reference consumed fewer draws than expected"
            outs.append(np.asarray(data["image"], dtype=np.float32))
    finally:
        np.random.uniform = orig
    return {"out": np.stack(outs), "params": params}


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference not found at %s" % REF)
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self  # AuxPath.__init__ calls .cuda() (aux_path_memory.py:44)
    torch.set_num_threads(8)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    if len(sys.argv) == 1:
        np.savez_compressed(os.path.join(out_dir, "loss_functions.npz"), **loss_function_vectors())
        print("wrote loss_functions.npz")
        np.savez_compressed(os.path.join(out_dir, "dice_metric.npz"), **dice_metric_vectors())
        print("wrote dice_metric.npz")
    if len(sys.argv) == 1 or "strong_augment" in sys.argv[1:]:
        np.savez_compressed(os.path.join(out_dir, "strong_augment.npz"), **strong_augment_vectors())
        print("wrote strong_augment.npz")
    only = [a for a in sys.argv[1:] if not a.startswith("-")]   # optional: regenerate the named cases only
    for name, case in CASES.items():
        if only and name not in only:
            continue
        rec = run_reference(name, case)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print("wrote", name, {k: float(v) for k, v in rec.items() if k.startswith("s0/loss")})


if __name__ == "__main__":
    main()
