"""Full-size parity machinery (BASELINE.json configs 1-4 at their real sizes, 12 slices of 256^2 / 224^2 per GPU).

One training step of the drop-in modules on cuda:0 (through the C ABI) is compared with the fp32 CPU oracle on the same
state dict and batch: every logits tensor the step returns, every loss term, every parameter gradient, the memory bank
after its update, and the arg-max pseudo-labels / predicted masks. Used by tests/test_gpu_fullsize.py; every comparison
appends its measured distances to gpurun_out/r02_parity_fullsize.txt (copied into profiles/ after a GPU run), so the
numbers behind the assertions are committed, not just the dots of `pytest -q`.

States:
  * "init": oracle.gen_golden.build_state (seeded He-initialised weights; what the golden fixtures use);
  * "trained": the same network after `steps` Adam steps of the SAME workload on the GPU (bf16 path, FlatAdam,
    lr 1e-3) on a pool of class-distinct synthetic slices (synth.make_batch(class_contrast=True)) — the state the
    north-star tolerances are meaningful on (at the seeded initial state the network amplifies
    a 1e-3 perturbation ~40x, so no bf16 implementation can land within 2e-2 of the fp32 logits there).
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pp_oracle as O  # noqa: E402
from oracle.gen_golden import build_state  # noqa: E402
from pacingpseudo_b200.synth import make_batch  # noqa: E402

REPORT = os.path.join(ROOT, "gpurun_out", "r02_parity_fullsize.txt")
LOSS_KEYS = ("loss_pce", "loss_ent", "loss_cr", "loss_aux_cls", "loss_memory")
LOGIT_KEYS = ("segmentation/logits", "segmentation/logits_strong", "logits_aux_cls")

# BASELINE.json configs at full size. kind: pacing = ConsistencyRegulr full step, baseline = UNet + pCE,
# upper = UNet + CE + Dice on dense labels (upper_bound_chaos.py:157-171).
CONFIGS = {
    "config1_baseline_256_C5": dict(kind="baseline", C=5, os=8, N=12, S=256),
    "config2_pacing_256_C5": dict(kind="pacing", C=5, os=8, N=12, S=256, cr="ce_loss", mode="cosine_similarity"),
    "config3_pacing_acdc_224_C4": dict(kind="pacing", C=4, os=8, N=12, S=224, cr="ce_loss", mode="cosine_similarity"),
    "config3_pacing_acdc_256_C4": dict(kind="pacing", C=4, os=8, N=12, S=256, cr="ce_loss", mode="cosine_similarity"),
    "config4_upper_256_C5": dict(kind="upper", C=5, os=8, N=12, S=256),
    "config5_pacing_lvsc_224_C2": dict(kind="pacing", C=2, os=8, N=12, S=224, cr="ce_loss", mode="cosine_similarity"),
}


def log(line):
    print(line, flush=True)
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(line + "\n")


def rel(a, b):
    a, b = a.detach().double().cpu().flatten(), b.detach().double().cpu().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def learnable_names(sd):
    return [k for k in sd if sd[k].is_floating_point() and "running" not in k and not k.endswith("memory_bank")]


def step_batch(cfg, seed, class_contrast):
    b = make_batch(cfg["N"], cfg["C"], cfg["S"], cfg["S"], seed=seed, class_contrast=class_contrast)
    return b


# ------------------------------------------------------------------------------------------------
# one step on the CPU oracle
# ------------------------------------------------------------------------------------------------
def oracle_step(sd, cfg, batch, bn_training, quant=False, epoch=40, step=40):
    """-> dict(logits..., losses..., grads {name: tensor}, bank, total). fp32 on the host cores."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    C = cfg["C"]
    s_ = {k: v.detach().clone() for k, v in sd.items()}
    names = learnable_names(s_)
    for k in names:
        s_[k].requires_grad_(True)
    rec = {}
    if cfg["kind"] == "pacing":
        scfg = O.StepConfig(num_classes=C, ignored_index=C, loss_cr_variants=cfg.get("cr", "ce_loss"),
                            ensemble_mode=cfg.get("mode", "cosine_similarity"), output_stride=cfg["os"], quant=quant)
        out = O.consistency_forward(s_, {k: v for k, v in batch.items() if k != "label"}, scfg, mode="train",
                                    step=step, training=bn_training)
        loss = O.total_loss(out, epoch=epoch)
        for k in LOSS_KEYS + LOGIT_KEYS:
            rec[k] = out[k].detach()
        rec["bank"] = s_["aux_path.memory_bank"].detach().reshape(C, -1).clone()
    else:
        z = O.unet_forward(s_, batch["image"], bn_training, output_stride=cfg["os"], quant=quant)["segmentation/logits"]
        rec["segmentation/logits"] = z.detach()
        if cfg["kind"] == "baseline":
            loss = O.partial_cross_entropy(z, batch["scribble"].argmax(1), C)
            rec["loss_pce"] = loss.detach()
        else:
            lce = O.partial_cross_entropy(z, batch["label"].argmax(1), C)
            ld = O.dice(z, batch["label"])
            rec["loss_ce"], rec["loss_dice"] = lce.detach(), ld.detach()
            loss = lce + ld
    rec["total"] = loss.detach()
    loss.backward()
    rec["grads"] = {k: s_[k].grad.detach() for k in names if s_[k].grad is not None}
    return rec


# ------------------------------------------------------------------------------------------------
# the drop-in modules on cuda:0
# ------------------------------------------------------------------------------------------------
def build_model(cfg, precision, sd=None, bn_training=True, device="cuda"):
    from pacingpseudo_b200.dropin import DROPIN_PATH
    if DROPIN_PATH not in sys.path:
        sys.path.insert(0, DROPIN_PATH)
    os.environ["PP_PRECISION"] = precision
    from models.unet import UNet
    from models.consistency_reglur_memory import ConsistencyRegulr
    C = cfg["C"]
    if cfg["kind"] == "pacing":
        ns = argparse.Namespace(ignored_index=C, do_loss_ent=True, do_decoder_consistency=True, detach_weak_cr=False,
                                loss_cr_variants=cfg.get("cr", "ce_loss"), do_aux_path=True, do_memory=True)
        model = ConsistencyRegulr(
            kwargs_unet=dict(input_ch=1, init_ch=32, max_ch=512, num_classes=C, output_stride=cfg["os"],
                             is_stride_conv=False, is_trans_conv=False, elab_end_points=True, precision=precision),
            kwargs_aux_path=dict(num_classes=C, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[512, 512],
                                 hid_ch=64, aux_drop_prob=0., do_memory=True, max_step=400, update_momentum=0.9,
                                 ensemble_mode=cfg.get("mode", "cosine_similarity")),
            args_parser=ns)
    else:
        model = UNet(1, 32, 512, C, cfg["os"], False, False, True, precision=precision)
    model.load_state_dict(sd if sd is not None else build_state(cfg), strict=True)
    return model.to(device).train(bn_training)


def model_loss(model, cfg, batch, epoch=40, step=40):
    """-> (total loss tensor, record of outputs) for one forward pass of the drop-in modules."""
    from losses import losses as DL
    C = cfg["C"]
    rec = {}
    if cfg["kind"] == "pacing":
        out = model({k: v for k, v in batch.items() if k != "label"}, mode="train", step=step)
        for k in LOSS_KEYS:
            rec[k] = out[k].detach().clone()   # before the caller-style in-place weighting below
        for k in LOGIT_KEYS:
            rec[k] = out[k].detach()
        loss = O.total_loss(out, epoch=epoch)
    else:
        z = model(batch["image"])["segmentation/logits"]
        rec["segmentation/logits"] = z.detach()
        if cfg["kind"] == "baseline":
            loss = DL.partial_cross_entropy_loss(z, batch["scribble"].argmax(1), C)
            rec["loss_pce"] = loss.detach().clone()
        else:
            lce = DL.partial_cross_entropy_loss(z, batch["label"].argmax(1), C)
            ld = DL.dice_loss_fn(z, batch["label"])
            rec["loss_ce"], rec["loss_dice"] = lce.detach().clone(), ld.detach().clone()
            loss = lce + ld
    rec["total"] = loss.detach().clone()
    return loss, rec


def cuda_step(sd, cfg, batch, bn_training, precision, device="cuda"):
    model = build_model(cfg, precision, sd, bn_training, device)
    dbatch = {k: v.to(device) for k, v in batch.items()}
    model.zero_grad(set_to_none=True)
    loss, rec = model_loss(model, cfg, dbatch)
    loss.backward()
    torch.cuda.synchronize()
    rec = {k: (v.float().cpu() if isinstance(v, torch.Tensor) else v) for k, v in rec.items()}
    rec["grads"] = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    if cfg["kind"] == "pacing":
        rec["bank"] = model.aux_path.memory_bank.detach().float().cpu().reshape(cfg["C"], -1)
    del model
    torch.cuda.empty_cache()
    return rec


def train_state(cfg, steps, lr=1e-3, pool=6, precision="bf16", device="cuda"):
    """`steps` Adam steps of the workload on the GPU from the seeded initial state -> CPU state dict in the reference's
    checkpoint format (model.state_dict(): same keys / shapes / dtypes as the reference modules)."""
    from pacingpseudo_b200.optim import FlatAdam
    model = build_model(cfg, precision, None, True, device)
    opt = FlatAdam(model.parameters(), lr=lr, weight_decay=3e-4)
    batches = [{k: v.to(device) for k, v in step_batch(cfg, 700 + i, True).items()} for i in range(pool)]
    t0 = time.time()
    first = last = None
    for it in range(steps):
        loss, _ = model_loss(model, cfg, batches[it % pool])
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it == 0:
            first = loss.item()
    last = loss.item()
    torch.cuda.synchronize()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    del model, opt
    torch.cuda.empty_cache()
    return sd, dict(steps=steps, lr=lr, loss_first=first, loss_last=last, seconds=time.time() - t0)


# ------------------------------------------------------------------------------------------------
# distances
# ------------------------------------------------------------------------------------------------
def distances(rec, ref, bn_training):
    """-> flat dict of the distances the north star names (all relative to `ref`)."""
    m = {}
    for k in LOGIT_KEYS:
        if k in ref:
            short = {"segmentation/logits": "weak", "segmentation/logits_strong": "strong", "logits_aux_cls": "aux"}[k]
            m["logits_" + short] = rel(rec[k], ref[k])
            same = rec[k].argmax(1).cpu() == ref[k].argmax(1).cpu()
            m["argmax_" + short] = float(same.float().mean())
            # pixels whose REFERENCE decision is not inside the logit tolerance band: top-2 margin > 2e-2 * max |z|
            # (a flip inside the band is what the logit tolerance itself permits; object boundaries always hold a few
            # 1e-4 of the pixels there, and 8x more for the aux logits, which are a bilinear x8 up-sampling)
            top2 = ref[k].float().topk(2, dim=1).values
            decided = (top2[:, 0] - top2[:, 1]) > 2e-2 * ref[k].abs().max()
            m["argmax_%s_decided" % short] = float(same[decided].float().mean()) if bool(decided.any()) else 1.0
            m["decided_frac_" + short] = float(decided.float().mean())
    for k in LOSS_KEYS + ("loss_ce", "loss_dice", "total"):
        if k in ref:
            # relative, with the same 1e-2 floor on the loss magnitude as tests/harness.py: a trained pCE of ~1e-3
            # carries ~1e-7 of fp32 summation noise, which is 1e-4 RELATIVE without saying anything about parity
            m[k] = abs(float(rec[k]) - float(ref[k])) / max(abs(float(ref[k])), 1e-2)
            m["abs_" + k] = abs(float(rec[k]) - float(ref[k]))   # reported; trained terms are also held absolutely
    if "total" in ref:
        m["ref_total"] = abs(float(ref["total"]))
    if "bank" in ref:
        m["bank"] = rel(rec["bank"], ref["bank"])
    g, g0 = rec["grads"], ref["grads"]
    names = [k for k in g0 if k in g]
    gmax = max(float(g0[k].norm()) for k in names)
    # a conv bias in front of a batch-statistics BatchNorm has a mathematically zero gradient (noise / noise)
    live = [k for k in names if float(g0[k].norm()) > 1e-6 * gmax and not (bn_training and k.endswith("conv.bias"))
            and not (bn_training and k.endswith("layer_bottleneck.1.bias"))]
    num = sum(float((g[k].double() - g0[k].double()).norm() ** 2) for k in live) ** 0.5
    den = sum(float(g0[k].double().norm() ** 2) for k in live) ** 0.5
    per = sorted((rel(g[k], g0[k]), k) for k in live)
    m["grad_all"] = num / den
    m["grad_median"] = per[len(per) // 2][0]
    m["grad_p90"] = per[int(0.9 * (len(per) - 1))][0]
    m["grad_worst"] = per[-1][0]
    m["grad_worst_name"] = per[-1][1]
    m["grad_missing"] = len([k for k in g0 if k not in g])
    return m


def fmt(m):
    return "  ".join("%s=%s" % (k, ("%.3e" % v) if isinstance(v, float) else v) for k, v in m.items())
