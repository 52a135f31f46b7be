"""Exploration (not a test): how the bf16-vs-fp32 distances of one step evolve with the amount of training.
Reference here = the library's own fp32 mode (pinned to the CPU oracle at ~1e-6 by test_gpu_fullsize.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import fullsize as FS

names = sys.argv[1].split(",")
steps_list = [int(x) for x in sys.argv[2].split(",")]
lr = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
for name in names:
    cfg = FS.CONFIGS[name]
    for steps in steps_list:
        sd, info = FS.train_state(cfg, steps, lr=lr)
        print("[train] %s %s" % (name, FS.fmt(info)), flush=True)
        for bname, batch in (("held-out", FS.step_batch(cfg, 911, True)), ("pool", FS.step_batch(cfg, 700, True))):
            for bn in (True, False):
                ref = FS.cuda_step(sd, cfg, batch, bn, "fp32")
                rec = FS.cuda_step(sd, cfg, batch, bn, "bf16")
                m = FS.distances(rec, ref, bn)
                keep = {k: v for k, v in m.items() if k.startswith(("logits_", "argmax_", "grad_all", "grad_median", "grad_worst", "loss_pce", "total"))}
                print("  steps=%d %s bn=%s: %s" % (steps, bname, "train" if bn else "eval", FS.fmt(keep)), flush=True)
