#!/usr/bin/env python
"""bench.py — PacingPseudo train-step throughput (BASELINE.json metric: train imgs/sec, 256x256 pacingpseudo step).

    python bench.py --gpus N --steps K --warmup W            # this repo (hand-written sm_100a path)
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference modules on the host cores
    python bench.py --workload {pacing,baseline,upperbound} --size 224 --classes 4 --batch 96 ...   # configs 1-5

Workloads (BASELINE.json configs): `pacing` (default, configs 2/3/5) = one full pacingpseudo iteration
(train_chaos.py:263-315): batched weak+strong UNet forward, partial CE + entropy + consistency + aux partial CE +
memory loss, memory-bank update, backward, gradient all-reduce (N > 1), Adam; one "image" = one weak+strong pair, as
batch_size counts them (train_chaos.py:237). `baseline` (config 1 on the GPU) = UNet + partial CE + Adam.
`upperbound` (config 4) = UNet + CE + Dice on dense labels (upper_bound_chaos.py:157-171). Prints ONE JSON line.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_RESULT_OUT = sys.stdout

GF_PER_PAIR = {  # algorithmic conv FLOPs per weak+strong pair, BASELINE.md section 3 (fwd F, bwd 2F)
    (256, 5): 348.25e9, (256, 4): 348.22e9, (224, 4): 266.61e9, (224, 2): 266.57e9,
}
GF_PER_IMAGE = {  # baseline / upperbound step = 3 F_unet per image
    (256, 5): 172.31e9, (256, 4): 172.30e9, (224, 4): 131.92e9, (224, 2): 131.90e9,
}
WORKLOADS = {
    "pacing": "pacingpseudo full step (--do_loss_ent --do_decoder_consistency --do_aux_path --do_memory)",
    "baseline": "baseline UNet + partial cross-entropy on scribbles",
    "upperbound": "upperbound UNet + cross-entropy + Dice on dense labels (upper_bound_chaos.py)",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pacing", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=12, help="weak+strong pairs per GPU (train_chaos.py:93)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--classes", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--bn", default="train", choices=["train", "eval"],
                    help="BatchNorm regime: batch statistics (epoch 0) or running statistics (epochs >= 1, SURVEY T2)")
    ap.add_argument("--epoch", type=int, default=40, help="epoch used for loss ramp-up weights and bank momentum")
    ap.add_argument("--unet-variant", default="maxpool", choices=["maxpool", "strided"],
                    help="strided: is_stride_conv + is_trans_conv (unet.py:113-116,141); not the headline config")
    ap.add_argument("--output-stride", type=int, default=8, choices=[8, 16, 32])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-profile-pass", action="store_true", help="skip the per-launch CUDA-event pass (ncu runs)")
    ap.add_argument("--no-same-box", action="store_true", help="skip the cuDNN same-box comparison (N=1 only)")
    ap.add_argument("--no-other-bn", action="store_true", help="skip the timing of the other BatchNorm regime")
    ap.add_argument("--ref-max-seconds", type=float, default=240.0,
                    help="reference arm: if K+W steps at the full per-GPU batch would exceed this, each step becomes a "
                         "bounded sample (fewer slices per step, stated in config / sample)")
    return ap.parse_args()


def load_traffic():
    """DRAM bytes per launch of the dominant kernel family from the committed ncu pass (profiles/r0N_conv_traffic.json)."""
    for name in ("r02_conv_traffic.json", "r01_conv_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return json.load(f)
        except Exception:
            continue
    return {}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline / same-box cuDNN: the UNMODIFIED reference modules (baseline/_ref, oracle/fetch_ref.py)
# ------------------------------------------------------------------------------------------------
def gaussian_ramp_up(t, base_value, max_t=80, scale=5.0):
    """utils/utils.py:53-65 (utils/ imports skimage, which is not installed: restated, 3 lines)."""
    return base_value * math.exp(-scale * (1 - t / max_t)) if t < max_t else base_value


def reference_available():
    try:
        from oracle.fetch_ref import fetch
        return fetch()
    except Exception:
        return False


class ReferenceStep:
    """One training step of the reference's own modules, as train_chaos.py:263-315 / upper_bound_chaos.py:152-171
    drive them (loss weighting with the in-place `*=` / `+=`, Adam lr 1e-4 wd 3e-4), on `device` ("cpu" or "cuda")."""

    def __init__(self, args, device, autocast=None, channels_last=False):
        import torch
        from oracle.fetch_ref import cpu_only, import_reference
        UNet, ConsistencyRegulr, RL = import_reference()
        self.torch, self.RL, self.args, self.device = torch, RL, args, device
        self.autocast, self.channels_last = autocast, channels_last
        C = args.classes
        torch.manual_seed(1)  # train_chaos.py:28,437
        kw_unet = dict(input_ch=1, init_ch=32, max_ch=512, num_classes=C, output_stride=args.output_stride,
                       is_stride_conv=args.unet_variant == "strided", is_trans_conv=args.unet_variant == "strided",
                       elab_end_points=True)
        if args.workload == "pacing":
            ns = argparse.Namespace(ignored_index=C, do_loss_ent=True, do_decoder_consistency=True, detach_weak_cr=False,
                                    loss_cr_variants="ce_loss", do_aux_path=True, do_memory=True)
            kw_aux = dict(num_classes=C, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[512, 512], hid_ch=64,
                          aux_drop_prob=0., do_memory=True, max_step=400, update_momentum=0.9,
                          ensemble_mode='cosine_similarity')
            if device == "cpu":
                with cpu_only():
                    model = ConsistencyRegulr(kwargs_unet=kw_unet, kwargs_aux_path=kw_aux, args_parser=ns)
            else:
                model = ConsistencyRegulr(kwargs_unet=kw_unet, kwargs_aux_path=kw_aux, args_parser=ns)
        else:
            model = UNet(**kw_unet)
        model = model.to(device)
        if channels_last:
            model = model.to(memory_format=torch.channels_last)
        model.train(args.bn == "train")
        self.model = model
        self.opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=3e-4)   # train_chaos.py:219
        self.w = gaussian_ramp_up(args.epoch, 1.0, scale=8.0)

    def __call__(self, batch):
        torch, args, RL, C = self.torch, self.args, self.RL, self.args.classes
        ctx = torch.autocast("cuda", dtype=self.autocast) if self.autocast is not None else _NullCtx()
        with ctx:
            if args.workload == "pacing":
                out = self.model(batch, mode='train', step=args.epoch)
                loss = out['loss_pce']
                loss_ent = out['loss_ent'] * self.w
                loss += loss_ent
                loss_cr = out['loss_cr'] * self.w
                loss += loss_cr
                loss_aux = out['loss_aux_cls']
                loss_aux *= 0.01
                loss += loss_aux
                loss_mem = out['loss_memory']
                loss_mem *= 1
                loss += loss_mem
            elif args.workload == "baseline":
                logits = self.model(batch['image'])['segmentation/logits']
                loss = RL.partial_cross_entropy_loss(logits, torch.argmax(batch['scribble'], dim=1), C)
            else:
                logits = self.model(batch['image'])['segmentation/logits']
                target = torch.argmax(batch['label'], dim=1).long()
                loss = RL.partial_cross_entropy_loss(logits, target, C)
                loss += RL.dice_loss_fn(logits, batch['label'])
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def workload_keys(workload):
    """Tensors the reference training loop moves to the GPU each step (train_chaos.py:264-269 deletes the labels;
    upper_bound_chaos.py:152-155 moves the whole batch of its single-stream dataset)."""
    if workload == "pacing":
        return ("image", "image_strong", "scribble", "scribble_strong", "valid_mask")
    if workload == "baseline":
        return ("image", "scribble", "valid_mask")
    return ("image", "label")


def cpu_reference_steps(args, steps, warmup, per_step, max_seconds=None):
    """-> (times, cores, slices per step, kind). kind 'reference' = the unmodified reference modules from
    baseline/_ref on the host cores; 'port' = the oracle restatement (only when baseline/_ref is absent)."""
    import torch
    from pacingpseudo_b200.synth import make_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    keys = workload_keys(args.workload)
    if reference_available():
        kind, stepper = "reference", ReferenceStep(args, "cpu")
    else:
        kind, stepper = "port", _OraclePortStep(args)
    times = []
    it = 0
    while it < warmup + steps:
        b = make_batch(per_step, args.classes, args.size, args.size, seed=1234 + it)
        batch = {k: b[k] for k in keys}
        t0 = time.perf_counter()
        stepper(batch)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
        elif it == 0 and max_seconds is not None and per_step > 1 and dt * (warmup + steps) > 1.5 * max_seconds:
            # bounded sample: the first (cold) step says the whole run would overshoot -> fewer slices per step
            per_step = max(1, int(per_step * max_seconds / (dt * (warmup + steps))))
        it += 1
    return times, cores, per_step, kind


class _OraclePortStep:
    """Fallback when baseline/_ref is absent: the oracle restatement of the same step (kind = 'port')."""

    def __init__(self, args):
        import torch
        from oracle import pp_oracle as O
        from oracle.gen_golden import build_state
        self.O, self.args, self.torch = O, args, torch
        kind = {"pacing": "pacing", "baseline": "baseline", "upperbound": "upper"}[args.workload]
        self.sd = build_state(dict(kind=kind, C=args.classes, os=args.output_stride))
        self.learn = [k for k in self.sd if self.sd[k].is_floating_point() and "running" not in k
                      and not k.endswith("memory_bank")]
        for k in self.learn:
            self.sd[k].requires_grad_(True)
        self.cfg = O.StepConfig(num_classes=args.classes, ignored_index=args.classes, output_stride=args.output_stride)
        self.state, self.it = {}, 0

    def __call__(self, batch):
        O, args, sd, torch = self.O, self.args, self.sd, self.torch
        for k in self.learn:
            sd[k].grad = None
        training = args.bn == "train"
        if args.workload == "pacing":
            out = O.consistency_forward(sd, batch, self.cfg, mode="train", step=args.epoch, training=training)
            loss = O.total_loss(out, args.epoch)
        else:
            z = O.unet_forward(sd, batch["image"], training, output_stride=args.output_stride)["segmentation/logits"]
            if args.workload == "baseline":
                loss = O.partial_cross_entropy(z, batch["scribble"].argmax(1), args.classes)
            else:
                loss = O.partial_cross_entropy(z, batch["label"].argmax(1), args.classes) + O.dice(z, batch["label"])
        loss.backward()
        self.it += 1
        with torch.no_grad():
            O.adam_step({k: sd[k] for k in self.learn}, {k: sd[k].grad for k in self.learn}, self.state, 1e-4, 3e-4, self.it)
        return loss


def metric_name(args):
    return "train imgs/sec (%d^2 %s step)" % (args.size, {"pacing": "pacingpseudo", "baseline": "baseline UNet+pCE",
                                                          "upperbound": "upperbound CE+Dice"}[args.workload])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    t_all = time.perf_counter()
    times, cores, per_step, kind = cpu_reference_steps(args, args.steps, args.warmup, args.batch, args.ref_max_seconds)
    total = sum(times)
    value = per_step * len(times) / total
    what = ("the UNMODIFIED reference modules (models/unet.py, consistency_reglur_memory.py, aux_path_memory.py, "
            "losses/losses.py from baseline/_ref) + torch.optim.Adam" if kind == "reference"
            else "oracle port (oracle/pp_oracle.py; baseline/_ref absent)")
    sample = "%s, torch %s CPU fp32, %d threads, %d slice(s) of %dx%d per step%s, %d timed steps (%.1f s)" % (
        what, torch.__version__, cores, per_step, args.size, args.size,
        "" if per_step == args.batch else " (bounded sample of the %d-slice batch)" % args.batch, len(times), total)
    line = {
        "impl": "reference", "metric": metric_name(args), "value": value, "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, pairs_override=per_step),
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def workload_config(args, pairs_override=None):
    unit = "pairs" if args.workload == "pacing" else "slices"
    per = pairs_override if pairs_override is not None else args.batch
    return {
        "workload": "%s, %s-shaped synthetic %dx%d 1-ch slices, %d classes (BASELINE.json configs[%s])" % (
            WORKLOADS[args.workload], {5: "CHAOS", 4: "ACDC", 2: "LVSC"}.get(args.classes, "CHAOS"), args.size,
            args.size, args.classes,
            {"pacing": {5: "1", 4: "2", 2: "4"}.get(args.classes, "1"), "baseline": "0", "upperbound": "3"}[args.workload]),
        "%s_per_gpu" % unit: per, "global_%s" % unit: per * max(1, args.gpus),
        "pairs_per_gpu": per, "global_pairs": per * max(1, args.gpus),
        "unet": "init_ch 32, max_ch 512, output_stride %d, %s" % (
            args.output_stride, "maxpool + bilinear" if args.unet_variant == "maxpool"
            else "stride-2 conv + ConvTranspose2d (is_stride_conv, is_trans_conv)"), "bn": args.bn,
        "loss_cr_variants": "ce_loss" if args.workload == "pacing" else None,
        "optimizer": "Adam lr 1e-4 wd 3e-4", "parallelism": "dp%d" % max(1, args.gpus),
        "l2": "per-step working set (~3.4 GB of activations per GPU) >> 126 MB L2; no explicit flush",
    }


def same_box_cudnn(args, dev, steps=5, warmup=2):
    """SURVEY 8(d) / BASELINE.md section 4 "same box, stronger competitor": the unmodified reference modules under stock
    torch eager + cuDNN on this B200 (fp32 without TF32, TF32, bf16 autocast + channels_last), same workload and batch,
    CUDA-event timed. Reported next to the result; not on the product path."""
    import torch
    from pacingpseudo_b200.synth import make_batch
    if not reference_available():
        return {"unavailable": "baseline/_ref absent (run oracle/fetch_ref.py where /root/reference exists)"}
    keys = workload_keys(args.workload)
    batches = [{k: v.to(dev) for k, v in make_batch(args.batch, args.classes, args.size, args.size, seed=4321 + i).items()
                if k in keys} for i in range(2)]
    out = {"how": "unmodified reference modules (baseline/_ref), stock torch %s eager + cuDNN %s on this GPU, %d slices "
                  "per step, %d timed steps after %d warm-up, CUDA events" % (
                      torch.__version__, torch.backends.cudnn.version(), args.batch, steps, warmup)}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32, ac, cl in (("fp32", False, None, False), ("tf32", True, None, False),
                                   ("bf16_autocast_channels_last", True, torch.bfloat16, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            try:
                stepper = ReferenceStep(args, dev, autocast=ac, channels_last=cl)
                bs = [{k: (v.contiguous(memory_format=torch.channels_last) if (cl and v.dim() == 4) else v)
                       for k, v in b.items()} for b in batches]
                for i in range(warmup):
                    stepper(bs[i % 2])
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    stepper(bs[i % 2])
                e1.record()
                torch.cuda.synchronize(dev)
                ms = e0.elapsed_time(e1) / steps
                out[name] = {"ms_per_step": ms, "value": args.batch / (ms / 1e3), "unit": "img/s"}
                del stepper
            except Exception as exc:   # a competitor that fails to run is reported, never fatal to the bench
                out[name] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
    return out


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (every ~2 ms; the timed
    region is only a few hundred ms long), falling back to `nvidia-smi -lms` when NVML is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
            0x80: "hw_power_brake_slowdown"}

    # Poller PROCESS (no GIL contention with the launch loop, which can starve an in-process thread down to a single
    # sample): prints "unix_time sm_mhz reasons_mask" every ~2 ms from before the warm-up; stop(t0, t1) keeps the
    # samples that fall inside the timed region. The in-process thread below stays as a fallback.
    POLLER = (
        "import sys, time, pynvml\n"
        "pynvml.nvmlInit()\n"
        "try:\n"
        "    h = pynvml.nvmlDeviceGetHandleByUUID(sys.argv[1])\n"
        "except Exception:\n"
        "    h = pynvml.nvmlDeviceGetHandleByIndex(int(sys.argv[2]))\n"
        "print('max', pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), flush=True)\n"
        "import os\n"
        "parent, t_end = os.getppid(), time.time() + 900\n"
        "while os.getppid() == parent and time.time() < t_end:   # never outlive the bench process\n"
        "    t = time.time()\n"
        "    sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)\n"
        "    try:\n"
        "        m = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)\n"
        "    except Exception:\n"
        "        m = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)\n"
        "    sys.stdout.write('%.6f %d %d\\n' % (t, sm, m))\n"
        "    sys.stdout.flush()\n"
        "    time.sleep(0.002)\n")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.samples, self.mask, self.stop_flag = None, None, [], 0, False
        self.poller, self.poller_lines, self.poller_thr = None, [], None

    def start_poller(self):
        """Call BEFORE the warm-up: the process needs ~0.1-0.3 s to import pynvml."""
        try:
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.poller = subprocess.Popen([sys.executable, "-c", self.POLLER, uuid, str(idx)], stdout=subprocess.PIPE,
                                           stderr=subprocess.DEVNULL, text=True)
            self.poller_thr = threading.Thread(target=self._drain, daemon=True)
            self.poller_thr.start()
        except Exception:
            self.poller = None

    def _drain(self):
        try:
            for line in self.poller.stdout:
                self.poller_lines.append(line)
        except Exception:
            pass

    def _poller_result(self, t0, t1):
        if self.poller is None:
            return None
        time.sleep(0.01)
        self.poller.terminate()
        try:
            self.poller.wait(timeout=5)
        except Exception:
            self.poller.kill()
        self.poller_thr.join(timeout=2)
        smax, sm, mask = None, [], 0
        for ln in list(self.poller_lines):
            f = ln.split()
            try:
                if f[0] == "max":
                    smax = float(f[1])
                elif t0 <= float(f[0]) <= t1:
                    sm.append(float(f[1]))
                    mask |= int(f[2])
            except (ValueError, IndexError):
                continue
        if len(sm) < 3:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smax,
                "reasons": sorted(v for k, v in self.BITS.items() if mask & k), "samples": len(sm),
                "source": "nvml (poller process, samples inside the timed region)"}

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(self.index).uuid))
        except Exception:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.smax = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    self.mask |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, t0=None, t1=None):
        if t0 is not None:
            res = self._poller_result(t0, t1)
            if res is not None:
                if self.nvml is not None:
                    self.stop_flag = True
                elif self.proc is not None:
                    self.proc.terminate()
                return res
        if self.nvml is not None:
            self.stop_flag = True
            self.thr.join(timeout=2)
            sm = sorted(self.samples)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax,
                    "reasons": sorted(v for k, v in self.BITS.items() if self.mask & k), "samples": len(sm),
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from pacingpseudo_b200 import dp
    from pacingpseudo_b200 import functional as PF
    from pacingpseudo_b200.data import (DevicePrefetcher, LossReader, compact_batch, sample_strong_params,
                                        strong_color_augment)
    from pacingpseudo_b200.dropin import DROPIN_PATH
    from pacingpseudo_b200.lib import get_lib
    from pacingpseudo_b200.optim import FlatAdam
    from pacingpseudo_b200.synth import make_batch
    from pacingpseudo_b200.schedules import loss_weight_ramp_up
    sys.path.insert(0, DROPIN_PATH)
    from models.consistency_reglur_memory import ConsistencyRegulr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the product path has no CPU fallback")
    rank, world, dev = dp.init_distributed()
    lib = get_lib()
    lib.ensure_init(dev.index)
    C, S, B = args.classes, args.size, args.batch

    torch.manual_seed(1)  # train_chaos.py:28,437
    pacing = args.workload == "pacing"
    kw_unet = dict(input_ch=1, init_ch=32, max_ch=512, num_classes=C, output_stride=args.output_stride,
                   is_stride_conv=args.unet_variant == "strided", is_trans_conv=args.unet_variant == "strided",
                   elab_end_points=True, precision=args.precision)
    if pacing:
        ns = argparse.Namespace(ignored_index=C, do_loss_ent=True, do_decoder_consistency=True, detach_weak_cr=False,
                                loss_cr_variants="ce_loss", do_aux_path=True, do_memory=True)
        model = ConsistencyRegulr(
            kwargs_unet=kw_unet,
            kwargs_aux_path=dict(num_classes=C, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[512, 512],
                                 hid_ch=64, aux_drop_prob=0., do_memory=True, max_step=400, update_momentum=0.9,
                                 ensemble_mode='cosine_similarity'),
            args_parser=ns).to(dev)
        backbone = model.backbone
    else:
        from models.unet import UNet
        from losses import losses as DL
        model = UNet(**kw_unet).to(dev)
        backbone = model
    model.train(args.bn == "train")
    opt = FlatAdam(model.parameters(), lr=1e-4, weight_decay=3e-4)
    reducer = dp.GradientAllReducer(opt.flat_grad, num_buckets=4, unet=backbone, optimizer=opt)
    opt.grad_scale = 1.0 / world
    if world > 1 and pacing:
        model.aux_path.bank_sync = dp.make_bank_sync(0)

    # a small pool of distinct per-rank batches; host copies pinned for the end-to-end leg
    keys = workload_keys(args.workload)  # what the reference training loop moves to the GPU each step
    pool_host = []
    for i in range(4):
        b = make_batch(B, C, S, S, seed=dp.shard_seed(1234, rank, i))
        pool_host.append({k: b[k].pin_memory() for k in keys})
    pool_dev = [{k: v.to(dev) for k, v in b.items()} for b in pool_host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in pool_host[0].values())
    # compact variant: uint8 scribble index map, no scribble_strong, image_strong made on the device
    pool_compact = []
    h2d_bytes_compact = None
    if pacing:
        for i, b in enumerate(pool_host):
            c = compact_batch(b, C)
            c["strong_params"] = sample_strong_params(B, 1.0, torch.Generator().manual_seed(dp.shard_seed(99, rank, i)))
            pool_compact.append({k: v.pin_memory() for k, v in c.items()})
        h2d_bytes_compact = sum(v.numel() * v.element_size() for v in pool_compact[0].values())
    w_ent = loss_weight_ramp_up(args.epoch, 1.0, scale=8.0)
    w_cr = loss_weight_ramp_up(args.epoch, 1.0, scale=8.0)

    loss_reader = LossReader(dev)
    prefetcher = DevicePrefetcher((), dev)

    def step(batch, read_back):
        if pacing:
            if "strong_params" in batch:   # compact host format: the strong branch's image is produced on the device
                batch = dict(batch, image_strong=strong_color_augment(batch["image"], batch["strong_params"]))
            out = model(batch, mode='train', step=args.epoch)
            loss = out['loss_pce']
            loss_ent = out['loss_ent'] * w_ent
            loss += loss_ent
            loss_cr = out['loss_cr'] * w_cr
            loss += loss_cr
            loss_aux = out['loss_aux_cls']
            loss_aux *= 0.01
            loss += loss_aux
            loss_mem = out['loss_memory']
            loss_mem *= 1
            loss += loss_mem
            scalars = [out['loss_pce'], loss_ent, loss_cr, loss_aux, loss_mem]
        elif args.workload == "baseline":   # config 1 on the GPU: UNet + partial CE on the scribbles
            logits = model(batch['image'])['segmentation/logits']
            loss = DL.partial_cross_entropy_loss(logits, PF.onehot_argmax(batch['scribble']), C)
            scalars = [loss]
        else:   # upper_bound_chaos.py:157-165: CE on argmax(label) + Dice on the one-hot label
            logits = model(batch['image'])['segmentation/logits']
            loss_ce = DL.partial_cross_entropy_loss(logits, PF.onehot_argmax(batch['label']), C)
            loss_dice = DL.dice_loss_fn(logits, batch['label'])
            scalars = [loss_ce.detach().clone(), loss_dice]
            loss = loss_ce
            loss += loss_dice
        opt.zero_grad()
        loss.backward()
        reducer.allreduce()
        opt.step()
        if read_back == "item":   # the blocking .item() reads of train_chaos.py:275-310
            return [t.item() for t in scalars]
        if read_back == "async":  # same scalars, non-blocking D2H, handed back one step later
            return loss_reader.push(scalars)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    n_scalars = {"pacing": 5, "baseline": 1, "upperbound": 2}[args.workload]
    step_ms = {}   # leg name -> per-step durations (CUDA events between consecutive steps), this rank

    def timed(nsteps, host_inputs, profile, pool=None, leg=None):
        pool = pool_host if pool is None else pool
        barrier()
        lib.cdll.pp_profile_reset()
        lib.cdll.pp_profile_enable(1 if profile else 0)
        l0 = lib.cdll.pp_launch_count()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(nsteps + 1)]
        marks[0].record()
        if host_inputs:   # pinned host batches, H2D on a copy stream one step ahead (pacingpseudo_b200/data.py)
            for i, batch in enumerate(prefetcher.reset(pool[i % len(pool)] for i in range(nsteps))):
                step(batch, host_inputs)
                marks[i + 1].record()
            if host_inputs == "async":
                assert len(loss_reader.flush()) == n_scalars
        else:
            for i in range(nsteps):
                step(pool_dev[i % len(pool_dev)], False)
                marks[i + 1].record()
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        barrier()
        lib.cdll.pp_profile_enable(0)
        ms = dp.max_over_ranks(marks[0].elapsed_time(e1), dev)
        if leg is not None and nsteps > 0:
            step_ms[leg] = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(nsteps))
        return ms, lib.cdll.pp_launch_count() - l0

    def median(leg):
        v = step_ms.get(leg)
        return dp.max_over_ranks(v[len(v) // 2], dev) if v else None

    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start_poller()
    for i in range(args.warmup):
        step(pool_dev[i % len(pool_dev)], False)
    if rank == 0:
        sampler.start()
    t_begin = time.time()
    ms, launches = timed(args.steps, host_inputs=False, profile=False, leg="device")
    ms_median = median("device")
    clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
    # the same K steps once more with every tcgen05 conv launch bracketed by CUDA events on its stream (roofline)
    ms_prof, _ = timed(0 if args.no_profile_pass else args.steps, host_inputs=False, profile=True)
    prof = {}
    for fam, name in ((0, "conv3x3_tc (fwd+dgrad)"), (1, "conv3x3_wgrad_tc"), (2, "scribble_loss (fwd+bwd)"), (-1, "all")):
        t, f, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        lib.call("pp_profile_collect", fam, ctypes.byref(t), ctypes.byref(f), ctypes.byref(n))
        prof[name] = dict(ms=t.value, flops=f.value, launches=n.value)
    prof_all = prof.pop("all")
    prof_loss = prof.pop("scribble_loss (fwd+bwd)")
    e2e = None
    if not args.no_e2e:
        for batch in prefetcher.reset(pool_host[i] for i in range(3)):   # allocates the staging buffers
            step(batch, "async")
        ms_e2e, _ = timed(args.steps, host_inputs="async", profile=False, leg="e2e")
        ms_e2e_median = median("e2e")
        ms_e2e_item, _ = timed(args.steps, host_inputs="item", profile=False)
        e2e = {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 4 * n_scalars, "ms_per_step": ms_e2e / args.steps,
               "ms_per_step_median": ms_e2e_median,
               "how": "pinned host batches -> DevicePrefetcher (H2D into persistent staging buffers on a copy stream, "
                      "one step ahead) -> module forward / backward / FlatAdam.step -> LossReader (the step's loss "
                      "scalars, non-blocking D2H to pinned memory every step, read one step later)",
               "ms_per_step_blocking_item_reads": ms_e2e_item / args.steps}
        if pacing:
            for batch in prefetcher.reset(pool_compact[i] for i in range(3)):   # staging buffers of the compact format
                step(batch, "async")
            ms_e2e_compact, _ = timed(args.steps, host_inputs="async", profile=False, pool=pool_compact)
            e2e["compact_input"] = {
                "value": B * world * args.steps / (ms_e2e_compact / 1e3), "unit": "img/s",
                "ms_per_step": ms_e2e_compact / args.steps, "h2d_bytes_per_step": h2d_bytes_compact,
                "how": "same loop from the compact host format (SURVEY 8f N3): image + uint8 scribble index map + "
                       "valid mask + 8 augmentation draws per slice; image_strong = pp_strong_color_augment(image) "
                       "on the device (different strong images than the reference-format leg, same work)"}
    # The other BatchNorm regime on the same model, same timing rules (all ranks): the reference trains epoch 0 with batch
    # statistics and every later epoch with running statistics (train_chaos.py:370 calls model.eval() for validation and
    # never model.train() again), so both are reported: the headline is --bn, the other goes to `extra`.
    other_bn = None
    if not args.no_other_bn:
        model.train(args.bn != "train")
        for i in range(max(3, args.warmup // 2)):
            step(pool_dev[i % len(pool_dev)], False)
        ms_o, _ = timed(args.steps, host_inputs=False, profile=False, leg="other_bn")
        other_bn = {"bn": "eval" if args.bn == "train" else "train", "value": B * world * args.steps / (ms_o / 1e3),
                    "unit": "img/s", "ms_per_step": ms_o / args.steps, "ms_per_step_median": median("other_bn"),
                    "steps": args.steps}
        model.train(args.bn == "train")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = B * world * args.steps / (ms / 1e3)
    # Dominant kernel: conv3x3_tc_kernel (forward and dgrad launches share it). Times are the device time during which
    # at least one launch of the family was running (union of the CUDA-event intervals: the weight-gradient kernels run
    # on a side stream and the two branches of the forward pass on two streams, so durations are not additive).
    dom = prof["conv3x3_tc (fwd+dgrad)"]
    steps_p = max(1, args.steps)
    tf = lambda p: (p["flops"] / (p["ms"] / 1e3) / 1e12) if p["ms"] > 0 else 0.0
    gf = (GF_PER_PAIR if pacing else GF_PER_IMAGE).get((S, C)) if (
        args.unet_variant == "maxpool" and args.output_stride == 8) else None
    traffic = load_traffic()
    line = {
        "metric": metric_name(args), "value": value, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_median": ms_median,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args), "clocks": clocks, "gpu_launches": int(launches),
        "e2e": e2e,
        "roofline": {
            "bound": "tensor", "kernel": "tcgen05 implicit-GEMM 3x3 conv, forward + dgrad launches (conv3x3_rows_tc_kernel / conv3x3_tc_kernel / "
                                   "conv3x3_halo_tc_kernel, chosen per shape)",
            "achieved": tf(dom), "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
            "frac": tf(dom) / peaks["tf_sustained"], "peak_source": peaks["src"] + " bf16 sustained",
            "traffic": traffic.get("conv3x3_tc_bytes_per_launch"), "traffic_source": traffic.get("source"),
            "launches": int(dom["launches"]), "flops_per_launch": dom["flops"] / max(1, dom["launches"]),
            "avg_launch_ms": dom["ms"] / max(1, dom["launches"]), "kernel_ms_per_step": dom["ms"] / steps_p,
            "share_of_step": dom["ms"] / ms_prof if ms_prof > 0 else None,
            "profiled_pass_ms_per_step": ms_prof / steps_p,
            "other_kernels": {
                "conv3x3_wgrad_tc_kernel": {"achieved": tf(prof["conv3x3_wgrad_tc"]),
                                            "frac": tf(prof["conv3x3_wgrad_tc"]) / peaks["tf_sustained"],
                                            "kernel_ms_per_step": prof["conv3x3_wgrad_tc"]["ms"] / steps_p,
                                            "launches": int(prof["conv3x3_wgrad_tc"]["launches"])},
                "all_conv_launches_union": {"achieved": tf(prof_all), "frac": tf(prof_all) / peaks["tf_sustained"],
                                            "kernel_ms_per_step": prof_all["ms"] / steps_p,
                                            "share_of_step": prof_all["ms"] / ms_prof if ms_prof > 0 else None},
                "scribble_loss_fwd+bwd (HBM bound)": {
                    "achieved": (prof_loss["flops"] / (prof_loss["ms"] / 1e3) / 1e9) if prof_loss["ms"] > 0 else 0.0,
                    "peak": peaks["hbm"], "unit": "GB/s",
                    "frac": ((prof_loss["flops"] / (prof_loss["ms"] / 1e3) / 1e9) / peaks["hbm"]) if prof_loss["ms"] > 0 else 0.0,
                    "bytes_per_step": prof_loss["flops"] / steps_p, "kernel_ms_per_step": prof_loss["ms"] / steps_p,
                    "note": "per-step working set (~50 MB) fits the 126 MB L2"},
            },
            "step_tensor_frac": (value / world * gf / 1e12 / peaks["tf_sustained"]) if gf else None,
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        # ~10-30 s of host work: the unmodified reference modules on this box's cores, one warm-up + 3 timed steps
        del pool_dev, pool_host, pool_compact
        torch.cuda.empty_cache()
        times, cores, per_step, kind = cpu_reference_steps(args, steps=3, warmup=1, per_step=B, max_seconds=30.0)
        v = per_step * len(times) / sum(times)
        line["cpu_baseline"] = {
            "value": v, "unit": "img/s", "cores": cores, "kind": kind,
            "sample": "%s on the host cores (%d threads): the same %s step, %d slice(s) of %dx%d per step, %d timed "
                      "steps (%.1f s)" % ("unmodified reference modules (baseline/_ref)" if kind == "reference" else
                                          "oracle port", cores, args.workload, per_step, S, S, len(times), sum(times))}
    line["extra"] = {}
    if other_bn is not None:
        line["extra"]["other_bn_regime"] = other_bn
    if world == 1 and not args.no_same_box:
        line["extra"]["cudnn_same_box"] = same_box_cudnn(args, dev)
    print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: library chatter written to file descriptor 1 (e.g. NCCL's version banner)
    # is sent to stderr for the whole run, and the result line goes to the saved descriptor.
    global _RESULT_OUT
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
