#!/usr/bin/env python
"""bench.py — PacingPseudo train-step throughput (BASELINE.json metric: train imgs/sec, 256x256 pacingpseudo step).

    python bench.py --gpus N --steps K --warmup W            # this repo (hand-written sm_100a path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm's CPU path (oracle port)

A step = one full pacingpseudo iteration (train_chaos.py:263-315): batched weak+strong UNet forward, partial CE +
entropy + consistency + aux partial CE + memory loss, memory-bank update, backward, gradient all-reduce (N > 1),
Adam. One "image" = one weak+strong pair, as batch_size counts them (train_chaos.py:237). Prints ONE JSON line.
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
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
    return ap.parse_args()


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu pass (profiles/r01_conv_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_conv_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the full pacingpseudo step on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_pacing_steps(args, steps, warmup, pairs=1):
    import torch
    from oracle import pp_oracle as O
    from oracle.gen_golden import build_state
    from pacingpseudo_b200.synth import make_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    case = dict(kind="pacing", C=args.classes, os=8)
    sd = build_state(case)
    learn = [k for k in sd if sd[k].is_floating_point() and "running" not in k and not k.endswith("memory_bank")]
    for k in learn:
        sd[k].requires_grad_(True)
    cfg = O.StepConfig(num_classes=args.classes, ignored_index=args.classes)
    state = {}
    times = []
    for it in range(warmup + steps):
        batch = make_batch(pairs, args.classes, args.size, args.size, seed=1234 + it)
        t0 = time.perf_counter()
        for k in learn:
            sd[k].grad = None
        out = O.consistency_forward(sd, batch, cfg, mode="train", step=args.epoch, training=(args.bn == "train"))
        loss = O.total_loss(out, args.epoch)
        loss.backward()
        with torch.no_grad():
            O.adam_step({k: sd[k] for k in learn}, {k: sd[k].grad for k in learn}, state, 1e-4, 3e-4, it + 1)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, cores, pairs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = time.perf_counter()
    times, cores, pairs = cpu_pacing_steps(args, args.steps, args.warmup, pairs=1)
    total = sum(times)
    value = pairs * len(times) / total
    sample = ("oracle port (oracle/pp_oracle.py, torch %s CPU fp32) of the full pacingpseudo step on %d weak+strong pair(s) "
              "of %dx%d per step, %d timed steps" % (__import__("torch").__version__, pairs, args.size, args.size, len(times)))
    line = {
        "impl": "reference", "metric": "train imgs/sec (256^2 pacingpseudo step)", "value": value, "unit": "img/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, pairs_override=pairs),
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def workload_config(args, pairs_override=None):
    return {
        "workload": "pacingpseudo full step (--do_loss_ent --do_decoder_consistency --do_aux_path --do_memory), "
                    "CHAOS-shaped synthetic %dx%d 1-ch slices, %d classes (BASELINE.json configs[1])" % (
                        args.size, args.size, args.classes),
        "pairs_per_gpu": pairs_override if pairs_override is not None else args.batch,
        "global_pairs": (pairs_override if pairs_override is not None else args.batch) * max(1, args.gpus),
        "unet": "init_ch 32, max_ch 512, output_stride %d, %s" % (
            args.output_stride, "maxpool + bilinear" if args.unet_variant == "maxpool"
            else "stride-2 conv + ConvTranspose2d (is_stride_conv, is_trans_conv)"), "bn": args.bn,
        "loss_cr_variants": "ce_loss", "optimizer": "Adam lr 1e-4 wd 3e-4", "parallelism": "dp%d" % max(1, args.gpus),
        "l2": "per-step working set (~3.4 GB of activations per GPU) >> 126 MB L2; no explicit flush",
    }


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
    ns = argparse.Namespace(ignored_index=C, do_loss_ent=True, do_decoder_consistency=True, detach_weak_cr=False,
                            loss_cr_variants="ce_loss", do_aux_path=True, do_memory=True)
    model = ConsistencyRegulr(
        kwargs_unet=dict(input_ch=1, init_ch=32, max_ch=512, num_classes=C, output_stride=args.output_stride,
                         is_stride_conv=args.unet_variant == "strided", is_trans_conv=args.unet_variant == "strided",
                         elab_end_points=True, precision=args.precision),
        kwargs_aux_path=dict(num_classes=C, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[512, 512],
                             hid_ch=64, aux_drop_prob=0., do_memory=True, max_step=400, update_momentum=0.9,
                             ensemble_mode='cosine_similarity'),
        args_parser=ns).to(dev)
    model.train(args.bn == "train")
    opt = FlatAdam(model.parameters(), lr=1e-4, weight_decay=3e-4)
    reducer = dp.GradientAllReducer(opt.flat_grad, num_buckets=4, unet=model.backbone, optimizer=opt)
    opt.grad_scale = 1.0 / world
    if world > 1:
        model.aux_path.bank_sync = dp.make_bank_sync(0)
        for p in opt.params:  # identical start on all ranks (same seed) — assert rather than broadcast
            pass

    # a small pool of distinct per-rank batches; host copies pinned for the end-to-end leg
    keys = ("image", "image_strong", "scribble", "scribble_strong", "valid_mask")  # what train_chaos.py:264-269 moves
    pool_host = []
    for i in range(4):
        b = make_batch(B, C, S, S, seed=dp.shard_seed(1234, rank, i))
        pool_host.append({k: b[k].pin_memory() for k in keys})
    pool_dev = [{k: v.to(dev) for k, v in b.items()} for b in pool_host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in pool_host[0].values())
    # compact variant: uint8 scribble index map, no scribble_strong, image_strong made on the device
    pool_compact = []
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
        opt.zero_grad()
        loss.backward()
        reducer.allreduce()
        opt.step()
        if read_back == "item":   # the five blocking .item() reads of train_chaos.py:275-310
            return [t.item() for t in (out['loss_pce'], loss_ent, loss_cr, loss_aux, loss_mem)]
        if read_back == "async":  # same five scalars, non-blocking D2H, handed back one step later
            return loss_reader.push([out['loss_pce'], loss_ent, loss_cr, loss_aux, loss_mem])
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(nsteps, host_inputs, profile, pool=None):
        pool = pool_host if pool is None else pool
        barrier()
        lib.cdll.pp_profile_reset()
        lib.cdll.pp_profile_enable(1 if profile else 0)
        l0 = lib.cdll.pp_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if host_inputs:   # pinned host batches, H2D on a copy stream one step ahead (pacingpseudo_b200/data.py)
            for batch in prefetcher.reset(pool[i % len(pool)] for i in range(nsteps)):
                step(batch, host_inputs)
            if host_inputs == "async":
                assert len(loss_reader.flush()) == 5
        else:
            for i in range(nsteps):
                step(pool_dev[i % len(pool_dev)], False)
        e1.record()
        barrier()
        lib.cdll.pp_profile_enable(0)
        ms = dp.max_over_ranks(e0.elapsed_time(e1), dev)
        return ms, lib.cdll.pp_launch_count() - l0

    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start_poller()
    for i in range(args.warmup):
        step(pool_dev[i % len(pool_dev)], False)
    if rank == 0:
        sampler.start()
    t_begin = time.time()
    ms, launches = timed(args.steps, host_inputs=False, profile=False)
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
        ms_e2e, _ = timed(args.steps, host_inputs="async", profile=False)
        ms_e2e_item, _ = timed(args.steps, host_inputs="item", profile=False)
        for batch in prefetcher.reset(pool_compact[i] for i in range(3)):   # staging buffers of the compact format
            step(batch, "async")
        ms_e2e_compact, _ = timed(args.steps, host_inputs="async", profile=False, pool=pool_compact)
        e2e = {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 20, "ms_per_step": ms_e2e / args.steps,
               "how": "pinned host batches -> DevicePrefetcher (H2D into persistent staging buffers on a copy stream, "
                      "one step ahead) -> ConsistencyRegulr.forward / backward / FlatAdam.step -> LossReader (the five "
                      "loss scalars, non-blocking D2H to pinned memory every step, read one step later)",
               "ms_per_step_blocking_item_reads": ms_e2e_item / args.steps,
               "compact_input": {
                   "value": B * world * args.steps / (ms_e2e_compact / 1e3), "unit": "img/s",
                   "ms_per_step": ms_e2e_compact / args.steps, "h2d_bytes_per_step": h2d_bytes_compact,
                   "how": "same loop from the compact host format (SURVEY 8f N3): image + uint8 scribble index map + "
                          "valid mask + 8 augmentation draws per slice; image_strong = pp_strong_color_augment(image) "
                          "on the device (different strong images than the reference-format leg, same work)"}}
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
    gf = GF_PER_PAIR.get((S, C)) if (args.unet_variant == "maxpool" and args.output_stride == 8) else None
    traffic = load_traffic()
    line = {
        "metric": "train imgs/sec (256^2 pacingpseudo step)", "value": value, "unit": "img/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args), "clocks": clocks, "gpu_launches": int(launches),
        "e2e": e2e,
        "roofline": {
            "bound": "tensor", "kernel": "conv3x3_tc_kernel: tcgen05 implicit-GEMM 3x3 conv, forward + dgrad launches",
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
        times, cores, pairs = cpu_pacing_steps(args, steps=12, warmup=1, pairs=2)   # ~10-15 s of host work
        v = pairs * len(times) / sum(times)
        line["cpu_baseline"] = {
            "value": v, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": "oracle port of the same pacingpseudo step, %d pair(s) of %dx%d per step, %d timed steps (%.1f s)" % (
                pairs, S, S, len(times), sum(times))}
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
