"""Concurrency-aware kernel timeline of one pacingpseudo step via torch.profiler (CUPTI): per kernel start offset,
duration, stream. Writes gpurun_out/timeline.txt and prints per-stream busy time and the biggest idle gaps."""
import argparse, os, sys
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pacingpseudo_b200.dropin import DROPIN_PATH
from pacingpseudo_b200.optim import FlatAdam
from pacingpseudo_b200.synth import make_batch
sys.path.insert(0, DROPIN_PATH)
from models.consistency_reglur_memory import ConsistencyRegulr

dev = torch.device("cuda", 0)
C, S, B = 5, 256, 12
torch.manual_seed(1)
ns = argparse.Namespace(ignored_index=C, do_loss_ent=True, do_decoder_consistency=True, detach_weak_cr=False,
                        loss_cr_variants="ce_loss", do_aux_path=True, do_memory=True)
model = ConsistencyRegulr(
    kwargs_unet=dict(input_ch=1, init_ch=32, max_ch=512, num_classes=C, output_stride=8, is_stride_conv=False,
                     is_trans_conv=False, elab_end_points=True),
    kwargs_aux_path=dict(num_classes=C, feat_stage=['encoder/stage6', 'encoder/stage5'], feat_ch=[512, 512],
                         hid_ch=64, aux_drop_prob=0., do_memory=True, max_step=400, update_momentum=0.9,
                         ensemble_mode='cosine_similarity'),
    args_parser=ns).to(dev)
model.train()
opt = FlatAdam(model.parameters(), lr=1e-4, weight_decay=3e-4)
devb = [{k: v.to(dev) for k, v in make_batch(B, C, S, S, seed=1234 + i).items()} for i in range(2)]

def step(b):
    out = model(b, mode='train', step=40)
    loss = out['loss_pce'] + out['loss_ent'] * 0.5 + out['loss_cr'] * 0.5 + out['loss_aux_cls'] * 0.01 + out['loss_memory']
    opt.zero_grad()
    loss.backward()
    opt.step()

for i in range(4):
    step(devb[i % 2])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(2):
        step(devb[i % 2])
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# keep the last step: kernels after the last adam_kernel-but-one
adam = [i for i, e in enumerate(evs) if "adam_kernel" in e.name]
first = adam[-2] + 1 if len(adam) >= 2 else 0
evs = evs[first:adam[-1] + 1]
t0 = evs[0].time_range.start
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
streams = {}
with open(os.path.join(ROOT, "gpurun_out", "timeline.txt"), "w") as f:
    for e in evs:
        st = getattr(e, "stream", None)
        if st is None:
            st = -1
        a, b = e.time_range.start - t0, e.time_range.end - t0
        streams.setdefault(st, []).append((a, b))
        f.write("%9.1f %8.1f  s%-3s %s\n" % (a, b - a, st, e.name[:110]))
end = max(b for v in streams.values() for _, b in v)
print("step span %.1f us, %d kernels" % (end, len(evs)))
allint = sorted(i for v in streams.values() for i in v)
busy, cur_a, cur_b, gaps = 0.0, None, None, []
for a, b in allint:
    if cur_b is None or a > cur_b:
        if cur_b is not None:
            busy += cur_b - cur_a
            gaps.append((a - cur_b, cur_b))
        cur_a, cur_b = a, b
    else:
        cur_b = max(cur_b, b)
busy += cur_b - cur_a
print("GPU busy (any stream) %.1f us, idle %.1f us" % (busy, end - busy))
for st, v in sorted(streams.items(), key=lambda kv: -sum(b - a for a, b in kv[1])):
    print("stream %s: %d kernels, busy %.1f us" % (st, len(v), sum(b - a for a, b in v)))
print("largest idle gaps (us, at):", [(round(g, 1), round(at, 1)) for g, at in sorted(gaps, reverse=True)[:8]])
