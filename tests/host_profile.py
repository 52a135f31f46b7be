"""Host-side cost of one pacingpseudo step (no GPU sync inside): where does the Python/launch time go?"""
import argparse, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pacingpseudo_b200.dropin import DROPIN_PATH
from pacingpseudo_b200.optim import FlatAdam
from pacingpseudo_b200.synth import make_batch
from pacingpseudo_b200.data import DevicePrefetcher
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
host = [{k: v.pin_memory() for k, v in make_batch(B, C, S, S, seed=1234 + i).items()} for i in range(2)]
devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
print("threads", torch.get_num_threads(), "OMP", os.environ.get("OMP_NUM_THREADS"))

def step(batch, acc, items):
    t0 = time.perf_counter()
    out = model(batch, mode='train', step=40)
    t1 = time.perf_counter()
    loss = out['loss_pce'] + out['loss_ent'] * 0.5 + out['loss_cr'] * 0.5 + out['loss_aux_cls'] * 0.01 + out['loss_memory']
    opt.zero_grad()
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    if items:
        vals = [out[k].item() for k in ('loss_pce', 'loss_ent', 'loss_cr', 'loss_aux_cls', 'loss_memory')]
    t5 = time.perf_counter()
    for k, v in zip(("fwd", "loss+zero", "bwd", "adam", "items"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        acc[k] = acc.get(k, 0.0) + v

import cProfile, pstats
MODES = sys.argv[1:] or ["dev-noitem", "dev-item", "h2d-item", "prefetch-item"]
for mode in MODES:
    prof = cProfile.Profile() if os.environ.get("PP_CPROFILE") else None
    for rep in range(2):
        acc = {}
        torch.cuda.synchronize()
        if prof and rep == 1: prof.enable()
        t0 = time.perf_counter()
        n = 10
        if mode == "prefetch-item":
            pf = DevicePrefetcher((host[i % 2] for i in range(n)), dev)
            while True:
                tn = time.perf_counter()
                try:
                    b = next(pf)
                except StopIteration:
                    break
                acc["next"] = acc.get("next", 0.0) + time.perf_counter() - tn
                tn = time.perf_counter()
                torch.cuda.current_stream().synchronize()
                acc["sync_after_next"] = acc.get("sync_after_next", 0.0) + time.perf_counter() - tn
                step(b, acc, True)
        else:
            for i in range(n):
                if mode == "h2d-item":
                    b = {k: v.to(dev, non_blocking=True) for k, v in host[i % 2].items()}
                else:
                    b = devb[i % 2]
                step(b, acc, mode != "dev-noitem")
        torch.cuda.synchronize()
        tot = time.perf_counter() - t0
        if prof and rep == 1:
            prof.disable()
            pstats.Stats(prof).sort_stats("tottime").print_stats(12)
    print(mode, "ms/step %.2f" % (1e3 * tot / n), {k: round(1e3 * v / n, 2) for k, v in acc.items()}, flush=True)
