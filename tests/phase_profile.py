"""GPU-side time of the phases of one pacingpseudo step (CUDA events on the main stream): forward (UNet + aux +
losses), backward, optimizer. Usage: python tests/phase_profile.py [steps]"""
import argparse, os, sys
import torch
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
acc = {}
for it in range(n + 3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    b = devb[it % 2]
    ev[0].record()
    x = torch.cat((b['image'], b['image_strong']), 0)
    logits_all, feats = model.backbone.run_native(x, groups=2, feat_names=tuple(model.aux_path.feat_stage))
    ev[1].record()
    model.backbone.end_points._set_native(feats)   # keep semantics irrelevant here; time the rest of forward
    out = model(b, mode='train', step=40) if False else None
    ev[2].record()
    out = model(b, mode='train', step=40)
    loss = out['loss_pce'] + out['loss_ent'] * 0.5 + out['loss_cr'] * 0.5 + out['loss_aux_cls'] * 0.01 + out['loss_memory']
    opt.zero_grad()
    ev[3].record()
    loss.backward()
    ev[4].record()
    opt.step()
    e5 = torch.cuda.Event(enable_timing=True); e5.record()
    torch.cuda.synchronize()
    if it >= 3:
        for k, (a, b2) in {"unet_fwd_only": (ev[0], ev[1]), "full_fwd(model())": (ev[2], ev[3]), "backward": (ev[3], ev[4]),
                           "adam": (ev[4], e5)}.items():
            acc[k] = acc.get(k, 0.0) + a.elapsed_time(b2)
print("overlap off" if os.environ.get("PP_NO_OVERLAP") == "1" else "overlap on", {k: round(v / n, 3) for k, v in acc.items()})
