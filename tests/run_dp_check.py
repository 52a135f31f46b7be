"""torchrun --nproc-per-node N tests/run_dp_check.py — multi-GPU check of the overlapped gradient exchange.

Every rank runs one full pacingpseudo step on its own batch, the bucketed all-reduce (events recorded inside
pp_unet_backward, NCCL on a side stream) sums the flat gradient buffer, and the result must equal the sum of the
per-rank gradients gathered separately (bit-exact for 2 ranks: a two-operand fp32 sum is order-independent). Also
checks the bank broadcast. Prints one OK/FAIL line from rank 0; exit code 1 on failure.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from pacingpseudo_b200 import dp
    from pacingpseudo_b200.dropin import DROPIN_PATH
    from pacingpseudo_b200.optim import FlatAdam
    from pacingpseudo_b200.synth import make_batch
    sys.path.insert(0, DROPIN_PATH)
    from models.consistency_reglur_memory import ConsistencyRegulr
    rank, world, dev = dp.init_distributed()
    C, S, B = 5, 128, 4
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
    reducer = dp.GradientAllReducer(opt.flat_grad, num_buckets=4, unet=model.backbone, optimizer=opt)
    model.aux_path.bank_sync = dp.make_bank_sync(0)
    ok = True
    names = {}
    for n_, p_ in model.named_parameters():
        for q, off in zip(opt.params, opt.offsets):
            if q is p_:
                names[n_] = (off, off + p_.numel())
    diag = os.environ.get("PP_DP_DIAG", "0") == "1"   # also split the error into repeatability and exchange parts
    for it in range(3):
        batch = {k: v.to(dev) for k, v in make_batch(B, C, S, S, seed=dp.shard_seed(1234, rank, it)).items()}
        def fwd_bwd():
            out = model(batch, mode='train', step=3)
            loss = out['loss_pce'] + out['loss_ent'] + out['loss_cr'] + 0.01 * out['loss_aux_cls'] + out['loss_memory']
            opt.zero_grad()
            loss.backward()

        # pass 1 (no exchange): this rank's own gradient; the state the forward pass mutates (BatchNorm running
        # statistics, memory bank) is restored so that pass 2 recomputes exactly the same gradient
        state = {k: v.clone() for k, v in model.state_dict().items()}
        fwd_bwd()
        torch.cuda.synchronize(dev)
        local = opt.flat_grad.clone()
        model.load_state_dict(state)
        # pass 2: backward with the bucketed all-reduce overlapping it
        fwd_bwd()
        local2 = None
        if diag:   # (synchronises: the exchange then no longer overlaps the backward pass)
            torch.cuda.synchronize(dev)
            local2 = opt.flat_grad.clone()
        reducer.allreduce()
        torch.cuda.synchronize(dev)
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        ref = gathered[0].double()
        for g in gathered[1:]:
            ref += g.double()
        err = float((opt.flat_grad.double() - ref).abs().max() / ref.abs().max())
        banks = [torch.empty_like(model.aux_path.memory_bank.data) for _ in range(world)]
        dist.all_gather(banks, model.aux_path.memory_bank.data)
        same_bank = all(torch.equal(banks[0], b) for b in banks[1:])
        nz = float(local.abs().max())
        if rank == 0 and (diag or err >= 1e-5):
            d = (opt.flat_grad.double() - ref).abs()
            worst = max(names.items(), key=lambda kv: float(d[kv[1][0]:kv[1][1]].max()))
            lo, hi = worst[1]
            print("    worst parameter %s: |diff| max %.3e, |ref| max %.3e" % (worst[0], float(d[lo:hi].max()),
                                                                            float(ref[lo:hi].abs().max())), flush=True)
        if local2 is not None:
            rep = (local2.double() - local.double()).abs()
            worst = max(names.items(), key=lambda kv: float(rep[kv[1][0]:kv[1][1]].max()))
            print("    rank %d pass 2 vs pass 1 (no exchange): max |diff| %.3e / |g|max %.3e, worst %s" % (
                rank, float(rep.max()), nz, worst[0]), flush=True)
        if rank == 0:
            print("step %d: buckets %s max rel err %.3e, |g|max %.3e, bank identical %s" % (
                it, [(fl, hi - lo) for fl, lo, hi in reducer.overlapped] + reducer.tail, err, nz, same_bank), flush=True)
        ok = ok and err < 1e-5 and same_bank and nz > 0
        opt.step()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP CHECK %s (world %d)" % ("OK" if flag.item() > 0 else "FAIL", world), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() > 0 else 1)


if __name__ == "__main__":
    main()
