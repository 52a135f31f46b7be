"""CPU, world_size 2, gloo: the host-side data-parallel logic (bucketed gradient all-reduce + averaging
convention, bank broadcast from rank 0, per-rank batch seeds, max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from pacingpseudo_b200 import dp
    r, w, dev = dp.init_distributed("gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    g = torch.Generator().manual_seed(dp.shard_seed(1234, rank, 0))
    flat = torch.randn(1000, generator=g)
    mine = flat.clone()
    red = dp.GradientAllReducer(flat, num_buckets=3)
    assert red.buckets[0][0] == 0 and red.buckets[-1][1] == 1000 and len(red.buckets) == 3
    red.allreduce()
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    ok_sum = torch.allclose(flat, sum(gathered))
    bank = torch.full((5, 64), float(rank + 1))
    dp.make_bank_sync(0)(bank)
    ok_bank = bool((bank == 1.0).all())
    t = dp.max_over_ranks(float(rank), dev)
    # BatchNorm buffers are per rank during training; before validation / checkpoints rank 0's win
    bn = torch.nn.Sequential(torch.nn.BatchNorm2d(4), torch.nn.BatchNorm2d(3))
    for m in bn:
        m.running_mean.fill_(float(rank + 1))
        m.running_var.fill_(10.0 * (rank + 1))
        m.num_batches_tracked.fill_(7 * (rank + 1))
    dp.sync_bn_buffers(bn, src=0)
    ok_bn = all(bool((m.running_mean == 1.0).all()) and bool((m.running_var == 10.0).all())
                and int(m.num_batches_tracked) == 7 and m.num_batches_tracked.dtype == torch.int64 for m in bn)
    ok_bank = ok_bank and ok_bn
    q.put((rank, ok_sum, ok_bank, t, dp.shard_seed(1234, rank, 7)))
    dist.destroy_process_group()


def test_dp_host_logic_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1:4] == (True, True, 1.0) and res[1][1:4] == (True, True, 1.0)
    assert res[0][4] == 1241 and res[1][4] == 2241


def test_layer_bucket_plan_follows_backward_order():
    """Buckets are cut on layer boundaries from the last layer down; the early layers form the final bucket."""
    from pacingpseudo_b200.dp import plan_layer_buckets
    sizes = [416, 9312, 18560, 36992, 73984, 147712, 295424, 590336, 1180672, 2360320, 2360320, 2360320, 4719616,
             2360320, 1770240, 590336, 442624, 147712, 110720, 36992, 27712, 9477]   # the default UNet, head folded in
    spans, off = [], 0
    for n in sizes:
        spans.append((off, off + n))
        off += n
    plan = plan_layer_buckets(spans, 4)
    assert [b[0] for b in plan] == [13, 11, 8, 0]
    assert plan[0][2] == off and plan[-1][1] == 0
    for (_, lo, hi), (_, lo2, hi2) in zip(plan, plan[1:]):   # contiguous, descending, no gaps
        assert lo == hi2 and lo2 < hi2
    assert all(lo == spans[fl][0] for fl, lo, _ in plan)
    assert plan_layer_buckets(spans, 1) == [(0, 0, off)]


def test_padded_unet_is_reduced_after_backward_not_overlapped():
    """A zero-padded UNet (max_ch=728) gets its padded layers' gradients from autograd after pp_unet_backward returned:
    GradientAllReducer must not arm the per-layer events for it (host logic, no GPU needed)."""
    import torch
    from pacingpseudo_b200 import dp

    class FakeUNet:
        _padded = True

        @property
        def engine(self):
            raise AssertionError("the executor must not be touched for a padded model")

    flat = torch.zeros(1000)
    red = dp.GradientAllReducer(flat, num_buckets=4, unet=FakeUNet(), optimizer=object())
    assert red.unet is None and sum(b - a for a, b in red.buckets) == 1000
