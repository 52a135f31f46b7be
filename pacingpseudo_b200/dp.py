"""Data-parallel layer (NEW functionality: the reference is single-GPU).

One process per GPU (torchrun). The batch is sharded by rank (each rank draws its own local batch: weak
scaling); BatchNorm statistics stay per rank, so every rank is exactly a reference single-GPU step on its
local batch; gradients are summed with NCCL all-reduce over bucket slices of the flat gradient buffer
(pacingpseudo_b200/optim.py) on a side stream and averaged by folding 1/world into the Adam kernel; the memory
bank, which the forward pass mutates from local sample 0, is taken from rank 0 (1.3 KB broadcast) so
loss_memory and its gradient agree on all ranks.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, device)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world)
    return rank, world, torch.device("cuda", local) if use_cuda else torch.device("cpu")


class GradientAllReducer:
    """Sum-all-reduce of a flat gradient buffer in `num_buckets` contiguous slices, on a side stream when on CUDA."""

    def __init__(self, flat_grad, num_buckets=4, group=None):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = flat_grad.numel()
        step = (n + num_buckets - 1) // num_buckets
        step = (step + 127) // 128 * 128
        self.buckets = [(o, min(o + step, n)) for o in range(0, n, step)]
        self.stream = torch.cuda.Stream(flat_grad.device) if flat_grad.is_cuda else None

    def allreduce(self):
        if self.world == 1:
            return
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.stream):
                works = [dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                         for a, b in self.buckets]
                for w in works:
                    w.wait()
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
        else:
            for a, b in self.buckets:
                dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group)


def make_bank_sync(src=0, group=None):
    """-> callable(bank) installed as AuxPath.bank_sync: rank `src`'s freshly updated bank wins."""
    def sync(bank):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(bank, src=src, group=group)
    return sync


def shard_seed(base_seed, rank, step):
    """Seed of rank `rank`'s local batch at `step` (1234 + 1000*rank + step)."""
    return base_seed + 1000 * rank + step


def average_gradients_emulated(per_rank_grads):
    """Single-process emulation of the exchange: list of {name: grad} -> {name: mean grad}."""
    world = len(per_rank_grads)
    return {k: sum(g[k] for g in per_rank_grads) / world for k in per_rank_grads[0]}


def max_over_ranks(value, device):
    """Max of a python float over all ranks (timings are reported as the max over ranks)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
