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
        # NCCL writes its banner / debug lines to stdout by default; keep stdout for the caller's own output
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world)
    return rank, world, torch.device("cuda", local) if use_cuda else torch.device("cpu")


def plan_layer_buckets(layer_spans, num_buckets):
    """Backward-ordered bucket plan over contiguous per-layer spans of the flat gradient buffer.

    layer_spans: [(lo, hi)] per conv layer in FORWARD order (+ the head folded into the last span), contiguous and
    increasing. The backward pass completes the layers last-to-first, so buckets are cut walking from the last
    layer down, each closed once it holds >= 1/num_buckets of the elements; the remaining early layers form the
    final bucket. Returns [(first_layer, lo, hi)]: bucket k = slice [lo, hi), complete once `first_layer`'s
    gradients are; first_layer == 0 marks the final bucket (ready only when the whole backward pass is).
    """
    total = layer_spans[-1][1] - layer_spans[0][0]
    target = max(1, total // max(1, num_buckets))
    plan, hi, acc = [], layer_spans[-1][1], 0
    for i in range(len(layer_spans) - 1, 0, -1):
        acc += layer_spans[i][1] - layer_spans[i][0]
        if acc >= target and len(plan) < num_buckets - 1:
            plan.append((i, layer_spans[i][0], hi))
            hi, acc = layer_spans[i][0], 0
    plan.append((0, layer_spans[0][0], hi))
    return plan


class GradientAllReducer:
    """Sum-all-reduce of a flat gradient buffer in contiguous bucket slices on a communication stream.

    With `unet` (the drop-in models.unet.UNet whose parameters live in `optimizer`'s flat buffer) the buckets are cut
    on layer boundaries in the order the backward pass completes them, and each bucket's all-reduce waits only for
    the event pp_unet_backward records when those layers' gradients are done (include/pacingpseudo_b200.h,
    pp_unet_set_grad_events): the exchange overlaps the remaining backward kernels. Gradients outside the UNet span
    (aux path) and the earliest layers go last, after the whole backward pass. Without `unet`: `num_buckets` equal
    slices after the backward pass (also the CPU / gloo path).
    """

    def __init__(self, flat_grad, num_buckets=4, group=None, unet=None, optimizer=None):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = flat_grad.numel()
        self.stream = torch.cuda.Stream(flat_grad.device) if flat_grad.is_cuda else None
        self.unet = None
        # A zero-padded UNet (max_ch = 728) hands padded temporaries to the kernels; autograd adds their sliced gradients
        # to the flat buffer only AFTER pp_unet_backward has returned, i.e. after the per-layer events fired: such a model
        # is reduced after the whole backward pass (no overlap) instead of racing the accumulation.
        if unet is not None and getattr(unet, "_padded", False):
            unet = None
        if unet is not None and optimizer is not None and flat_grad.is_cuda and self.world > 1:
            # (a single process needs no events: pp_unet_backward then stays eligible for CUDA-graph replay)
            spans = self._layer_spans(unet, optimizer)
            plan = plan_layer_buckets(spans, num_buckets)
            self.overlapped = plan[:-1]                        # [(first_layer, lo, hi)] each with its own event
            lo_tail, hi_tail = plan[-1][1], plan[-1][2]
            self.tail = [(a, b) for a, b in ((0, hi_tail), (spans[-1][1], n)) if b > a]   # after the whole backward
            assert lo_tail == spans[0][0]
            self.buckets = [(lo, hi) for _, lo, hi in self.overlapped] + self.tail
            self.unet = unet
            import ctypes
            layers = [fl for fl, _, _ in self.overlapped]
            arr = (ctypes.c_int * max(1, len(layers)))(*layers)
            unet.engine.lib.call("pp_unet_set_grad_events", unet.engine.handle, len(layers), arr)
        else:
            step = (n + num_buckets - 1) // num_buckets
            step = (step + 127) // 128 * 128
            self.buckets = [(o, min(o + step, n)) for o in range(0, n, step)]

    @staticmethod
    def _layer_spans(unet, optimizer):
        where = {id(p): (off, p.numel()) for p, off in zip(optimizer.params, optimizer.offsets)}
        spans = []
        mods = unet._layer_modules()
        for i, m in enumerate(mods):
            ps = ([m.weight] if not hasattr(m, "conv") else   # ConvTranspose2d of the trans-conv variant
                  [m.conv.weight, m.conv.bias, m.norm_op.weight, m.norm_op.bias])
            if i == len(mods) - 1:
                ps += [unet.final_conv.weight, unet.final_conv.bias]
            locs = [where[id(p)] for p in ps]
            spans.append((min(o for o, _ in locs), max(o + k for o, k in locs)))
        for a, b in zip(spans, spans[1:]):   # the flat buffer follows module registration order == layer order
            if a[1] > b[0]:
                raise RuntimeError("GradientAllReducer: UNet parameters are not laid out in layer order")
        spans = [(lo, nxt[0]) for (lo, _), nxt in zip(spans, spans[1:])] + [spans[-1]]   # absorb alignment padding
        return spans

    def allreduce(self):
        """Call right after loss.backward() (all backward kernels enqueued on the current stream)."""
        if self.world == 1:
            return
        if self.stream is None:
            for a, b in self.buckets:
                dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group)
            return
        cur = torch.cuda.current_stream(self.flat.device)
        works = []
        with torch.cuda.stream(self.stream):
            if self.unet is not None:
                import ctypes
                eng = self.unet.engine
                st = ctypes.c_void_p(self.stream.cuda_stream)
                for i, (_, a, b) in enumerate(self.overlapped):
                    eng.lib.call("pp_unet_wait_grad_event", eng.handle, i, st)
                    works.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                tail = self.tail
            else:
                tail = self.buckets
            self.stream.wait_stream(cur)
            for a, b in tail:
                works.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for w in works:
                w.wait()
        cur.wait_stream(self.stream)


def make_bank_sync(src=0, group=None):
    """-> callable(bank) installed as AuxPath.bank_sync: rank `src`'s freshly updated bank wins."""
    def sync(bank):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(bank, src=src, group=group)
    return sync


def sync_bn_buffers(model, src=0, group=None):
    """Call before validation / checkpointing under data parallelism: BatchNorm statistics are per rank during
    training (each rank is exactly a reference single-GPU step), so running_mean / running_var /
    num_batches_tracked are taken from rank `src` — eval-mode forward passes and checkpoints then agree on every rank.
    One broadcast over a flat copy of all floating-point buffers plus one for the integer counters."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    bufs = [b for b in model.buffers()]
    for kind in (True, False):
        sel = [b for b in bufs if b.is_floating_point() == kind]
        if not sel:
            continue
        flat = torch.cat([b.detach().reshape(-1).to(torch.float32 if kind else torch.int64) for b in sel])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for b in sel:
            n = b.numel()
            b.data.copy_(flat[off:off + n].view_as(b).to(b.dtype))
            off += n


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
