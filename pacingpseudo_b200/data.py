"""Host <-> device plumbing around the training step (train_chaos.py:263-310 moves five tensors per step with
blocking `.cuda()` calls and reads five losses with blocking `.item()` calls).

`DevicePrefetcher` wraps any iterable of host batches (dicts of CPU tensors, pinned for true asynchrony) and copies
batch i+1 into one of two persistent device staging buffers on a dedicated copy stream while step i computes, so
the 47 MB/step of the reference's input format (fp32 one-hot scribbles) hide behind the kernels. No allocator
traffic in steady state: the staging buffers are allocated once per (key, shape, dtype); a buffer is overwritten
only after the step that consumed it has been fully enqueued AND finished on the device (event recorded on the
consumer's stream when the next batch is requested).

`LossReader` copies the step's loss scalars to pinned host memory without blocking and hands them back one step
later, so logging does not drain the GPU queue every iteration.
"""
import torch


class DevicePrefetcher:
    def __init__(self, batches, device):
        self.it = iter(batches)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher: target must be a CUDA device")
        self.stream = torch.cuda.Stream(self.device)
        self.slots = [{}, {}]            # persistent device staging buffers, reused every other batch
        self.ready = [None, None]        # copy-stream event: slot filled
        self.released = [None, None]     # consumer-stream event: slot no longer read
        self.pending = None              # slot index of the batch copied ahead
        self.current = None              # slot index handed to the consumer
        self.count = 0
        self.h2d_bytes = 0

    def reset(self, batches):
        """Re-arm with a new iterable (next epoch); stream and staging buffers are kept, so no allocation happens."""
        if self.current is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self.released[self.current] = ev
        self.it = iter(batches)
        self.pending, self.current, self.count = None, None, 0
        return self

    def _issue(self, slot):
        try:
            host = next(self.it)
        except StopIteration:
            return False
        bufs = self.slots[slot]
        with torch.cuda.stream(self.stream):
            if self.released[slot] is not None:
                self.stream.wait_event(self.released[slot])
            out = {}
            for k, v in host.items():
                if not torch.is_tensor(v):
                    out[k] = v
                    continue
                b = bufs.get(k)
                if b is None or b.shape != v.shape or b.dtype != v.dtype:
                    b = bufs[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                b.copy_(v, non_blocking=True)
                out[k] = b
                self.h2d_bytes += v.numel() * v.element_size()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready[slot] = (out, ev)
        return True

    def __iter__(self):
        return self

    def __next__(self):
        cur = torch.cuda.current_stream(self.device)
        if self.current is not None:     # everything that read the previous batch is enqueued by now
            ev = torch.cuda.Event()
            ev.record(cur)
            self.released[self.current] = ev
        if self.count == 0:
            self.pending = 0 if self._issue(0) else None
        if self.pending is None:
            raise StopIteration
        slot = self.pending
        out, ev = self.ready[slot]
        self.current = slot
        self.count += 1
        self.pending = (1 - slot) if self._issue(1 - slot) else None   # copy the next batch while this one computes
        cur.wait_event(ev)
        return out


class LossReader:
    """push(list of 0-dim CUDA tensors) -> the PREVIOUS push's values as Python floats (None the first time);
    flush() -> the last push's values. One small async D2H per step into a pinned ring, no queue drain."""

    def __init__(self, device, width=8):
        self.device = torch.device(device)
        self.host = [torch.zeros(width, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.dev = [torch.zeros(width, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.events = [None, None]
        self.n = [0, 0]
        self.i = 0
        self.d2h_bytes = 0

    def _collect(self, slot):
        if self.events[slot] is None:
            return None
        self.events[slot].synchronize()
        vals = self.host[slot][:self.n[slot]].tolist()
        self.events[slot] = None
        return vals

    def push(self, tensors):
        slot = self.i & 1
        prev = self._collect(1 - slot)
        n = len(tensors)
        torch.stack([t.detach().reshape(()) for t in tensors], out=self.dev[slot][:n])
        self.host[slot][:n].copy_(self.dev[slot][:n], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.events[slot], self.n[slot] = ev, n
        self.d2h_bytes += 4 * n
        self.i += 1
        return prev

    def flush(self):
        return self._collect((self.i - 1) & 1) if self.i else None


# ------------------------------------------------------------------------------------------------
# Compact input format + device-side strong augmentation
# ------------------------------------------------------------------------------------------------
def sample_strong_params(n, strength=1.0, generator=None):
    """Per-image parameters of TransformsColor.get_strong_transforms(strength) (chaos_aug_configs.py:63-86), drawn with
    the reference's probabilities and ranges: Brightness U(-0.8 s, 0.8 s) p=0.8; Contrast U(max(0, 1-0.8 s), 1+0.8 s)
    p=0.8; Gamma p=0.8, with probability 1/2 from [lo, 1) when lo < 1, else from [max(1, lo), hi]
    (augmentations.py:150-153). -> CPU fp32 tensor [n, 8] for strong_color_augment."""
    u = lambda *shape: torch.rand(*shape, generator=generator)
    lo, hi = max(0.0, 1 - strength * 0.8), 1 + strength * 0.8
    p = torch.zeros(n, 8)
    p[:, 0] = (u(n) < 0.8).float()
    p[:, 1] = (u(n) * 2 - 1) * strength * 0.8
    p[:, 2] = (u(n) < 0.8).float()
    p[:, 3] = lo + u(n) * (hi - lo)
    p[:, 4] = (u(n) < 0.8).float()
    low_half = (u(n) < 0.5) & (lo < 1.0)
    g_lo = lo + u(n) * (1.0 - lo)
    g_hi = max(1.0, lo) + u(n) * (hi - max(1.0, lo))
    p[:, 5] = torch.where(low_half, g_lo, g_hi)
    return p


def strong_color_augment(image, params):
    """image_strong from the weak image ON THE DEVICE: Brightness -> Contrast -> GammaAugmentation(retain_stats)
    (datasets/augmentations.py:98-166) with the given per-image parameters ([N, 8], see sample_strong_params).
    image: CUDA fp32 (N, 1, H, W). One hand-written kernel (pp_strong_color_augment); no CPU path."""
    from .lib import current_stream, get_lib, ptr, require_cuda
    require_cuda(image, "image")
    image = image.contiguous().float()
    n = image.shape[0]
    hw = image[0].numel()
    params = params.to(device=image.device, dtype=torch.float32).contiguous()
    if tuple(params.shape) != (n, 8):
        raise ValueError("strong_color_augment: params must be [N, 8], got %r" % (tuple(params.shape),))
    out = torch.empty_like(image)
    with torch.cuda.device(image.device):
        get_lib().call("pp_strong_color_augment", ptr(image), ptr(params), ptr(out), n, hw, current_stream(image.device))
    return out


def compact_batch(batch, num_classes):
    """Host-side: the reference batch (fp32 one-hot `scribble` (N, C+1, H, W), train_chaos.py:264-269) in the compact
    format the drop-in modules also accept: `scribble` as a uint8 class-index map (N, H, W); `image_strong` and the
    unused `scribble_strong` dropped (the strong image is produced on the device by strong_color_augment)."""
    out = {"image": batch["image"], "scribble": batch["scribble"].argmax(1).to(torch.uint8)}
    if "valid_mask" in batch:
        out["valid_mask"] = batch["valid_mask"]
    return out
