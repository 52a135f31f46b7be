"""Deterministic synthetic CHAOS/ACDC/LVSC-shaped batches: single-channel slices made of
soft ellipses, dense labels, one-pixel-wide scribbles, a strong (colour-jittered) view and a valid mask,
in exactly the tensor formats `*TwoStream.__getitem__` hands the training loop
(/root/reference/datasets/chaos/chaos_dataset.py:58-90): fp32 images (N,1,H,W), one-hot fp32 scribble
(N,C+1,H,W) whose last channel is the ignore class, one-hot fp32 label (N,C,H,W), fp32 valid_mask (N,1,H,W).
Generated on the CPU with a seeded torch.Generator so every rank / test / oracle sees identical data.
"""
import math

import torch


def make_batch(n, num_classes, height, width, seed, absent_class_in_sample0=None, class_contrast=False):
    """class_contrast=True: the k-th foreground class gets its own intensity band (amp = 0.35 + 0.45 k with 10 %
    jitter, sharper edges) so that class identity is learnable from appearance — used by the trained-state parity
    tests, where near-tied logits between visually identical classes would make the argmax comparison meaningless.
    The default (False) draws every amplitude from the same range and is what the golden fixtures were made with."""
    g = torch.Generator().manual_seed(int(seed))
    C = num_classes
    yy, xx = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32),
                            indexing='ij')
    image = torch.zeros(n, 1, height, width)
    label = torch.zeros(n, height, width, dtype=torch.long)
    scrib = torch.full((n, height, width), C, dtype=torch.long)
    mask = torch.ones(n, 1, height, width)
    for i in range(n):
        img = torch.zeros(height, width)
        k_fore = C - 1
        geo = []
        for k in range(1, k_fore + 1):
            cx = (0.2 + 0.6 * torch.rand((), generator=g).item()) * width
            cy = (0.2 + 0.6 * torch.rand((), generator=g).item()) * height
            a = (0.08 + 0.12 * torch.rand((), generator=g).item()) * width
            b = (0.06 + 0.10 * torch.rand((), generator=g).item()) * height
            th = math.pi * torch.rand((), generator=g).item()
            amp = 0.5 + torch.rand((), generator=g).item()
            if class_contrast:
                amp = (0.35 + 0.45 * k) * (0.95 + 0.1 * (amp - 0.5))
            geo.append((cx, cy, a, b, th, amp))
        for k, (cx, cy, a, b, th, amp) in enumerate(geo, start=1):
            if i == 0 and absent_class_in_sample0 == k:
                continue
            ct, st = math.cos(th), math.sin(th)
            u = ((xx - cx) * ct + (yy - cy) * st) / a
            v = (-(xx - cx) * st + (yy - cy) * ct) / b
            r2 = u * u + v * v
            img = img + amp * torch.exp(-1.5 * r2)
            inside = r2 < 1.0
            label[i][inside] = k
        # scribbles: a polyline along the major axis of every region that survived occlusion, plus background
        for k, (cx, cy, a, b, th, amp) in enumerate(geo, start=1):
            if i == 0 and absent_class_in_sample0 == k:
                continue
            t = torch.linspace(-0.6, 0.6, steps=int(2 * a) + 8)
            px = (cx + t * a * math.cos(th)).round().long().clamp(0, width - 1)
            py = (cy + t * a * math.sin(th)).round().long().clamp(0, height - 1)
            ok = label[i, py, px] == k
            scrib[i, py[ok], px[ok]] = k
        if not (i == 0 and absent_class_in_sample0 == 0):
            x0 = int(torch.randint(0, width, (1,), generator=g))
            t = torch.arange(height)
            px = (x0 + (0.3 * t).long()) % width
            ok = label[i, t, px] == 0
            scrib[i, t[ok], px[ok]] = 0
        img = img + 0.1 * torch.randn(height, width, generator=g)
        img = (img - img.mean()) / (img.std() + 1e-8)  # MeanStdNorm, chaos_aug_configs.py:54
        if torch.rand((), generator=g).item() < 0.25:  # padded / rotated border: zero image, invalid mask
            m = torch.zeros(height, width)
            y0 = int(torch.randint(0, height // 8 + 1, (1,), generator=g))
            x0 = int(torch.randint(0, width // 8 + 1, (1,), generator=g))
            m[y0:height - y0 // 2, x0:width - x0 // 2] = 1
            img = img * m
            mask[i, 0] = m
            scrib[i][m == 0] = C
        image[i, 0] = img
    # strong view: contrast / brightness / gamma jitter of the weak view (chaos_aug_configs.py:70-85)
    strong = torch.empty_like(image)
    for i in range(n):
        a = 0.2 + 1.6 * torch.rand((), generator=g).item()
        b = -0.8 + 1.6 * torch.rand((), generator=g).item()
        w = image[i]
        mu = w.mean()
        s = (w - mu) * a + mu + b
        if torch.rand((), generator=g).item() < 0.5:
            gam = 0.2 + 1.6 * torch.rand((), generator=g).item()
            lo, hi = s.min(), s.max()
            s = ((s - lo) / (hi - lo + 1e-8)).pow(gam) * (hi - lo) + lo
        strong[i] = s * mask[i]
    scribble_1h = torch.nn.functional.one_hot(scrib, C + 1).permute(0, 3, 1, 2).float().contiguous()
    label_1h = torch.nn.functional.one_hot(label, C).permute(0, 3, 1, 2).float().contiguous()
    return {
        'image': image.contiguous(), 'image_strong': strong.contiguous(), 'scribble': scribble_1h,
        'scribble_strong': scribble_1h.clone(), 'label': label_1h, 'valid_mask': mask.contiguous(),
    }


def to_device(batch, device, non_blocking=False):
    return {k: (v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
