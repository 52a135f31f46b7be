"""CPU: the inverse index map used by the bilinear-upsample backward kernel (csrc/ops.cu: lerp_first_out / up_taps,
restated here in numpy float32 arithmetic) reproduces, for every input index, exactly the set of outputs and weights
with which the forward map (ATen's align_corners=True source index rule, aux_path_memory.py:52 / unet.py:144) reads it."""
import numpy as np
import pytest

f32 = np.float32


def lerp_src(dst, in_size, scale):
    r = f32(scale) * f32(dst)
    i0 = int(r)
    i1 = i0 + (1 if i0 < in_size - 1 else 0)
    w1 = f32(r) - f32(i0)
    return i0, i1, f32(1) - w1, w1


def first_out(i, in_size, out_size, scale):
    if i <= 0:
        return 0
    if i >= in_size or scale <= 0:
        return out_size
    a = max(0, min(out_size, int(f32(i) / f32(scale))))
    while a > 0 and int(f32(scale) * f32(a - 1)) >= i:
        a -= 1
    while a < out_size and int(f32(scale) * f32(a)) < i:
        a += 1
    return a


def taps(i, in_size, out_size, scale):
    a0, a1, a2 = (first_out(k, in_size, out_size, scale) for k in (i - 1, i, i + 1))
    lo = a1 if i == 0 else a0
    w = []
    for o in range(lo, a2):
        i0, i1, w0, w1 = lerp_src(o, in_size, scale)
        w.append((w0 if i0 == i else f32(0)) + (w1 if i1 == i else f32(0)))
    return lo, w


@pytest.mark.parametrize("sizes", [(32, 64), (64, 128), (128, 256), (28, 56), (14, 28), (112, 224), (4, 8), (1, 2),
                                   (2, 4), (3, 9), (32, 256), (5, 5), (7, 20), (16, 17)])
def test_inverse_tap_map_matches_forward_map(sizes):
    in_size, out_size = sizes
    scale = f32(in_size - 1) / f32(out_size - 1) if out_size > 1 else f32(0)
    W = np.zeros((in_size, out_size))
    for o in range(out_size):
        i0, i1, w0, w1 = lerp_src(o, in_size, scale)
        W[i0, o] += w0
        W[i1, o] += w1
    for i in range(in_size):
        lo, w = taps(i, in_size, out_size, scale)
        T = np.zeros(out_size)
        T[lo:lo + len(w)] = w
        assert np.array_equal(T, W[i]), (sizes, i)
    assert np.allclose(W.sum(0), 1.0, atol=1e-6)
