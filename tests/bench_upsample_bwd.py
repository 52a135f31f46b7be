"""Time pp_upsample_nhwc_bwd ALONE at the UNet's three scale-2 decoder shapes (24 images), CUDA events, L2 flushed.
PP_UPSAMPLE_STRIP=0 selects the per-pixel gather kernel, default the strip kernel. Usage: python tests/bench_upsample_bwd.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pacingpseudo_b200 import lib as pplib  # noqa: E402


def main():
    L = pplib.get_lib()
    L.ensure_init(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    N = 24
    print("env PP_UPSAMPLE_STRIP=%s" % os.environ.get("PP_UPSAMPLE_STRIP"))
    tot = 0.0
    for h, C in ((128, 64), (64, 128), (32, 256)):
        gy = torch.randn(N, 2 * h, 2 * h, C, device="cuda").bfloat16()
        gx = torch.empty(N, h, h, C, device="cuda", dtype=torch.bfloat16)
        ts = []
        for it in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.call("pp_upsample_nhwc_bwd", pplib.BF16, p(gy), p(gx), N, h, h, 2 * h, 2 * h, C, 0, st)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        t = ts[len(ts) // 2]
        tot += t
        nbytes = 2 * (gy.numel() + gx.numel())
        print("%3d->%3d C=%-3d %6.1f us  %5.0f GB/s algorithmic" % (h, 2 * h, C, t * 1e3, nbytes / t / 1e6))
    print("sum %.1f us" % (tot * 1e3))


if __name__ == "__main__":
    main()
