"""Time pp_bn_bwd (reduce + apply launches) ALONE at the UNet's layer shapes (24 images = 12 weak + 12 strong, two
statistics groups), CUDA events, L2 flushed between repetitions. Prints algorithmic GB/s: reads da, y twice (reduce and
apply) and writes dy once = 10 B per bf16 element. Usage: python tests/bench_bn_kernels.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pacingpseudo_b200 import lib as pplib  # noqa: E402

SHAPES = [(256, 32), (128, 64), (64, 128), (32, 256), (32, 512)]


def main():
    L = pplib.get_lib()
    L.ensure_init(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    G, N = 2, 24
    print("env PP_BN_RED_BPS=%s" % os.environ.get("PP_BN_RED_BPS"))
    total = 0.0
    for hw, C in SHAPES:
        Pg = (N // G) * hw * hw
        da = torch.randn(N, hw, hw, C, device="cuda").bfloat16()
        y = torch.randn(N, hw, hw, C, device="cuda").bfloat16()
        dy = torch.empty_like(da)
        coef = torch.rand(G * 4 * C, device="cuda") + 0.5
        bsums = torch.zeros(2 * G * C, dtype=torch.float64, device="cuda")
        bcoef = torch.zeros(2 * G * C, device="cuda")
        dg, db, dbias = (torch.zeros(C, device="cuda") for _ in range(3))
        ts = []
        for it in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.call("pp_bn_bwd", pplib.BF16, p(da), p(y), p(coef), p(bsums), p(bcoef), p(dg), p(db), p(dbias), p(dy), G, Pg,
                   C, 1, 0.01, st)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        t = ts[len(ts) // 2]
        total += t
        nbytes = 10 * N * hw * hw * C
        print("%3dx%-3d C=%-3d  %6.1f us  %5.0f GB/s (memset + reduce + apply)" % (hw, hw, C, t * 1e3, nbytes / t / 1e6))
    print("sum %.1f us" % (total * 1e3))


if __name__ == "__main__":
    main()
