"""Time the fused scribble-loss kernels ALONE through the C ABI (CUDA events, L2 flushed between repetitions) and
print achieved algorithmic GB/s against MEASURED_PEAKS.json. Usage: python tests/bench_loss_kernels.py [N C H W]..."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pacingpseudo_b200 import lib as pplib  # noqa: E402


def run(N, C, H, W, reps=20):
    L = pplib.get_lib()
    L.ensure_init(0)
    g = torch.Generator(device="cuda").manual_seed(1)
    zw, zs, za = (torch.randn(N, C, H, W, device="cuda", generator=g) for _ in range(3))
    target = torch.full((N, H, W), C, dtype=torch.uint8, device="cuda")
    lab = torch.rand(N, H, W, device="cuda", generator=g) < 0.01
    target[lab] = torch.randint(0, C, (int(lab.sum()),), device="cuda", dtype=torch.uint8)
    mask = (torch.rand(N, 1, H, W, device="cuda", generator=g) < 0.8).float()
    acc = torch.zeros(8, dtype=torch.float64, device="cuda")
    outs = [torch.zeros((), device="cuda") for _ in range(4)]
    gs = [torch.ones((), device="cuda") for _ in range(4)]
    d = [torch.empty_like(zw) for _ in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    P = N * H * W
    bytes_f = P * (3 * 4 * C + 1 + 4)
    bytes_b = bytes_f + P * 3 * 4 * C

    def fwd():
        L.call("pp_scribble_loss_fwd", p(zw), p(zs), p(za), p(target), p(mask), p(acc), *[p(o) for o in outs], N, C, H * W,
               C, 1, 1, st)

    def bwd():
        L.call("pp_scribble_loss_bwd", p(zw), p(zs), p(za), p(target), p(mask), p(acc), *[p(x) for x in gs],
               *[p(x) for x in d], N, C, H * W, C, 1, 1, 0, st)

    res = {}
    for name, fn, nbytes in (("fwd", fwd, bytes_f), ("bwd", bwd, bytes_b)):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        res[name] = (ts[len(ts) // 2] * 1e3, nbytes / (ts[len(ts) // 2] * 1e-3) / 1e9)
    return res


if __name__ == "__main__":
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    args = [int(a) for a in sys.argv[1:]] or [12, 5, 256, 256, 48, 5, 256, 256, 96, 2, 224, 224, 96, 4, 224, 224]
    print("env PP_LOSS_V=%s PP_LOSS_GENERIC=%s" % (os.environ.get("PP_LOSS_V"), os.environ.get("PP_LOSS_GENERIC")))
    for i in range(0, len(args), 4):
        N, C, H, W = args[i:i + 4]
        r = run(N, C, H, W)
        print("N=%d C=%d %dx%d: fwd %.1f us %.0f GB/s | bwd %.1f us %.0f GB/s (cold L2; fwd includes the 1-thread finalize launch)"
              % (N, C, H, W, r["fwd"][0], r["fwd"][1], r["bwd"][0], r["bwd"][1]))
