"""The narrow-layer weight-gradient launches of the default UNet timed ALONE at the bench's shapes (24 slices): CUDA
events around 20 back-to-back launches after 3 warm-ups, L2 not flushed (operands are 100-300 MB, larger than L2).
Prints algorithmic HBM traffic (X and dY read once) and TF/s. PP_WGRAD_ROWS=1 selects the round-1 row kernel (horizontal
taps as A-operand row offsets, three N = Cout MMAs per 16 pixels), the default (2) the kernel with the horizontal taps
packed into N. Usage: python tests/bench_wgrad_narrow.py [tag]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from pacingpseudo_b200 import lib as pplib  # noqa: E402
from pacingpseudo_b200 import functional as PF  # noqa: E402

L = pplib.get_lib()
L.ensure_init(0)
torch.cuda.set_device(0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)  # noqa: E731
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())  # noqa: E731
# (name, N, H, W, C0, C1, Cout)
LAYERS = [("enc1b / dec1b 32->32 @256", 24, 256, 256, 32, 0, 32), ("dec1a 64+32->32 @256", 24, 256, 256, 64, 32, 32),
          ("enc2a 32->64 @128", 24, 128, 128, 32, 0, 64), ("enc2b / dec2b 64->64 @128", 24, 128, 128, 64, 0, 64),
          ("dec2a 128+64->64 @128", 24, 128, 128, 128, 64, 64)]
tag = sys.argv[1] if len(sys.argv) > 1 else "rows=%s" % os.environ.get("PP_WGRAD_ROWS", "2")
tot = 0.0
for name, N, H, W, C0, C1, Co in LAYERS:
    x0 = torch.randn(N, H, W, C0, device="cuda").bfloat16()
    x1 = torch.randn(N, H, W, C1, device="cuda").bfloat16() if C1 else None
    dy = torch.randn(N, H, W, Co, device="cuda").bfloat16()
    dwp = torch.zeros(9 * Co * (C0 + C1), device="cuda")

    def run():
        L.call("pp_conv3x3_wgrad", PF.BF16, p(dy), Co, p(x0), C0, p(x1), C1, p(dwp), N, H, W, 1, st())
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * N * H * W * 9 * (C0 + C1) * Co
    nbytes = 2.0 * N * H * W * (C0 + C1 + Co)
    tot += us
    print("%-8s %-28s %8.1f us  %6.0f GB/s  %6.1f TF/s" % (tag, name, us, nbytes / us / 1e3, fl / us / 1e6), flush=True)
print("%-8s total %.1f us" % (tag, tot))
