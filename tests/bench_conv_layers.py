"""Per-layer timing of the bf16 forward / dgrad convolution launches of the default UNet at the bench's shapes
(12 slices per forward branch, 24 in the backward pass): CUDA events around 20 back-to-back launches after 3 warm-ups.
Usage: python tests/bench_conv_layers.py [tag]   (environment selects kernels, e.g. PP_CONV_ROWS=0)"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pacingpseudo_b200 import lib as pplib
from pacingpseudo_b200 import functional as PF
L = pplib.get_lib(); L.ensure_init(0); torch.cuda.set_device(0)
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
# (name, N, H, W, C0, C1, Cout, dil)
LAYERS = [
    ("enc3a 64->128 @64", 12, 64, 64, 64, 0, 128, 1), ("enc3b 128->128 @64", 12, 64, 64, 128, 0, 128, 1),
    ("enc4a 128->256 @32", 12, 32, 32, 128, 0, 256, 1), ("enc4b 256->256 @32", 12, 32, 32, 256, 0, 256, 1),
    ("enc5a 256->512 d2", 12, 32, 32, 256, 0, 512, 2), ("enc5b 512->512 d2", 12, 32, 32, 512, 0, 512, 2),
    ("enc6 512->512 d4", 12, 32, 32, 512, 0, 512, 4), ("dec5a 1024->512", 12, 32, 32, 512, 512, 512, 1),
    ("dec5b 512->512", 12, 32, 32, 512, 0, 512, 1), ("dec4a 768->256", 12, 32, 32, 512, 256, 256, 1),
    ("dec4b 256->256", 12, 32, 32, 256, 0, 256, 1), ("dec3a 384->128 @64", 12, 64, 64, 256, 128, 128, 1),
    ("dec3b 128->128 @64", 12, 64, 64, 128, 0, 128, 1), ("dec2a 192->64 @128", 12, 128, 128, 128, 64, 64, 1),
    # dgrad launches (full batch of 24): input = dY (Cout channels), outputs = the sources
    ("dgrad dec5b 512->512", 24, 32, 32, 512, 0, 512, 1), ("dgrad dec5a 512->1024", 24, 32, 32, 512, 0, 1024, 1),
    ("dgrad enc6 512->512 d4", 24, 32, 32, 512, 0, 512, 4), ("dgrad enc5b d2", 24, 32, 32, 512, 0, 512, 2),
    ("dgrad dec4a 256->768", 24, 32, 32, 256, 0, 768, 1), ("dgrad dec3a 128->384 @64", 24, 64, 64, 128, 0, 384, 1),
    ("dgrad enc3b 128->128 @64", 24, 64, 64, 128, 0, 128, 1), ("dgrad enc4b 256->256", 24, 32, 32, 256, 0, 256, 1),
]
tag = sys.argv[1] if len(sys.argv) > 1 else ""
if os.environ.get("PP_LAYERS"):
    LAYERS = [l for l in LAYERS if any(k in l[0] for k in os.environ["PP_LAYERS"].split(","))]
tot = 0.0
for name, N, H, W, C0, C1, Co, dil in LAYERS:
    x0 = torch.randn(N, H, W, C0, device="cuda").bfloat16()
    x1 = torch.randn(N, H, W, C1, device="cuda").bfloat16() if C1 else None
    w = torch.randn(Co, C0 + C1, 3, 3, device="cuda") / (3 * (C0 + C1) ** 0.5)
    wf = torch.empty(9 * Co * (C0 + C1) * 2, dtype=torch.uint8, device="cuda"); wd = torch.empty_like(wf)
    L.call("pp_pack_weights", PF.BF16, p(w), p(wf), p(wd), Co, C0 + C1, st())
    y = torch.empty(N, H, W, Co, device="cuda", dtype=torch.bfloat16)
    b = torch.zeros(Co, device="cuda")
    def run():
        L.call("pp_conv3x3", PF.BF16, p(x0), C0, p(x1), C1, p(wf), p(b), p(y), Co, 0, None, 0, 0, N, H, W, dil, st())
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * N * H * W * 9 * (C0 + C1) * Co
    tot += us
    print("%-6s %-28s %8.1f us  %7.1f TF/s" % (tag, name, us, fl / us / 1e6), flush=True)
print("%-6s total %.1f us" % (tag, tot))
