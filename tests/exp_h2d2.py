import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacingpseudo_b200.synth import make_batch
dev = torch.device("cuda", 0)
host = [{k: v.pin_memory() for k, v in make_batch(12, 5, 256, 256, seed=i).items()} for i in range(2)]
side = torch.cuda.Stream(dev)
big = torch.randn(8192, 8192, device=dev)
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
slots = [{}, {}]
released = [None, None]
for it in range(12):
    if it == 2: T.clear()
    slot = it & 1
    hb = host[it & 1]
    cur = torch.cuda.current_stream(dev)
    t0 = time.perf_counter(); ev = torch.cuda.Event(); ev.record(cur); released[1 - slot] = ev; tick("release", t0)
    with torch.cuda.stream(side):
        t0 = time.perf_counter()
        if released[slot] is not None: side.wait_event(released[slot])
        tick("wait_event", t0)
        for k, v in hb.items():
            b = slots[slot].get(k)
            if b is None:
                b = slots[slot][k] = torch.empty(v.shape, dtype=v.dtype, device=dev)
            t0 = time.perf_counter(); b.copy_(v, non_blocking=True); tick("copy_" + k, t0)
        t0 = time.perf_counter(); e2 = torch.cuda.Event(); e2.record(side); tick("record", t0)
    t0 = time.perf_counter(); cur.wait_event(e2); tick("cur.wait", t0)
    for _ in range(3): torch.mm(big, big)
    t0 = time.perf_counter(); torch.cuda.synchronize(); tick("sync", t0)
print({k: round(1e3 * v / 10, 3) for k, v in T.items()})
