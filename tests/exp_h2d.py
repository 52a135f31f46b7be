import time, torch
dev = torch.device("cuda", 0)
host = {"a": torch.randn(12, 6, 256, 256).pin_memory(), "b": torch.randn(12, 6, 256, 256).pin_memory(),
        "c": torch.randn(12, 1, 256, 256).pin_memory()}
print("pinned", [v.is_pinned() for v in host.values()])
side = torch.cuda.Stream(dev)
bufs = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
big = torch.randn(8192, 8192, device=dev)
def busy():
    for _ in range(4):
        torch.mm(big, big)
def t(label, fn, n=5):
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        busy()
        t0 = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t0))
        torch.cuda.synchronize()
    print("%-40s host ms: %s" % (label, " ".join("%.2f" % x for x in ts)), flush=True)
t("to() current stream", lambda: [v.to(dev, non_blocking=True) for v in host.values()])
t("copy_ current stream", lambda: [bufs[k].copy_(v, non_blocking=True) for k, v in host.items()])
def side_copy():
    with torch.cuda.stream(side):
        for k, v in host.items():
            bufs[k].copy_(v, non_blocking=True)
t("copy_ side stream", side_copy)
def side_to():
    with torch.cuda.stream(side):
        return [v.to(dev, non_blocking=True) for v in host.values()]
t("to() side stream", side_to)
def side_copy_ev():
    with torch.cuda.stream(side):
        side.wait_event(ev0)
        for k, v in host.items():
            bufs[k].copy_(v, non_blocking=True)
        e = torch.cuda.Event(); e.record(side)
    return e
ev0 = torch.cuda.Event(); ev0.record(torch.cuda.current_stream())
t("copy_ side stream + events", side_copy_ev)
# idle GPU (no busy work queued)
torch.cuda.synchronize()
for lab, fn in (("idle: copy_ side", side_copy), ("idle: to() current", lambda: [v.to(dev, non_blocking=True) for v in host.values()])):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); ts.append(1e3 * (time.perf_counter() - t0))
    print("%-40s host ms: %s" % (lab, " ".join("%.2f" % x for x in ts)))
