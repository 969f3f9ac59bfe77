"""H2D copy rate of a pinned block of the e2e transport size, alone and split over two streams, next to the e2e step
(bench.py e2e leg): tells whether the e2e step is the PCIe copy or something around it.  Run under gpurun."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 25_780_256
dev = torch.device("cuda:0")
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
out = {}
def rate(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = rate(lambda: d.copy_(h, non_blocking=True))
out["h2d_one_stream"] = {"ms": ms, "gbs": nbytes / ms / 1e6}
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
half = nbytes // 2
def two():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): d[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2): d[half:].copy_(h[half:], non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
ms = rate(two)
out["h2d_two_streams"] = {"ms": ms, "gbs": nbytes / ms / 1e6}
big = torch.empty(1 << 30, dtype=torch.uint8).pin_memory(); dbig = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
ms = rate(lambda: dbig.copy_(big, non_blocking=True), 5)
out["h2d_1GiB"] = {"ms": ms, "gbs": (1 << 30) / ms / 1e6}
ms = rate(lambda: big.copy_(dbig, non_blocking=True), 5)
out["d2h_1GiB"] = {"ms": ms, "gbs": (1 << 30) / ms / 1e6}
print(json.dumps(out))
