"""Pure-write / pure-read / copy bandwidth of this GPU (torch kernels, CUDA events, best of 10):
context for kernels whose traffic is not half reads, half writes like the copy that defines
MEASURED_PEAKS.json's hbm_gbs."""
import json
import torch

dev = torch.device("cuda:0")
out = {}
for mb in (200, 2000):
    n = mb * 1000 * 1000 // 4
    a = torch.empty(n, dtype=torch.int32, device=dev)
    b = torch.empty(n, dtype=torch.int32, device=dev)
    a.zero_(); b.zero_()
    def best(fn, bytes_):
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return bytes_ / min(ts) / 1e6
    out["%dMB" % mb] = {
        "write_only_gbs(zero_)": best(lambda: a.zero_(), n * 4),
        "read_only_gbs(sum)": best(lambda: a.sum(), n * 4),
        "copy_gbs(read+write)": best(lambda: b.copy_(a), 2 * n * 4),
    }
print(json.dumps(out))
