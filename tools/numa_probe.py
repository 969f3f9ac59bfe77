"""8-GPU host->device copy bandwidth and where pinned memory lives: topology as the container sees it, then the aggregate
rate of concurrent 14 MB H2D copies (one per GPU) with pinned buffers allocated under the default memory policy and under
MPOL_BIND / MPOL_PREFERRED to each GPU's own NUMA node.  Run under gpurun --gpus 8."""
import ctypes, json, os, subprocess, sys, time
import torch
out = {}
def sh(c):
    try: return subprocess.run(c, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e: return "ERR %s" % e
out["topo"] = sh("nvidia-smi topo -m | head -14")
out["mems_allowed"] = sh("grep -i 'mems_allowed_list\\|cpus_allowed_list' /proc/self/status")
out["nodes"] = sh("ls /sys/devices/system/node/ | grep node; cat /sys/devices/system/node/online")
n = torch.cuda.device_count()
gpu_node = []
for i in range(n):
    bdf = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    gpu_node.append(None)
q = sh("nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader")
out["gpus"] = q
for line in q.splitlines():
    idx, bdf = [x.strip() for x in line.split(",")]
    bdf = bdf.lower()
    if bdf.startswith("0000"): bdf = bdf[4:]
    node = sh("cat /sys/bus/pci/devices/%s/numa_node" % bdf)
    try: gpu_node[int(idx)] = int(node)
    except Exception: pass
out["gpu_numa_node"] = gpu_node
libc = ctypes.CDLL(None, use_errno=True)
def set_mempolicy(mode, node):
    if node is None or node < 0:
        return libc.syscall(238, 0, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(238, mode, ctypes.byref(mask), 64)
NB = 14_000_000
def run(policy):
    bufs, devs, streams = [], [], []
    for i in range(n):
        rc = 0
        if policy == "bind": rc = set_mempolicy(2, gpu_node[i])
        elif policy == "preferred": rc = set_mempolicy(1, gpu_node[i])
        err = ctypes.get_errno() if rc != 0 else 0
        h = torch.empty(NB, dtype=torch.uint8).pin_memory(); h.fill_(1)
        set_mempolicy(0, None)
        bufs.append(h); devs.append(torch.empty(NB, dtype=torch.uint8, device="cuda:%d" % i)); streams.append(torch.cuda.Stream(device=i))
        if rc != 0: return {"error": "set_mempolicy rc=%d errno=%d" % (rc, err)}
    res = {}
    for active in ([0], list(range(n))):
        for _ in range(3):
            for i in active:
                with torch.cuda.stream(streams[i]): devs[i].copy_(bufs[i], non_blocking=True)
        for i in active: streams[i].synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            for i in active:
                with torch.cuda.stream(streams[i]): devs[i].copy_(bufs[i], non_blocking=True)
        for i in active: streams[i].synchronize()
        dt = (time.perf_counter() - t0) / 50
        res["%d_gpus" % len(active)] = {"ms_per_round": dt * 1e3, "aggregate_gbs": NB * len(active) / dt / 1e9}
    # per-GPU alone
    per = []
    for i in range(n):
        streams[i].synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            with torch.cuda.stream(streams[i]): devs[i].copy_(bufs[i], non_blocking=True)
        streams[i].synchronize(); per.append(round(NB * 20 / (time.perf_counter() - t0) / 1e9, 1))
    res["alone_gbs"] = per
    return res
for pol in ("default", "bind", "preferred"):
    out[pol] = run(pol)
print(json.dumps(out, indent=1))
