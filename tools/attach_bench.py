"""(from the repo root: python tools/attach_bench.py [clients]; XLATING_B200_REBUILD_TIMING=1 logs every re-layout)
attach churn: N clients static vs one attach (+ one detach) per block; step time ratio"""
import os, sys, importlib, time
import numpy as np
sys.path.insert(0, os.getcwd())
import bench, torch
pkg = importlib.import_module("sdr-server_b200")
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
wl = bench.workload("c1000", "297")
plan = wl["plan"][:C]
NB = 32
host_blocks = bench.synth_blocks(wl["fmt"], NB, wl["block_elems"], seed=3)
dev = torch.from_numpy(host_blocks.view(np.uint8).reshape(NB, -1)).cuda()
fmt = pkg.FMT[wl["fmt"]]
tapsets = {}
def taps_of(p):
    key = (p["cutoff"], p["tw"])
    if key not in tapsets:
        tapsets[key] = pkg.create_low_pass_filter(1.0, wl["fs"], p["cutoff"], p["tw"])
    return tapsets[key]
g = pkg.Group(wl["fs"], wl["block_elems"], flags=pkg.XLG_OUT_DEVICE | pkg.XLG_SM_PARTITION)
ids = [g.add_client(p["decimation"], taps_of(p), p["center"]) for p in plan]
g.reserve(int(sum(wl["block_elems"] // 2 // p["decimation"] + 2 for p in plan) * 1.2))
def run(k, churn):
    last = -1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(k):
        if churn:
            j = i % len(ids)
            g.remove_client(ids[j])
            ids[j] = g.add_client(plan[j]["decimation"], taps_of(plan[j]), plan[j]["center"])
        last = g.submit_ptr(fmt, dev.data_ptr() + (i % NB) * dev.stride(0), wl["block_elems"], pkg.XLG_INPUT_DEVICE)
    g.wait(last); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e6
run(32, False)
s0 = run(128, False)
c0 = run(100, True)
s1 = run(128, False)
kinds = {}
for c in ids:
    k = g.client_info(c)[1]; kinds[k] = kinds.get(k, 0) + 1
print(f"clients {C}: static {s0:.1f} us/block, one detach+attach per block {c0:.1f} us/block (x{c0/s0:.1f}), static again {s1:.1f}; kinds {kinds}", flush=True)
g.close()
