"""steady-state CTA timeline of the tiled kernel: submit blocks back to back, destroy -> trace file.
usage (from the repo root): XLATING_B200_TRACE=1 XLATING_B200_TRACE_FILE=out.bin [XLATING_B200_TIMELINE=tl.txt] python tools/trace_run.py cfg2 [taps_mode]"""
import os, sys, importlib, time
import numpy as np
sys.path.insert(0, os.getcwd())
import bench
import torch
pkg = importlib.import_module("sdr-server_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = bench.workload(name, sys.argv[2] if len(sys.argv) > 2 else "server")
g = pkg.Group(wl["fs"], wl["block_elems"], flags=pkg.XLG_OUT_DEVICE | pkg.XLG_SM_PARTITION)
tapsets = {}
for p in wl["plan"]:
    key = (p["cutoff"], p["tw"])
    if key not in tapsets:
        tapsets[key] = pkg.create_low_pass_filter(1.0, wl["fs"], p["cutoff"], p["tw"])
    g.add_client(p["decimation"], tapsets[key], p["center"])
NB = 64
host_blocks = bench.synth_blocks(wl["fmt"], NB, wl["block_elems"], seed=3)
dev = torch.from_numpy(host_blocks.view(np.uint8).reshape(NB, -1)).cuda()
fmt = pkg.FMT[wl["fmt"]]
last = -1
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(256):
        last = g.submit_ptr(fmt, dev.data_ptr() + (i % NB) * dev.stride(0), wl["block_elems"], pkg.XLG_INPUT_DEVICE)
    g.wait(last)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {256 * wl['block_elems'] / 2 / dt / 1e6:.1f} MS/s wall, {dt / 256 * 1e6:.2f} us/block", flush=True)
g.close()
