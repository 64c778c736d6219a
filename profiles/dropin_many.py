"""Throughput of the UNMODIFIED reference model on the drop-in ABI: one filter + one
dsp thread per client, every thread processing its own copy of the same block
sequence (src/dsp_worker.c:41-88).  No batch binding: every call stages its own block."""
import ctypes as C
import importlib
import sys
import threading
import time

import numpy as np

sys.path.insert(0, ".")
pkg = importlib.import_module("sdr-server_b200")
L = pkg.lib()
fs, n_clients, n_blocks = 2016000, int(sys.argv[1]) if len(sys.argv) > 1 else 64, 40
plan = pkg.client_plan(fs, [48000 if c % 2 == 0 else 96000 for c in range(n_clients)])
filters = []
for p in plan:
    taps = pkg.create_low_pass_filter(1.0, fs, p["cutoff"], p["tw"])
    filters.append(pkg.XlatingFilter(p["decimation"], taps, p["center"], fs, 262144))
blocks = [np.random.default_rng(i).integers(0, 256, 262144, dtype=np.uint8) for i in range(4)]


def dsp_thread(f):
    out, n = C.c_void_p(), C.c_size_t(0)
    for b in range(n_blocks):
        x = blocks[b % 4]
        L.process_native_cu8_cf32(x.ctypes.data, x.size, C.byref(out), C.byref(n), f._h)


for f in filters[:4]:
    dsp_thread(f)  # warm-up
threads = [threading.Thread(target=dsp_thread, args=(f,)) for f in filters]
t0 = time.perf_counter()
for t in threads:
    t.start()
for t in threads:
    t.join()
dt = time.perf_counter() - t0
print(f"dropin thread-per-client: {n_clients} filters x {n_blocks} blocks in {dt:.3f} s -> "
      f"{n_blocks * 131072 / dt / 1e6:.1f} MS/s in (all clients), {n_clients * n_blocks / dt:.0f} calls/s")
