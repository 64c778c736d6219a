import importlib, sys, time, numpy as np
sys.path.insert(0,'.')
pkg=importlib.import_module("sdr-server_b200")
import ctypes as C
fs=2016000
for tw,name in [(9600,'505'),(2000,'2429')]:
    taps=pkg.create_low_pass_filter(1.0,fs,24000,tw)
    f=pkg.XlatingFilter(42,taps,-312000,fs,262144)
    x=np.random.default_rng(0).integers(0,256,262144,dtype=np.uint8)
    L=pkg.lib(); out=C.c_void_p(); n=C.c_size_t(0)
    for _ in range(20): L.process_native_cu8_cf32(x.ctypes.data,x.size,C.byref(out),C.byref(n),f._h)
    t0=time.perf_counter(); K=300
    for _ in range(K): L.process_native_cu8_cf32(x.ctypes.data,x.size,C.byref(out),C.byref(n),f._h)
    dt=(time.perf_counter()-t0)/K
    print(f"dropin process_native_cu8_cf32 taps={name}: {dt*1e6:.1f} us/call -> {131072/dt/1e6:.1f} MS/s (n_out={n.value})")
    f.close()
