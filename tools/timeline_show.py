import sys, numpy as np
a = np.loadtxt(sys.argv[1], comments="#")
a = a[np.argsort(a[:, 0])]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 600
print("ticket | conv_ready conv_done | phase_start phase_done | fir_ready fir_done | fir_ready-prev_fir_ready, fir_dur  (us, relative to first row)")
base = None; prev = None
for r in a:
    if r[0] < lo or r[0] >= lo + 24: continue
    t = r[1:] * 1e3
    if base is None: base = t[0]
    t = t - base
    d = (t[4] - prev) if prev is not None else 0.0
    prev = t[4]
    print(f"{int(r[0]):5d} | {t[0]:8.1f} {t[1]:8.1f} | {t[2]:8.1f} {t[3]:8.1f} | {t[4]:8.1f} {t[5]:8.1f} | {d:6.1f} {t[5]-t[4]:6.1f}")
