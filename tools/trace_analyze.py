import sys, numpy as np
raw = np.fromfile(sys.argv[1], dtype=np.int64)
REC, CTAS, NL, launches = raw[:4]
n_of = raw[4:4 + NL]
t = raw[4 + NL:].reshape(NL, CTAS, REC)
print("launches so far", launches, "ctas per slot", n_of[:NL].tolist())
order = [(launches - NL + i) % NL for i in range(NL)]  # oldest .. newest
M48 = (1 << 48) - 1
recs = []
for li, slot in enumerate(order):
    n = n_of[slot]
    r = t[slot, :n]
    c0 = r[:, 0]; c1 = r[:, 1]
    c2 = (r[:, 2] & M48) | (c1 & ~M48)
    c3 = r[:, 3]
    smid = (r[:, 2] >> 48) & 0xffff
    gt = r[:, 4]
    ci = r[:, 5] >> 32; wid = (r[:, 5] >> 16) & 0xffff; L = r[:, 5] & 0xffff
    for i in range(n):
        recs.append((li, int(smid[i]), int(c0[i]), int(c1[i]), int(c2[i]), int(c3[i]), int(gt[i]), int(ci[i]), int(wid[i]), int(L[i])))
recs = np.array(recs, dtype=np.int64)
li, smid, c0, c1, c2, c3, gt, ci, wid, L = recs.T
print("SMs used", len(np.unique(smid)), "smid range", smid.min(), smid.max())
dur = c3 - c0
for k in np.unique(L):
    m = (L == k) & (li >= 3) & (li <= NL - 4)
    print(f"L={k}: n={m.sum()} dur mean {dur[m].mean():.0f} cyc (p10 {np.percentile(dur[m],10):.0f} p90 {np.percentile(dur[m],90):.0f}) stage {np.mean(c1[m]-c0[m]):.0f} loop {np.mean(c2[m]-c1[m]):.0f} epi {np.mean(c3[m]-c2[m]):.0f}  in-loop FFMA issue/clk/warp {np.mean(k*128.0/(c2[m]-c1[m])):.3f}")
# per-launch span in ns (globaltimer at CTA end)
for l in range(NL):
    m = li == l
    print(f"launch {l}: end-gt span {(gt[m].max()-gt[m].min())/1e3:.1f} us, first end at {(gt[m].min()-gt.min())/1e3:.1f} us")
# residency per SM within the steady window
res = []; ffma_rate = []; smsp_hist = np.zeros(4)
for s in np.unique(smid):
    m = smid == s
    w0 = np.median(c0[m & (li == 3)]) if (m & (li == 3)).any() else None
    w1 = np.median(c3[m & (li == NL - 4)]) if (m & (li == NL - 4)).any() else None
    if w0 is None or w1 is None or w1 <= w0: continue
    a = np.clip(c0[m], w0, w1); b = np.clip(c3[m], w0, w1)
    res.append((b - a).sum() / (w1 - w0))
    # FFMA warp-instr issued inside window (assume uniform over loop span)
    la = np.clip(c1[m], w0, w1); lb = np.clip(c2[m], w0, w1)
    frac = (lb - la) / np.maximum(c2[m] - c1[m], 1)
    ffma = (frac * L[m] * 128 * 2).sum()          # 2 warps per CTA
    ffma_rate.append(ffma / (w1 - w0) / 4)         # per SMSP per clk
    for x in wid[m]: smsp_hist[x % 4] += 1
res = np.array(res); ffma_rate = np.array(ffma_rate)
print(f"steady window: resident CTAs per SM mean {res.mean():.2f} (min {res.min():.2f} max {res.max():.2f}); FFMA issue per SMSP per clk mean {ffma_rate.mean():.3f} (min {ffma_rate.min():.3f} max {ffma_rate.max():.3f})")
print("warp0 hardware warp slot % 4 histogram", smsp_hist.tolist())
st = gt - dur / 1.965
for l in range(NL):
    m = li == l
    print(f"launch {l}: CTA starts from {(st[m].min()-gt.min())/1e3:.1f} to {(st[m].max()-gt.min())/1e3:.1f} us; ends {(gt[m].min()-gt.min())/1e3:.1f} .. {(gt[m].max()-gt.min())/1e3:.1f} us")
# overlap between consecutive launches: fraction of CTAs of launch l starting before the last CTA of l-1 ended (same SM clock not comparable across SMs -> use gt)
for l in range(1, NL):
    a = gt[li == l - 1].max(); 
    b = gt[li == l]
    print(f"launch {l}: {np.mean(b < a)*100:.0f}% of its CTAs END before launch {l-1}'s last CTA ends")
