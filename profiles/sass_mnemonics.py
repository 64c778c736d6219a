"""per-kernel SASS mnemonic counts of the built library (run from the repo root after `make -C sdr-server_b200`)"""
import collections, re, subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else "sdr-server_b200/lib/libxlating_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
KEYS = ["FFMA2", "FFMA", "FMUL", "FADD", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "BAR", "IMAD",
        "MUFU", "HMMA", "UTCMMA", "SHFL"]
print("# SASS evidence (cuobjdump -sass %s, sm_100a), per kernel: instruction mnemonic counts that matter for the" % obj)
print("# claims in DESIGN.md section 5: UBLKCP = cp.async.bulk (1-D TMA bulk copy), UTMALDG = cp.async.bulk.tensor (TMA")
print("# tensor copy), SYNCS = mbarrier ops, LDGSTS = cp.async, FFMA / FFMA2 = fp32 FMA (scalar / packed pair), LDS = shared")
print("# loads; no HMMA / UTC*MMA anywhere (the path is fp32 CUDA-core math by contract).\n")
cur, cnt, tot = None, None, 0
def flush():
    if cur is not None:
        print(cur)
        print("    " + "  ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]) + f"  (total {tot})")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        cur, cnt, tot = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1).split(".")[0]
        tot += 1
        for k in KEYS:
            if op == k:
                cnt[k] += 1
flush()
