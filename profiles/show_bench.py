import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
        r = d["roofline"]
        print(f, "value=%.1f ms/step=%.4f e2e=%s fp32frac=%.3f steady=%.3f kern=%s host=%s lat_us=%s" % (
            d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), r["fp32"]["frac"],
            r["fp32"].get("steady_state", {}).get("frac", 0),
            {k: round(v, 4) for k, v in r["step_kernels_ms"].items()},
            {k: round(v, 1) for k, v in d.get("host", {}).items()}, (d.get("block_latency_us") or {}).get("median")))
    except Exception as e:
        print(f, "ERR", e)
