import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        print(f, "value=%.1f ms/step=%.4f e2e=%s fp32frac=%.3f kern=%s"%(d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d["roofline"]["fp32"]["frac"], {k:round(v,4) for k,v in d["roofline"]["step_kernels_ms"].items()}))
    except Exception as e:
        print(f, "ERR", e)
