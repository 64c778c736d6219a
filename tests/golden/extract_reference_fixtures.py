#!/usr/bin/env python3
"""Extract the known-answer vectors the reference's own tests hold for the
xlating / lpf path into tests/golden/reference_fixtures.json.

Run in the dev container only (it reads /root/reference, which does not exist on
the GPU box); the JSON it writes is committed and is what the tests read.

Sources (numbers only -- no reference code is copied):
  test/test_xlating.c  :24-37  full 2000-byte block -> 200 cf32 / 200 cs16
                       :39-61  2 x 200-byte calls   -> 20+20 cf32 / cs16 (state carry-over)
                       :63-81  198 B then 2 B       -> 0 outputs
  test/test_lpf.c      :25-39  39 taps for (fs 8000, cutoff 1750, tw 500)
  test/test_tcp_server.c :154-248  cu8 / cs16 / cs8 ramps of 200 elements through
                       the 61-tap server-default filter (48 kHz -> 9.6 kHz)
Tolerance semantics of the reference's asserts: test/utils.c:176-196
(cf32: (int32)(x*10000) equality; cs16: exact).
"""
import json
import os
import re
import sys

REF = os.environ.get("REF_DIR", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_fixtures.json")

ARRAY_RE = re.compile(r"const\s+(float|int16_t)\s+(\w+)\[\]\s*=\s*\{([^}]*)\}\s*;", re.S)
FUNC_RE = re.compile(r"^void\s+(test_\w+)\s*\(\s*\)\s*\{", re.M)


def arrays_by_function(path):
    src = open(path).read()
    funcs = [(m.start(), m.group(1)) for m in FUNC_RE.finditer(src)]
    out = {}
    for m in ARRAY_RE.finditer(src):
        owner = None
        for pos, name in funcs:
            if pos < m.start():
                owner = name
        vals = [v.strip() for v in m.group(3).replace("\n", " ").split(",") if v.strip()]
        if m.group(1) == "float":
            vals = [float(v.rstrip("fF")) for v in vals]
        else:
            vals = [int(v) for v in vals]
        out.setdefault(owner, {})[m.group(2)] = vals
    return out


def main():
    x = arrays_by_function(os.path.join(REF, "test/test_xlating.c"))
    l = arrays_by_function(os.path.join(REF, "test/test_lpf.c"))
    t = arrays_by_function(os.path.join(REF, "test/test_tcp_server.c"))
    fixtures = {
        "_source": "dernasherbrezon/sdr-server test/test_xlating.c, test/test_lpf.c, test/test_tcp_server.c",
        "_assert_semantics": "cf32/taps: (int32)(x*10000) equality (test/utils.c:176-182,191-196); cs16: exact (:184-189)",
        "xlating": {
            "setup": {"sampling_freq": 48000, "target_freq": 9600, "decimation": 5,
                      "lpf": {"gain": 1.0, "cutoff": 4800, "transition_width": 2000},
                      "center_freq": -12000, "ntaps": 57,
                      "input": "cu8 ramp: (uint8)(offset+i)  (test/utils.c:137-145)"},
            "max_input_buffer_size": {"input_len": 2000, "max_input": 2000,
                                      "cf32": x["test_max_input_buffer_size"]["expected_cf32"],
                                      "cs16": x["test_max_input_buffer_size"]["expected_cs16"]},
            "partial_input_buffer_size": {"input_len": 200, "max_input": 2000,
                                          "cf32": x["test_partial_input_buffer_size"]["expected_cf32"],
                                          "cs16": x["test_partial_input_buffer_size"]["expected_cs16"],
                                          "next_cf32": x["test_partial_input_buffer_size"]["expected_next_cf32"],
                                          "next_cs16": x["test_partial_input_buffer_size"]["expected_next_cs16"]},
            "small_input_data": {"first_len": 198, "second_len": 2, "max_input": 2000, "expected_outputs": 0},
        },
        "lpf": {"args": {"gain": 1.0, "sampling_freq": 8000, "cutoff": 1750, "transition_width": 500},
                "ntaps": 39, "taps": l["test_lowpassTaps"]["expected_taps"],
                "bad_args": [[0, 1750, 500], [8000, 5000, 500], [8000, 1750, 0]]},
        "tcp_server": {
            "setup": {"band_sampling_rate": 48000, "client_rate": 9600, "decimation": 5,
                      "lpf": {"gain": 1.0, "cutoff": 4800, "transition_width": 1920}, "ntaps": 61,
                      "center_offset": -12000, "input_elements": 200, "buffer_size": 131072,
                      "inputs": {"cu8": "(uint8)(i)", "cs8": "(int8)(i)", "cs16": "(int16)(i) - 100"}},
            "rtlsdr_cu8": t["test_rtlsdr"]["expected"],
            "airspy_cs16": t["test_airspy"]["expected"],
            "hackrf_cs8": t["test_hackrf"]["expected"],
        },
    }
    with open(OUT, "w") as f:
        json.dump(fixtures, f, indent=1)
    n = sum(len(v) for v in (fixtures["xlating"]["max_input_buffer_size"]["cf32"],
                             fixtures["lpf"]["taps"], fixtures["tcp_server"]["rtlsdr_cu8"]))
    print(f"wrote {OUT} ({n} numbers in the three headline vectors)")


if __name__ == "__main__":
    sys.exit(main())
