"""The reference's own programs, compiled UNMODIFIED where they lie under /root/reference
and linked against libxlating_b200.so instead of src/xlating.c + src/lpf.c (recipe:
oracle/Makefile -> oracle/_ref/perf_xlating_b200).

test/perf_xlating.c creates one 2429-tap filter and calls process_native_cu8_cf32,
process_optimized_cu8_cf32, process_native_cu8_cs16 and process_optimized_cu8_cs16 a
thousand times each on 200 000-byte blocks."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PERF_B200 = os.path.join(ROOT, "oracle", "_ref", "perf_xlating_b200")

needs_binary = pytest.mark.skipif(not os.path.exists(PERF_B200), reason="oracle/_ref/perf_xlating_b200 not built")


@needs_binary
@pytest.mark.skipif(torch.cuda.is_available(), reason="this is the no-GPU behaviour")
def test_reference_perf_program_links_and_fails_loudly_without_a_gpu():
    """It links (every symbol it needs is exported with the reference's signature) and,
    with no device, create_frequency_xlating_filter fails -- the program exits with
    EXIT_FAILURE instead of silently running some CPU path."""
    r = subprocess.run([PERF_B200], capture_output=True, text=True, timeout=120)
    assert "SIMD optimization: CUDA sm_100a" in r.stdout
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and r.stderr.startswith("<3>")
    assert "cu8_cf32" not in r.stdout  # never got as far as processing anything


@needs_binary
@pytest.mark.gpu
def test_reference_perf_program_runs_on_the_gpu():
    r = subprocess.run([PERF_B200], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-400:]
    assert "SIMD optimization: CUDA sm_100a" in r.stdout
    found = re.findall(r"^(native|optimized)\s+(cu8_cf32|cu8_cs16): ([0-9.]+) seconds$", r.stdout, flags=re.M)
    times = {(variant, kind): float(t) for variant, kind, t in found}
    assert set(times) == {("native", "cu8_cf32"), ("optimized", "cu8_cf32"), ("native", "cu8_cs16"),
                          ("optimized", "cu8_cs16")}, r.stdout
    # clock() sums CPU time over threads; the reference's own AVX build needs 1.6 ms per call here
    assert all(0.0 < t < 0.005 for t in times.values()), times
    print(r.stdout)
