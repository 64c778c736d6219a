"""The reference's own programs, compiled UNMODIFIED where they lie under /root/reference
and linked against libxlating_b200.so instead of src/xlating.c + src/lpf.c (recipe:
oracle/Makefile -> oracle/_ref/perf_xlating_b200).

test/perf_xlating.c creates one 2429-tap filter and calls process_native_cu8_cf32,
process_optimized_cu8_cf32, process_native_cu8_cs16 and process_optimized_cu8_cs16 a
thousand times each on 200 000-byte blocks.

test/test_xlating.c and test/test_lpf.c are the reference's Unity tests of the path; they
are built with the reference's Unity and oracle/ref_test_support.c (a stand-in for the few
helpers of test/utils.c, which itself needs libpng/zlib) -- once against the reference's
own sources (*_ref: proves the stand-in) and once against this library (*_b200)."""
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PERF_B200 = os.path.join(ROOT, "oracle", "_ref", "perf_xlating_b200")

needs_binary = pytest.mark.skipif(not os.path.exists(PERF_B200), reason="oracle/_ref/perf_xlating_b200 not built")


def ref_program(name):
    path = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(path):
        pytest.skip(f"oracle/_ref/{name} not built (no /root/reference when oracle/ was made)")
    return path


def unity_summary(stdout):
    m = re.search(r"^(\d+) Tests (\d+) Failures (\d+) Ignored", stdout, flags=re.M)
    assert m, stdout[-400:]
    return tuple(int(g) for g in m.groups())


@pytest.mark.parametrize("name,tests", [("test_xlating_ref", 3), ("test_lpf_ref", 4)])
def test_reference_unit_tests_pass_on_the_reference_with_our_stand_in(name, tests):
    r = subprocess.run([ref_program(name)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and unity_summary(r.stdout) == (tests, 0, 0), r.stdout[-400:]


def test_reference_lpf_unit_test_passes_on_this_library():
    """test/test_lpf.c:7-52 against libxlating_b200.so: the tap designer stays on the host,
    so the reference's own test of it runs -- and passes -- without a GPU."""
    r = subprocess.run([ref_program("test_lpf_b200")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and unity_summary(r.stdout) == (4, 0, 0), r.stdout[-400:]
    assert r.stderr.count("<3>") == 3  # the three bounds tests log like the reference (src/lpf.c:12-29)


@pytest.mark.skipif(torch.cuda.is_available(), reason="this is the no-GPU behaviour")
def test_reference_xlating_unit_test_fails_loudly_without_a_gpu():
    r = subprocess.run([ref_program("test_xlating_b200")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "Expected 0 Was -19" in r.stdout  # create_frequency_xlating_filter -> -ENODEV
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="added after the round's GPU budget was spent: first run on a B200 is the "
                                        "driver's; the same vectors with the same (int32)(x*10000) comparison pass "
                                        "in tests/test_gpu_parity.py::test_fixture_*")
def test_reference_xlating_unit_test_passes_on_the_gpu():
    """test/test_xlating.c:24-81, unmodified: full block, partial blocks with carried
    history and phase, too-short input; float path at 4 decimals, Q15 path exactly."""
    r = subprocess.run([ref_program("test_xlating_b200")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and unity_summary(r.stdout) == (3, 0, 0), r.stdout[-600:]


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


@pytest.mark.skipif(not os.path.exists("/root/reference/src/xlating.h"), reason="needs /root/reference")
def test_our_headers_declare_the_reference_prototypes(tmp_path):
    """One translation unit including the reference's headers AND ours: C rejects a second
    declaration of a function with a different type, so this compiles only if every
    prototype in include/xlating.h and include/lpf.h equals the reference's
    (src/xlating.h:8-38, src/lpf.h:6)."""
    src = tmp_path / "both.c"
    src.write_text('#include <stdlib.h>\n#include "/root/reference/src/lpf.h"\n#include "/root/reference/src/xlating.h"\n'
                   f'#include "{ROOT}/include/lpf.h"\n#include "{ROOT}/include/xlating.h"\n'
                   "int main(void) { return 0; }\n")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-c", str(src), "-o", str(tmp_path / "both.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
