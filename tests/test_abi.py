"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/*.h declares, its host-side code (tap designer, per-client
constants) is bit-identical to the oracle, and -- with no GPU in this container --
the compute entry points fail loudly instead of falling back to a CPU path."""
import ctypes as C
import errno
import os
import re

import numpy as np
import pytest
import torch

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\(", src)
    return sorted({n for n in names if n.startswith(("xlg_", "process_", "create_", "destroy_"))})


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    declared = set()
    for h in ("xlating.h", "xlating_group.h", "lpf.h"):
        fns = declared_functions(h)
        assert fns, h
        declared.update(fns)
    assert len(declared) >= 14 + 16 + 1
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    for name in pkg.REFERENCE_SYMBOLS + pkg.GROUP_SYMBOLS:
        assert name in declared or name == "SIMD_STATUS"
    assert pkg.simd_status() == "CUDA sm_100a"


def test_library_does_not_link_the_oracle(pkg):
    """The product must not route through oracle/ (or any CPU implementation)."""
    import subprocess
    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out
    syms = subprocess.run(["nm", "-D", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in syms


@pytest.mark.parametrize("fs,cutoff,tw", [(8000, 1750, 500), (48000, 4800, 2000), (48000, 4800, 1920),
                                          (2016000, 24000, 16400), (2016000, 24000, 9600), (2016000, 48000, 19200),
                                          (2016000, 24000, 2000), (10000000, 125000, 20060),
                                          (10000000, 125000, 50000), (61440000, 24000, 9600)])
def test_host_tap_designer_bit_exact(pkg, fs, cutoff, tw):
    a = pkg.create_low_pass_filter(1.0, fs, cutoff, tw)
    b = po.lpf_design(1.0, fs, cutoff, tw)
    assert a.tobytes() == b.tobytes()


def test_host_tap_designer_known_lengths(pkg):
    # the tap counts the BASELINE configs quote
    assert len(pkg.create_low_pass_filter(1.0, 2016000, 24000, 16400)) == 297
    assert len(pkg.create_low_pass_filter(1.0, 2016000, 24000, 9600)) == 505
    assert len(pkg.create_low_pass_filter(1.0, 2016000, 48000, 19200)) == 253
    assert len(pkg.create_low_pass_filter(1.0, 2016000, 24000, 2000)) == 2429
    assert len(pkg.create_low_pass_filter(1.0, 10000000, 125000, 20060)) == 1201


def test_host_tap_designer_errors(pkg, fixtures):
    for fs, cutoff, tw in fixtures["lpf"]["bad_args"]:
        with pytest.raises(ValueError) as e:
            pkg.create_low_pass_filter(1.0, fs, cutoff, tw)
        assert e.value.args[0] == -1  # test/test_lpf.c:7-23


class Consts(C.Structure):
    _fields_ = [("rev_cf32", C.POINTER(C.c_float)), ("rev_q15", C.POINTER(C.c_int16)),
                ("incr_re", C.c_float), ("incr_im", C.c_float), ("qincr_re", C.c_int16), ("qincr_im", C.c_int16)]


@pytest.mark.parametrize("fs,rate,tw,center", [(48000, 9600, 2000, -12000), (2016000, 48000, 16400, -312000),
                                               (2016000, 96000, 19200, 960000), (10000000, 250000, 50000, -4875000),
                                               (2016000, 48000, 9600, 0)])
def test_host_client_constants_bit_exact(pkg, fs, rate, tw, center):
    """csrc/taps_host.c == oracle create (src/xlating.c:519-549): reversed band-pass
    taps, Q15 taps and the oscillator step."""
    lib = pkg.lib()
    taps = po.lpf_design(1.0, fs, rate // 2, tw)
    D = fs // rate
    k = Consts()
    lib.xl_client_consts_build.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_uint32, C.c_int32, C.c_uint32,
                                           C.POINTER(Consts)]
    lib.xl_client_consts_build.restype = C.c_int
    assert lib.xl_client_consts_build(taps.ctypes.data_as(C.POINTER(C.c_float)), len(taps), D, center, fs,
                                      C.byref(k)) == 0
    rev = np.ctypeslib.as_array(k.rev_cf32, shape=(2 * len(taps),)).copy().view(np.complex64)
    o = po.OracleFilter(D, taps, center, fs, 1000)
    assert rev.tobytes() == o.rev_taps.tobytes()
    # oscillator step: history is T-1 zeros, so ONE sample yields one output and,
    # without renormalisation, phase = 1*incr = incr exactly
    assert len(o.process_cf32("cu8", np.zeros(2, dtype=np.uint8), renorm=False)) == 1
    assert o.phase == complex(k.incr_re, k.incr_im)
    q = po.RefFilter(D, taps, center, fs, 1000) if po.ref_available() else None
    qtaps = np.ctypeslib.as_array(k.rev_q15, shape=(2 * len(taps),)).copy()
    np.testing.assert_array_equal(qtaps, (rev.view(np.float32) * np.float32(32768)).astype(np.int16))
    assert k.qincr_re == np.int16(np.float32(k.incr_re) * np.float32(32767))
    assert k.qincr_im == np.int16(np.float32(k.incr_im) * np.float32(32767))
    del q
    lib.xl_client_consts_free.argtypes = [C.POINTER(Consts)]
    lib.xl_client_consts_free(C.byref(k))


def test_even_length_taps_keep_reference_quirk(pkg):
    """src/xlating.c:530-534 leaves the middle pair of an even-length filter
    un-reversed; the host constants reproduce that."""
    lib = pkg.lib()
    taps = np.arange(1, 9, dtype=np.float32)
    k = Consts()
    lib.xl_client_consts_build.argtypes = [C.POINTER(C.c_float), C.c_size_t, C.c_uint32, C.c_int32, C.c_uint32,
                                           C.POINTER(Consts)]
    lib.xl_client_consts_build.restype = C.c_int
    assert lib.xl_client_consts_build(taps.ctypes.data_as(C.POINTER(C.c_float)), 8, 2, 0, 48000, C.byref(k)) == 0
    rev = np.ctypeslib.as_array(k.rev_cf32, shape=(16,)).copy().view(np.complex64)
    o = po.OracleFilter(2, taps, 0, 48000, 100)
    assert rev.tobytes() == o.rev_taps.tobytes()
    assert list(rev.real) == [8, 7, 6, 4, 5, 3, 2, 1]


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly(pkg, capfd):
    """No CPU fallback: without a device create returns an error and logs '<3>'."""
    with pytest.raises(RuntimeError):
        pkg.Group(2016000, 262144)
    with pytest.raises(ValueError) as e:
        pkg.XlatingFilter(5, np.ones(57, dtype=np.float32), -12000, 48000, 2000)
    assert e.value.args[0] in (-errno.ENODEV, -errno.EIO)
    assert "<3>" in capfd.readouterr().err
    # the reference's argument check still comes first (src/xlating.c:496)
    with pytest.raises(ValueError) as e:
        pkg.XlatingFilter(5, np.zeros(0, dtype=np.float32), -12000, 48000, 2000)
    assert e.value.args[0] == -1


def test_host_oscillator_is_the_reference_recursion(pkg):
    """csrc/taps_host.c: xl_osc_chain_cf32 (the per-filter engine walks the float oscillator on
    the calling thread) against a numpy float32 restatement of src/xlating.c:70-73: every
    product and sum rounded to float, hypotf = (float)sqrt of the exact double sum."""
    import math
    fn = pkg.lib().xl_osc_chain_cf32
    fn.argtypes = [C.POINTER(C.c_float)] * 2 + [C.c_float] * 2 + [C.POINTER(C.c_float), C.c_int]
    fn.restype = None
    f = np.float32

    def restated(pr, pi, ir, ii, n):
        table = []
        for k in range(n):
            if k % 2 == 0:
                table += [pr, pi]
            pr, pi = f(f(pr * ir) - f(pi * ii)), f(f(pr * ii) + f(pi * ir))
        if n > 0:
            mag = f(math.sqrt(float(pr) * float(pr) + float(pi) * float(pi)))
            pr, pi = f(pr / mag), f(pi / mag)
        return np.array(table, dtype=np.float32), pr, pi

    for center, D in ((-312000, 42), (400000, 21), (7, 3)):
        w = f(2 * math.pi * center / 2016000)
        step = np.exp(np.complex64(1j) * f(-w * f(D)))
        ir, ii = f(step.real), f(step.imag)
        pr, pi = C.c_float(1.0), C.c_float(0.0)
        epr, epi = f(1.0), f(0.0)
        for n in (0, 1, 2, 7, 3121, 3120, 1):  # consecutive calls: the state carries over
            table = (C.c_float * (n + 2))()
            fn(C.byref(pr), C.byref(pi), ir, ii, table, n)
            etable, epr, epi = restated(epr, epi, ir, ii, n)
            assert np.array(table[:len(etable)], dtype=np.float32).tobytes() == etable.tobytes(), (center, n)
            assert f(pr.value).tobytes() == epr.tobytes() and f(pi.value).tobytes() == epi.tobytes(), (center, n)
