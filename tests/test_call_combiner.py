"""CPU tests of the drop-in engine's call combiner (sdr-server_b200/csrc/
call_combiner.h): the reference's one-dsp-thread-per-client calls
(src/dsp_worker.c:41-88) are served in batches; every call is served exactly once,
lanes are exclusive, nobody is left asleep."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("cc") / "libcc_shim.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-pthread",
                    "-I" + os.path.join(ROOT, "sdr-server_b200", "csrc"),
                    os.path.join(ROOT, "tests", "call_combiner_shim.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.cc_stress.restype = C.c_long
    L.cc_stress.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_long)] * 4
    return L


def stress(L, threads, calls, lanes, max_batch, work_us, fail_every=0):
    out = [C.c_long(0) for _ in range(4)]
    bad = L.cc_stress(threads, calls, lanes, max_batch, work_us, fail_every, *[C.byref(o) for o in out])
    return bad, [o.value for o in out]


def test_single_caller_is_its_own_leader(shim):
    bad, (batches, served, biggest, failed) = stress(shim, 1, 500, 2, 64, 0)
    assert bad == 0 and batches == 500 and served == 500 and biggest == 1 and failed == 0


@pytest.mark.timeout(120)
@pytest.mark.parametrize("threads,lanes,max_batch,work_us", [(8, 1, 64, 20), (64, 2, 64, 50), (200, 4, 1024, 30),
                                                             (100, 3, 7, 10), (32, 8, 2, 0)])
def test_many_callers_are_combined_and_all_served(shim, threads, lanes, max_batch, work_us):
    calls = 150
    bad, (batches, served, biggest, failed) = stress(shim, threads, calls, lanes, max_batch, work_us)
    assert bad == 0
    assert served == threads * calls           # nobody lost, nobody served twice
    assert failed == 0
    assert biggest <= max_batch
    if work_us >= 20 and threads >= 64:
        assert batches < served                # calls that arrive together share a batch


@pytest.mark.timeout(120)
def test_batch_status_reaches_every_caller_in_the_batch(shim):
    bad, (batches, served, biggest, failed) = stress(shim, 48, 100, 2, 32, 20, fail_every=5)
    assert bad == 0 and served == 4800
    assert failed > 0                          # the calls of every 5th batch saw its error code
