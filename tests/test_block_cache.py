"""CPU tests of the drop-in engine's content-addressed input cache
(sdr-server_b200/csrc/block_cache.h): callers that submit identical bytes -- the
reference's per-client copies of one SDR block, src/queue.c:114 -- share one
entry, different bytes never do, entries in use are never recycled."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("bc") / "libbc_shim.so")
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-pthread",
                    "-I" + os.path.join(ROOT, "sdr-server_b200", "csrc"),
                    os.path.join(ROOT, "tests", "block_cache_shim.cpp"), "-o", so], check=True)
    L = C.CDLL(so)
    L.bc_new.restype = C.c_void_p
    L.bc_new.argtypes = [C.c_void_p]
    L.bc_delete.argtypes = [C.c_void_p]
    L.bc_acquire.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.bc_release.argtypes = [C.c_void_p, C.c_int]
    L.bc_matches.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.bc_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.bc_referenced.argtypes = [C.c_void_p]
    L.bc_uploads.restype = C.c_long
    L.bc_slots.restype = C.c_int
    L.bc_key.restype = C.c_uint64
    L.bc_key.argtypes = [C.c_void_p, C.c_size_t]
    L.bc_stress.restype = C.c_long
    L.bc_stress.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    return L


def acquire(L, c, a):
    return L.bc_acquire(c, a.ctypes.data, a.nbytes)


def stats(L, c):
    h, p = C.c_uint64(0), C.c_uint64(0)
    L.bc_stats(c, C.byref(h), C.byref(p))
    return h.value, p.value


def test_key_is_a_function_of_length_and_sampled_bytes(shim):
    rng = np.random.default_rng(1)
    for n in (1, 7, 31, 32, 33, 100, 511, 512, 4096, 262144):
        a = rng.integers(0, 256, n, dtype=np.uint8)
        k = shim.bc_key(a.ctypes.data, n)
        assert k == shim.bc_key(a.copy().ctypes.data, n)
        for pos in {0, n - 1}:                       # first and last byte are always sampled
            b = a.copy()
            b[pos] ^= 1
            assert shim.bc_key(b.ctypes.data, n) != k, (n, pos)
        if n < 512:                                  # short blocks are keyed on every byte
            b = a.copy()
            b[n // 2] ^= 1
            assert shim.bc_key(b.ctypes.data, n) != k
        if n > 1:
            assert shim.bc_key(a.ctypes.data, n - 1) != k


def test_same_key_different_bytes_are_told_apart(shim):
    """Blocks that differ only where the key does not look get the same key; the byte
    comparison keeps them apart, and each is still shared with its own duplicates."""
    c = shim.bc_new(None)
    rng = np.random.default_rng(4)
    x = rng.integers(0, 256, 262144, dtype=np.uint8)
    variants = []
    for v in range(5):
        y = x.copy()
        y[100] = v                                    # byte 100 is not sampled for this length
        assert shim.bc_key(y.ctypes.data, y.nbytes) == shim.bc_key(x.ctypes.data, x.nbytes)
        variants.append(y)
    slots = [acquire(shim, c, y) for y in variants]
    assert len(set(slots)) == 5 and min(slots) >= 0
    again = [acquire(shim, c, y.copy()) for y in reversed(variants)]
    assert again == list(reversed(slots))
    for s_, y in zip(slots, variants):
        assert shim.bc_matches(c, s_, y.ctypes.data, y.nbytes)
    assert stats(shim, c) == (5, 5)
    for s_ in slots + again:
        shim.bc_release(c, s_)
    assert shim.bc_referenced(c) == 0
    shim.bc_delete(c)


def test_identical_blocks_share_one_entry(shim):
    c = shim.bc_new(None)
    rng = np.random.default_rng(2)
    x = rng.integers(0, 256, 65536, dtype=np.uint8)
    u0 = shim.bc_uploads()
    slots = [acquire(shim, c, x.copy()) for _ in range(10)]
    assert slots[0] >= 0 and len(set(slots)) == 1
    assert shim.bc_uploads() - u0 == 1           # crossed "PCIe" once
    assert stats(shim, c) == (9, 1)
    assert shim.bc_matches(c, slots[0], x.ctypes.data, x.nbytes)
    y = x.copy()
    y[-1] ^= 0x80
    sy = acquire(shim, c, y)
    assert sy >= 0 and sy != slots[0]            # one differing byte -> its own entry
    assert shim.bc_matches(c, sy, y.ctypes.data, y.nbytes)
    assert shim.bc_referenced(c) == 11
    for s in slots + [sy]:
        shim.bc_release(c, s)
    assert shim.bc_referenced(c) == 0
    shim.bc_delete(c)


def test_referenced_entries_are_never_recycled_and_lru_is(shim):
    c = shim.bc_new(None)
    rng = np.random.default_rng(3)
    K = shim.bc_slots()
    blocks = [rng.integers(0, 256, 4096 + 16 * i, dtype=np.uint8) for i in range(K + 4)]
    held = [acquire(shim, c, b) for b in blocks[:K]]
    assert sorted(held) == list(range(K))
    assert acquire(shim, c, blocks[K]) == -1     # every entry referenced: the caller goes private
    for s, b in zip(held, blocks):
        assert shim.bc_matches(c, s, b.ctypes.data, b.nbytes)
    shim.bc_release(c, held[2])
    sk = acquire(shim, c, blocks[K])             # recycles the only unreferenced entry (and regrows it)
    assert sk == held[2] and shim.bc_matches(c, sk, blocks[K].ctypes.data, blocks[K].nbytes)
    for i in range(K):
        if i != 2:
            shim.bc_release(c, held[i])
    shim.bc_release(c, sk)
    # block 0 was published first but touched again now -> block 1 is the LRU victim
    s0 = acquire(shim, c, blocks[0])
    assert s0 == held[0]
    shim.bc_release(c, s0)
    s9 = acquire(shim, c, blocks[K + 1])
    assert s9 == held[1]
    shim.bc_release(c, s9)
    assert shim.bc_referenced(c) == 0
    shim.bc_delete(c)


def test_failed_upload_falls_back_to_private(shim):
    flag = C.c_int(1)
    c = shim.bc_new(C.addressof(flag))
    x = np.arange(8192, dtype=np.uint8)
    assert acquire(shim, c, x) == -1
    assert shim.bc_referenced(c) == 0
    flag.value = 0
    s = acquire(shim, c, x)
    assert s >= 0
    shim.bc_release(c, s)
    shim.bc_delete(c)


@pytest.mark.parametrize("threads,pool", [(32, 3), (64, 16), (96, 80)])
def test_thread_per_client_stress(shim, threads, pool):
    c = shim.bc_new(None)
    shared, priv = C.c_long(0), C.c_long(0)
    bad = shim.bc_stress(c, threads, 200, pool, 32768, C.byref(shared), C.byref(priv))
    assert bad == 0
    assert shared.value + priv.value == threads * 200
    assert shared.value > 0
    hits, pubs = stats(shim, c)
    assert hits + pubs == shared.value
    assert shim.bc_referenced(c) == 0
    shim.bc_delete(c)
