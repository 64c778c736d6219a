"""GPU tests of the host-side server model (sdr-server_b200/host): ingest thread ->
one GPU submit per block -> per-client dsp threads -> socket/file, checked against
the reference's own end-to-end goldens (test/test_tcp_server.c:154-248) and the
oracle.  This is SURVEY 8f rows 1-2: the dsp_worker/sdr_callback integration on top
of the batch ABI, with the reference's thread-per-client model."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from util import assert_cf32_close, ramp, rand_block, trunc4

pytestmark = pytest.mark.gpu

FMT_CODE = {"cu8": 0, "cs8": 1, "cs16": 2}


def make_stream(pkg, tmp_path, fmt, band_rate, buffer_size, queue_size=8, use_gzip=0):
    H = pkg.host_lib()
    cfg = pkg.XlStreamConfig(FMT_CODE[fmt], band_rate, buffer_size, queue_size, 5, str(tmp_path).encode(), 0, use_gzip)
    s = C.c_void_p()
    code = H.xl_stream_create(C.byref(cfg), C.byref(s))
    assert code == 0, code
    return H, s, cfg


@pytest.mark.parametrize("fmt,key", [("cu8", "rtlsdr_cu8"), ("cs16", "airspy_cs16"), ("cs8", "hackrf_cs8")])
def test_reference_end_to_end_goldens(pkg, fixtures, tmp_path, fmt, key):
    """test_rtlsdr / test_airspy / test_hackrf: two clients on the same band, one to a
    socket (a pipe here), one to a file; 200 input elements -> 20 cf32 each."""
    s_ = fixtures["tcp_server"]["setup"]
    H, st, _ = make_stream(pkg, tmp_path, fmt, s_["band_sampling_rate"], s_["buffer_size"])
    r, w = os.pipe()
    band = 460100200
    c0 = pkg.XlClientConfig(band + s_["center_offset"] & 0xFFFFFFFF, s_["client_rate"], band, 1, w, 0)
    c1 = pkg.XlClientConfig(band + s_["center_offset"] & 0xFFFFFFFF, s_["client_rate"], band, 0, -1, 1)
    assert H.xl_stream_add_client(st, C.byref(c0)) == 0
    assert H.xl_stream_add_client(st, C.byref(c1)) == 0
    assert H.xl_stream_client_count(st) == 2
    x = ramp(fmt, 0, s_["input_elements"])
    assert H.xl_stream_push(st, x.ctypes.data, x.nbytes) == 0
    H.xl_stream_flush(st)
    exp = np.array(fixtures["tcp_server"][key], dtype=np.float32)
    got_sock = np.frombuffer(os.read(r, exp.nbytes), dtype=np.float32)
    H.xl_stream_destroy(st)
    os.close(r)
    got_file = np.fromfile(os.path.join(tmp_path, "1.cf32"), dtype=np.float32)
    for got in (got_sock, got_file):
        assert got.shape == exp.shape
        bad = np.nonzero(trunc4(got) != trunc4(exp))[0]
        for i in bad:
            assert abs(float(got[i]) - float(exp[i])) <= 1e-6


def test_many_clients_many_blocks_files_match_oracle(pkg, tmp_path):
    """40 file clients (mixed 48/96 ksps) + attach/detach while streaming; every
    client's file must equal the oracle's stream for the blocks it was attached for."""
    fs, buf = 2016000, 65536
    H, st, _ = make_stream(pkg, tmp_path, "cu8", fs, buf, queue_size=16)
    rng = np.random.default_rng(31)
    band = 100000000
    plan = pkg.client_plan(fs, [48000 if c % 2 == 0 else 96000 for c in range(40)])
    oracles = {}
    for cid, p in enumerate(plan):
        cc = pkg.XlClientConfig((band + p["center"]) & 0xFFFFFFFF, p["rate"], band, 0, -1, cid)
        assert H.xl_stream_add_client(st, C.byref(cc)) == 0
        taps = po.lpf_design(1.0, fs, p["rate"] // 2, p["rate"] // 5)
        oracles[cid] = [po.OracleFilter(fs // p["rate"], taps, p["center"], fs, buf), []]
    for blk in range(10):
        if blk == 4:  # one client leaves, one joins (zero history from here on)
            assert H.xl_stream_remove_client(st, 3) == 0
            del oracles[3]
            p = {"rate": 48000, "center": 54321}
            cc = pkg.XlClientConfig(band + p["center"], p["rate"], band, 0, -1, 99)
            assert H.xl_stream_add_client(st, C.byref(cc)) == 0
            taps = po.lpf_design(1.0, fs, 24000, 9600)
            oracles[99] = [po.OracleFilter(42, taps, p["center"], fs, buf), []]
        x = rand_block(rng, "cu8", buf if blk != 6 else 10000)
        assert H.xl_stream_push(st, x.ctypes.data, x.nbytes) == 0
        for cid, (o, acc) in oracles.items():
            acc.append(o.process_cf32("cu8", x))
    H.xl_stream_flush(st)
    H.xl_stream_destroy(st)
    for cid, (o, acc) in oracles.items():
        ref = np.concatenate(acc)
        got = np.fromfile(os.path.join(tmp_path, f"{cid}.cf32"), dtype=np.complex64)
        assert_cf32_close(got, ref, f"client {cid}")


def test_airspy_golden_through_the_gzip_destination(pkg, fixtures, tmp_path):
    """test/test_tcp_server.c:188-220 (test_airspy runs with use_gzip = true): the file
    client's samples are written with gzwrite (src/dsp_worker.c:10-26, :126-133)."""
    import gzip
    s_ = fixtures["tcp_server"]["setup"]
    H, st, _ = make_stream(pkg, tmp_path, "cs16", s_["band_sampling_rate"], s_["buffer_size"], use_gzip=1)
    band = 460100200
    c1 = pkg.XlClientConfig(band + s_["center_offset"] & 0xFFFFFFFF, s_["client_rate"], band, 0, -1, 1)
    assert H.xl_stream_add_client(st, C.byref(c1)) == 0
    x = ramp("cs16", 0, s_["input_elements"])
    assert H.xl_stream_push(st, x.ctypes.data, x.nbytes) == 0
    assert H.xl_stream_flush(st) == 0
    H.xl_stream_destroy(st)
    assert not os.path.exists(os.path.join(tmp_path, "1.cf32"))
    with gzip.open(os.path.join(tmp_path, "1.cf32.gz"), "rb") as f:
        got = np.frombuffer(f.read(), dtype=np.float32)
    exp = np.array(fixtures["tcp_server"]["airspy_cs16"], dtype=np.float32)
    assert got.shape == exp.shape
    for i in np.nonzero(trunc4(got) != trunc4(exp))[0]:
        assert abs(float(got[i]) - float(exp[i])) <= 1e-6


def test_slow_socket_client_never_receives_a_torn_block(pkg, tmp_path):
    """a consumer blocked in write() while the producer laps the result
    ring must lose WHOLE blocks (like the reference's overwrite-newest queue,
    src/queue.c:90-94, with its detached node protected, :150-158), never a block whose
    bytes were replaced under it.  One socket client is not drained while 24 blocks are
    pushed through a 4-deep queue / result ring; afterwards everything it did send must be
    a concatenation of complete, correct blocks in stream order, and a fast file client
    on the same stream must have lost nothing."""
    import threading
    fs, buf = 2016000, 65536
    H, st, _ = make_stream(pkg, tmp_path, "cu8", fs, buf, queue_size=4)
    rng = np.random.default_rng(83)
    band = 100000000
    r, w = os.pipe()
    slow = pkg.XlClientConfig(band + 250000, 96000, band, 1, w, 0)   # socket destination
    fast = pkg.XlClientConfig(band - 300000, 48000, band, 0, -1, 1)  # file destination
    assert H.xl_stream_add_client(st, C.byref(slow)) == 0
    assert H.xl_stream_add_client(st, C.byref(fast)) == 0
    o_slow = po.OracleFilter(21, po.lpf_design(1.0, fs, 48000, 19200), 250000, fs, buf)
    o_fast = po.OracleFilter(42, po.lpf_design(1.0, fs, 24000, 9600), -300000, fs, buf)
    ref_slow, ref_fast = [], []
    import time
    for blk in range(24):
        x = rand_block(rng, "cu8", buf)
        assert H.xl_stream_push(st, x.ctypes.data, x.nbytes) == 0
        ref_slow.append(o_slow.process_cf32("cu8", x))
        ref_fast.append(o_fast.process_cf32("cu8", x))
        time.sleep(0.002)  # the slow client's thread is stuck in write() after ~5 blocks (64 KiB pipe)
    got = bytearray()

    def drain():
        while True:
            chunk = os.read(r, 1 << 20)
            if not chunk:
                return
            got.extend(chunk)

    t = threading.Thread(target=drain)
    t.start()
    assert H.xl_stream_flush(st) == 0
    H.xl_stream_destroy(st)  # closes nothing of ours: the write end is ours to close
    os.close(w)
    t.join()
    os.close(r)
    stream = np.frombuffer(bytes(got), dtype=np.complex64)
    pos, matched = 0, []
    for b, ref in enumerate(ref_slow):
        n = len(ref)
        if pos + n <= len(stream) and np.max(np.abs(stream[pos:pos + n] - ref)) <= 1e-5 * np.max(np.abs(ref)):
            matched.append(b)
            pos += n
    assert pos == len(stream), f"{len(stream) - pos} samples after block {matched[-1] if matched else None} match no whole block"
    assert 0 in matched and len(matched) >= 4          # it did deliver what it could ...
    assert len(matched) < 24                            # ... and it did lose blocks (the test is meaningful)
    got_fast = np.fromfile(os.path.join(tmp_path, "1.cf32"), dtype=np.complex64)
    assert_cf32_close(got_fast, np.concatenate(ref_fast), "fast client")
