"""GPU parity at BASELINE.json's full sizes, EVERY client checked, and of the exact
configuration bench.py times ("Next round" item 1).

* the timed configuration: XLG_OUT_DEVICE | XLG_SM_PARTITION group fed XLG_INPUT_DEVICE
  pointers, >= 64 blocks submitted back to back without waiting (the device pipeline is
  XLG_SLOTS deep, two alternating compute streams inside an SM partition);
* configs[1] (256 clients), configs[2] (64 clients, 1201 taps), the per-GPU shape of
  configs[3] (512 clients), the 1000-client target row (297 taps) and the reference's own
  perf_xlating.c filter (2429 taps), all clients against the oracle;
* >= 300 consecutive blocks through the BATCH path (oscillator pre-pass kernel with its
  double-precision hypotf emulation): a closed form would drift out of tolerance by
  block ~10;
* the split-K long-filter kernel (T = 15419) on 64 clients x 6 blocks.

The oracle runs on a thread pool (tests/util.py: oracle_stream); reference lines:
src/xlating.c:52-83 (loop), test/test_xlating.c:24-81 (state carry-over).
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from util import assert_cf32_close, oracle_stream, rand_block

pytestmark = pytest.mark.gpu


def build(pkg, fs, max_in, plan, flags=0, host_ring=0):
    g = pkg.Group(fs, max_in, flags=flags, host_ring=host_ring)
    tapsets, ids, oracles = {}, [], []
    for p in plan:
        key = (p["cutoff"], p["tw"])
        if key not in tapsets:
            tapsets[key] = pkg.create_low_pass_filter(1.0, fs, p["cutoff"], p["tw"])
        ids.append(g.add_client(p["decimation"], tapsets[key], p["center"]))
        oracles.append(po.OracleFilter(p["decimation"], tapsets[key], p["center"], fs, max_in))
    return g, ids, oracles, tapsets


def check_all(g, ticket, ids, refs, b, what, read=None):
    worst = 0.0
    for i, cid in enumerate(ids):
        y = read(ticket, cid) if read else g.output(ticket, cid)
        worst = max(worst, assert_cf32_close(y, refs[i][b], f"{what}: block {b} client {i}"))
    return worst


def cfg2_plan(pkg, n=256):
    return pkg.client_plan(2016000, [48000 if c % 2 == 0 else 96000 for c in range(n)], tw=None)


def test_timed_configuration_every_client(pkg):
    """What bench.py's `value` leg runs: outputs stay in HBM, the oscillator pre-pass in
    its own 8-SM green context, inputs already on the device, blocks submitted without
    waiting.  Every client of the last XLG_SLOTS blocks of a 64-block burst is compared
    (their history and oscillator depend on all 64), then 8 more blocks one pipeline
    depth at a time, every block of every client."""
    import torch
    rng = np.random.default_rng(101)
    fs, max_in = 2016000, 262144
    flags = pkg.XLG_OUT_DEVICE | pkg.XLG_SM_PARTITION
    g, ids, oracles, _ = build(pkg, fs, max_in, cfg2_plan(pkg), flags)
    n_burst, n_tail = 64, 8
    blocks = [rand_block(rng, "cu8", max_in) for _ in range(n_burst + n_tail)]
    dev = torch.from_numpy(np.stack(blocks)).cuda()
    keep = list(range(n_burst - pkg.XLG_SLOTS, n_burst + n_tail))
    refs = oracle_stream(oracles, "cu8", blocks, keep=keep)

    def submit(b):
        return g.submit("cu8", (dev[b].data_ptr(), max_in), flags=pkg.XLG_INPUT_DEVICE)

    tickets = [submit(b) for b in range(n_burst)]  # never waits explicitly
    worst = 0.0
    for b in range(n_burst - pkg.XLG_SLOTS, n_burst):
        worst = max(worst, check_all(g, tickets[b], ids, refs, b, "burst", read=g.read_output))
    for b0 in range(n_burst, n_burst + n_tail, pkg.XLG_SLOTS):
        ts = [(b, submit(b)) for b in range(b0, min(b0 + pkg.XLG_SLOTS, n_burst + n_tail))]
        for b, t in ts:
            worst = max(worst, check_all(g, t, ids, refs, b, "tail", read=g.read_output))
    assert all(g.client_info(c)[1] == 1 for c in ids)
    assert worst < 1e-5
    g.close()


def test_cfg2_full_size_all_clients(pkg):
    """BASELINE configs[1]: 256 clients, mixed 48/96 ksps, host buffers (the e2e leg)."""
    rng = np.random.default_rng(103)
    fs, max_in = 2016000, 262144
    g, ids, oracles, _ = build(pkg, fs, max_in, cfg2_plan(pkg))
    blocks = [rand_block(rng, "cu8", max_in) for _ in range(4)]
    refs = oracle_stream(oracles, "cu8", blocks)
    tickets = [g.submit("cu8", x) for x in blocks]  # pipelined, XLG_SLOTS in flight
    for b, t in enumerate(tickets):
        g.wait(t)
        check_all(g, t, ids, refs, b, "cfg2")
    assert all(g.client_info(c)[1] == 1 for c in ids)
    g.close()


def test_cfg3_full_size_all_clients(pkg):
    """BASELINE configs[2]: 64 clients at 250 ksps, 10 Msps cs16, 1201 taps (tw = 20060)."""
    rng = np.random.default_rng(107)
    fs, max_in = 10000000, 131072
    plan = pkg.client_plan(fs, [250000] * 64, tw=20060)
    g, ids, oracles, tapsets = build(pkg, fs, max_in, plan)
    assert [len(t) for t in tapsets.values()] == [1201]
    blocks = [rand_block(rng, "cs16", max_in) for _ in range(3)]
    refs = oracle_stream(oracles, "cs16", blocks)
    for b, x in enumerate(blocks):
        t = g.submit("cs16", x)
        g.wait(t)
        check_all(g, t, ids, refs, b, "cfg3")
    assert all(g.client_info(c)[1] == 1 for c in ids)
    g.close()


@pytest.mark.parametrize("n_clients,tw,taps_len", [(512, None, 505), (1000, 16400, 297)])
def test_many_clients_full_size_all_clients(pkg, n_clients, tw, taps_len):
    """configs[3]'s per-GPU shape (512 x 48 ksps, server-default 505 taps) and the
    north_star's target row (1000 x 48 ksps, BASELINE's 297 taps), every client."""
    rng = np.random.default_rng(109 + n_clients)
    fs, max_in = 2016000, 262144
    plan = pkg.client_plan(fs, [48000] * n_clients, tw=tw)
    g, ids, oracles, tapsets = build(pkg, fs, max_in, plan)
    assert [len(t) for t in tapsets.values()] == [taps_len]
    blocks = [rand_block(rng, "cu8", max_in) for _ in range(3)]
    refs = oracle_stream(oracles, "cu8", blocks)
    tickets = [g.submit("cu8", x) for x in blocks]
    for b, t in enumerate(tickets):
        g.wait(t)
        check_all(g, t, ids, refs, b, f"c{n_clients}")
    assert all(g.client_info(c)[1] == 1 for c in ids)
    g.close()


def test_perf_xlating_filter_2429_taps(pkg):
    """The reference's own benchmark filter (test/perf_xlating.c:21-27: tw = 2000 -> 2429
    taps, D = 42, 200 000-byte blocks): 24 clients in the tiled kernel plus the
    benchmark's own lone client (offset -12 kHz) in the generic kernel."""
    rng = np.random.default_rng(113)
    fs, max_in = 2016000, 200000
    taps = pkg.create_low_pass_filter(1.0, fs, 24000, 2000)
    assert len(taps) == 2429
    centers = [-12000] + [int(-900000 + 75000 * c) for c in range(24)]
    g = pkg.Group(fs, max_in)
    lone_taps = pkg.create_low_pass_filter(1.0, fs, 24000, 2001)  # its own (D, T) class: stays generic
    ids = [g.add_client(42, lone_taps if i == 0 else taps, c) for i, c in enumerate(centers)]
    oracles = [po.OracleFilter(42, lone_taps if i == 0 else taps, c, fs, max_in) for i, c in enumerate(centers)]
    blocks = [rand_block(rng, "cu8", max_in) for _ in range(4)]
    refs = oracle_stream(oracles, "cu8", blocks)
    for b, x in enumerate(blocks):
        t = g.submit("cu8", x)
        g.wait(t)
        check_all(g, t, ids, refs, b, "perf filter")
    kinds = [g.client_info(c)[1] for c in ids]
    assert kinds[0] == 0 and set(kinds[1:]) == {1}
    g.close()


def test_batch_path_320_blocks_no_drift(pkg):
    """320 consecutive blocks through the batch engine (phase_cf32_kernel's recursion and
    renormalisation, 48 mixed-rate clients in two tiled classes + 2 generic ones)."""
    rng = np.random.default_rng(127)
    fs, max_in = 2016000, 65536
    plan = cfg2_plan(pkg, 48) + [{"rate": 252000, "decimation": 8, "center": 100000, "cutoff": 100000, "tw": 60000},
                                 {"rate": 16000, "decimation": 126, "center": 777777, "cutoff": 8000, "tw": 3200}]
    g, ids, oracles, _ = build(pkg, fs, max_in, plan)
    n_blocks = 320
    blocks = [rand_block(rng, "cu8", max_in) for _ in range(n_blocks)]
    refs = oracle_stream(oracles, "cu8", blocks)
    worst, pend = 0.0, []
    for b, x in enumerate(blocks):
        pend.append((b, g.submit("cu8", x)))
        if len(pend) == pkg.XLG_SLOTS - 1:
            bb, t = pend.pop(0)
            g.wait(t)
            worst = max(worst, check_all(g, t, ids, refs, bb, "drift"))
    for bb, t in pend:
        g.wait(t)
        worst = max(worst, check_all(g, t, ids, refs, bb, "drift"))
    assert worst < 1e-5, worst
    g.close()


def test_long_filter_split_k_margin_64_clients(pkg):
    """configs[4] shape on 64 clients x 6 blocks.  Split-K changes the order of 15419 fp32
    additions (121 segments of 128 taps), and at this length the reference's OWN sequential
    fp32 sum carries ~sqrt(T) * 6e-8 = 7e-6 of rounding noise, so the distance to it is
    close to the 1e-5 contract by nature (measured: 7.8e-6 worst over these 384
    client-blocks).  Two assertions: the contract itself, and -- against an exact float64
    evaluation of the same dot products -- that the GPU result is at least as close to the
    exact value as the reference's own arithmetic is."""
    rng = np.random.default_rng(131)
    fs, max_in = 61440000, 131072
    taps = pkg.create_low_pass_filter(1.0, fs, 24000, 9600)
    assert len(taps) == 15419
    centers = [int(-30000000 + c * 930000) for c in range(64)]
    g = pkg.Group(fs, max_in)
    ids = [g.add_client(1280, taps, c) for c in centers]
    oracles = [po.OracleFilter(1280, taps, c, fs, max_in) for c in centers]
    blocks = [rand_block(rng, "cs16", max_in) for _ in range(6)]
    refs = oracle_stream(oracles, "cs16", blocks)
    worst = 0.0
    got = [[] for _ in ids]
    for b, x in enumerate(blocks):
        t = g.submit("cs16", x)
        g.wait(t)
        for i, cid in enumerate(ids):
            y = g.output(t, cid)
            got[i].append(y)
            worst = max(worst, assert_cf32_close(y, refs[i][b], f"long: block {b} client {i}"))
    assert {g.client_info(c)[1] for c in ids} == {2}
    assert worst < 1e-5, worst
    print(f"split-K vs the reference's sequential fp32 sum: worst norm-wise distance {worst:.2e}")
    # exact magnitudes: |y_k| = |sum_j xpad[k*D + j] * rev[j]| (the oscillator has modulus 1 to ~1e-7)
    stream = np.concatenate(blocks).astype(np.float64).reshape(-1, 2) / 32768.0
    xpad = np.concatenate([np.zeros(len(taps) - 1, dtype=np.complex128), stream[:, 0] + 1j * stream[:, 1]])
    closer = 0
    for i in (0, 21, 42, 63):
        rev = oracles[i].rev_taps.astype(np.complex128)
        y_gpu, y_ref = np.concatenate(got[i]), np.concatenate(refs[i])
        exact = np.array([np.abs(np.dot(xpad[k * 1280:k * 1280 + len(rev)], rev)) for k in range(len(y_ref))])
        scale = exact.max()
        e_gpu = np.max(np.abs(np.abs(y_gpu.astype(np.complex128)) - exact)) / scale
        e_ref = np.max(np.abs(np.abs(y_ref.astype(np.complex128)) - exact)) / scale
        print(f"client {i}: |y| error vs float64  GPU {e_gpu:.2e}  reference arithmetic {e_ref:.2e}")
        assert e_gpu < 1e-5
        closer += e_gpu <= e_ref + 2e-7
    assert closer == 4, "the split-K sum should be at least as close to exact math as the sequential fp32 sum"
    g.close()


def test_twenty_distinct_classes_all_on_the_tiled_kernel(pkg):
    """sdr-server accepts any client rate that divides the band rate (src/tcp_server.c:101):
    20 different (decimation, taps) classes of 8 clients each in ONE launch -- the class
    table holds 40 -- every client on the tiled kernel, every client against the oracle."""
    rng = np.random.default_rng(137)
    fs, max_in = 2016000, 65536
    rates = [48000, 96000, 24000, 32000, 16000, 8000, 12000, 56000, 72000, 28000,
             36000, 18000, 14400, 9600, 42000, 63000, 84000, 126000, 144000, 168000]
    assert all(fs % r == 0 for r in rates) and len(set(fs // r for r in rates)) == 20
    plan = []
    for r in rates:
        for c in range(8):
            center = int(-fs / 2 + r / 2 + (len(plan) + 0.5) * (fs - r) / 160)
            plan.append({"rate": r, "decimation": fs // r, "center": center, "cutoff": r // 2, "tw": r // 5})
    g, ids, oracles, tapsets = build(pkg, fs, max_in, plan)
    assert len(tapsets) == 20
    blocks = [rand_block(rng, "cu8", max_in) for _ in range(3)]
    refs = oracle_stream(oracles, "cu8", blocks)
    for b, x in enumerate(blocks):
        t = g.submit("cu8", x)
        g.wait(t)
        check_all(g, t, ids, refs, b, "20 classes")
    assert all(g.client_info(c)[1] == 1 for c in ids)
    g.close()
