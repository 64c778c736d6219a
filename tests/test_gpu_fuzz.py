"""Randomised scenarios through the batch C ABI against the oracle, client by client and block by block.

Each seed draws a stream (sampling rate, input format), a client population that mixes every kernel family
(tiled classes with natural and skewed input layouts, classes too small for a tile, odd decimations, a long
split-K class in some seeds), ragged block lengths (odd, tiny and empty ones, so the window starts and the ring
position change parity and the ring wraps), clients that attach and detach mid-stream, and a random number of
blocks in flight.  Every output of every client of every block is compared (count per block, values over the
client's whole stream); nothing is sampled.
"""
import numpy as np
import pytest

from oracle import pyoracle as po
from util import assert_cf32_close, rand_block

pytestmark = pytest.mark.gpu

DECIMS = [5, 8, 16, 21, 40, 42, 48, 63]


def draw_clients(rng, fs, n, taps_by_d):
    out = []
    for _ in range(n):
        d = int(rng.choice(DECIMS))
        center = int(rng.integers(-fs // 2 + 1000, fs // 2 - 1000))
        out.append((d, taps_by_d[d], center))
    return out


@pytest.mark.parametrize("seed", list(range(16)))
def test_random_scenarios(pkg, seed):
    rng = np.random.default_rng(9000 + seed)
    fmt = ["cu8", "cs16", "cs8"][seed % 3]
    fs = int(rng.choice([2016000, 2400000, 10000000]))
    max_in = int(rng.choice([16384, 32768, 65536]))
    # one tap set per decimation (so that clients of one decimation form a class); lengths from 31 to ~1200, odd and even
    taps_by_d = {}
    for d in DECIMS:
        if rng.random() < 0.3:
            # arbitrary (also even) lengths, but never shorter than the decimation: with D > T the reference's
            # history_offset underflows (src/xlating.c:76, undefined behaviour -- the oracle, a faithful
            # restatement, crashes there too; lpf.c never produces such taps)
            t = int(rng.integers(max(31, d + 1), 400))
            taps_by_d[d] = (rng.standard_normal(t) * 0.05).astype(np.float32)
        else:
            taps_by_d[d] = pkg.create_low_pass_filter(1.0, fs, fs // d // 2, max(fs // d // int(rng.integers(3, 12)), 50))
    long_d = None
    if seed % 4 == 1:  # a class whose window does not fit a shared-memory tile -> split-K kernels
        long_d = 320
        taps_by_d[long_d] = (rng.standard_normal(4001) * 0.01).astype(np.float32)

    g = pkg.Group(fs, max_in, flags=pkg.XLG_SM_PARTITION if seed % 2 else 0)
    live = {}  # cid -> oracle

    def attach(d, taps, center):
        cid = g.add_client(d, taps, center)
        live[cid] = po.OracleFilter(d, taps, center, fs, max_in)

    for d, taps, center in draw_clients(rng, fs, int(rng.integers(20, 70)), taps_by_d):
        attach(d, taps, center)
    if long_d:
        for c in range(10):
            attach(long_d, taps_by_d[long_d], int(-fs // 3 + c * (fs // 16)))

    pending = []  # (ticket, {cid: expected})
    depth = int(rng.integers(1, pkg.XLG_SLOTS + 1))

    got_all, ref_all = {}, {}  # per client: the whole stream (the contract's tolerance is norm-wise: a block of one or
                               # two outputs that happen to be small has no meaningful scale of its own)

    def drain_one():
        t, exp = pending.pop(0)
        g.wait(t)
        for cid, ref in exp.items():
            y = np.array(g.output(t, cid), dtype=np.complex64)
            assert y.shape == np.asarray(ref).shape, f"seed {seed} ticket {t} client {cid}: {y.shape} outputs, oracle {np.asarray(ref).shape}"
            got_all.setdefault(cid, []).append(y)
            ref_all.setdefault(cid, []).append(np.array(ref, dtype=np.complex64))

    for blk in range(14):
        r = rng.random()
        half = max_in // 2  # complex samples in a full block
        if r < 0.12:
            ns = 0
        elif r < 0.3:
            ns = int(rng.integers(1, 400))
        elif r < 0.55:
            ns = int(rng.integers(half // 4, half))
        else:
            ns = half
        if ns > 1 and rng.random() < 0.35:  # an odd number of complex samples: stream position and window starts flip parity
            ns = ns - 1 if ns % 2 == 0 else ns
        n = 2 * ns
        x = rand_block(rng, fmt, n)
        # churn between blocks: results of tickets in flight must survive (reserve below) or be consumed first
        if blk in (4, 9) and live:
            while pending:
                drain_one()
            for cid in list(rng.choice(sorted(live), size=min(3, len(live)), replace=False)):
                g.remove_client(int(cid))
                del live[int(cid)]
            for d, taps, center in draw_clients(rng, fs, int(rng.integers(1, 6)), taps_by_d):
                attach(d, taps, center)
        t = g.submit(fmt, x)
        pending.append((t, {cid: o.process_cf32(fmt, x) for cid, o in live.items()}))
        while len(pending) > depth - 1:
            drain_one()
    while pending:
        drain_one()
    for cid in ref_all:
        assert_cf32_close(np.concatenate(got_all[cid]), np.concatenate(ref_all[cid]), f"seed {seed} client {cid}")
    kinds = {g.client_info(c)[1] for c in live}
    assert kinds <= {0, 1, 2}
    if long_d:
        assert 2 in kinds
    g.close()
