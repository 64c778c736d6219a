"""GPU tests of the per-filter drop-in ABI when its filters form one band: the engine finds
out that the dsp threads hold the same block sequence and serves them with ONE submit to
the batch engine per block (csrc/stream_overlay.h, csrc/xlating_dropin.cu).  Whatever the
threads do -- drop blocks, lag, attach late, belong to another source -- every
process_* call must return what a reference filter fed the same bytes returns
(src/xlating.c:52-83), within the float contract."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(scenario, clients=24, blocks=16, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_dropin_overlay_worker.py"), scenario,
                        str(clients), str(blocks)], capture_output=True, text=True, timeout=600, env=e)
    assert r.stdout.strip(), r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0 and not line["errors"], (line, r.stderr[-1500:])
    assert line["worst"] < 1e-5
    return line["stream"]


def test_same_stream_is_served_by_one_batch_group():
    st = run("steady", clients=32, blocks=16)
    total = 32 * 16
    assert st["joins"] == 32 and st["desyncs"] == 0
    assert st["served_by_group"] >= 0.8 * total          # all but the first block or two of each filter
    assert st["published"] <= 16 + 2                       # each block submitted ONCE (plus bootstrap races)


def test_filters_that_drop_blocks_fall_out_and_rejoin():
    st = run("drops")
    assert st["desyncs"] >= 1 and st["joins"] > 24 and st["served_by_group"] > 0


def test_late_attachers_join_the_running_stream():
    st = run("late")
    # (a filter that attaches while the others run ahead joins once it has caught up with the
    # newest block: with 16 back-to-back blocks one or two may still be private at the end)
    assert st["joins"] >= 18 and st["desyncs"] == 0


def test_a_second_source_with_the_same_band_parameters_is_not_mixed_in():
    st = run("two_sources")
    assert st["served_by_group"] > 0 and st["joins"] <= 12 + 1


def test_a_client_lagging_a_whole_ring_is_served_privately():
    st = run("lag", env={"XLATING_B200_STREAM_RING": "4"})
    assert st["desyncs"] >= 1


def test_overlay_off_switch():
    st = run("steady", clients=8, blocks=4, env={"XLATING_B200_STREAM": "0"})
    assert st["served_by_group"] == 0 and st["joins"] == 0
