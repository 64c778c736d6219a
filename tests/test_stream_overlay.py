"""CPU model check of the drop-in engine's stream overlay
(sdr-server_b200/csrc/stream_overlay.h): the overlay's bookkeeping -- who is a member of the
band's batch group, which block a filter expects, when it publishes, desyncs and rejoins --
driven by 12 threads with the ORACLE standing in for both GPU back ends
(tests/stream_overlay_shim.cpp).  Every call of every filter must return exactly what an
independent oracle filter returns for that filter's own input sequence, in every
scenario: the reference's queue dropping blocks (src/queue.c:90-94), a client lagging a
whole ring, clients attaching late (src/tcp_server.c:301-384), two SDR sources with the
same band parameters, ragged block sizes (test/test_xlating.c:39-81)."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    exe = str(tmp_path_factory.mktemp("so") / "stream_overlay_model")
    subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "sdr-server_b200", "csrc"),
                    "-I" + os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "stream_overlay_shim.cpp"),
                    "-o", exe, "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                    "-Wl,-rpath," + os.path.join(ROOT, "oracle")], check=True)
    return exe


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("scenario", ["steady", "drops", "lag", "two_sources", "late_joiners", "ragged"])
def test_every_call_equals_the_filters_own_oracle(model, scenario, seed):
    r = subprocess.run([model, scenario, str(seed)], capture_output=True, text=True, timeout=300)
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0 and line["mismatching_calls"] == 0, line
    # the scenarios do exercise what they are named after
    assert line["joins"] >= 6 and line["served_by_group"] > line["served_privately"] // 4
    if scenario == "steady":
        assert line["desyncs"] == 0 and line["served_by_group"] >= 0.85 * (line["served_by_group"] + line["served_privately"])
    if scenario in ("drops", "lag"):
        assert line["desyncs"] >= 1  # filters fell out of step, were served privately, and the results still match
    if scenario == "two_sources":
        # one source's filters form the batch; the other source's filters never push a block
        # between the members and their next block (no member ever falls out of step)
        assert line["joins"] == 6 and line["desyncs"] == 0
