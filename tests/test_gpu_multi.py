"""Multi-GPU data path of BASELINE configs[4] (one wideband stream NCCL-broadcast, clients
sharded c mod N): 2 ranks over NCCL, every client of every rank against the oracle.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a 1-GPU box.  The host-side partitioning
logic is covered on the CPU (gloo) by tests/test_multi_gpu_host.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_broadcast_sharded_clients_match_the_oracle_on_two_gpus(tmp_path):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "_nccl_cfg5_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = [json.load(open(tmp_path / f"rank{i}.json")) for i in range(2)]
    assert sorted(res[0]["clients"] + res[1]["clients"]) == list(range(48))
    assert set(res[0]["clients"]).isdisjoint(res[1]["clients"])
    for x in res:
        assert x["worst"] < 1e-5 and x["kinds"] == [2]  # every client on the split-K long-filter kernel
