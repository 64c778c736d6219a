"""CPU test (gloo, world_size 2) of the N>1 host logic in bench.py: one process
per GPU, every rank decimates its OWN stream (weak scaling, no data-path
collective), the timed region is reduced with MAX over ranks and the job value is
the aggregate of all ranks' samples over that time."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gloo_aggregation(tmp_path):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29591", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29591",
           os.path.join(ROOT, "tests", "_gloo_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = [json.load(open(tmp_path / f"rank{i}.json")) for i in range(2)]
    for x in res:
        assert x["ms_max"] == 11.0  # slowest rank
        # 2 ranks x 100 steps x 131072 samples in 11 ms
        assert abs(x["value"] - 2 * 100 * 131072 / 11e-3 / 1e6) < 1e-6
        assert x["clients"] == 256
    assert res[0]["digests"] == res[1]["digests"]
    assert res[0]["digests"][0] != res[0]["digests"][1]  # independent streams


def test_reference_arm_only_rank0(tmp_path):
    """--impl reference under a 2-rank launch: rank 1 exits 0 without work (contract)."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0
    assert r.stdout.strip() == ""


def test_client_sharding_covers_every_client_once():
    """BASELINE configs[4] partitioning (bench.shard_clients: client c -> rank c mod N): the
    shards are disjoint, cover all 4096 clients and differ in size by at most one."""
    sys.path.insert(0, ROOT)
    import bench
    for world in (1, 2, 3, 4, 8):
        shards = [bench.shard_clients(4096, r, world) for r in range(world)]
        flat = sorted(c for s in shards for c in s)
        assert flat == list(range(4096))
        assert max(map(len, shards)) - min(map(len, shards)) <= 1
        assert all(c % world == r for r, s in enumerate(shards) for c in s)


def test_sharded_clients_over_gloo_equal_the_unsharded_oracle(tmp_path):
    """Host logic of the broadcast path with 2 gloo ranks: rank 0 broadcasts each block, every
    rank runs the ORACLE for its shard; gathered, the shards' outputs are exactly what one
    process computes for all clients (nothing is lost or duplicated by the partitioning)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29593", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29593",
           os.path.join(ROOT, "tests", "_gloo_worker.py"), str(tmp_path), "broadcast"]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.load(open(tmp_path / "broadcast.json"))
    assert res["clients_checked"] == 12 and res["bit_identical"] is True
