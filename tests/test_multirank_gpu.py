"""GPU, >= 2 devices: the processes-per-GPU data path (torchrun + NCCL) end to end at a small size.

Launches bench.py under torch.distributed.run with two ranks (200 k-point shards): the run itself asserts that the packed
all_gather + merge kernel output equals a numpy merge of the separately gathered per-shard rows and that the exact ground
truth agrees with the unmodified reference's BruteforceSearch per shard; here the emitted line is checked.  Skipped on
a single-GPU box (tests/test_sharded_gpu.py covers the single-process sharded path there)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_sharded_search_line(lib):
    if lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, B200HNSW_CACHE="/tmp/b200hnsw_cache_test")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--points", "200000", "--nq", "2000",
           "--steps", "4", "--warmup", "3", "--batches", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["n_gpus"] == 2 and j["unit"] == "shard-searches/s" and j["total_points"] == 400000
    assert abs(j["merged_qps"] * 2 - j["value"]) < 1e-6 * j["value"]
    mc = j["config"]["merge_check"]
    assert mc["equal_to_numpy_merge_of_gathered_shard_rows"] and mc["rows_checked"] == 2000
    assert abs(sum(mc["result_share_per_shard"]) - 1.0) < 1e-3        # both shards contribute to the merged top-k
    assert "identical to the reference BruteforceSearch" in j["config"]["ground_truth"]
    assert j["config"]["recall_at_10"] >= 0.95
