"""N > 1 plumbing of bench.py on CPU: two ranks over gloo (no GPU).  The data path shards frames across ranks
with no collective; torch.distributed only carries the barrier and the max-over-ranks of the timing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_gloo():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29613", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29613", os.path.join(ROOT, "tests", "_rank_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    import re
    recs = [json.loads(m) for m in re.findall(r"\{[^{}]*\}", out.stdout)]      # ranks may interleave on one line
    assert sorted(r["rank"] for r in recs) == [0, 1]
    assert all(r["world"] == 2 for r in recs)
    assert recs[0]["digest"] != recs[1]["digest"]                  # disjoint shards
    assert all(r["slowest"] == 2.0 for r in recs)                   # max over ranks
    assert all(abs(r["fps"] - 2 * 100 / 2.0) < 1e-9 for r in recs)  # whole-job aggregate over the slowest rank


def test_reference_arm_other_ranks_exit_quietly():
    """--impl reference under torchrun: rank != 0 prints nothing and exits 0."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
