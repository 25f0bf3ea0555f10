"""Worker for tests/test_multirank_cpu.py: exercises bench.py's multi-rank plumbing with the gloo backend."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    D = bench.Dist(world, backend="gloo")
    D.barrier()
    # each rank owns different frames (weak scaling, no data exchange)
    frames = bench.synthetic_frames(2, 256, 64, seed=bench.shard_seed(rank))
    digest = int(np.sum(frames[0].astype(np.uint64) * np.arange(frames[0].size, dtype=np.uint64).reshape(frames[0].shape) % 1000003))
    seconds = 1.0 + rank            # pretend rank r needed 1 + r seconds
    slowest = D.max(seconds)
    fps = bench.aggregate_fps(world, 100, slowest)
    D.barrier()
    print(json.dumps({"rank": rank, "world": world, "digest": digest, "slowest": slowest, "fps": fps}), flush=True)
    D.close()


if __name__ == "__main__":
    main()
