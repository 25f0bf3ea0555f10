#!/bin/bash
# round 2: where the SDK arm's time goes at 1080p with 16 pool threads (plan creation, GPU host calls, sparse VLC walk)
mkdir -p gpurun_out
cd integration/_build
for f in yuy2 2vuy yu64; do CFHD_B200_STATS=1 ./sdk_roundtrip 1920 1080 31 16 24 0 $f 2>&1 | tail -2 | cut -c1-700; done > ../../gpurun_out/r02_sdk_1080.txt 2>&1
cat ../../gpurun_out/r02_sdk_1080.txt
cd ../..
timeout 600 python -m pytest tests/test_sdk_integration_gpu.py -m gpu -q 2>&1 | tail -4
