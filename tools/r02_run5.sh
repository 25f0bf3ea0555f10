#!/bin/bash
# round 2: full GPU suite (sparse v2, VLC hand-over in the SDK shim), inverse occupancy A/B, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_tests_b.log; tail -15 gpurun_out/r02_tests_b.log
for v in r1 r1b5; do
  CFB_INV422=$v timeout 120 python tools/kernel_ab.py --level 1 --dir inv 2>&1 | tail -1 | tee -a gpurun_out/r02_ab_inv422.txt
done
timeout 600 python bench.py > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -c 1500 gpurun_out/r02_bench_b.json; tail -3 gpurun_out/r02_bench_b.err
