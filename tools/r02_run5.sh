#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_output16.py tests/test_inverse_gpu.py tests/test_sparse.py tests/test_pool_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/e2e_trace.py 8 1 2>&1 | tee gpurun_out/r02_e2e_trace.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 3000 gpurun_out/r02_bench_a.json; tail -3 gpurun_out/r02_bench_a.err
timeout 600 python bench.py --no-cpu-baseline --config rgb444 > gpurun_out/r02_bench_rgb_a.json 2> gpurun_out/r02_bench_rgb_a.err; tail -c 2500 gpurun_out/r02_bench_rgb_a.json; tail -3 gpurun_out/r02_bench_rgb_a.err
timeout 600 python bench.py --no-cpu-baseline --config bayer8k > gpurun_out/r02_bench_bayer_a.json 2> gpurun_out/r02_bench_bayer_a.err; tail -c 2500 gpurun_out/r02_bench_bayer_a.json; tail -3 gpurun_out/r02_bench_bayer_a.err
