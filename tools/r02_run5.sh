#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_output16.py tests/test_sparse.py tests/test_pool_gpu.py tests/test_config_sizes_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/e2e_trace.py 8 1 2>&1 | tee gpurun_out/r02_e2e_trace.txt
for cfg in "8 1" "12 1" "8 2"; do set -- $cfg; timeout 200 python tools/e2e_probe.py $1 $2; done 2>&1 | tee gpurun_out/r02_e2e_sweep4.txt
