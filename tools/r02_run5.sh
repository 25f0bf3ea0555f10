#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_output16.py tests/test_inverse_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/e2e_trace.py 8 1 2>&1 | tee gpurun_out/r02_e2e_trace.txt
timeout 300 python tools/e2e_trace.py 8 2 2>&1 | tee -a gpurun_out/r02_e2e_trace.txt
