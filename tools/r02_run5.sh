#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/e2e_trace.py 8 1 2>&1 | tee gpurun_out/r02_e2e_trace.txt
timeout 300 python tools/e2e_trace.py 8 2 2>&1 | tee -a gpurun_out/r02_e2e_trace.txt
