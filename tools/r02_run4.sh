#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_full.log 2>&1; tail -4 gpurun_out/r02_tests_full.log
for cfg in "4 1" "4 2" "8 1" "8 2" "6 4" "12 2"; do set -- $cfg; timeout 200 python tools/e2e_probe.py $1 $2; done 2>&1 | tee gpurun_out/r02_e2e_sweep.txt
timeout 200 python tools/e2e_probe.py 8 2 fwd 2>&1 | tee -a gpurun_out/r02_e2e_sweep.txt
