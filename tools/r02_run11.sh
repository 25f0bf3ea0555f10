#!/bin/bash
# round 2: e2e timeline, full GPU suite, TestCFHD -E (encoder pool speed test over its format table) with and without the GPU
mkdir -p gpurun_out
timeout 300 python tools/e2e_trace.py 8 1 > gpurun_out/r02_e2e_trace_c.txt 2>&1; tail -20 gpurun_out/r02_e2e_trace_c.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_tests_g.log; tail -15 gpurun_out/r02_tests_g.log
cd integration/_build
(CFHD_B200_STATS=1 timeout 900 ./TestCFHD -E) > ../../gpurun_out/r02_testcfhd_E_gpu.txt 2>&1
(timeout 900 ./TestCFHD_ref -E) > ../../gpurun_out/r02_testcfhd_E_ref.txt 2>&1
cd ../..
grep -i "fps" gpurun_out/r02_testcfhd_E_gpu.txt | head -30; echo ---; grep -i fps gpurun_out/r02_testcfhd_E_ref.txt | head -30; tail -2 gpurun_out/r02_testcfhd_E_gpu.txt
