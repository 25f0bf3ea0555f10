#!/bin/bash
# round 2: e2e timeline with the one-pass sparse kernels; SDK shim: every wired source format
mkdir -p gpurun_out
timeout 300 python tools/e2e_trace.py 8 1 > gpurun_out/r02_e2e_trace_b.txt 2>&1; tail -22 gpurun_out/r02_e2e_trace_b.txt
timeout 1500 python -m pytest tests/test_sdk_integration_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_tests_f.log; tail -15 gpurun_out/r02_tests_f.log
cd integration/_build
for f in yu64 rg48 ar10 byr4; do CFHD_B200_STATS=1 ./sdk_roundtrip 3840 2160 6 8 24 0 $f 2>&1 | tail -2; ./sdk_roundtrip_ref 3840 2160 6 8 24 0 $f 2>&1 | tail -1; done > ../../gpurun_out/r02_sdk_4k_formats.txt 2>&1
cat ../../gpurun_out/r02_sdk_4k_formats.txt | cut -c1-330
