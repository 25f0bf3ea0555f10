#!/bin/bash
# round 2: SDK shim tests, TestCFHD -E with the GPU interposed (format table, encoder pool speed test)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sdk_integration_gpu.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_tests_h.log; tail -12 gpurun_out/r02_tests_h.log
cd integration/_build
(CFHD_B200_STATS=1 timeout 900 ./TestCFHD -E) > ../../gpurun_out/r02_testcfhd_E_gpu.txt 2>&1
cd ../..
grep -i "fps" gpurun_out/r02_testcfhd_E_gpu.txt | head -30; tail -2 gpurun_out/r02_testcfhd_E_gpu.txt
cd integration/_build
for f in yuy2 byr4; do CFHD_B200_STATS=1 ./sdk_roundtrip 4096 2304 6 8 24 0 $f 2>&1 | tail -2; ./sdk_roundtrip_ref 4096 2304 6 8 24 0 $f 2>&1 | tail -1; done > ../../gpurun_out/r02_sdk_4k_b.txt 2>&1
cat ../../gpurun_out/r02_sdk_4k_b.txt | cut -c1-330
