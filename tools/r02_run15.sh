#!/bin/bash
# round 2: TestCFHD -E with plans pre-created at CFHD_PrepareEncoderPool time; SDK shim tests
mkdir -p gpurun_out
cd integration/_build
(CFHD_B200_STATS=1 timeout 900 ./TestCFHD -E) > ../../gpurun_out/r02_testcfhd_E_gpu.txt 2>&1
cd ../..
grep -i "fps" gpurun_out/r02_testcfhd_E_gpu.txt | awk '{print NR": "$0}' | paste - - - - ; tail -1 gpurun_out/r02_testcfhd_E_gpu.txt | cut -c1-600
timeout 900 python -m pytest tests/test_sdk_integration_gpu.py tests/test_pool_gpu.py tests/test_forward_gpu.py -m gpu -q 2>&1 | tail -4
