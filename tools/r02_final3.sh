#!/bin/bash
# round 2, last GPU call (about 2 GPU-minutes left): the B64A output added after the budget was spent, and the SDK byte-identity
# test on the rebuilt host walker
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_z_output_b64a.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_final3_b64a.txt
timeout 45 python -m pytest tests/test_sdk_integration_gpu.py -m gpu -x -q -k byte_identical 2>&1 | tail -4 | tee gpurun_out/r02_final3_sdk.txt
