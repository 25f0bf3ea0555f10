#!/bin/bash
# round 2: what is left of the GPU budget (about half a minute) for the 10-bit RGB output tests
mkdir -p gpurun_out
timeout 28 python -m pytest tests/test_zz_output_rgb30.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_final4_rgb30.txt
