#!/bin/bash
# One GPU pass: all parity tests, timings of the default and the widened paths.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 200 python tools/microbench.py --inverse 2>&1 | grep -v "^launches"
echo "--- interlaced"; timeout 200 python tools/microbench.py --inverse --interlaced 2>&1 | grep -v "^launches"
echo "--- half"; timeout 200 python tools/microbench.py --inverse --resolution 2 2>&1 | grep inverse
echo "--- quarter"; timeout 200 python tools/microbench.py --inverse --resolution 3 2>&1 | grep inverse
