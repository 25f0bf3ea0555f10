#!/bin/bash
# round 2: TMA-fed final inverse level -- parity first, then the A/B of every variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_inverse_gpu.py tests/test_output16.py tests/test_golden.py tests/test_config_sizes_gpu.py tests/test_pool_gpu.py tests/test_sparse.py -m gpu -x -q 2>&1 | tail -5
for v in r1 tma24 tma23 tma22 tma16 tma18 tma44 tma43 tma42; do
  CFB_INV422=$v timeout 120 python tools/kernel_ab.py --level 1 --dir inv 2>&1 | tail -1 | tee -a gpurun_out/r02_ab_inv422.txt
done
for th in 8 12 24 32; do
  CFB_TH=$th CFB_INV422=tma24 timeout 120 python tools/kernel_ab.py --level 1 --dir inv 2>&1 | tail -1 | tee -a gpurun_out/r02_ab_inv422.txt
done
