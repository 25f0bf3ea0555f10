#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_byr4.py tests/test_rg48.py tests/test_config_sizes_gpu.py tests/test_forward_gpu.py tests/test_ragged_gpu.py tests/test_gop2.py -m gpu -x -q 2>&1 | tail -3
( CFB_FWDPLANE=r1 python tools/kernel_ab.py --level 1 --dir fwd --format BYR4 --batch 4 --width 7680 --height 4320
  python tools/kernel_ab.py --level 1 --dir fwd --format BYR4 --batch 4 --width 7680 --height 4320
  python tools/kernel_ab.py --level 1 --dir fwd --format RG48 --batch 8 ) 2>&1 | tee gpurun_out/r02_ab_fwdplane2.txt
