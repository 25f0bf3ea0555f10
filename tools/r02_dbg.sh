#!/bin/bash
mkdir -p gpurun_out
CFB_INV422=tma24 timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/kernel_ab.py --level 1 --dir inv --width 1920 --height 1080 --batch 1 --iters 1 2>&1 | grep -v "^$" | head -60 > gpurun_out/r02_dbg_inv_tma.txt
tail -50 gpurun_out/r02_dbg_inv_tma.txt
