#!/bin/bash
mkdir -p gpurun_out
for v in tma3 tma5 tma6; do for th in 8 12 16 24; do CFB_TH=$th CFB_FWD422=$v python tools/kernel_ab.py --level 1 --dir fwd; done; done 2>&1 | tee gpurun_out/r02_ab_fwd422_c.txt
