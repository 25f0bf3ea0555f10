#!/bin/bash
mkdir -p gpurun_out
CFB_FWD422=r1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_422 -s 8 -c 1 -f -o gpurun_out/r02_prof_fwd422_r1 \
    python tools/kernel_ab.py --level 1 --dir fwd --iters 3 > /dev/null 2>&1
CFB_FWD422=tma3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_422 -s 8 -c 1 -f -o gpurun_out/r02_prof_fwd422_tma \
    python tools/kernel_ab.py --level 1 --dir fwd --iters 3 > /dev/null 2>&1
ls -la gpurun_out/r02_prof_fwd422*
