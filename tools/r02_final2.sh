#!/bin/bash
# round 2, last GPU call: the whole GPU suite on the final tree + ncu --set full of the final inverse level and the level-2 forward
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_tests_final.log; tail -6 gpurun_out/r02_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inv_422 -s 6 -c 1 -f -o gpurun_out/r02_prof_inv422 \
    python tools/kernel_ab.py --level 1 --dir inv --iters 3 > gpurun_out/r02_prof_inv422.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fwd_plane -s 6 -c 1 -f -o gpurun_out/r02_prof_fwdplane_nn \
    python tools/kernel_ab.py --level 2 --dir fwd --iters 3 > gpurun_out/r02_prof_fwdplane_nn.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4; tail -2 gpurun_out/r02_prof_inv422.log
