#!/bin/bash
# round 2: non-negative prescale taps at level 2 (A/B), SDK shim byte identity, full GPU suite, benches of configs 3-5
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02_tests_c.log; tail -8 gpurun_out/r02_tests_c.log
for nn in 0 1; do
  CFB_FWDPLANE_NN=$nn timeout 120 python tools/kernel_ab.py --level 2 --dir fwd 2>&1 | tail -1 | tee -a gpurun_out/r02_ab_fwdplane_nn.txt
  CFB_FWDPLANE_NN=$nn timeout 120 python tools/kernel_ab.py --level 2 --dir fwd --format RG48 --batch 8 2>&1 | tail -1 | tee -a gpurun_out/r02_ab_fwdplane_nn.txt
  CFB_FWDPLANE_NN=$nn timeout 120 python tools/kernel_ab.py --level 3 --dir fwd --format RG48 --batch 8 2>&1 | tail -1 | tee -a gpurun_out/r02_ab_fwdplane_nn.txt
done
for cfg in yuv422 rgb444 bayer8k; do
  timeout 600 python bench.py --config $cfg > gpurun_out/r02_bench_c_$cfg.json 2> gpurun_out/r02_bench_c_$cfg.err; tail -c 600 gpurun_out/r02_bench_c_$cfg.json; echo
done
