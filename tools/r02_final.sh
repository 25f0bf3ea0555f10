#!/bin/bash
# round 2, end-of-round evidence on one GPU: parity, smoke, bench lines of the three BASELINE configs + reference arm, ncu launch
# list of the bench command, ncu --set full of the final inverse level and of the sparse kernels, TestCFHD -E
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_tests_final.log; tail -12 gpurun_out/r02_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -c 500 gpurun_out/r02_bench_reference.json; echo
for cfg in yuv422 rgb444 bayer8k; do
  timeout 600 python bench.py --config $cfg > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err; tail -c 300 gpurun_out/r02_bench_$cfg.json; echo
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-seconds 0.2 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inv_422 -s 8 -c 1 -f -o gpurun_out/r02_prof_inv422 \
    python tools/kernel_ab.py --level 1 --dir inv --iters 3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sparse_ -s 4 -c 2 -f -o gpurun_out/r02_prof_sparse_b \
    python tools/e2e_probe.py 4 2 > /dev/null 2>&1
cd integration/_build
(CFHD_B200_STATS=1 timeout 900 ./TestCFHD -E) > ../../gpurun_out/r02_testcfhd_E_gpu.txt 2>&1
cd ../..
grep -c fps gpurun_out/r02_testcfhd_E_gpu.txt; tail -1 gpurun_out/r02_testcfhd_E_gpu.txt; ls -la gpurun_out/*.ncu-rep | tail -3
