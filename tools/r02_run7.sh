#!/bin/bash
# round 2: full GPU suite (no -x), ncu launch list of the bench command, ncu --set full of the sparse kernels and the final inverse level
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_tests_d.log; tail -25 gpurun_out/r02_tests_d.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-seconds 0.2 --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sparse_ -s 4 -c 2 -f -o gpurun_out/r02_prof_sparse \
    python tools/e2e_probe.py 4 2 > gpurun_out/r02_prof_sparse.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
