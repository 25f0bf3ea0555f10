#!/bin/bash
# End-of-round evidence run on the GPU box (one GPU): parity, both bench arms, ncu launch list of the bench command,
# and (FULL=1) one `--set full` capture of each dominant kernel.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
if [ -n "$FULL" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_422 -s 3 -c 1 -f -o gpurun_out/prof_fwd422 \
    python tools/microbench.py --iters 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inv_422 -s 3 -c 1 -f -o gpurun_out/prof_inv422 \
    python tools/microbench.py --inverse --iters 2 > /dev/null 2>&1
fi
ls -la gpurun_out | tail -8
