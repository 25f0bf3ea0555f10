timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 900 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
