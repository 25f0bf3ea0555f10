#!/bin/bash
# round 2, 8 GPUs of one box: the scaling bench at N = 8 (one rank per GPU) and ONE pool over 1 / 4 / 8 devices in one process
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r02_topo.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 \
    > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -c 1800 gpurun_out/r02_bench_8gpu.json; echo
for nd in 1 4 8; do timeout 300 python tools/pool_multi_probe.py $nd 2>&1 | tail -1 | tee -a gpurun_out/r02_pool_multi.txt; done
