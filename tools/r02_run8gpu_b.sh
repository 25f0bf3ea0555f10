#!/bin/bash
# round 2, 8 GPUs: ONE pool over 8 devices in one process with NUMA-local host buffers
mkdir -p gpurun_out
for nd in 8; do timeout 300 python tools/pool_multi_probe.py $nd 2>&1 | tail -1 | tee -a gpurun_out/r02_pool_multi.txt; done
