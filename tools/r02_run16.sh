#!/bin/bash
# round 2: e2e overlap experiments -- hardware queue count (CUDA_DEVICE_MAX_CONNECTIONS) and slot count
mkdir -p gpurun_out
for conn in 8 32; do for slots in 8 12; do
  echo -n "CUDA_DEVICE_MAX_CONNECTIONS=$conn: " | tee -a gpurun_out/r02_e2e_sweep5.txt
  CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 200 python tools/e2e_probe.py $slots 1 2>&1 | tail -1 | tee -a gpurun_out/r02_e2e_sweep5.txt
done; done
