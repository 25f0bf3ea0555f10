#!/bin/bash
# round 2, GPU call 1: new config-size parity tests + ncu --set full of the level-2/3 kernels + PCIe measurement
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -2
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
timeout 1500 python -m pytest tests/test_config_sizes_gpu.py tests/test_sparse.py \
    "tests/test_sdk_integration_gpu.py::test_public_api_roundtrip_height_not_multiple_of_8" -m gpu -x -q > gpurun_out/r02_tests1.log 2>&1
tail -15 gpurun_out/r02_tests1.log
timeout 300 python tools/pcie_bw.py > gpurun_out/r02_pcie.log 2>&1; cat gpurun_out/r02_pcie.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_plane -s 6 -c 2 -f -o gpurun_out/r02_prof_fwdplane \
    python tools/microbench.py --iters 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inv_plane -s 6 -c 2 -f -o gpurun_out/r02_prof_invplane \
    python tools/microbench.py --inverse --iters 2 > /dev/null 2>&1
ls -la gpurun_out | tail -8
