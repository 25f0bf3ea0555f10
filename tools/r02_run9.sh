#!/bin/bash
# round 2: sparse kernels with 256-thread CTAs -- parity, per-launch times, SDK shim formats
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_sparse.py tests/test_pool_gpu.py tests/test_config_sizes_gpu.py tests/test_sdk_integration_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_tests_e.log; tail -15 gpurun_out/r02_tests_e.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_sparse_ -s 40 -c 60 --csv --log-file gpurun_out/r02_sparse_launches.csv \
    python tools/e2e_probe.py 4 2 > gpurun_out/r02_sparse_launches.log 2>&1
grep -c k_sparse gpurun_out/r02_sparse_launches.csv; grep k_sparse_pack gpurun_out/r02_sparse_launches.csv | tail -3; grep k_sparse_unpack gpurun_out/r02_sparse_launches.csv | tail -3
timeout 300 python tools/e2e_probe.py 8 1 2>&1 | tail -1
cd integration/_build
for f in yuy2 rg48; do CFHD_B200_STATS=1 ./sdk_roundtrip 3840 2160 6 8 24 0 $f 2>&1 | tail -2; ./sdk_roundtrip_ref 3840 2160 6 8 24 0 $f 2>&1 | tail -1; done > ../../gpurun_out/r02_sdk_4k.txt 2>&1
CFHD_B200_DENSE=1 CFHD_B200_STATS=1 ./sdk_roundtrip 3840 2160 6 8 24 0 yuy2 2>&1 | tail -2 >> ../../gpurun_out/r02_sdk_4k.txt
cat ../../gpurun_out/r02_sdk_4k.txt | cut -c1-420
