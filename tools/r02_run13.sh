#!/bin/bash
# round 2: Bayer sample diff (GPU arm vs reference arm), TestCFHD -E with the plan pool
mkdir -p gpurun_out
cd integration/_build
CFHD_DUMP_SAMPLE=/tmp/b_gpu.bin ./sdk_roundtrip 2048 1152 1 0 24 0 byr4 > /dev/null 2>&1
CFHD_DUMP_SAMPLE=/tmp/b_ref.bin ./sdk_roundtrip_ref 2048 1152 1 0 24 0 byr4 > /dev/null 2>&1
cd ../..
python - <<'PY' > gpurun_out/r02_byr4_diff.txt 2>&1
import numpy as np
a=np.fromfile('/tmp/b_gpu.bin',np.uint8); b=np.fromfile('/tmp/b_ref.bin',np.uint8)
print('sizes', a.size, b.size)
n=min(a.size,b.size); d=np.nonzero(a[:n]!=b[:n])[0]
print('differing bytes', d.size, 'first', d[:60].tolist(), 'last', d[-10:].tolist())
for o in d[:10]:
    print('  @',o, 'gpu', a[max(0,o-12):o+12].tolist(), 'ref', b[max(0,o-12):o+12].tolist())
PY
cat gpurun_out/r02_byr4_diff.txt | cut -c1-400
cd integration/_build
(CFHD_B200_STATS=1 timeout 900 ./TestCFHD -E) > ../../gpurun_out/r02_testcfhd_E_gpu.txt 2>&1
cd ../..
grep -i "fps" gpurun_out/r02_testcfhd_E_gpu.txt | head -30; tail -1 gpurun_out/r02_testcfhd_E_gpu.txt
timeout 900 python -m pytest tests/test_sdk_integration_gpu.py -m gpu -q 2>&1 | tail -6
