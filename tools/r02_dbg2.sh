#!/bin/bash
mkdir -p gpurun_out
cd integration/_build
CFHD_DUMP_SAMPLE=/tmp/s_gpu.bin CFHD_B200_DENSE=1 ./sdk_roundtrip 640 96 1 0 > /dev/null 2>&1
CFHD_DUMP_SAMPLE=/tmp/s_gpu_sparse.bin ./sdk_roundtrip 640 96 1 0 > /dev/null 2>&1
CFHD_DUMP_SAMPLE=/tmp/s_ref.bin ./sdk_roundtrip_ref 640 96 1 0 > /dev/null 2>&1
cd ../..
python - <<'PY' > gpurun_out/r02_sample_diff.txt 2>&1
import numpy as np
a=np.fromfile('/tmp/s_gpu.bin',np.uint8); b=np.fromfile('/tmp/s_ref.bin',np.uint8); c=np.fromfile('/tmp/s_gpu_sparse.bin',np.uint8)
print('sizes', a.size, b.size, c.size)
for name,x in (('dense',a),('sparse',c)):
    n=min(x.size,b.size); d=np.nonzero(x[:n]!=b[:n])[0]
    print(name,'differing bytes', d.size, 'first', d[:40].tolist())
    for o in d[:12]:
        print('  @',o, 'gpu', x[max(0,o-8):o+8].tolist(), 'ref', b[max(0,o-8):o+8].tolist())
PY
cat gpurun_out/r02_sample_diff.txt
timeout 600 python -m pytest tests/test_range_audit.py tests/test_config_sizes_gpu.py::test_pool_sparse_4k_interleaved_bitexact tests/test_gop2.py -m gpu -x -q 2>&1 | tail -15
