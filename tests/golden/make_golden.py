"""Generates the golden fixtures in this directory FROM THE REFERENCE ITSELF (oracle/_ref =
the unmodified reference compiled in place).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Each .npz holds a Qbist input frame (TestCFHD's generator, seed 50) and every wavelet band the
reference's own EncodeSample produced for it (transform[c]->wavelet[k]->band[b]), plus the
quantisation tables it used.  tests/test_golden.py (CPU, oracle) and tests/test_forward_gpu.py
(GPU, CUDA path) compare against these files, so the GPU box needs neither the reference tree
nor oracle/_ref."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import parity_util as pu  # noqa: E402


def main():
    ref_lib = ol.load_ref()
    for (w, h, frame_no, quality) in [(256, 64, 1, 4), (512, 128, 2, 4), (704, 96, 1, 3)]:
        frame = pu.qbist_yuy2(ref_lib, w, h, frame_no)
        bands, div, prescale, sample = pu.ref_encode_frame(ref_lib, frame, w, h, pu.COLOR_FORMAT_YUYV, 0, 3, quality)
        arrays = {"frame": frame, "divisors": np.array(div, np.int32), "prescale": np.array(prescale[0], np.int32),
                  "quality": np.array(quality), "sample_size": np.array(sample.size)}
        for (c, lvl, name), a in bands.items():
            arrays[f"b_{c}_{lvl}_{name}"] = a
        path = os.path.join(HERE, f"qbist_yuy2_{w}x{h}_f{frame_no}_q{quality}.npz")
        np.savez_compressed(path, **arrays)
        print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
