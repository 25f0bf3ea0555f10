"""Oracle vs the committed golden vectors (generated from the reference itself by
tests/golden/make_golden.py).  Runs on CPU, needs neither /root/reference nor oracle/_ref."""
import glob
import os

import numpy as np
import pytest

import oracle_lib as ol
import parity_util as pu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def load_golden(path):
    z = np.load(path)
    bands = {}
    for k in z.files:
        if k.startswith("b_"):
            _, c, lvl, name = k.split("_")
            bands[(int(c), int(lvl), name)] = z[k]
    return z["frame"], z["divisors"].tolist(), tuple(int(v) for v in z["prescale"]), int(z["quality"]), bands


def test_golden_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    frame, div, prescale, quality, bands = load_golden(path)
    pyr = pu.forward_pyramid_422(ol.oracle(), frame, div, prescale, fmt=0)
    for key, want in bands.items():
        assert np.array_equal(pyr[key], want), f"band {key}"
