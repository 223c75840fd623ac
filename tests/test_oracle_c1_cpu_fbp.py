"""BASELINE.json config 1: the reference's CPU direct method for 2-D data -- ``RecToolsDIR(device_projector="cpu")
.FBP`` = ``_filtersinc2D`` (methodsDIR.py:295-320) + ASTRA's CPU ``BP`` with the ``line`` projector
(methodsDIR.py:161-168) -- restated in oracle/oracle.py::fbp2d_cpu + oracle/fbp2d_oracle.c and pinned on the
reference's own golden for that path (tests/test_RecToolsDIR.py:198-218; the reference's eps is 1e-6, the
restatement lands within 2.1e-6 / 0.9e-6)."""

import os

import numpy as np
from numpy.testing import assert_allclose

from golden_cases import normaliser_mean
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _slice60():
    raw = np.load(os.path.join(GOLD, "tomo_standard.npz"))
    angles = np.load(os.path.join(GOLD, "normalised_data.npz"))["angles"]
    norm = normaliser_mean(raw["data"], raw["flats"], raw["darks"]).astype(np.float32)
    return norm[:, 60, :], angles


def test_cpu_fbp2d_golden():
    data2d, angles = _slice60()
    rec = O.fbp2d_cpu(data2d, angles, data2d.shape[1])
    assert rec.dtype == np.float32 and rec.shape == (160, 160)
    assert_allclose(rec.min(), -0.010723082, rtol=4e-6)
    assert_allclose(rec.max(), 0.030544952, rtol=4e-6)


def test_cpu_fbp2d_all_cores_equals_single_thread():
    data2d, angles = _slice60()
    one = O.fbp2d_cpu(data2d, angles, 160, threads=1)
    many = O.fbp2d_cpu(data2d, angles, 160, threads=max(2, O.threads()))
    assert np.abs(many - one).max() <= 2e-6 * np.abs(one).max()  # private images summed in another order


def test_cpu_fbp2d_is_the_vertical_flip_of_the_3d_path():
    """the 2-D CPU class is y-up (row 0 on top), the 3-D path is row <-> +y."""
    data2d, angles = _slice60()
    cpu = O.fbp2d_cpu(data2d, angles, 160)
    gpu_model = O.RecDIR(160, 0, 1, 0.0, angles, 160).FBP(data2d[:, None, :], cutoff_freq=1.1)[0]
    c_flip = np.corrcoef(cpu[::-1].ravel(), gpu_model.ravel())[0, 1]
    c_asis = np.corrcoef(cpu.ravel(), gpu_model.ravel())[0, 1]
    assert c_flip > 0.99 and c_flip > c_asis + 0.05
