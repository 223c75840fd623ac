"""Regenerates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists).  The GPU box has no /root/reference, so everything the `-m gpu`
tests need is committed here.

  normalised_data.npz : the reference's own test scan (tests/test_data/normalised_data.npz,
                        tests/conftest.py:110-121), recompressed; data_norm (180,128,160) f32
                        [angles, detY, detX], angles (180,) f32.
"""
import os
import numpy as np

REF = "/root/reference/tests/test_data"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    d = np.load(os.path.join(REF, "normalised_data.npz"))
    np.savez_compressed(os.path.join(HERE, "normalised_data.npz"),
                        data_norm=d["data_norm"], angles=d["angles"])
    print("wrote normalised_data.npz")
