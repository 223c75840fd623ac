"""Regenerates the committed fixtures under tests/golden/ (run in the build container, where
/root/reference exists).  The GPU box has no /root/reference, so everything the `-m gpu`
tests need is committed here.

  normalised_data.npz : the reference's own test scan (tests/test_data/normalised_data.npz,
                        tests/conftest.py:110-121), recompressed; data_norm (180,128,160) f32
                        [angles, detY, detX], angles (180,) f32.
  tomo_standard.npz   : the raw uint16 scan + flats + darks of the same dataset
                        (tests/test_data/tomo_standard.npz, tests/conftest.py:85-106), which the PWLS
                        goldens normalise with supp/suppTools.py:187-264 first.
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
    r = np.load(os.path.join(REF, "tomo_standard.npz"))
    np.savez_compressed(os.path.join(HERE, "tomo_standard.npz"), data=r["data"], flats=r["flats"], darks=r["darks"])
    print("wrote tomo_standard.npz")
