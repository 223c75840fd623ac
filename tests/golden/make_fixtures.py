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


def reference_host_fixtures():
    """Outputs of the reference's OWN pure-numpy host functions, imported file by file from
    /root/reference (the package itself cannot be imported here: cupy / astra are missing):
      supp/suppTools.py : normaliser (mean / median), apply_circular_mask
      supp/funcs.py     : _vec_geom_init3D, _swap_data_axes_to_accepted
      fourier.py        : calc_filter (numpy fallback; tomobar.cuda_kernels stubbed)
    -> reference_host_fixtures.npz"""
    import importlib.util
    import itertools
    import sys
    import types

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join("/root/reference/tomobar", rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    out = {}
    st = load("ref_suppTools", "supp/suppTools.py")
    r = np.load(os.path.join(REF, "tomo_standard.npz"))
    data, flats, darks = (np.float32(r[k]) for k in ("data", "flats", "darks"))
    for method in ("mean", "median"):
        norm = st.normaliser(data.copy(), flats.copy(), darks.copy(), method=method)
        out[f"norm_{method}_sub"] = norm[::9, ::8, ::8].copy()           # strided sub-sample
        out[f"norm_{method}_stats"] = np.array([norm.min(), norm.max(), norm.mean(dtype=np.float64)])
    lin = st.normaliser(data.copy(), flats.copy(), darks.copy(), log=False)
    out["norm_nolog_sub"] = lin[::9, ::8, ::8].copy()
    # circular masks: (n, radius) -> kept pixels
    cases = [(16, 1.0), (17, 1.0), (33, 0.7), (64, 0.95), (160, 0.7), (160, 2.0), (50, 1.3)]
    out["mask_cases"] = np.array(cases, dtype=np.float64)
    for i, (n, rad) in enumerate(cases):
        out[f"mask_{i}"] = np.packbits(st.apply_circular_mask(np.ones((n, n), np.float32), rad) > 0)
    fn = load("ref_funcs", "supp/funcs.py")
    ang32 = np.linspace(0, np.pi, 7, endpoint=False).astype(np.float32)
    ang64 = np.linspace(-0.3, 6.0, 9)
    out["geom_angles32"], out["geom_angles64"] = ang32, ang64
    out["geom_vec32_cor0"] = fn._vec_geom_init3D(ang32, 1.0, 1.0, 0.0)
    out["geom_vec64_cor"] = fn._vec_geom_init3D(ang64, 1.0, 1.0, 3.25)
    want3, want2 = ["detY", "angles", "detX"], ["angles", "detX"]
    swaps = []
    for labels in itertools.permutations(want3):
        s = fn._swap_data_axes_to_accepted(list(labels), want3)
        swaps.append("|".join(labels) + "=" + repr(tuple(s)))
    for labels in itertools.permutations(want2):
        s = fn._swap_data_axes_to_accepted(list(labels), want2)
        swaps.append("|".join(labels) + "=" + repr(tuple(s)))
    out["axis_swaps"] = np.array(swaps)
    # fourier.py with its package-level import stubbed
    pkg = types.ModuleType("tomobar")
    ck = types.ModuleType("tomobar.cuda_kernels")
    ck.load_cuda_module = lambda *a, **k: None
    sys.modules.setdefault("tomobar", pkg)
    sys.modules["tomobar.cuda_kernels"] = ck
    fo = load("ref_fourier", "fourier.py")
    for name in ("none", "ramp", "shepp", "cosine", "cosine2", "hamming", "hann", "parzen"):
        for n in (100, 128, 4096):  # the reference needs n//2 + 1 >= 40
            for cut in (1.0, 0.35):
                out[f"filt_{name}_{n}_{cut}"] = np.asarray(fo.calc_filter(n, name, cut), dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "reference_host_fixtures.npz"), **out)
    print("wrote reference_host_fixtures.npz with", len(out), "arrays")


if __name__ == "__main__":
    reference_host_fixtures()
