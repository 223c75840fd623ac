"""ctypes binding of libtmb.so (include/tmb.h).  There is no fallback: if the CUDA library is
missing the import fails, and every call raises on a non-zero return code."""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TMB_LIB selects another build of the same library (A/B timing of kernel changes on one GPU box)
LIB_PATH = os.environ.get("TMB_LIB") or os.path.join(_HERE, "libtmb.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make` at the repository root "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). tomobar_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

_vp, _fp, _ip, _dp = C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
_i, _f, _sz = C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/tmb.h one to one
SIGNATURES = {
    "tmb_version": (_i, []),
    "tmb_last_error": (C.c_char_p, []),
    "tmb_geom_create": (_vp, [_i, _i, _i, _i, _dp, _dp, _dp, _i, _i]),
    "tmb_geom_destroy": (None, [_vp]),
    "tmb_geom_subset_size": (_i, [_vp, _i]),
    "tmb_geom_subset_row": (_i, [_vp, _i, _ip]),
    "tmb_geom_table": (_i, [_vp, C.POINTER(C.c_float)]),
    "tmb_geom_fp_launches": (_i, [_vp, _i]),
    "tmb_geom_fp_group": (_i, [_vp, _i]),
    "tmb_geom_workspace_bytes": (_sz, [_vp]),
    "tmb_fp_set_kernel": (_i, [_i]),
    "tmb_fp_set_segment": (_i, [_i]),
    "tmb_fp3d": (_i, [_vp, _i, _fp, _fp, _vp, _vp]),
    "tmb_bp3d": (_i, [_vp, _i, _fp, _fp, _vp, _vp]),
    "tmb_grad": (_i, [_vp, _i, _i, _fp, _fp, _fp, _fp, _vp, _vp]),
    "tmb_grad_ext": (_i, [_vp, _i, _fp, _fp, _fp, _i, _f, _f, _fp, _f, _f, _fp, _fp, _vp, _vp]),
    "tmb_tv_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "tmb_pd_tv": (_i, [_fp, _fp, _i, _i, _i, _f, _i, _i, _i, _f, _i, _vp, _vp]),
    "tmb_rof_tv": (_i, [_fp, _fp, _i, _i, _i, _f, _i, _f, _i, _vp, _vp]),
    "tmb_pd_tv_iter": (_i, [_fp, _fp, _fp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _i, _f, _i, _i, _i,
                            _fp, _vp, _vp, _vp, _fp, _vp]),
    "tmb_pd_tv_iter2": (_i, [_fp] * 9 + [_i, _i, _i, _f, _i, _i, _f, _i, _i] + [_fp] * 10 + [_vp]),
    "tmb_rof_tv_iter": (_i, [_fp, _fp, _fp, _i, _i, _i, _f, _f, _i, _i, _i, _fp, _fp, _vp]),
    "tmb_tv_set_simple_kernels": (_i, [_i]),
    "tmb_tv_set_f2t": (_i, [_i, _i]),
    "tmb_pd_tv_launches": (_i, [_i, _i, _i, _i, _i]),
    "tmb_fista_grad_step": (_i, [_fp, _fp, _fp, _sz, _f, _i, _vp]),
    "tmb_fista_momentum": (_i, [_fp, _fp, _fp, _sz, _f, _vp]),
    "tmb_admm_z_step": (_i, [_fp, _fp, _fp, _fp, _fp, _fp, _sz, _f, _f, _i, _i, _f, _vp]),
    "tmb_admm_u_step": (_i, [_fp, _fp, _fp, _sz, _vp]),
    "tmb_axpy": (_i, [_f, _fp, _fp, _sz, _i, _vp]),
    "tmb_sinc_filter": (_i, [_f, _fp, _i, _f, _vp]),
    "tmb_apply_filter": (_i, [_fp, _fp, _sz, _i, _vp]),
    "tmb_edge_pad": (_i, [_fp, _fp, _sz, _i, _i, _i, _vp]),
    "tmb_circular_mask": (_i, [_fp, _i, _i, _f, _vp]),
    "tmb_normalise": (_i, [_vp, _i, _fp, _fp, _fp, _i, _i, _i, _i, _i, _vp]),
    "tmb_fi_pack": (_i, [_fp, _fp, _i, _i, _i, _vp]),
    "tmb_fi_pack_rows": (_i, [_fp, _sz, _sz, _fp, _i, _i, _i, _vp]),
    "tmb_edge_pad_pair": (_i, [_fp, _fp, _i, _sz, _i, _i, _i, _vp]),
    "tmb_fi_crop_sign": (_i, [_fp, _sz, _fp, _i, _sz, _vp]),
    "tmb_fi_scale_sign": (_i, [_fp, _f, _i, _i, _i, _vp]),
    "tmb_fi_scale_sign_pairs": (_i, [_fp, _fp, _f, _i, _i, _i, _vp]),
    "tmb_fi_gather_pairs": (_i, [_fp, _fp, _fp, _fp, _vp, _i, _f, _i, _i, _i, _vp]),
    "tmb_fi_set_gather": (_i, [_i]),
    "tmb_fi_set_slices_per_thread": (_i, [_i]),
    "tmb_fi_gather": (_i, [_fp, _fp, _fp, _fp, _vp, _i, _f, _i, _i, _i, _vp]),
    "tmb_fi_gather_center": (_i, [_fp, _fp, _fp, _fp, _vp, _i, _f, _i, _i, _i, _i, _vp]),
    "tmb_fi_scatter": (_i, [_fp, _fp, _fp, _i, _f, _i, _i, _i, _i, _vp]),
    "tmb_fi_sign2d": (_i, [_fp, _i, _i, _vp]),
    "tmb_fi_unpad": (_i, [_fp, _fp, _f, _f, _i, _i, _i, _i, _i, _i, _vp]),
    "tmb_fp3d_host": (_i, [_vp, _i, _fp, _fp]),
    "tmb_bp3d_host": (_i, [_vp, _i, _fp, _fp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    if os.environ.get("TMB_LIB") and not hasattr(lib, _name):
        continue  # an older build loaded for A/B timing may lack newer entry points
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


class TmbError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib.tmb_last_error()
        raise TmbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
