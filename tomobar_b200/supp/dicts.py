"""Validation / default population of the ``_data_``, ``_algorithm_`` and ``_regularisation_``
dictionaries.  Keys, defaults, side effects on the reconstruction object and error types follow
tomobar/supp/dicts.py:6-184; the defaults are kept in tables instead of an if-chain."""

from __future__ import annotations

from typing import Optional, Tuple

import torch

from tomobar_b200._tensors import as_cuda_f32
from tomobar_b200.supp.funcs import _data_dims_swapper

LABELS_3D = ["detY", "angles", "detX"]
LABELS_2D = ["angles", "detX"]

# method -> default outer iterations as (classical, ordered-subsets)   (dicts.py:102-137)
_DEFAULT_ITERATIONS = {
    "SIRT": (200, 200),
    "CGLS": (30, 30),
    "power": (15, 15),
    "Landweber": (1500, 1500),
    "OSEM": (300, 15),
    "FISTA": (400, 20),
    "ADMM": (400, 10),
}
_NO_LIPSCHITZ = {"SIRT", "CGLS", "power", "Landweber", "OSEM"}
_NO_OS = {"SIRT", "CGLS", "Landweber"}

# dicts.py:138-156
_ALGORITHM_DEFAULTS = {
    "initialise": None,
    "nonnegativity": False,
    "recon_mask_radius": 1.0,
    "tolerance": 0.0,
    "verbose": False,
}
# dicts.py:162-183
_REGULARISATION_DEFAULTS = {
    "regul_param": 0.001,
    "iterations": 150,
    "tolerance": 0.0,
    "time_marching_step": 0.005,
    "PD_LipschitzConstant": 12.0,
    "methodTV": 0,
    "device_regulariser": 0,
}


def dicts_check(
    self,
    _data_: dict,
    _algorithm_: Optional[dict] = None,
    _regularisation_: Optional[dict] = None,
    method_run: str = "FISTA",
) -> Tuple[dict, dict, dict]:
    """Populate the three parameter dictionaries (in place, like the reference) and set
    ``self.data_fidelity`` / ``self.nonneg_regul``."""
    if _data_ is None:
        raise NameError("The data dictionary must be always provided")
    if _data_.get("projection_data") is None:
        raise NameError("'projection_data' needs to be provided")

    device = getattr(getattr(self, "Atools", None), "device", None)
    data = as_cuda_f32(_data_["projection_data"], device, "projection data")
    is2d = data.ndim == 2

    labels = _data_.get("data_axes_labels_order")
    _data_.setdefault("data_axes_labels_order", None)
    if labels is not None:
        data = _data_dims_swapper(data, labels, LABELS_2D if is2d else LABELS_3D)
        # the swap has been applied; do not swap again inside the method (dicts.py:83-84)
        _data_["data_axes_labels_order"] = None
    if is2d:
        data = data.unsqueeze(0)
    _data_["projection_data"] = data

    if _data_.get("data_fidelity") is None:
        _data_["data_fidelity"] = "LS"
    if _data_["data_fidelity"] not in {"LS", "PWLS", "KL", "SWLS"}:
        raise ValueError("_data_['data_fidelity'] should be provided as 'LS', 'PWLS', 'KL' (or 'SWLS').")
    self.data_fidelity = _data_["data_fidelity"]
    # robust / ring-artefact extensions: absent from this reference snapshot's dicts.py, keys and
    # defaults as in its legacy demos (Demos/methods_IR_legacy/DemoFISTA_artifacts2D.py:197,243,307-309)
    _data_.setdefault("huber_threshold", None)
    _data_.setdefault("studentst_threshold", None)
    _data_.setdefault("ringGH_lambda", None)
    _data_.setdefault("ringGH_accelerate", 50)
    _data_.setdefault("beta_SWLS", 0.1)
    if _data_["data_fidelity"] == "SWLS" and method_run != "FISTA":
        raise ValueError("The SWLS data term is available in FISTA only")

    use_os = self.OS_number > 1
    if use_os and method_run in _NO_OS:
        raise NameError(
            "There is no ordered-subsets implementation for this reconstruction method, please set OS_number=None"
        )

    # ---- _algorithm_ -------------------------------------------------------------------------
    if _algorithm_ is None:
        _algorithm_ = {}
    if method_run in _NO_LIPSCHITZ:
        _algorithm_["lipschitz_const"] = 0  # bypasses the power method
        if method_run != "OSEM":
            _algorithm_.setdefault("tau_step_lanweber", 1e-05)
            if _algorithm_["tau_step_lanweber"] is None:
                _algorithm_["tau_step_lanweber"] = 1e-05
    if _algorithm_.get("iterations") is None and method_run in _DEFAULT_ITERATIONS:
        _algorithm_["iterations"] = _DEFAULT_ITERATIONS[method_run][1 if use_os else 0]
    if method_run == "ADMM":
        _algorithm_.setdefault("ADMM_rho_const", 1.0)
        _algorithm_.setdefault("ADMM_relax_par", 1.6)
    for key, value in _ALGORITHM_DEFAULTS.items():
        _algorithm_.setdefault(key, value)
    if _algorithm_["nonnegativity"] not in [True, False]:
        raise ValueError("_algorithm_['nonnegativity'] should be set to True or False.")
    self.nonneg_regul = 1 if _algorithm_["nonnegativity"] else 0

    # ---- _regularisation_ --------------------------------------------------------------------
    if _regularisation_ is None:
        _regularisation_ = {}
    if not _regularisation_:
        _regularisation_["method"] = None
    if method_run in {"FISTA", "ADMM", "OSEM"}:
        for key, value in _REGULARISATION_DEFAULTS.items():
            _regularisation_.setdefault(key, value)
    return (_data_, _algorithm_, _regularisation_)
