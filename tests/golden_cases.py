"""Table of the reference's pinned iterative-reconstruction goldens (tests/test_RecToolsIRCuPy.py of
the reference, line numbers in each entry) and a runner that executes them through the public
classes of tomobar_b200.  Shared by tests/test_gpu_goldens_ir.py and tools/golden_report.py."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
LABELS = ["angles", "detY", "detX"]

ROF10 = {"method": "ROF_TV", "regul_param": 0.0005, "iterations": 10, "time_marching_step": 0.001,
         "device_regulariser": 0}
PD10 = {"method": "PD_TV", "regul_param": 0.0005, "iterations": 10, "device_regulariser": 0}


def normaliser_mean(data, flats, darks):
    """Restatement of supp/suppTools.py:187-264 (method "mean", log=True) for the PWLS goldens."""
    flats = np.mean(flats, 0)
    darks = np.mean(darks, 0)
    denom = flats - darks
    denom[denom <= 0.0] = 1.0
    nomin = data - darks
    nomin[nomin < 0.0] = 1.0
    out = np.true_divide(nomin, denom)
    pos = out > 0.0
    out[pos] = -np.log(out[pos])
    out[out < 0.0] = 0.0
    return out


def load_scan():
    d = np.load(os.path.join(GOLDEN, "normalised_data.npz"))
    scan = {"data": d["data_norm"], "angles": d["angles"]}
    raw_path = os.path.join(GOLDEN, "tomo_standard.npz")
    if os.path.exists(raw_path):
        r = np.load(raw_path)
        scan["raw"] = (r["data"], r["flats"], r["darks"])  # uint16
    return scan


# name -> dict(method, ctor kwargs, data kwargs, algorithm, regularisation, expect{min,max,lc,mean}, rtol/atol of
# the reference test, ref = line numbers in the reference's tests/test_RecToolsIRCuPy.py)
CASES = {
    "landweber_pad1_mean": dict(method="Landweber", pad=1, alg={"iterations": 5}, expect={"mean": 0.0015990591},
                                atol=1e-3, ref="72-96"),
    # the two goldens that encode the reference's non-contiguous-view bug: reproduced
    # with compat_view_bug=True (the first back-projection of CGLS reads the swapped view's raw buffer)
    "cgls_x15_viewbug": dict(method="CGLS", compat=True, alg={"iterations": 15},
                             expect={"min": -0.0039929836, "max": 0.024821747}, rtol=1e-4, ref="128-155"),
    "cgls_x3_after_sirt_viewbug": dict(method="CGLS", compat=True, alg={"iterations": 3},
                                       expect={"min": -0.0030896277, "max": 0.022553273}, rtol=1e-4, ref="190-218"),
    "cgls_pad50_mask2": dict(method="CGLS", pad=50, alg={"iterations": 15, "recon_mask_radius": 2.0},
                             expect={"min": -0.011976417, "max": 0.0382089}, rtol=1e-4, ref="156-187"),
    "fista_2d_x50": dict(method="FISTA", two_d=True, power=True, alg={"iterations": 50},
                         expect={"min": -0.010516173, "max": 0.03179016}, rtol=1e-6, ref="358-388"),
    "fista_pad60_x20": dict(method="FISTA", pad=60, power=True, alg={"iterations": 20, "recon_mask_radius": 2.0},
                            expect={"min": -0.004563322, "max": 0.026597505}, rtol=1e-4, ref="391-421"),
    "fista_pdtv_3d": dict(method="FISTA", power=True, alg={"iterations": 10}, reg=dict(PD10),
                          expect={"min": -0.0003926696, "max": 0.022365307}, rtol=1e-4, ref="424-460"),
    "fista_pdtv_2d": dict(method="FISTA", two_d=True, power=True, alg={"iterations": 100},
                          reg=dict(PD10, iterations=50),
                          expect={"min": -6.906301e-05, "max": 0.019546613}, rtol=1e-4, ref="463-503"),
    "fista_roftv_3d": dict(method="FISTA", power=True, alg={"iterations": 50}, reg=dict(ROF10, iterations=50),
                           expect={"min": -0.0006241638, "max": 0.023243543}, rtol=1e-4, ref="506-543"),
    "fista_os5_3d": dict(method="FISTA", os=5, power=True, alg={"iterations": 10},
                         expect={"lc": 5510.867, "min": -0.01763365, "max": 0.046532914}, rtol=1e-4, ref="546-577"),
    "fista_os5_2d": dict(method="FISTA", os=5, two_d=True, power=True, alg={"iterations": 10},
                         expect={"min": -0.010578496, "max": 0.03182499}, rtol=1e-6, ref="580-611"),
    "fista_os5_pad60": dict(method="FISTA", os=5, pad=60, power=True,
                            alg={"iterations": 10, "recon_mask_radius": 2.0},
                            expect={"lc": 9644.283, "min": -0.011405378, "max": 0.03799749}, rtol=1e-4, ref="614-645"),
    "fista_os5_pdtv_3d": dict(method="FISTA", os=5, power=True, alg={"iterations": 10}, reg=dict(PD10),
                              expect={"lc": 5510.867, "min": -0.00024514267, "max": 0.02189674}, rtol=1e-4,
                              ref="648-687"),
    "fista_os6_pdtv_2d": dict(method="FISTA", os=6, two_d=True, power=True, alg={"iterations": 20},
                              reg=dict(PD10, iterations=30),
                              expect={"min": -9.581739e-05, "max": 0.019569699}, rtol=1e-4, ref="690-729"),
    "fista_os5_roftv_3d": dict(method="FISTA", os=5, power=True, alg={"iterations": 10},
                               reg=dict(ROF10, iterations=20),
                               expect={"lc": 5510.867, "min": -0.006529817, "max": 0.03582852}, rtol=1e-4,
                               ref="732-775"),
    "fista_os6_pwls_pdtv": dict(method="FISTA", os=6, power=True, raw=True, fidelity="PWLS",
                                alg={"iterations": 10}, reg=dict(PD10, regul_param=0.00001),
                                expect={"max": 0.03565}, rtol=1e-3, ref="778-819"),
    "fista_os5_pwls_roftv": dict(method="FISTA", os=5, power=True, raw=True, fidelity="PWLS",
                                 alg={"iterations": 10}, reg=dict(ROF10),
                                 expect={"max": 0.032535}, rtol=1e-3, ref="822-865"),
    "admm_none": dict(method="ADMM", alg={"iterations": 2, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.6},
                      expect={"min": -0.00019439, "max": 0.01522996}, atol=1e-6, ref="887-934"),
    "admm_roftv": dict(method="ADMM", alg={"iterations": 2, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.6},
                       reg=dict(ROF10), expect={"min": -0.0001876, "max": 0.01522454}, atol=1e-6, ref="887-934"),
    "admm_pdtv": dict(method="ADMM", alg={"iterations": 2, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.6},
                      reg=dict(PD10), expect={"min": -0.00014054, "max": 0.0150647}, atol=1e-6, ref="887-934"),
    "admm_os2_none": dict(method="ADMM", os=2, alg={"iterations": 2, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.6},
                          expect={"min": -0.00033778, "max": 0.01923883}, atol=1e-6, ref="937-985"),
    "admm_os2_roftv": dict(method="ADMM", os=2, alg={"iterations": 2, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.6},
                           reg=dict(ROF10), expect={"min": -0.0003317, "max": 0.01923325}, atol=1e-6, ref="937-985"),
    "admm_os2_pdtv": dict(method="ADMM", os=2, alg={"iterations": 2, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.6},
                          reg=dict(PD10), expect={"min": -0.00019503, "max": 0.01889719}, atol=1e-6, ref="937-985"),
    "admm_os24_pwls_warm_pad17": dict(method="ADMM", os=24, pad=17, objsize=128, fidelity="PWLS", warm=True,
                                      alg={"iterations": 8, "ADMM_rho_const": 1.0, "ADMM_relax_par": 1.6},
                                      reg={"method": "PD_TV", "regul_param": 0.001, "iterations": 10},
                                      expect={"max": 0.031305}, rtol=1e-3, shape=(128, 128, 128), ref="988-1030"),
}


def run_case(case, scan):
    import torch

    from tomobar_b200.methodsIR_CuPy import RecToolsIRCuPy

    dev = torch.device("cuda", 0)
    if case.get("raw"):
        # the PWLS goldens start from the raw scan (reference tests :778-780): fused normalisation kernel
        from tomobar_b200.supp.suppTools import normaliser

        data_t = normaliser(*scan["raw"])
    else:
        data_t = torch.from_numpy(scan["data"]).to(dev)
    angles = scan["angles"]
    detX, detY = data_t.shape[2], data_t.shape[1]
    two_d = case.get("two_d", False)
    rec = RecToolsIRCuPy(DetectorsDimH=detX, DetectorsDimH_pad=case.get("pad", 0),
                         DetectorsDimV=None if two_d else detY, CenterRotOffset=0.0, AnglesVec=angles,
                         ObjSize=case.get("objsize", detX), device_projector=0, OS_number=case.get("os"))
    rec.compat_view_bug = bool(case.get("compat", False))
    if two_d:
        _data_ = {"data_fidelity": "LS", "projection_data": data_t[:, 64, :].contiguous(),
                  "data_axes_labels_order": ["angles", "detX"]}
    else:
        _data_ = {"projection_data": data_t, "data_axes_labels_order": list(LABELS)}
    if case.get("fidelity"):
        _data_["data_fidelity"] = case["fidelity"]
    alg = dict(case.get("alg", {}))
    out = {}
    if case.get("power"):
        lc = rec.powermethod(_data_)
        alg["lipschitz_const"] = lc
        out["lc"] = lc
    if case.get("warm"):
        pad = case.get("pad", 0)
        alg["initialise"] = torch.zeros((detY, detX + 2 * pad, detX + 2 * pad), dtype=torch.float32, device=dev)
    reg = dict(case["reg"]) if case.get("reg") else None
    fn = getattr(rec, case["method"])
    res = fn(_data_, alg, reg) if case["method"] in ("FISTA", "ADMM", "OSEM") else fn(_data_, alg)
    if two_d:
        res = res[0]
    res = res.float()
    out.update({"min": float(res.min()), "max": float(res.max()), "mean": float(res.mean()),
                "shape": tuple(res.shape), "dtype": res.dtype})
    return out
