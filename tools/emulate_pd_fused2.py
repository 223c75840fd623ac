"""CPU emulation of the data flow of k_pd_tv3d_f2 (tomobar_b200/csrc/tmb_tv.cu).

The fused two-iteration PD_TV kernel is a per-warp state machine (plane lag, row sweep, lane-private
shared-memory slots, shuffles, clamped loads, unstored window edges).  This script replays exactly that
state machine in numpy -- one "warp" at a time, lanes as an array axis, shared-memory slots poisoned
with NaN -- and compares what it stores with two applications of a plain whole-volume iteration.
It checks the INDEX LOGIC of the kernel without a GPU (the arithmetic is the same numpy expression on
both sides, so agreement is bit-exact when the logic is right).

    python tools/emulate_pd_fused2.py
"""

import itertools
import sys

import numpy as np

F32 = np.float32
S, WARPS, OUT = 4, 4, 120
ROWS = S + 4
UA, UA2, PA, P3A6, IN = 0, 6, 12, 27, 28


def dual_step(p1, p2, p3, d1, d2, d3, sigma, aniso):
    p1 = p1 + sigma * d1
    p2 = p2 + sigma * d2
    p3 = p3 + sigma * d3
    if aniso:
        return np.clip(p1, -1, 1), np.clip(p2, -1, 1), np.clip(p3, -1, 1)
    den = p1 * p1 + p2 * p2 + p3 * p3
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.where(den > 1, F32(1) / np.sqrt(den), F32(1)).astype(F32)
    return p1 * s, p2 * s, p3 * s


def primal(u, q1, p1m, q2, p2m, q3, p3m, inp, tau, lt, theta, nonneg):
    ub = np.maximum(u, 0) if nonneg else u
    div = -(q1 - p1m) + -(q2 - p2m) + -(q3 - p3m)
    nu = (ub - tau * div + lt * inp) / (F32(1) + lt)
    return (nu + theta * (nu - ub)).astype(F32)


def iterate_plain(inp, U, P, sigma, tau, lt, theta, nonneg, aniso):
    """One iteration on the whole volume (semantics of k_pd_tv)."""
    dz, dy, dx = U.shape

    def fwd(a, axis):
        n = a.shape[axis]
        idx = np.arange(n) + 1
        idx[-1] = n - 2  # the last index uses its backward neighbour
        return np.take(a, idx, axis=axis)

    q = dual_step(P[0], P[1], P[2], fwd(U, 2) - U, fwd(U, 1) - U, fwd(U, 0) - U, sigma, aniso)

    def bwd0(a, axis):
        out = np.zeros_like(a)
        sl = [slice(None)] * 3
        sr = [slice(None)] * 3
        sl[axis] = slice(1, None)
        sr[axis] = slice(0, -1)
        out[tuple(sl)] = a[tuple(sr)]
        return out

    Un = primal(U, q[0], bwd0(q[0], 2), q[1], bwd0(q[1], 1), q[2], bwd0(q[2], 0), inp, tau, lt, theta, nonneg)
    return Un, [a.astype(F32) for a in q]


def emulate(inp, U, P, sigma, tau, lt, theta, nonneg, aniso, zrun):
    dz, dy, dx = U.shape
    Uo = np.full_like(U, np.nan)
    Q = [np.full_like(U, np.nan) for _ in range(3)]
    stores = np.zeros(U.shape, dtype=np.int32)
    lanes = np.arange(32)
    gx, gy, gz = -(-dx // OUT), -(-dy // (S * WARPS)), -(-dz // zrun)

    def shfl_down(v):  # value of lane + 1 (lane 31 keeps its own)
        return np.concatenate([v[1:], v[-1:]])

    def shfl_up(v):
        return np.concatenate([v[:1], v[:-1]])

    for bx, by, bz, warp in itertools.product(range(gx), range(gy), range(gz), range(WARPS)):
        x0 = bx * OUT - 4
        xa = x0 + 4 * lanes
        y0 = (by * WARPS + warp) * S
        za, zb = bz * zrun, min(dz, bz * zrun + zrun)
        if y0 >= dy or za >= zb:
            continue
        firstx, lastx = xa == 0, xa + 4 == dx
        st_lane = (lanes >= 1) & (lanes <= 30) & (xa < dx)
        xl = np.clip(xa, 0, dx - 4)
        cols = xl[:, None] + np.arange(4)[None, :]  # [lane, 4]
        rows = np.clip(y0 - 2 + np.arange(ROWS), 0, dy - 1)
        sm = np.full((32, 32, 4), np.nan, dtype=F32)  # [slot, lane, component]

        def ldv4(a, z, k):
            return a[z, rows[k]][cols].astype(F32)

        def load_packet(z, k):
            pk = {"un": ldv4(U, z - 1 if z == dz - 1 else z + 1, k)}
            if k <= S + 2:
                pk["p1"], pk["p2"], pk["p3"] = ldv4(P[0], z, k), ldv4(P[1], z, k), ldv4(P[2], z, k)
                if k >= 1:
                    pk["in"] = ldv4(inp, z, k)
            return pk

        def dual_row(p1, p2, p3, u, uy, un, lastx):
            ux3 = shfl_down(u[:, 0])
            ux3 = np.where(lastx, u[:, 2], ux3)
            uxp = np.stack([u[:, 1], u[:, 2], u[:, 3], ux3], axis=1)
            return dual_step(p1, p2, p3, uxp - u, uy - u, un - u, sigma, aniso)

        def primal_row(u, q1, q2, q3, pmy, pmz, inn):
            pm = shfl_up(q1[:, 3])
            pm = np.where(firstx, F32(0), pm)
            p1m = np.stack([pm, q1[:, 0], q1[:, 1], q1[:, 2]], axis=1)
            return primal(u, q1, p1m, q2, pmy, q3, pmz, inn, tau, lt, theta, nonneg)

        zs, zB0 = max(za - 2, 0), max(za - 1, 0)
        uc = [ldv4(U, zs, k) for k in range(ROWS)]
        p3b = [np.zeros((32, 4), F32) for _ in range(S)]
        nxt = load_packet(zs, 0)
        zero4 = np.zeros((32, 4), F32)
        for z in range(zs, zb + 1):
            doA, doB, emit, hasz = z < dz, z - 1 >= zB0, z - 1 >= za, z > 0
            more = z + 1 <= zb and z + 1 < dz
            ua_dst = UA2 if z == dz - 1 else UA
            cen_src = UA if doA else UA2
            p2a = p2b = cen_prev = un_saved = None
            for k in range(ROWS):
                cur = nxt
                if doA:
                    if k < S + 3:
                        nxt = load_packet(z, k + 1)
                    elif more:
                        nxt = load_packet(z + 1, 0)
                    else:
                        nxt = None
                y = y0 - 2 + k
                hasy, lasty = y > 0, y == dy - 1
                qa = ua = None
                if doA and k <= S + 2:
                    u = uc[k]
                    uy = uc[k - 1] if (k > 0 and lasty) else uc[k + 1]
                    qa = dual_row(cur["p1"], cur["p2"], cur["p3"], u, uy, cur["un"], lastx)
                    if k >= 1:
                        pmy = p2a if hasy else zero4
                        pmz = sm[PA + 3 * (k - 1) + 2 if k <= S + 1 else P3A6] if hasz else zero4
                        ua = primal_row(u, qa[0], qa[1], qa[2], pmy, pmz, cur["in"])
                    p2a = qa[1]
                if doB and 1 <= k <= S + 1:
                    cen, cnx = sm[cen_src + k - 1].copy(), sm[cen_src + k].copy()
                    fw = ua if doA else sm[UA + k - 1].copy()
                    r = [sm[PA + 3 * (k - 1) + c].copy() for c in range(3)]
                    uy = cen_prev if lasty else cnx
                    r = dual_row(r[0], r[1], r[2], cen, uy, fw, lastx)
                    if k >= 2:
                        pmy = p2b if hasy else zero4
                        o4 = primal_row(cen, r[0], r[1], r[2], pmy, p3b[k - 2], sm[IN + k - 2].copy())
                        if emit and y < dy:
                            for lane in np.nonzero(st_lane)[0]:
                                c = cols[lane]
                                Uo[z - 1, rows[k], c] = o4[lane]
                                for comp in range(3):
                                    Q[comp][z - 1, rows[k], c] = r[comp][lane]
                                stores[z - 1, rows[k], c] += 1
                        p3b[k - 2] = r[2]
                    p2b = r[1]
                    cen_prev = cen
                if doA:
                    if 1 <= k <= S + 2:
                        sm[ua_dst + k - 1] = ua
                        if k <= S + 1:
                            for c in range(3):
                                sm[PA + 3 * (k - 1) + c] = qa[c]
                        else:
                            sm[P3A6] = qa[2]
                        if 2 <= k <= S + 1:
                            sm[IN + k - 2] = cur["in"]
                    if k >= 1:
                        uc[k - 1] = un_saved
                    un_saved = cur["un"]
            if doA:
                uc[S + 3] = un_saved
    return Uo, Q, stores


def run_case(shape, zrun, nonneg, aniso, seed):
    rng = np.random.default_rng(seed)
    dz, dy, dx = shape
    inp = rng.standard_normal(shape).astype(F32)
    U = (inp + 0.3 * rng.standard_normal(shape)).astype(F32)
    P = [(0.7 * rng.standard_normal(shape)).astype(F32) for _ in range(3)]
    sigma, tau, lt, theta = F32(0.9), F32(0.05), F32(0.37), F32(1.0)
    U1, P1 = iterate_plain(inp, U, P, sigma, tau, lt, theta, nonneg, aniso)
    U2, P2 = iterate_plain(inp, U1, P1, sigma, tau, lt, theta, nonneg, aniso)
    Uo, Q, stores = emulate(inp, U, P, sigma, tau, lt, theta, nonneg, aniso, zrun)
    ok = np.array_equal(stores, np.ones_like(stores)) and np.array_equal(Uo, U2)
    ok = ok and all(np.array_equal(Q[c], P2[c]) for c in range(3))
    print(f"shape={shape} zrun={zrun} nonneg={nonneg} aniso={aniso}: "
          f"{'OK' if ok else 'MISMATCH'}  (stores min/max {stores.min()}/{stores.max()}, "
          f"bad U {np.count_nonzero(Uo != U2)}, bad P {[int(np.count_nonzero(Q[c] != P2[c])) for c in range(3)]})")
    if not ok:
        bad = np.argwhere(Uo != U2)
        print("   first bad U voxels (z, y, x):", bad[:8].tolist())
    return ok


def main():
    ok = True
    ok &= run_case((2, 3, 8), 2, False, False, 0)
    ok &= run_case((5, 9, 124), 5, True, False, 1)
    ok &= run_case((7, 18, 132), 3, False, False, 2)
    ok &= run_case((9, 21, 244), 4, True, True, 3)
    ok &= run_case((6, 16, 120), 2, False, False, 4)
    ok &= run_case((4, 5, 4), 1, False, False, 5)
    ok &= run_case((3, 2, 12), 3, False, False, 6)
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
