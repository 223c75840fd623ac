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


def emulate(inp, U, P, sigma, tau, lt, theta, nonneg, aniso, zrun, below=None, above=None):
    """below / above: (inp, U, P) of the neighbouring z-shards, or None at the ends of the volume.  With a
    neighbour the kernel reads two ghost planes of U (one of P and Input) on that side straight from the
    neighbour's arrays -- plane -1 is the neighbour's last plane, plane dz its first."""
    dz, dy, dx = U.shape
    lo, hi = below is not None, above is not None

    def plane(name, comp, z):
        """Plane z of array `name` (\"in\", \"U\", \"P\") as the kernel addresses it."""
        src = {"in": lambda t: t[0], "U": lambda t: t[1], "P": lambda t: t[2][comp]}[name]
        if z < 0:
            assert lo and z >= -2 and (name != "in" or z == -1)
            return src(below)[z]           # python's negative index: -1 = last plane of the shard below
        if z >= dz:
            assert hi and z - dz <= (1 if name == "U" else 0)
            return src(above)[z - dz]
        return src((inp, U, P))[z]
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

        def ldv4(name, comp, z, k):
            return plane(name, comp, z)[rows[k]][cols].astype(F32)

        def load_packet(z, k):
            pk = {"un": ldv4("U", 0, z - 1 if (z == dz - 1 and not hi) else z + 1, k)}
            if k <= S + 2:
                pk["p1"], pk["p2"], pk["p3"] = ldv4("P", 0, z, k), ldv4("P", 1, z, k), ldv4("P", 2, z, k)
                if k >= 1:
                    pk["in"] = ldv4("in", 0, max(z, -1), k)  # plane -2 of Input is never needed: clamped
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
        if lo:
            zs, zB0 = za - 2, za - 1  # planes below 0 exist: they are the neighbour's
        zlast = zb if hi else min(zb, dz - 1)  # last plane of iteration A (plane dz is the neighbour's)
        uc = [ldv4("U", 0, zs, k) for k in range(ROWS)]
        p3b = [np.zeros((32, 4), F32) for _ in range(S)]
        nxt = load_packet(zs, 0)
        zero4 = np.zeros((32, 4), F32)
        steps = list(range(zs, zlast + 1)) + ([dz] if (zb == dz and not hi) else [])  # + the tail step
        for z in steps:
            doA, doB, emit, hasz = z <= zlast, z - 1 >= zB0, z - 1 >= za, (z > 0 or lo)
            more = z + 1 <= zlast
            ua_dst = UA2 if (z == dz - 1 and not hi) else UA
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


def run_sharded_case(shape, cuts, zrun, nonneg, aniso, seed):
    """The volume cut into z-shards, every shard emulated on its own with ghost reads from its
    neighbours' arrays; the assembled result must equal two plain iterations of the whole volume."""
    rng = np.random.default_rng(seed)
    inp = rng.standard_normal(shape).astype(F32)
    U = (inp + 0.3 * rng.standard_normal(shape)).astype(F32)
    P = [(0.7 * rng.standard_normal(shape)).astype(F32) for _ in range(3)]
    sigma, tau, lt, theta = F32(0.9), F32(0.05), F32(0.37), F32(1.0)
    U1, P1 = iterate_plain(inp, U, P, sigma, tau, lt, theta, nonneg, aniso)
    U2, P2 = iterate_plain(inp, U1, P1, sigma, tau, lt, theta, nonneg, aniso)
    bounds = list(zip([0] + list(cuts), list(cuts) + [shape[0]]))
    shards = [(inp[a:b], U[a:b], [c[a:b] for c in P]) for a, b in bounds]
    outs, ok = [], True
    for i, sh in enumerate(shards):
        Uo, Q, stores = emulate(sh[0], sh[1], sh[2], sigma, tau, lt, theta, nonneg, aniso, zrun,
                                below=shards[i - 1] if i > 0 else None,
                                above=shards[i + 1] if i + 1 < len(shards) else None)
        ok = ok and np.array_equal(stores, np.ones_like(stores))
        outs.append((Uo, Q))
    Ug = np.concatenate([o[0] for o in outs], axis=0)
    Qg = [np.concatenate([o[1][c] for o in outs], axis=0) for c in range(3)]
    ok = ok and np.array_equal(Ug, U2) and all(np.array_equal(Qg[c], P2[c]) for c in range(3))
    print(f"sharded shape={shape} cuts={cuts} zrun={zrun} nonneg={nonneg} aniso={aniso}: {'OK' if ok else 'MISMATCH'} "
          f"(bad U {np.count_nonzero(Ug != U2)}, bad P {[int(np.count_nonzero(Qg[c] != P2[c])) for c in range(3)]})")
    return ok


def main():
    ok = True
    ok &= run_sharded_case((8, 9, 124), [4], 4, True, False, 10)
    ok &= run_sharded_case((9, 6, 12), [2, 5], 2, False, False, 11)
    ok &= run_sharded_case((10, 18, 132), [3, 7], 8, False, True, 12)
    ok &= run_sharded_case((6, 5, 8), [2, 4], 1, True, False, 13)
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
