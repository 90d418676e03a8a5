"""numpy prototype of the soft-floor form of the direct two-term columns (rvl_cmi_epilogue_f, pass C): for a density
column in which only one y-class is visible, Q_i = g X_i + E and

    sum_i Q_i log Q_i = g (log g SX + LX) + E (G log g + LLX)          closed form from pass A's per-k sums
                        + E ln2 sum_i (1 + t_i) log2(1 + 1/t_i),      t_i = g X_i / E

The second line is O(E) per cell whatever t is, so FLOAT arithmetic carries it: with d = log2 t (from float copies of
log X and log g), a = |d|, e = 2^-a, P(e) = log2(1+e)/e (degree-3 minimax polynomial, relative error 3.9e-4):
    d > 0:   (1 + t) log2(1 + 1/t) = (1 + e) P(e)
    d <= 0:  (1 + t) log2(1 + 1/t) = (1 + e) (a + e P(e))
Checks the whole L3 = sum Q log Q of the cube against a long-double literal evaluation."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from epilogue_c import setup, literal

F32 = np.float32
PC = [F32(v) for v in (1.4421281814575195, -0.701747477054596, 0.3664810359477997, -0.10725461691617966)]
LOG2E = 1.4426950408889634


def corr_f32(lx2f, lg2f):
    """sum over cells of (1+t) log2(1+1/t) in float32; lx2f: float32 log2 X_i, lg2f: float32 log2(g/E)."""
    d = (lx2f + lg2f).astype(F32)
    a = np.abs(d)
    e = np.exp2(-a.astype(np.float64)).astype(F32)
    p = np.full_like(e, PC[-1])
    for c in PC[-2::-1]:
        p = (p * e + c).astype(F32)
    r = np.where(d > 0, p, (e * p + a).astype(F32)).astype(F32)
    v = ((F32(1) + e) * r).astype(F32)
    acc = F32(0)
    for x in v:
        acc = F32(acc + x)
    return float(acc)


def l3_softfloor(A, g0y, g0z, epsq, hy, am, theta=1e-14, stats=None):
    G = len(g0y)
    gz = np.stack([g0z, g0z[::-1]])
    M0, M1, m0, m1 = am * 2, am * 2 + 1, (1 - am) * 2, (1 - am) * 2 + 1
    V = A[:, M0][:, None] * gz[0][None, :] + A[:, M1][:, None] * gz[1][None, :]
    W = A[:, m0][:, None] * gz[0][None, :] + A[:, m1][:, None] * gz[1][None, :]
    t = np.arange(G) / (G - 1) / hy
    LG = -0.5 * t * t
    E, lE = epsq, np.log(epsq)
    Stot = A.sum() * g0y.sum() * g0z.sum() + G ** 3 * epsq
    dead_thr = max(E * 2.0 ** -55, theta * Stot / G ** 3)
    tot = 0.0
    st = dict(dead=0, fact=0, soft=0, four=0)
    for k in range(G):
        X = [V[:, k], W[:, k]]
        lX = [np.log(np.maximum(x, 1e-290)) for x in X]
        SX = [x.sum() for x in X]
        LX = [(x * l).sum() for x, l in zip(X, lX)]
        LLX = [l.sum() for l in lX]
        lXf = [(l * LOG2E).astype(F32) for l in lX]
        mx = [x.max() for x in X]
        mn = [x.min() for x in X]
        for jp in range(G):
            g = [g0y[jp], g0y[G - 1 - jp]]
            lg = [LG[jp], LG[G - 1 - jp]]
            vis = [g[c] * mx[c] >= 2.0 ** -54 * E for c in (0, 1)]
            alive = any(g[c] * mx[c] >= 0.5 * dead_thr for c in (0, 1))
            if not alive:
                tot += G * E * lE
                st["dead"] += 1
                continue
            if vis[0] and vis[1]:
                q = g[0] * X[0] + g[1] * X[1] + E
                tot += (q * np.log(q)).sum()
                st["four"] += 1
                continue
            c = 0 if vis[0] else 1
            if not vis[c]:   # alive but neither visible cannot happen (dead_thr >= 2^-55 E); keep exact
                tot += G * E * lE
                continue
            closed = g[c] * (lg[c] * SX[c] + LX[c]) + E * (G * lg[c] + LLX[c])
            if g[c] * mn[c] >= 16 * E:
                tot += closed + E * G
                st["fact"] += 1
            else:
                lgf = F32((lg[c] - lE) * LOG2E)
                tot += closed + E * np.log(2.0) * corr_f32(lXf[c], lgf)
                st["soft"] += 1
    if stats is not None:
        stats.update(st)
    return tot, Stot


def column_errors(A, g0y, g0z, epsq, hy, am, theta=1e-14):
    """Error of every column kind against a long-double evaluation of the same column, summed per kind, in units of S
    (i.e. as it enters CMI)."""
    ld = np.longdouble
    G = len(g0y)
    gz = np.stack([g0z, g0z[::-1]])
    M0, M1, m0, m1 = am * 2, am * 2 + 1, (1 - am) * 2, (1 - am) * 2 + 1
    V = A[:, M0][:, None] * gz[0][None, :] + A[:, M1][:, None] * gz[1][None, :]
    W = A[:, m0][:, None] * gz[0][None, :] + A[:, m1][:, None] * gz[1][None, :]
    t = np.arange(G) / (G - 1) / hy
    LG = -0.5 * t * t
    E, lE = epsq, np.log(epsq)
    Stot = A.sum() * g0y.sum() * g0z.sum() + G ** 3 * epsq
    dead_thr = max(E * 2.0 ** -55, theta * Stot / G ** 3)
    err = dict(dead=0.0, fact=0.0, soft=0.0, four=0.0)
    cnt = dict(dead=0, fact=0, soft=0, four=0)
    for k in range(G):
        X = [V[:, k], W[:, k]]
        lX = [np.log(np.maximum(x, 1e-290)) for x in X]
        SX = [x.sum() for x in X]
        LX = [(x * l).sum() for x, l in zip(X, lX)]
        LLX = [l.sum() for l in lX]
        lXf = [(l * LOG2E).astype(F32) for l in lX]
        mx = [x.max() for x in X]
        mn = [x.min() for x in X]
        for jp in range(G):
            g = [g0y[jp], g0y[G - 1 - jp]]
            lg = [LG[jp], LG[G - 1 - jp]]
            q = ld(g[0]) * X[0].astype(ld) + ld(g[1]) * X[1].astype(ld) + ld(E)
            exact = (q * np.log(q)).sum()
            vis = [g[c] * mx[c] >= 2.0 ** -54 * E for c in (0, 1)]
            alive = any(g[c] * mx[c] >= 0.5 * dead_thr for c in (0, 1))
            if not alive:
                kind, val = "dead", G * E * lE
            elif vis[0] and vis[1]:
                qq = g[0] * X[0] + g[1] * X[1] + E
                kind, val = "four", (qq * np.log(qq)).sum()
            else:
                c = 0 if vis[0] else 1
                closed = g[c] * (lg[c] * SX[c] + LX[c]) + E * (G * lg[c] + LLX[c])
                if g[c] * mn[c] >= 16 * E:
                    kind, val = "fact", closed + E * G
                else:
                    kind, val = "soft", closed + E * np.log(2.0) * corr_f32(lXf[c], F32((lg[c] - lE) * LOG2E))
            err[kind] += float(val - exact)
            cnt[kind] += 1
    return {k: v / Stot for k, v in err.items()}, cnt


if __name__ == "__main__":
    from revealer_b200 import synth
    worst = 0
    for n, F in ((1000, 40), (500, 20), (82, 20), (200, 20)):
        w = synth.make_workload(n=n, F=F)
        x, z = w["target"], w["seed"].astype(float)
        rng = np.random.default_rng(n)
        rows = [r for r in w["m2"]] + [(rng.random(n) < p).astype(np.uint8) for p in (0.1, 0.3, 0.6, 0.9)]
        for y in rows:
            y = y.astype(float)
            if y.sum() in (0, n):
                continue
            A, g0y, g0z, epsq, hy, am, rho = setup(x, y, z)
            ref, Q = literal(A, g0y, g0z, epsq, am)
            L3ref = float((Q * np.log(Q)).sum())
            st = {}
            L3, Stot = l3_softfloor(A, g0y, g0z, epsq, hy, am, stats=st)
            err = (L3 - L3ref) / Stot
            worst = max(worst, abs(err))
            print(n, int(y.sum()), "cmi %.4e dL3/S %.2e" % (ref, err), st)
    print("worst |dL3|/S", worst)
