"""numpy prototype of the cell-regime entropy epilogue (rvl_cmi_epilogue_c): validates the closed forms and the
direct-item bookkeeping against the literal sum over the 25^3 cube, before the CUDA version is written."""
import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O

T_HI, T_LO = 16.0, 0.5
TWO_PI = 2 * np.pi
EPS = 2.220446049250313e-16


def setup(x, y, z, G=25):
    n = len(x)
    bx, by, bz = O.bcv(x), O.bcv(y), O.bcv(z)
    rho = O.cor(x, y)
    c = 1 - 0.75 * max(rho, 0)
    hx, hy, hz = 0.25 * bx * c, 0.25 * by * c, 0.25 * bz * c
    dx = (x.max() - x.min()) / (G - 1)
    u = (x - x.min()) / dx
    beta = 0.5 * (dx / hx) ** 2
    A = np.zeros((G, 4))
    for a in (0, 1):
        for b in (0, 1):
            m = (y == a) & (z == b)
            A[:, a * 2 + b] = np.exp(-beta * (np.arange(G)[:, None] - u[m][None, :]) ** 2).sum(axis=1)
    t = np.arange(G) / (G - 1)
    g0y = np.exp(-0.5 * (t / hy) ** 2)
    g0z = np.exp(-0.5 * (t / hz) ** 2)
    epsq = EPS * (n * TWO_PI * np.sqrt(TWO_PI) * hx * hy * hz)
    am = int(y.sum() > n - y.sum())
    return A, g0y, g0z, epsq, hy, am, rho


def literal(A, g0y, g0z, epsq, am, dtype=np.longdouble):
    G = len(g0y)
    A = A.astype(dtype); gy = np.stack([g0y, g0y[::-1]]).astype(dtype); gz = np.stack([g0z, g0z[::-1]]).astype(dtype)
    e = dtype(epsq)
    D = np.zeros((G, G, G), dtype)
    for a in (0, 1):
        for b in (0, 1):
            D += A[:, a * 2 + b][:, None, None] * gy[a][None, :, None] * gz[b][None, None, :]
    Q = D + e
    S = Q.sum()
    L = lambda q: (q * np.log(q)).sum()
    return float((L(Q) - L(Q.sum(axis=1)) - L(Q.sum(axis=0)) + L(Q.sum(axis=(0, 1)))) / S), Q


def cells(XM, Xm, g0, LG, E, stats):
    """Sum over cells and over j of Q log Q, Q = g0[dM] XM + g0[dm] Xm + E, dM + dm = G-1 (flat arrays XM, Xm)."""
    G = len(g0)
    PG, PGL, PL = np.cumsum(g0), np.cumsum(g0 * LG), np.cumsum(LG)
    SG1 = np.append(np.cumsum(g0[::-1])[::-1], 0.0)
    SG2 = np.append(np.cumsum((g0 ** 2)[::-1])[::-1], 0.0)
    SG3 = np.append(np.cumsum((g0 ** 3)[::-1])[::-1], 0.0)
    kap = -LG[1]
    inv = np.float32(1.0 / kap)
    lE = np.log(E)
    Lhi, Llo, Ld = np.log(T_HI * E), np.log(T_LO * E), np.log(2.0 ** -54 * E)
    tot = 0.0
    ndead = 0
    items = 0
    for XMv, Xmv in zip(XM, Xm):
        lam = [np.log(max(XMv, 1e-290)), np.log(max(Xmv, 1e-290))]
        X = [XMv, Xmv]
        a1, a2, a3 = [], [], []
        for c in (0, 1):
            def bound(L):
                xb = np.float32(lam[c] - L) * inv
                return -1 if xb < 0 else min(G - 1, int(np.sqrt(np.float32(xb))))
            a1.append(bound(Lhi)); a2.append(bound(Llo)); a3.append(bound(Ld))
        mask = (1 << G) - 1           # bit j' (M-distance) set = direct
        def rng(lo, hi):
            return (((1 << (hi + 1)) - 1) & ~((1 << lo) - 1)) if lo <= hi and hi >= 0 else 0
        for c in (0, 1):
            lim = G - 2 - a3[1 - c]
            b = min(a1[c], lim)
            if b >= 0:
                tot += X[c] * (PGL[b] + lam[c] * PG[b]) + E * (PL[b] + (b + 1) * (lam[c] + 1))
                r = rng(0, b)
                mask &= ~(r if c == 0 else int(bin(r)[2:].zfill(G)[::-1][:G][::-1], 2) and sum(1 << (G - 1 - j) for j in range(G) if (r >> j) & 1))
            s1, s2 = a2[c] + 1, min(a3[c], lim)
            if s1 <= s2:
                S1, S2, S3 = SG1[s1] - SG1[s2 + 1], SG2[s1] - SG2[s2 + 1], SG3[s1] - SG3[s2 + 1]
                v = X[c]
                tot += v * S1 * (lE + 1) + v * v * S2 / (2 * E) - v * v * v * S3 / (6 * E * E) + (s2 - s1 + 1) * E * lE
                r = rng(s1, s2)
                mask &= ~(r if c == 0 else sum(1 << (G - 1 - j) for j in range(G) if (r >> j) & 1))
        lo, hi = a3[0] + 1, G - 2 - a3[1]
        if lo <= hi:
            ndead += hi - lo + 1
            mask &= ~rng(lo, hi)
        for j in range(G):
            if (mask >> j) & 1:
                q = g0[j] * XMv + g0[G - 1 - j] * Xmv + E
                tot += q * np.log(q)
                items += 1
    stats["items"] = stats.get("items", 0) + items
    stats["dead"] = stats.get("dead", 0) + ndead
    return tot + ndead * E * lE


def regime(A, g0y, g0z, epsq, hy, am):
    G = len(g0y)
    gz = np.stack([g0z, g0z[::-1]])
    M0, M1, m0, m1 = am * 2, am * 2 + 1, (1 - am) * 2, (1 - am) * 2 + 1
    V = A[:, M0][:, None] * gz[0][None, :] + A[:, M1][:, None] * gz[1][None, :]       # [i][k]
    W = A[:, m0][:, None] * gz[0][None, :] + A[:, m1][:, None] * gz[1][None, :]
    t = np.arange(G) / (G - 1) / hy
    LG = -0.5 * t * t
    stats = {}
    L3 = cells(V.ravel(), W.ravel(), g0y, LG, epsq, stats)
    sa = A.sum(axis=0)
    Gy, Gz = g0y.sum(), g0z.sum()
    gE = G * epsq
    sM = sa[M0] * gz[0] + sa[M1] * gz[1]
    sm = sa[m0] * gz[0] + sa[m1] * gz[1]
    st2 = {}
    Lyz = cells(sM, sm, g0y, LG, gE, st2)
    qxz = Gy * (V + W) + gE
    Lxz = (qxz * np.log(qxz)).sum()
    qz = (sa[0] + sa[2]) * Gy * gz[0] + (sa[1] + sa[3]) * Gy * gz[1] + G * G * epsq
    Lz = (qz * np.log(qz)).sum()
    Stot = sa.sum() * Gy * Gz + G ** 3 * epsq
    return (L3 - Lxz - Lyz + Lz) / Stot, stats, st2


if __name__ == "__main__":
    from revealer_b200 import synth
    worst = 0
    for n, F in ((1000, 60), (500, 40), (82, 40), (200, 40)):
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
            got, st, st2 = regime(A, g0y, g0z, epsq, hy, am)
            worst = max(worst, abs(got - ref))
            print(n, int(y.sum()), "cmi %.6e err %.2e items %d dead %d | yz items %d" % (ref, got - ref, st["items"], st["dead"], st2["items"]))
    print("worst |dCMI|", worst)


def diag(n=200, row=3):
    from revealer_b200 import synth
    w = synth.make_workload(n=n, F=40)
    x, z = w["target"], w["seed"].astype(float)
    y = w["m2"][row].astype(float)
    A, g0y, g0z, epsq, hy, am, rho = setup(x, y, z)
    G = 25
    gz = np.stack([g0z, g0z[::-1]])
    M0, M1, m0, m1 = am * 2, am * 2 + 1, (1 - am) * 2, (1 - am) * 2 + 1
    V = (A[:, M0][:, None] * gz[0][None, :] + A[:, M1][:, None] * gz[1][None, :]).ravel()
    W = (A[:, m0][:, None] * gz[0][None, :] + A[:, m1][:, None] * gz[1][None, :]).ravel()
    ld = np.longdouble
    D = g0y[None, :].astype(ld) * V[:, None].astype(ld) + g0y[::-1][None, :].astype(ld) * W[:, None].astype(ld)
    Q = D + ld(epsq)
    exact = Q * np.log(Q)
    t = D / ld(epsq)
    big = t >= T_HI
    small = t <= T_LO
    first = D * np.log(D) + ld(epsq) * (np.log(D) + 1)
    S = (A.sum() * g0y.sum() * g0z.sum() + G ** 3 * epsq)
    print("epsq", epsq, "S", S, "eps/S", epsq / S)
    print("big cells", big.sum(), "first-order err sum / S", float((exact - first)[big].sum() / S))
    sm3 = ld(epsq) * np.log(ld(epsq)) + D * (np.log(ld(epsq)) + 1) + D * D / (2 * ld(epsq)) - D ** 3 / (6 * ld(epsq) ** 2)
    print("small cells", small.sum(), "third-order err sum / S", float((exact - sm3)[small].sum() / S))


if len(sys.argv) > 1 and sys.argv[1] == "diag":
    diag()
