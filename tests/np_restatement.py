"""Second, independent restatement (numpy/scipy, vectorised) of the reference arithmetic, used only to
guard the C oracle (SURVEY 8c mitigation 1).  Written from the R call sites, not from the C file:
  mutual.inf.v2  LIB:2524-2561, cond.mutual.inf LIB:2564-2614, MASS::bcv / kde2d, misc3d::kde3d.
"""
import numpy as np
from scipy import optimize, stats


def bcv_objective(x, nb=1000):
    x = np.asarray(x, float)
    n = len(x)
    d = (x.max() - x.min()) * 1.01 / nb
    b = np.trunc(x / d).astype(np.int64)
    diff = np.abs(b[:, None] - b[None, :])[np.triu_indices(n, 1)]
    cnt = np.bincount(diff, minlength=nb)[:nb].astype(float)
    idx = np.arange(nb)

    def f(h):
        hh = h / 4.0
        delta = (idx * d / hh) ** 2
        keep = np.cumprod(delta < 1000).astype(bool)
        term = np.exp(-delta[keep] / 4) * (delta[keep] ** 2 - 12 * delta[keep] + 12)
        s = np.sum(term * cnt[keep])
        return 1 / (2 * n * hh * np.sqrt(np.pi)) + s / (64.0 * n * n * hh * np.sqrt(np.pi))
    hmax = 1.144 * np.sqrt(np.var(x, ddof=1)) * n ** (-0.2) * 4
    return f, 0.1 * hmax, hmax


def bcv_scipy(x):
    """bcv with scipy's fminbound (same Forsythe-Malcolm-Moler routine as R's optimize)."""
    f, lo, hi = bcv_objective(x)
    return optimize.fminbound(f, lo, hi, xtol=0.1 * lo)


def mutual_inf_v2(x, y, delta, n_grid=25):
    x, y = np.asarray(x, float), np.asarray(y, float)
    rho = np.corrcoef(x, y)[0, 1]
    h = np.asarray(delta, float) * (1 - 0.75 * abs(rho)) / 4.0
    gx = np.linspace(x.min(), x.max(), n_grid)
    gy = np.linspace(y.min(), y.max(), n_grid)
    ax = stats.norm.pdf((gx[:, None] - x[None, :]) / h[0])
    ay = stats.norm.pdf((gy[:, None] - y[None, :]) / h[1])
    z = ax @ ay.T / (len(x) * h[0] * h[1])
    F = z + np.finfo(float).eps
    dx, dy = gx[1] - gx[0], gy[1] - gy[0]
    P = F / (F.sum() * dx * dy)
    PX, PY = P.sum(1) * dy, P.sum(0) * dx
    MI = np.sum(P * np.log(P / np.outer(PX, PY))) * dx * dy
    return dict(MI=MI, IC=np.sign(rho) * np.sqrt(1 - np.exp(-2 * MI)), rho=rho)


def cond_mutual_inf(x, y, z, delta, n_grid=25):
    x, y, z = (np.asarray(v, float) for v in (x, y, z))
    rho = np.corrcoef(x, y)[0, 1]
    h = np.asarray(delta, float) * (1 - 0.75 * max(rho, 0.0))
    g = [np.linspace(v.min(), v.max(), n_grid) for v in (x, y, z)]
    m = [stats.norm.pdf(g[a][:, None], loc=v[None, :], scale=h[a]) for a, v in enumerate((x, y, z))]
    d = np.einsum("is,js,ks->ijk", m[0], m[1], m[2], optimize=True) / len(x)
    P = d + np.finfo(float).eps
    dx, dy, dz = (gg[1] - gg[0] for gg in g)
    P = P / (P.sum() * dx * dy * dz)
    PXZ, PYZ, PZ = P.sum(1) * dy, P.sum(0) * dx, P.sum((0, 1)) * dx * dy

    def H(p, vol):
        return -np.sum(p * np.log(p)) * vol
    CMI = H(PXZ, dx * dz) + H(PYZ, dy * dz) - H(P, dx * dy * dz) - H(PZ, dz)
    return dict(CMI=CMI, CIC=np.sign(rho) * np.sqrt(1 - np.exp(-2 * CMI)), rho=rho)
