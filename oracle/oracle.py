"""ctypes bridge to the CPU oracle (oracle/revealer_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from revealer_b200/.  PARITY UNPINNED (see the C
file's header): this restates MASS::bcv/kde2d, misc3d::kde3d and stats::optimize/cor/var.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class MI(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("MI", "SMI", "HXY", "HX", "HY", "NMI", "IC")]


class CMI(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("CMI", "MI", "SCMI", "SMI", "HXY", "HXYZ", "IC", "CIC")]


def build(force=False, name="liboracle.so"):
    so = os.path.join(_HERE, name)
    srcs = [os.path.join(_HERE, f) for f in ("revealer_oracle.c", "oracle_hp.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.orc_var.restype = C.c_double
        L.orc_var.argtypes = [dp, C.c_int]
        L.orc_cor.restype = C.c_double
        L.orc_cor.argtypes = [dp, dp, C.c_int]
        L.orc_bcv.restype = C.c_double
        L.orc_bcv.argtypes = [dp, C.c_int]
        L.orc_assoc.restype = C.c_double
        L.orc_assoc.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_cond_assoc.restype = C.c_double
        L.orc_cond_assoc.argtypes = [dp, dp, dp, C.c_int, C.c_int]
        L.orc_mutual_inf_v2.restype = C.c_int
        L.orc_mutual_inf_v2.argtypes = [dp, dp, C.c_int, C.c_int, dp, C.POINTER(MI)]
        L.orc_cond_mutual_inf.restype = C.c_int
        L.orc_cond_mutual_inf.argtypes = [dp, dp, dp, C.c_int, C.c_int, dp, C.POINTER(CMI)]
        L.orc_score_rows.restype = None
        L.orc_score_rows.argtypes = [dp, dp, C.c_int64, C.c_int, dp, C.c_int, C.c_int, C.c_int, dp]
        L.orc_top1.restype = C.c_int64
        L.orc_top1.argtypes = [dp, C.c_int64, C.c_int]
        L.orc_order.restype = None
        L.orc_order.argtypes = [dp, C.c_int64, C.c_int, ip]
        L.orc_search.restype = None
        L.orc_search.argtypes = [dp, dp, C.c_int64, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, dp, ip, dp, dp, dp]
        L.orc_pvals.restype = None
        L.orc_pvals.argtypes = [dp, C.c_int64, dp, C.c_int64, dp]
        L.orc_max_threads.restype = C.c_int
        L.orc_score_rows_info.restype = None
        L.orc_score_rows_info.argtypes = [dp, dp, C.c_int64, C.c_int, dp, C.c_int, C.c_int, dp, dp]
        _LIB = L
    return _LIB


_HP = {}


def hp_lib(quad=False):
    """The extended-precision evaluation of the same estimators (oracle_hp.c): long double, or
    __float128 with quad=True.  Noise-band measurements only."""
    if quad not in _HP:
        L = C.CDLL(build(name="liboracle_q.so" if quad else "liboracle_hp.so"))
        dp = C.POINTER(C.c_double)
        L.orchp_bits.restype = C.c_int
        L.orchp_score_rows.restype = None
        L.orchp_score_rows.argtypes = [dp, dp, C.c_int64, C.c_int, dp, C.c_int, C.c_int, dp, dp]
        _HP[quad] = L
    return _HP[quad]


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def var(x):
    x, px = _d(x)
    return lib().orc_var(px, len(x))


def cor(x, y):
    x, px = _d(x)
    y, py = _d(y)
    return lib().orc_cor(px, py, len(x))


def bcv(x):
    x, px = _d(x)
    return lib().orc_bcv(px, len(x))


def assoc(x, y, n_grid=25):
    x, px = _d(x)
    y, py = _d(y)
    return lib().orc_assoc(px, py, len(x), n_grid)


def cond_assoc(x, y, z, n_grid=25):
    x, px = _d(x)
    y, py = _d(y)
    z, pz = _d(z)
    return lib().orc_cond_assoc(px, py, pz, len(x), n_grid)


def mutual_inf_v2(x, y, n_grid=25, delta=None):
    x, px = _d(x)
    y, py = _d(y)
    out = MI()
    pd = None
    if delta is not None:
        delta, pd = _d(delta)
    lib().orc_mutual_inf_v2(px, py, len(x), n_grid, pd, C.byref(out))
    return {k: getattr(out, k) for k, _ in MI._fields_}


def cond_mutual_inf(x, y, z, n_grid=25, delta=None):
    x, px = _d(x)
    y, py = _d(y)
    z, pz = _d(z)
    out = CMI()
    pd = None
    if delta is not None:
        delta, pd = _d(delta)
    lib().orc_cond_mutual_inf(px, py, pz, len(x), n_grid, pd, C.byref(out))
    return {k: getattr(out, k) for k, _ in CMI._fields_}


def score_rows(target, m2_rows, seed, n_grid=25, hoist=True, nthreads=0):
    """m2_rows: F x n (row k = m.2[k,]).  Returns F scores = cond.assoc(target, m.2[k,], seed)."""
    t, pt = _d(target)
    m, pm = _d(m2_rows)
    s, ps = _d(seed)
    F, n = m.shape
    out = np.empty(F, dtype=np.float64)
    lib().orc_score_rows(pt, pm, F, n, ps, n_grid, int(hoist), nthreads,
                         out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def score_rows_info(target, m2_rows, seed, n_grid=25, nthreads=0):
    """(scores, info): info[k] = the MI / CMI score k was formed from (double-precision oracle, hoisted bcv)."""
    t, pt = _d(target)
    m, pm = _d(m2_rows)
    s, ps = _d(seed)
    F, n = m.shape
    out, info = np.empty(F), np.empty(F)
    dp = C.POINTER(C.c_double)
    lib().orc_score_rows_info(pt, pm, F, n, ps, n_grid, nthreads, out.ctypes.data_as(dp), info.ctypes.data_as(dp))
    return out, info


def score_rows_hp(target, m2_rows, seed, n_grid=25, nthreads=0, quad=False):
    """(scores, info) from the extended-precision evaluation (same inputs, same bcv bandwidths)."""
    t, pt = _d(target)
    m, pm = _d(m2_rows)
    s, ps = _d(seed)
    F, n = m.shape
    out, info = np.empty(F), np.empty(F)
    dp = C.POINTER(C.c_double)
    hp_lib(quad).orchp_score_rows(pt, pm, F, n, ps, n_grid, nthreads, out.ctypes.data_as(dp), info.ctypes.data_as(dp))
    return out, info


def top1(v, negative=False):
    v, pv = _d(v)
    return lib().orc_top1(pv, len(v), int(negative))


def order(v, negative=False):
    v, pv = _d(v)
    idx = np.empty(len(v), dtype=np.int64)
    lib().orc_order(pv, len(v), int(negative), idx.ctypes.data_as(C.POINTER(C.c_int64)))
    return idx


def search(target, m2_rows, seed, iters, negative=False, n_grid=25, hoist=True, nthreads=0):
    t, pt = _d(target)
    m, pm = _d(m2_rows)
    s = np.array(seed, dtype=np.float64, copy=True)
    F, n = m.shape
    cmi = np.empty((iters, F), dtype=np.float64)  # cmi[it, k]  (R: cmi[k, it], column-major)
    top = np.empty(iters, dtype=np.int64)
    seed_iter = np.empty((iters, n), dtype=np.float64)
    cmi_seed = np.empty(iters, dtype=np.float64)
    ic0 = C.c_double()
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
    lib().orc_search(pt, pm, F, n, s.ctypes.data_as(dp), iters, int(negative), n_grid, int(hoist),
                     nthreads, cmi.ctypes.data_as(dp), top.ctypes.data_as(ip),
                     seed_iter.ctypes.data_as(dp), cmi_seed.ctypes.data_as(dp), C.byref(ic0))
    return dict(cmi=cmi, top=top, seed_iter=seed_iter, cmi_seed=cmi_seed, ic_seed0=ic0.value)


def pvals(cmi, cmi_rand):
    c, pc = _d(cmi)
    r, pr = _d(np.ravel(cmi_rand))
    out = np.empty(len(c), dtype=np.float64)
    lib().orc_pvals(pc, len(c), pr, len(r), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def max_threads():
    return lib().orc_max_threads()
