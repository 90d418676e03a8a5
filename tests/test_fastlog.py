"""CPU emulation of rvl_log_pos (revealer_b200/csrc/rvl_common.cuh): the branch-free log the entropy
epilogue uses for positive normal doubles.  Same bit manipulations, same operation order (numpy float64
has no FMA, which only makes this emulation slightly *less* accurate than the device code); the MUFU
reciprocal seed is modelled by a float32 reciprocal (2^-24; the PTX rcp.approx.ftz.f64 guarantees 2^-23,
and two Newton steps square that twice).  Must stay within 4 ulp of numpy's log over the range of Q."""
import numpy as np


def rvl_log_pos(q):
    bits = q.view(np.uint64)
    hi = (bits >> np.uint64(32)).astype(np.int64)
    lo = bits & np.uint64(0xFFFFFFFF)
    hi = hi + (0x3FF00000 - 0x3FE6A09E)
    k = (hi >> 20) - 0x3FF
    hi = (hi & 0x000FFFFF) + 0x3FE6A09E
    m = ((hi.astype(np.uint64) << np.uint64(32)) | lo).view(np.float64)
    f, t = m - 1.0, m + 1.0
    r = (np.float32(1.0) / t.astype(np.float32)).astype(np.float64)
    e = 1.0 - t * r
    r = r + r * e
    e = 1.0 - t * r
    r = r + r * e
    s = f * r
    z = s * s
    lg = (6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01,
          1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01)
    p = np.full_like(z, lg[6])
    for c in lg[5::-1]:
        p = p * z + c
    return k * 0.6931471805599453094 + ((s * z) * p + (s + s))


def rvl_exp_neg(a):
    """Emulation of rvl_exp_neg (rvl_common.cuh): exp(a) for a <= 0."""
    import math
    magic = 6755399441055744.0
    t = a * 1.4426950408889634074 + magic
    k = (t.view(np.int64) & 0xFFFFFFFF).astype(np.int64)
    k = np.where(k >= 2 ** 31, k - 2 ** 32, k)
    kd = t - magic
    r = a - kd * 6.93147180369123816490e-01
    r = r - kd * 1.90821492927058770002e-10
    c = [1.0 / math.factorial(i) for i in range(13)]
    q = np.full_like(r, c[12])
    for i in range(11, 1, -1):
        q = q * r + c[i]
    p = (r * r) * q + r + 1.0
    return np.where(a < -707.0, 0.0, np.ldexp(p, k.astype(np.int32)))


def test_fast_exp_within_a_few_ulp():
    rng = np.random.default_rng(1)
    a = -np.concatenate([rng.uniform(0, 707, 1_000_000), rng.uniform(0, 3, 300_000), 10 ** rng.uniform(-12, 0, 100_000),
                         np.array([0.0, 706.999, 1e-300])])
    got, ref = rvl_exp_neg(a), np.exp(a)
    assert np.max(np.abs(got - ref) / ref) <= 4 * 2.2204e-16
    assert rvl_exp_neg(np.array([-708.0, -1e4]))[0] == 0.0 and rvl_exp_neg(np.array([0.0]))[0] == 1.0


def test_fast_log_within_a_few_ulp():
    rng = np.random.default_rng(0)
    q = np.concatenate([np.exp(rng.uniform(-200, 40, 1_000_000)), np.exp(rng.uniform(-1e-3, 1e-3, 200_000)),
                        np.array([1.0, 2.0, 0.5, np.sqrt(0.5), np.sqrt(2.0), 2.220446049250313e-16, 1e-290])])
    got, ref = rvl_log_pos(q), np.log(q)
    ulp = np.spacing(np.maximum(np.abs(ref), 1e-300))
    big = np.abs(ref) > 1e-3
    assert np.max(np.abs(got - ref)[big] / ulp[big]) <= 4.0
    assert np.max(np.abs(got - ref)[~big]) <= 1e-18
    assert rvl_log_pos(np.array([1.0]))[0] == 0.0
