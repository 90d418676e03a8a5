"""CPU emulation of rvl_log_pos / rvl_exp_neg (revealer_b200/csrc/rvl_common.cuh): the branch-free log and
exp the score kernel uses.  Same bit manipulations and operation order.  numpy float64 has no FMA: the one
FMA whose single rounding matters (r = fma(m, rc, -1)) is emulated through an 80-bit product, the others
only make this emulation slightly *less* accurate than the device code.  The log must stay within 2 ulp
of the correctly rounded value for |log q| > 0.5 and within 1e-16 absolute below."""
import numpy as np

LD = np.longdouble


def rvl_log_table():
    """(rc_i, -log rc_i) as rvl_upload_log_table (score.cu) builds them."""
    idx = np.arange(256, dtype=np.uint64)
    a_hi = np.uint64(0x3FE6A09E) + idx * np.uint64(4096)
    a = (a_hi << np.uint64(32)).view(np.float64)
    b = ((a_hi + np.uint64(4096)) << np.uint64(32)).view(np.float64)
    rc = 1.0 / (0.5 * (a + b))
    return rc, (-np.log(rc.astype(LD))).astype(np.float64)


_RC, _LC = rvl_log_table()


def rvl_log_pos(q):
    bits = q.view(np.uint64)
    hi = (bits >> np.uint64(32)).astype(np.int64)
    lo = bits & np.uint64(0xFFFFFFFF)
    hi = hi + (0x3FF00000 - 0x3FE6A09E)
    k = (hi >> 20) - 0x3FF
    off = hi & 0x000FFFFF
    m = (((off + 0x3FE6A09E).astype(np.uint64) << np.uint64(32)) | lo).view(np.float64)
    rc, lc = _RC[off >> 12], _LC[off >> 12]
    r = (m.astype(LD) * rc.astype(LD) - 1).astype(np.float64)
    r2 = r * r
    pa = r * (1.0 / 3) + (-0.5)
    pb = r * (1.0 / 5) + (-0.25)
    p = pb * r2 + pa
    lm = r2 * p + r
    return (k * 0.6931471805599453094 + lc) + lm


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
                        np.exp(rng.uniform(-0.8, 0.8, 500_000)),
                        np.array([1.0, 2.0, 0.5, np.sqrt(0.5), np.sqrt(2.0), 2.220446049250313e-16, 1e-290])])
    got, ref = rvl_log_pos(q), np.log(q.astype(LD))
    err = np.abs(got.astype(LD) - ref).astype(np.float64)
    refd = ref.astype(np.float64)
    ulp = np.spacing(np.maximum(np.abs(refd), 1e-300))
    big = np.abs(refd) > 0.5
    assert np.max(err[big] / ulp[big]) <= 2.0
    assert np.max(err[~big]) <= 1e-16


def test_log_table_intervals():
    """|r| = |m * rc_i - 1| <= 2^-9 over every sub-interval (the bound the degree-5 series relies on)."""
    idx = np.arange(256, dtype=np.uint64)
    a_hi = np.uint64(0x3FE6A09E) + idx * np.uint64(4096)
    a = (a_hi << np.uint64(32)).view(np.float64)
    b = ((a_hi + np.uint64(4096)) << np.uint64(32)).view(np.float64)
    assert np.all(np.abs(a * _RC - 1) <= 2.0 ** -9) and np.all(np.abs(b * _RC - 1) <= 2.0 ** -9)
    assert a[0] < np.sqrt(0.5) < a[1] and b[-1] == 2.0 * a[0]  # the fold covers exactly one octave
