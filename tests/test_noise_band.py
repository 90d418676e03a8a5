"""The rounding-noise band of the reference's own double arithmetic, and the GPU's position inside it.

CPU part (`-m "not gpu"`): the double-precision oracle against the __float128 evaluation of the same
formulas from the same inputs and bandwidths (oracle/oracle_hp.c).  |I_double - I_quad| (I = MI or CMI) is
what any double evaluation of LIB:2524-2614 -- R's included -- cannot resolve; tests/parity_tol.py derives
KAPPA from the figure printed here.  (The long double build is shown too: R accumulates `sum()` in long double,
which does not help -- the noise is in the rounded P log P terms, not in the accumulator.)

GPU part: the CUDA scores against the same __float128 values; the GPU must sit inside KAPPA as well, i.e. be
as close to the exact value of the reference's formulas as the reference's own arithmetic is, up to the factor
parity_tol.py states.
"""
import numpy as np
import pytest

from parity_tol import KAPPA, error_report, info_of, score_close

CASES = [  # (n, F, null_seed): CIC at three sample counts, IC (NULLSEED) at one
    (82, 96, False), (500, 64, False), (1000, 48, False), (257, 64, True)]


def _case(n, F, null_seed):
    from revealer_b200 import synth
    w = synth.make_workload(n=n, F=F, null_seed=null_seed)
    return w["target"], w["m2"], w["seed"]


@pytest.fixture(scope="module")
def quad_refs(oracle):
    out = {}
    for n, F, ns in CASES:
        x, m2, seed = _case(n, F, ns)
        out[(n, F, ns)] = oracle.score_rows_hp(x, m2.astype(float), seed.astype(float), quad=True)
    return out


def test_double_oracle_noise_band_against_float128(oracle, quad_refs):
    worst = 0.0
    for n, F, ns in CASES:
        x, m2, seed = _case(n, F, ns)
        sd, idb = oracle.score_rows_info(x, m2.astype(float), seed.astype(float))
        sl, il = oracle.score_rows_hp(x, m2.astype(float), seed.astype(float), quad=False)
        sq, iq = quad_refs[(n, F, ns)]
        band = float(np.abs(idb - iq).max())
        band_ld = float(np.abs(il - iq).max())
        rep = error_report(sd, sq, iq)
        print("n=%d F=%d nullseed=%s: |I_double - I_quad| max %.3g  (long double: %.3g)  %s" % (n, F, ns, band, band_ld, rep))
        worst = max(worst, band)
        ok, err = score_close(sd, sq)
        assert ok, err
    # the figure KAPPA is derived from (KAPPA = 15 x 3.4e-15): KAPPA must stay >= 10 x the measured band
    assert worst <= KAPPA / 10, worst
    assert worst >= 1e-17          # a band of exactly 0 would mean the two evaluations are not independent


@pytest.mark.gpu
def test_gpu_scores_sit_inside_the_reference_noise_band(oracle, quad_refs):
    import revealer_b200 as rb
    for n, F, ns in CASES:
        x, m2, seed = _case(n, F, ns)
        with rb.RevealerContext(0) as ctx:
            ctx.load_matrix(m2)
            ctx.set_targets(x)
            ctx.set_seed(seed)
            g = ctx.score_once()[0]
            ctx.set_dense_x(True)          # the literal form (no x-axis table, no factorised columns)
            gl = ctx.score_once()[0]
        sq, iq = quad_refs[(n, F, ns)]
        for name, v in (("production", g), ("literal", gl)):
            rep = error_report(v, sq, iq)
            print("GPU %s n=%d F=%d nullseed=%s vs float128: %s" % (name, n, F, ns, rep))
            assert rep["max_abs_dI"] <= KAPPA, (name, n, rep)
            ok, err = score_close(v, sq)
            assert ok, (name, n, err)
        big = np.abs(sq) > 1e-3
        assert np.all(np.abs(g - sq)[big] <= 1e-9 * np.abs(sq[big]))
