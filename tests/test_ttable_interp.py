"""CPU check of the bandwidth-interpolation identity the score kernel relies on (score.cu, "x-axis
sums"): the per-class total  T(c)[i] = sum_s exp(-beta0 (i - u_s)^2 / c^2)  is reproduced by 8-point
Lagrange interpolation in tau = -2 ln c on nodes spaced 1/128 to ~1e-16 of the class size, for the
whole range c in [0.25, 1] of the reference's shrink factor (LIB:2571-2573) -- i.e. to rounding error.
Pure numpy restatement of rvl_tt_lookup / k_ttable; the CUDA side is checked against its own literal
mode in tests/test_gpu_parity.py::test_table_mode_equals_dense_mode."""
import numpy as np
import pytest

DT, K, PAD = 1.0 / 128, 8, 4


def lagrange_weights(x):
    offs = np.arange(K) - (K // 2 - 1)
    w = np.ones(K)
    for q in range(K):
        for r in range(K):
            if r != q:
                w[q] *= (x - offs[r]) / (offs[q] - offs[r])
    return offs, w


def lookup(f, c, nodes):
    pos = -2.0 * np.log(c) / DT
    fl = np.floor(pos)
    m0 = int(fl) + PAD - (K // 2 - 1)
    m0 = max(0, min(m0, nodes - K))
    x = pos - (m0 - PAD + (K // 2 - 1))
    offs, w = lagrange_weights(x)
    return sum(w[q] * f((m0 + q - PAD) * DT) for q in range(K))


@pytest.mark.parametrize("n,frac,beta0", [(82, 0.1, 0.5), (82, 0.9, 12.0), (1000, 0.02, 3.0), (1000, 0.98, 12.0),
                                          (10000, 0.01, 0.5), (10000, 0.5, 30.0)])
def test_interpolation_error_is_rounding_level(n, frac, beta0):
    rng = np.random.default_rng(n + int(beta0 * 10))
    u = np.sort(rng.standard_normal(n))
    u = (u - u.min()) / (u.max() - u.min()) * 24
    us = u[rng.random(n) < frac]
    d2 = (np.arange(25)[:, None] - us[None, :]) ** 2
    nodes = int(np.ceil(-2 * np.log(0.25) / DT)) + 2 * PAD + 1

    def f(tau):
        return np.exp(-beta0 * np.exp(tau) * d2).sum(axis=1)
    worst = 0.0
    for c in np.concatenate([rng.uniform(0.25, 1.0, 40), [0.25, 0.2500001, 0.9999999, 0.5]]):
        est = lookup(f, c, nodes)
        exact = f(-2 * np.log(c))
        worst = max(worst, float(np.max(np.abs(est - exact))) / max(1, len(us)))
    assert worst < 2e-15, worst


def test_weights_sum_to_one_and_reproduce_polynomials():
    for x in (0.0, 0.3, 0.999):
        offs, w = lagrange_weights(x)
        assert abs(w.sum() - 1) < 1e-14
        for deg in range(1, K):
            assert abs(np.sum(w * offs.astype(float) ** deg) - x ** deg) < 1e-9
