"""CPU check of the algebra behind k_score's soft cells (revealer_b200/csrc/score.cu, rvl_soft_cell): for a density
column in which one y-class is visible,
    sum_i Q_i log Q_i = g (log g SX + LX) + E (G log g + LLX) + E ln2 sum_i (1 + t_i) log2(1 + 1/t_i),   t_i = g X_i / E,
with the last sum carried in float32 through the kernel's degree-3 polynomial.  The numpy restatement of that epilogue
(tools/proto/epilogue_softfloor.py) is held against a long-double evaluation of the literal 25^3 cube."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "proto"))


def test_polynomial_is_the_kernels_and_accurate():
    import epilogue_softfloor as E
    src = open(os.path.join(ROOT, "revealer_b200", "csrc", "score.cu")).read()
    for c in E.PC:                                   # the prototype and the kernel use the same coefficients
        assert ("%.17g" % abs(float(c)))[:12] in src, float(c)
    e = np.linspace(0.0, 1.0, 20001)[1:]
    p = np.full_like(e, float(E.PC[-1]))
    for c in E.PC[-2::-1]:
        p = p * e + float(c)
    rel = np.abs(p * e - np.log2(1 + e)) / np.log2(1 + e)
    assert rel.max() < 4.5e-4, rel.max()


def test_soft_floor_column_sums_match_long_double_columns():
    import epilogue_softfloor as E
    from epilogue_c import setup
    from revealer_b200 import synth
    kinds = {"soft": 0, "fact": 0, "four": 0, "dead": 0}
    worst = {"soft": 0.0, "fact": 0.0, "four": 0.0, "dead": 0.0}
    for n, F in ((200, 6), (82, 4)):
        w = synth.make_workload(n=n, F=F)
        x, z = w["target"], w["seed"].astype(float)
        rng = np.random.default_rng(n)
        rows = [r for r in w["m2"]] + [(rng.random(n) < p).astype(np.uint8) for p in (0.3, 0.9)]
        for y in rows:
            y = y.astype(float)
            if y.sum() in (0, n):
                continue
            A, g0y, g0z, epsq, hy, am, rho = setup(x, y, z)
            err, cnt = E.column_errors(A, g0y, g0z, epsq, hy, am)
            for k in kinds:
                kinds[k] += cnt[k]
                worst[k] = max(worst[k], abs(err[k]))
    assert kinds["soft"] > 500 and kinds["four"] > 100 and kinds["fact"] > 50, kinds   # every column kind was exercised
    # per evaluation, as the error enters CMI.  One soft column's correction is ~1e-16 .. 1e-14 of S, so a wrong or missing
    # correction term shows; the closed form + float32 correction is closer to the long-double column than the double logs.
    assert worst["soft"] < 5e-17, worst
    assert worst["fact"] < 5e-15 and worst["four"] < 5e-15 and worst["dead"] < 5e-14, worst
