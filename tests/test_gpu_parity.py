"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle on the
same seeded inputs.  The oracle restates R + MASS + misc3d (PARITY UNPINNED against real R; see
oracle/revealer_oracle.c).

Tolerances (north_star: scores within <= 1e-6 relative of the double-precision reference; selected
features and seed merges bit-exact):
  * IC / CIC scores: tests/parity_tol.py -- 1e-6 relative, or the underlying MI / CMI within KAPPA of the
    reference's (20x the rounding noise of the reference's own double arithmetic, measured against a
    __float128 evaluation in tests/test_noise_band.py).
  * winners, seed merges, tie-breaks, guards (exact 0.0): bit-exact.
  * bandwidths (bcv): 1e-12 relative (observed: bit-identical).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


from parity_tol import score_close  # noqa: E402  (tolerance and its justification: tests/parity_tol.py)


def _ctx(**kw):
    import revealer_b200 as rb
    return rb.RevealerContext(0, **kw)


def _workload(n, F, **kw):
    from revealer_b200 import synth
    return synth.make_workload(n=n, F=F, **kw)


# ------------------------------------------------------------------ bandwidths

@pytest.mark.parametrize("n", [20, 82, 200, 513])
def test_bandwidths_match_oracle(oracle, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    with _ctx() as ctx:
        ctx.set_targets(np.stack([x, np.sort(x), x * 3 + 1]))
        for t, v in enumerate((x, np.sort(x), x * 3 + 1)):
            assert abs(ctx.target_bandwidth(t) - oracle.bcv(v)) <= 1e-12 * oracle.bcv(v)
        tab = ctx.binary_bandwidth_table()
    assert tab[0] == 0.0 and tab[n] == 0.0
    for n1 in sorted({1, 2, 3, n // 7 + 1, n // 2, n - 2, n - 1}):
        y = np.zeros(n)
        y[:n1] = 1
        assert abs(tab[n1] - oracle.bcv(y)) <= 1e-12 * oracle.bcv(y), n1


# ------------------------------------------------------------------ scoring

@pytest.mark.parametrize("n,F", [(82, 300), (200, 400), (500, 300)])
def test_cic_scores_match_oracle(oracle, n, F):
    w = _workload(n, F)
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        g = ctx.score_once()[0]
    o = oracle.score_rows(w["target"], w["m2"].astype(float), w["seed"].astype(float))
    ok, err = score_close(g, o)
    assert ok, err


@pytest.mark.parametrize("n,F", [(82, 300), (257, 300)])
def test_ic_scores_match_oracle_nullseed(oracle, n, F):
    w = _workload(n, F, null_seed=True)
    assert w["seed"].sum() == 0
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        g = ctx.score_once()[0]
        ctx.set_seed(np.ones(n, dtype=np.uint8))        # constant seed of ones: also the IC path
        g1 = ctx.score_once()[0]
    o = oracle.score_rows(w["target"], w["m2"].astype(float), np.zeros(n))
    ok, err = score_close(g, o)
    assert ok, err
    assert np.array_equal(g, g1)


@pytest.mark.parametrize("n", [3, 5, 63, 64, 65, 127, 128, 129, 1000])
def test_ragged_sample_counts(oracle, n):
    rng = np.random.default_rng(n + 7)
    F = 40
    x = rng.standard_normal(n)
    m2 = (rng.random((F, n)) < 0.3).astype(np.uint8)
    seed = (rng.random(n) < 0.3).astype(np.uint8)
    seed[0], seed[-1] = 1, 0
    with _ctx() as ctx:
        ctx.load_matrix(m2)
        ctx.set_targets(x)
        ctx.set_seed(seed)
        g = ctx.score_once()[0]
    o = oracle.score_rows(x, m2.astype(float), seed.astype(float))
    ok, err = score_close(g, o)
    assert ok, (n, err)


def test_guards_are_exact_zero(oracle):
    n, F = 100, 8
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n)
    m2 = (rng.random((F, n)) < 0.2).astype(np.uint8)
    m2[2] = 0
    m2[5] = 1
    seed = (rng.random(n) < 0.2).astype(np.uint8)
    with _ctx() as ctx:
        ctx.load_matrix(m2)
        ctx.set_targets(np.stack([x, np.full(n, 1.5)]))   # second target is constant
        ctx.set_seed(seed)
        g = ctx.score_once()
    assert g[0, 2] == 0.0 and g[0, 5] == 0.0                # constant y -> 0      LIB:2507
    assert np.all(g[1] == 0.0)                              # constant x -> 0
    assert g[0, 0] != 0.0


def test_duplicate_rows_bit_identical_and_lowest_index_wins(oracle):
    w = _workload(150, 600)
    m2 = w["m2"].copy()
    # make the best row appear three times, far apart (different warps / blocks)
    with _ctx() as ctx:
        ctx.load_matrix(m2)
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        g = ctx.score_once()[0]
        best = int(np.nanargmax(g))
        m2[[17, 311, 599]] = m2[best]
        ctx.load_matrix(m2)
        g2 = ctx.score_once()[0]
        res = ctx.search(1)
    assert g2[17] == g2[311] == g2[599] == g2[best]
    assert res["top"][0, 0] == min(17, best)
    dup = np.flatnonzero((m2[1:] == m2[:-1]).all(axis=1)) + 1
    assert len(dup) > 20
    assert np.array_equal(g2[dup], g2[dup - 1])


# ------------------------------------------------------------------ search

@pytest.mark.parametrize("n,F,iters,match", [(82, 500, 3, "positive"), (200, 400, 4, "positive"),
                                             (120, 300, 3, "negative")])
def test_search_matches_oracle(oracle, n, F, iters, match):
    w = _workload(n, F)
    x = w["target"] if match == "positive" else w["target"][::-1].copy()
    m2 = w["m2"] if match == "positive" else w["m2"][:, ::-1].copy()
    seed = w["seed"] if match == "positive" else w["seed"][::-1].copy()
    neg = match == "negative"
    with _ctx(negative=neg) as ctx:
        ctx.load_matrix(m2)
        ctx.set_targets(x)
        ctx.set_seed(seed)
        res = ctx.search(iters)
    ref = oracle.search(x, m2.astype(float), seed.astype(float), iters, negative=neg)
    assert list(res["top"][0]) == list(ref["top"])                       # selected-feature sequence
    assert np.array_equal(res["seed_iter"][0], ref["seed_iter"].astype(np.uint8))   # seed merges
    ok, err = score_close(res["cmi"][0], ref["cmi"])
    assert ok, err
    ok, err = score_close(res["cmi_seed"][0], ref["cmi_seed"])
    assert ok, err
    ok, err = score_close(res["ic_seed0"], [ref["ic_seed0"]])
    assert ok, err
    assert np.array_equal(res["top_score"][0], res["cmi"][0][np.arange(iters), res["top"][0]])


def test_search_nullseed_first_iteration_uses_ic(oracle):
    w = _workload(100, 300, null_seed=True)
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        res = ctx.search(3)
    ref = oracle.search(w["target"], w["m2"].astype(float), np.zeros(100), 3)
    assert res["ic_seed0"][0] == 0.0                                       # assoc(target, zeros) = 0
    assert list(res["top"][0]) == list(ref["top"])
    ok, err = score_close(res["cmi"][0], ref["cmi"])
    assert ok, err


def test_reference_shaped_functions(oracle):
    import revealer_b200 as rb
    rng = np.random.default_rng(5)
    n = 90
    x = rng.standard_normal(n)
    y = (rng.random(n) < 0.2).astype(float)
    z = (rng.random(n) < 0.3).astype(float)
    assert abs(rb.assoc(x, y, "IC") - oracle.assoc(x, y)) <= 1e-9
    assert abs(rb.cond_assoc(x, y, z, "IC") - oracle.cond_assoc(x, y, z)) <= 1e-9
    assert rb.cond_assoc(x, np.zeros(n), z, "IC") == 0.0
    assert abs(rb.cond_assoc(x, y, np.zeros(n), "IC") - oracle.assoc(x, y)) <= 1e-9
    with pytest.raises(rb.RevealerError):
        rb.assoc(x, y, "COR")


def test_revealer_v1_entry_point_end_to_end(oracle):
    """REVEALER_v1 with the reference's formals on in-memory GCT-shaped inputs: preprocessing (host) +
    search (CUDA) against the oracle run on the same prepared inputs; seed baselines as LIB:1817-1833."""
    import revealer_b200 as rb
    from revealer_b200.api import prepare_inputs
    w = _workload(90, 260)
    rng = np.random.default_rng(6)
    perm = rng.permutation(90)                                     # unsorted samples: REVEALER.v1 sorts them
    names = ["S%03d" % i for i in range(90)]
    rows = ["F%03d" % i for i in range(260)]
    d1 = dict(ds=w["target"][perm][None, :], row_names=["TARGET"], names=[names[i] for i in perm])
    d2 = dict(ds=w["m2"][:, perm].astype(float), row_names=rows, names=[names[i] for i in perm])
    kw = dict(seed_names=["F023", "F047"], count_thres_low=3, count_thres_high=40)
    out = rb.REVEALER_v1(d1, "TARGET", "positive", d2, max_n_iter=3, **kw)
    p = prepare_inputs(d1, "TARGET", "positive", d2, **kw)
    ref = oracle.search(p["target"], p["m2"], p["seed"], 3)
    assert list(out["top"]) == list(ref["top"])
    assert out["seed_names_iter"] == [p["feature_names"][i] for i in ref["top"]]
    assert "F023" not in out["feature_names"] and "F047" not in out["feature_names"]
    ok, err = score_close(out["cmi"], ref["cmi"].T)
    assert ok, err
    ok, err = score_close(out["cmi_seed"], ref["cmi_seed"])
    assert ok, err
    ok, err = score_close(out["cmi_orig_seed"], [oracle.assoc(p["target"], v) for v in p["seed_vectors"]])
    assert ok, err
    med = np.median(p["target"])
    want_pct = ref["seed_iter"][:, p["target"] > med].sum(axis=1) / (p["target"] > med).sum()
    assert np.allclose(out["pct_explained_seed"], want_pct)


def test_oracle_golden_fixture_through_cuda(oracle):
    import revealer_b200 as rb
    g = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))
    for c in g["cases"]:
        x, y, z = (np.array(c[k], dtype=float) for k in ("x", "y", "z"))
        assert abs(rb.assoc(x, y) - c["IC"]) <= 1e-9
        assert abs(rb.cond_assoc(x, y, z) - c["CIC"]) <= 1e-9
        with rb.RevealerContext(0) as ctx:
            ctx.set_targets(x)
            assert abs(ctx.target_bandwidth(0) - c["bcv_x"]) <= 1e-12 * c["bcv_x"]
            tab = ctx.binary_bandwidth_table()
            assert abs(tab[int(y.sum())] - c["bcv_y"]) <= 1e-12 * c["bcv_y"]
    s = g["search"]
    out = rb.revealer_search(np.array(s["target"]), np.array(s["m2"], dtype=np.uint8), np.array(s["seed"]),
                             max_n_iter=s["iters"])
    assert list(out["top"]) == s["top"]
    assert np.allclose(out["cmi_seed"], s["cmi_seed"], rtol=1e-9, atol=0)
    assert np.allclose(out["top_score"], s["top_score"], rtol=1e-9, atol=0)


# ------------------------------------------------------------------ inputs and errors

def test_r_golden_vectors_through_cuda_if_present():
    """Real-R fixtures (tools/make_golden.R on a machine with R + MASS + misc3d) against the CUDA path itself:
    bandwidths, IC, CIC, and the 3-iteration search (winners and seeds exact, scores within tolerance).  The file
    cannot be produced in this environment (no R): until someone drops it in, parity with R stays unpinned and this
    test says so by skipping."""
    import revealer_b200 as rb
    path = os.path.join(HERE, "golden", "r_golden.json")
    if not os.path.exists(path):
        pytest.skip("no R-generated golden file (tests/golden/r_golden.json): parity with R stays unpinned")
    g = json.load(open(path))
    if g.get("provenance", "").startswith("STAND-IN") and not os.environ.get("RVL_ALLOW_STANDIN"):
        pytest.skip("r_golden.json is the oracle-made stand-in of tools/dryrun_r_golden.py, not R output: parity stays unpinned")
    for c in g["cases"]:
        x, y, z = (np.array(c[k], dtype=float) for k in ("x", "y", "z"))
        with _ctx() as ctx:
            ctx.set_targets(x)
            assert abs(ctx.target_bandwidth(0) - c["bcv_x"]) <= 1e-9 * c["bcv_x"]
            tab = ctx.binary_bandwidth_table()
            assert abs(tab[int(y.sum())] - c["bcv_y"]) <= 1e-9 * c["bcv_y"]
            assert abs(tab[int(z.sum())] - c["bcv_z"]) <= 1e-9 * c["bcv_z"]
        assert score_close([rb.assoc(x, y)], [c["IC"]])[0]
        assert score_close([rb.cond_assoc(x, y, z)], [c["CIC"]])[0]
    if "search" in g:
        s = g["search"]
        out = rb.revealer_search(np.array(s["target"], dtype=float), np.array(s["m2"], dtype=np.uint8),
                                 np.array(s["seed"], dtype=np.uint8), max_n_iter=int(s["iters"]))
        assert list(out["top"]) == [int(v) for v in s["top"]]
        assert np.array_equal(out["seed_iter"], np.array(s["seed_iter"], dtype=np.uint8))
        assert score_close(out["cmi"].T, np.array(s["cmi"], dtype=float))[0]
        assert score_close(out["cmi_seed"], np.array(s["cmi_seed"], dtype=float))[0]
        assert score_close([out["ic_seed0"]], [float(s["ic_seed0"])])[0]


def test_device_topk_is_r_order(oracle):
    """rvl_topk == the head of R's order(): stable among exact ties (duplicate rows), NaN last, -0 == +0, and the
    increasing order for target.match = "negative"."""
    import revealer_b200 as rb
    w = _workload(120, 700)
    for match in ("positive", "negative"):
        out = rb.revealer_search(w["target"], w["m2"], w["seed"], max_n_iter=2, target_match=match, n_markers=40)
        for it in range(2):
            want = oracle.order(out["cmi"][:, it], negative=(match == "negative"))[:40]
            assert np.array_equal(out["top_markers"][it], want), (match, it)
            assert out["top_markers"][it][0] == out["top"][it]
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"][:50])
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        ctx.search(1)
        rows, sc = ctx.topk(0, 1000)                                   # k clipped to the row count
        assert len(rows) == 50 and sorted(rows.tolist()) == list(range(50)) and np.all(np.diff(sc[~np.isnan(sc)]) <= 0)


def test_input_layouts_agree():
    import revealer_b200 as rb
    w = _workload(130, 200)
    outs = []
    with _ctx() as ctx:
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        ctx.load_matrix(w["m2"])                                   # uint8 row-major
        outs.append(ctx.score_once())
        assert np.array_equal(ctx.row_counts(), w["m2"].sum(axis=1))
        ctx.load_matrix(w["m2"].astype(np.float64))                # R layout: f64 column-major
        outs.append(ctx.score_once())
        bits, words = rb.pack_bits(w["m2"])
        ctx.load_matrix_bits(bits, 130)                            # bit-packed
        outs.append(ctx.score_once())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])


def test_invalid_inputs_raise_no_fallback():
    import revealer_b200 as rb
    from revealer_b200 import _lib
    w = _workload(64, 20)
    with _ctx() as ctx:
        bad = w["m2"].astype(np.float64)
        bad[3, 5] = 0.5
        with pytest.raises(rb.RevealerError) as e:
            ctx.load_matrix(bad)
        assert e.value.code == _lib.RVL_ERR_NOT_BINARY
        bad[3, 5] = np.nan
        with pytest.raises(rb.RevealerError):
            ctx.load_matrix(bad)
        bad8 = w["m2"].copy()
        bad8[0, 0] = 2
        with pytest.raises(rb.RevealerError):
            ctx.load_matrix(bad8)
        with pytest.raises(rb.RevealerError) as e:
            ctx.score_once()                                        # nothing loaded
        assert e.value.code == _lib.RVL_ERR_STATE
        ctx.load_matrix(w["m2"])
        t = w["target"].copy()
        t[4] = np.nan
        with pytest.raises(rb.RevealerError) as e:
            ctx.set_targets(t)
        assert e.value.code == _lib.RVL_ERR_NONFINITE
        with pytest.raises(rb.RevealerError):
            ctx.set_targets(w["target"][:-1])                       # length mismatch
        ctx.set_targets(w["target"])
        with pytest.raises(rb.RevealerError):
            ctx.set_seed(np.zeros(63, dtype=np.uint8))
        bits, words = rb.pack_bits(w["m2"])
        bits[0, -1] = 1                                             # padding bit set
        with pytest.raises(rb.RevealerError):
            ctx.load_matrix_bits(bits, 64)
        with pytest.raises(rb.RevealerError):
            ctx.set_options(n_grid=64)
    with pytest.raises(rb.RevealerError):
        rb.RevealerContext(99)                                      # no such device -> error, not CPU


def test_sample_count_limit_is_an_error_not_a_crash():
    import revealer_b200 as rb
    from revealer_b200 import _lib
    n = 1_600_000                                                # one bit row (n / 8 bytes) > the shared-memory budget of a k_score worker warp
    with _ctx() as ctx:
        with pytest.raises(rb.RevealerError) as e:
            ctx.load_matrix(np.zeros((2, n), dtype=np.uint8))
        assert e.value.code == _lib.RVL_ERR_ARG
        with pytest.raises(rb.RevealerError):
            ctx.set_targets(np.arange(n, dtype=np.float64))


def test_many_targets_and_single_row():
    """Degenerate shapes: one feature row against 300 targets; 2 samples."""
    rng = np.random.default_rng(31)
    with _ctx() as ctx:
        ctx.load_matrix((rng.random((1, 64)) < 0.3).astype(np.uint8))
        ctx.set_targets(rng.standard_normal((300, 64)))
        ctx.set_seed((rng.random(64) < 0.3).astype(np.uint8))
        s = ctx.score_once()
        assert s.shape == (300, 1) and np.isfinite(s).all()
        r = ctx.search(2)
        assert np.all(r["top"] == 0)
    with _ctx() as ctx:                                           # n = 2: y or z is constant or trivially split
        ctx.load_matrix(np.array([[0, 1], [1, 1], [0, 0]], dtype=np.uint8))
        ctx.set_targets(np.array([0.5, -1.0]))
        ctx.set_seed(np.array([1, 0], dtype=np.uint8))
        s = ctx.score_once()[0]
        assert s[1] == 0.0 and s[2] == 0.0 and np.isfinite(s[0]) or np.isnan(s[0])


def test_empty_matrix():
    w = _workload(64, 20)
    with _ctx() as ctx:
        ctx.load_matrix(np.zeros((0, 64), dtype=np.uint8))
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        assert ctx.score_once().shape == (1, 0)


def test_grid_and_bandwidth_options(oracle):
    import np_restatement as NP
    rng = np.random.default_rng(9)
    n = 80
    x = rng.standard_normal(n)
    y = (rng.random(n) < 0.2).astype(float)
    z = (rng.random(n) < 0.3).astype(float)
    import revealer_b200 as rb
    for G in (9, 25, 32):
        got = rb.cond_assoc(x, y, z, n_grid=G)
        delta = 0.25 * np.array([oracle.bcv(x), oracle.bcv(y), oracle.bcv(z)])
        want = NP.cond_mutual_inf(x, y, z, delta, n_grid=G)["CIC"]
        assert abs(got - want) <= 1e-8, (G, got, want)


@pytest.mark.parametrize("n,F,dense_rows", [(200, 600, False), (1000, 4000, False), (300, 200, True)])
def test_table_mode_equals_dense_mode(n, F, dense_rows):
    """The production x-axis path (bandwidth-interpolation table + minority-class visit) against the
    kernel's literal mode (every sample visited per feature), CIC and IC paths, sparse and dense rows."""
    w = _workload(n, F)
    m2 = w["m2"].copy()
    if dense_rows:                                         # dense rows exercise the complement visit
        rng = np.random.default_rng(4)
        m2 = (rng.random((F, n)) < rng.uniform(0.05, 0.95, size=(F, 1))).astype(np.uint8)
    with _ctx() as ctx:
        ctx.load_matrix(m2)
        ctx.set_targets(w["target"])
        for seed in (w["seed"], np.zeros(n, dtype=np.uint8)):
            ctx.set_seed(seed)
            ctx.set_dense_x(False)
            a = ctx.score_once()[0]
            ctx.set_dense_x(True)
            b = ctx.score_once()[0]
            ctx.set_dense_x(False)
            ok, err = score_close(a, b)
            assert ok, err
            big = np.abs(b) > 1e-3
            assert np.max(np.abs(a[big] - b[big]) / np.abs(b[big])) < 1e-10


def test_duplicate_row_grouping_is_exact():
    """Scoring only the first row of each group of identical rows (default) must give bit-identical
    scores, winners and seeds to scoring every row; also with scattered (non-adjacent) duplicates and
    with the all-zero / all-one guard rows."""
    w = _workload(160, 900)
    m2 = w["m2"].copy()
    rng = np.random.default_rng(8)
    src = rng.integers(0, 900, 60)
    dst = rng.integers(0, 900, 60)
    m2[dst] = m2[src]
    uniq = len({r.tobytes() for r in m2})
    outs = []
    with _ctx() as ctx:
        ctx.set_targets(np.stack([w["target"], rng.standard_normal(160)]))
        ctx.set_seed(w["seed"])
        for on in (True, False):
            ctx.set_dedup(on)
            ctx.load_matrix(m2)
            assert ctx.unique_rows() == uniq
            outs.append((ctx.score_once(), ctx.search(3)))
    (sa, ra), (sb, rb_) = outs
    assert np.array_equal(sa, sb)
    for k in ("cmi", "top", "top_score", "seed_iter", "cmi_seed"):
        assert np.array_equal(ra[k], rb_[k]), k


def test_production_epilogue_equals_literal_cube(oracle):
    """The production entropy epilogue (cell regimes: closed forms for the big / small densities, a log only
    for the few in between) against the literal cell-by-cell cube of the validation mode, on the same x-axis
    sums: CMI may differ by rounding only.  The literal mode's mass-based column pruning (rvl_set_prune_theta)
    is checked against its bound theta*(|log eps'|+1) as well."""
    for n, F in ((400, 1500), (82, 600), (1000, 800)):
        w = _workload(n, F)
        rng = np.random.default_rng(n)
        m2 = w["m2"].copy()
        for r, p in zip(range(5, 5 + 6), (0.1, 0.25, 0.5, 0.75, 0.9, 0.97)):     # dense rows: both y-classes visible
            m2[r] = rng.random(n) < p
        with _ctx() as ctx:
            ctx.set_profiling(True)
            ctx.load_matrix(m2)
            ctx.set_targets(w["target"])
            ctx.set_seed(w["seed"])
            a = ctx.score_once()[0]
            ca = ctx.counters(reset=True)
            ctx.set_dense_x(True)
            ctx.set_prune_theta(0.0)
            b = ctx.score_once()[0]
            cb = ctx.counters(reset=True)
            ctx.set_prune_theta(1e-14)
            c = ctx.score_once()[0]
            with pytest.raises(Exception):
                ctx.set_prune_theta(1e-3)
        assert ca["logs"] < cb["logs"]
        cmi_a, cmi_b, cmi_c = (-0.5 * np.log1p(-v * v) for v in (a, b, c))
        assert np.max(np.abs(cmi_a - cmi_b)) <= 2e-14, (n, np.max(np.abs(cmi_a - cmi_b)))
        assert np.max(np.abs(cmi_c - cmi_b)) <= 1e-12
        assert np.array_equal(np.sign(a), np.sign(b))
        rows = np.arange(0, F, 10)
        o = oracle.score_rows(w["target"], m2[rows].astype(float), w["seed"].astype(float))
        ok, err = score_close(a[rows], o)
        assert ok, err


@pytest.mark.parametrize("G", [8, 17, 32])
def test_production_epilogue_other_grids_and_exact_pruning(G):
    """The column intervals, the float table's row stride (multiple of 4 floats with an odd quarter: 12 / 20 / 36 for these
    grids) and the soft cells at n_grid != 25, and with the mass-based pruning switched off in BOTH modes (theta = 0: a
    column is dead only where neither class can change fl(Q) -- the interval ends then meet the visibility bounds)."""
    n, F = 300, 700
    w = _workload(n, F)
    rng = np.random.default_rng(G)
    m2 = w["m2"].copy()
    for r, p in zip(range(3, 3 + 5), (0.15, 0.4, 0.6, 0.85, 0.98)):              # dense and complemented rows too
        m2[r] = rng.random(n) < p
    with _ctx() as ctx:
        ctx.set_options(n_grid=G)
        ctx.load_matrix(m2)
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        a = ctx.score_once()[0]
        ctx.set_prune_theta(0.0)
        a0 = ctx.score_once()[0]
        ctx.set_dense_x(True)
        b0 = ctx.score_once()[0]
    cmi_a, cmi_a0, cmi_b0 = (-0.5 * np.log1p(-v * v) for v in (a, a0, b0))
    assert np.max(np.abs(cmi_a0 - cmi_b0)) <= 2e-14, (G, np.max(np.abs(cmi_a0 - cmi_b0)))
    assert np.max(np.abs(cmi_a - cmi_b0)) <= 1e-12
    assert np.array_equal(np.sign(a0), np.sign(b0))


# ------------------------------------------------------------------ target batches and the permutation null

def test_target_batch_equals_individual_runs():
    w = _workload(100, 200)
    rng = np.random.default_rng(11)
    targets = np.stack([w["target"], rng.standard_normal(100), rng.standard_normal(100)])
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(targets)
        ctx.set_seed(w["seed"])
        batch = ctx.search(3)
        singles = []
        for t in range(3):
            ctx.set_targets(targets[t])
            ctx.set_seed(w["seed"])
            singles.append(ctx.search(3))
    for t in range(3):
        assert np.array_equal(batch["cmi"][t], singles[t]["cmi"][0])
        assert np.array_equal(batch["top"][t], singles[t]["top"][0])
        assert np.array_equal(batch["seed_iter"][t], singles[t]["seed_iter"][0])


def test_permutation_null_and_pvalues(oracle):
    import revealer_b200 as rb
    w = _workload(90, 150)
    rng = np.random.default_rng(12)
    perm = np.stack([rng.permutation(w["target"]) for _ in range(4)])
    out = rb.revealer_search(w["target"], w["m2"], w["seed"], max_n_iter=2, target_rand=perm)
    ref = oracle.search(w["target"], w["m2"].astype(float), w["seed"].astype(float), 2)
    assert list(out["top"]) == list(ref["top"])
    # null scores of iteration 1 use the seed after iteration 0's merge (LIB:1853-1855)
    seeds = [w["seed"].astype(float), ref["seed_iter"][0]]
    for it in range(2):
        for j in range(4):
            o = oracle.score_rows(perm[j], w["m2"].astype(float), seeds[it])
            ok, err = score_close(out["cmi_rand"][it][:, j], o)
            assert ok, (it, j, err)
        want = oracle.pvals(out["cmi"][:, it], out["cmi_rand"][it])
        assert np.array_equal(out["p_val"][it], want)
    # the stand-alone call on host arrays gives the same numbers; a NaN score gives NaN (R: NA), NaN nulls are not counted
    with _ctx() as ctx:
        it = 1
        assert np.array_equal(ctx.pvalues(out["cmi"][:, it], out["cmi_rand"][it]), want)
        c2 = out["cmi"][:, it].copy()
        c2[3] = np.nan
        r2 = out["cmi_rand"][it].copy()
        r2[0, 0] = np.nan
        got = ctx.pvalues(c2, r2)
        assert np.isnan(got[3]) and np.array_equal(got, oracle.pvals(c2, r2), equal_nan=True)
    assert out["FDR"].shape == out["p_val"].shape and np.all((out["FDR"] >= 0) & (out["FDR"] <= 1))
    assert np.all(out["FDR"] >= np.where(out["cmi"].T < 0, 1 - out["p_val"], out["p_val"]) - 1e-15)


def test_assess_features_matches_oracle(oracle):
    """REVEALER_assess_features.v1's numbers (LIB:2875-2890): per-feature IC and its permutation p-value.
    Few rows x many permuted targets: the core takes the literal x-axis pass here (no table)."""
    import revealer_b200 as rb
    rng = np.random.default_rng(21)
    n, F, n_perm = 120, 5, 40
    x = rng.standard_normal(n)
    feats = (rng.random((F, n)) < 0.15).astype(np.uint8)
    feats[1, :30] = (x[:30] > 0)                               # one feature that tracks the target
    perms = np.stack([rng.permutation(x) for _ in range(n_perm)])
    out = rb.assess_features(x, feats, target_rand=perms)
    ic = np.array([oracle.mutual_inf_v2(x, feats[i].astype(float))["IC"] for i in range(F)])
    ok, err = score_close(out["IC"], ic)
    assert ok, err
    null = np.array([[oracle.mutual_inf_v2(perms[h], feats[i].astype(float))["IC"] for i in range(F)]
                     for h in range(n_perm)])
    ok, err = score_close(out["null_IC"], null)
    assert ok, err
    want_p = np.where(ic >= 0, (null >= ic).sum(0), (null <= ic).sum(0)) / n_perm
    assert np.allclose(out["p_val"], want_p)


def test_permuted_targets_share_bandwidth_and_are_verified():
    import revealer_b200 as rb
    from revealer_b200 import _lib
    rng = np.random.default_rng(22)
    x = rng.standard_normal(200)
    perms = np.stack([x] + [rng.permutation(x) for _ in range(5)])
    with _ctx() as ctx:
        ctx.set_targets(perms)
        a = [ctx.target_bandwidth(t) for t in range(6)]
        ctx.set_targets(perms, permuted=True)
        b = [ctx.target_bandwidth(t) for t in range(6)]
        # bcv is permutation-invariant; computed independently the variance's last bit may differ
        assert len(set(b)) == 1 and b[0] == a[0]
        assert max(abs(v - a[0]) for v in a) <= 1e-13 * a[0]
        bad = perms.copy()
        bad[3, 7] += 1e-9
        with pytest.raises(rb.RevealerError) as e:
            ctx.set_targets(bad, permuted=True)
        assert e.value.code == _lib.RVL_ERR_ARG


# ------------------------------------------------------------------ sharded form on one GPU

def test_two_shards_equal_single_shard():
    """Feature-row sharding emulated with two contexts on one GPU: candidate records are
    exchanged through torch tensors exactly as the NCCL all-gather would fill them."""
    import torch
    import revealer_b200 as rb
    w = _workload(140, 501)
    F = 501
    iters = 4
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        single = ctx.search(iters)
    world = 2
    ctxs, sends, recvs = [], [], []
    stream = torch.cuda.current_stream().cuda_stream
    for r in range(world):
        lo, cnt = rb.shard_rows(F, world, r)
        c = _ctx()
        c.set_stream(stream)
        c.set_shard(r, world, lo, F)
        c.load_matrix(w["m2"][lo:lo + cnt])
        c.set_targets(w["target"])
        c.set_seed(w["seed"])
        nb = c.candidate_bytes()
        s = torch.zeros(nb, dtype=torch.uint8, device="cuda")
        rv = torch.zeros(nb * world, dtype=torch.uint8, device="cuda")
        c.set_exchange_buffers(s.data_ptr(), rv.data_ptr())
        c.search_begin(iters)
        ctxs.append(c), sends.append(s), recvs.append(rv)
    for it in range(iters):
        for c in ctxs:
            c.iter_score(it)
        gathered = torch.cat(sends)
        for c, rv in zip(ctxs, recvs):
            rv.copy_(gathered)
            c.iter_merge(it)
    outs = [c.search_fetch(iters) for c in ctxs]
    for r, o in enumerate(outs):
        lo, cnt = rb.shard_rows(F, world, r)
        assert np.array_equal(o["top"], single["top"])
        assert np.array_equal(o["seed_iter"], single["seed_iter"])
        assert np.array_equal(o["cmi_seed"], single["cmi_seed"])
        assert np.array_equal(o["cmi"][0], single["cmi"][0][:, lo:lo + cnt])
    for c in ctxs:
        c.close()


# ------------------------------------------------------------------ one persistent launch == split-phase launches

@pytest.mark.parametrize("n,F,iters,T,kw", [(140, 900, 5, 1, {}), (82, 300, 3, 1, {"null_seed": True}),
                                            (1000, 3000, 4, 3, {}), (64, 50, 2, 2, {}), (300, 40, 3, 1, {})])
def test_fused_search_kernel_equals_split_phase_launches(n, F, iters, T, kw):
    """rvl_search runs the whole search as ONE cooperative launch (k_search).  It must reproduce the split-phase
    launches (k_score / k_argmax / k_advance per iteration) bit for bit: same functions, same order.  Covered:
    single and several targets, NULLSEED (IC dispatch -> CIC after the first merge), target.match = negative,
    the literal x-axis pass chosen for few rows (F <= 64), permutation-null targets, repeated searches."""
    w = _workload(n, F, **kw)
    rng = np.random.default_rng(F)
    targets = np.stack([w["target"]] + [rng.standard_normal(n) for _ in range(T - 1)])
    for negative in (False, True):
        res = {}
        for fused in (True, False):
            with _ctx(negative=negative) as ctx:
                ctx.set_fused(fused)
                ctx.load_matrix(w["m2"])
                ctx.set_targets(targets)
                ctx.set_seed(w["seed"])
                a = ctx.search(iters)
                b = ctx.search(iters)                  # a second search on the same context (odd iters: slot parity)
                l0 = ctx.launch_count()
                ctx.search(iters, fetch=False)
                res[fused] = (a, b, ctx.launch_count() - l0)
        for k in ("cmi", "top", "top_score", "seed_iter", "cmi_seed", "ic_seed0"):
            assert np.array_equal(res[True][0][k], res[False][0][k], equal_nan=True), (k, negative)
            assert np.array_equal(res[True][1][k], res[True][0][k], equal_nan=True), (k, negative)
        assert res[True][2] < res[False][2]            # 1 + a few launches instead of 3 per iteration
    # permutation-null targets ride along (scored against target 0's seed, never selected)
    perm = np.stack([w["target"]] + [rng.permutation(w["target"]) for _ in range(3)])
    out = {}
    for fused in (True, False):
        with _ctx() as ctx:
            ctx.set_fused(fused)
            ctx.load_matrix(w["m2"])
            ctx.set_targets(perm, first_null=1, permuted=True)
            ctx.set_seed(w["seed"])
            out[fused] = ctx.search(iters)
    for k in ("cmi", "top", "seed_iter", "cmi_seed"):
        assert np.array_equal(out[True][k][:1] if k != "cmi" else out[True][k], out[False][k][:1] if k != "cmi" else out[False][k],
                              equal_nan=True), k


# ------------------------------------------------------------------ config C1's shape (gpunit example: 82 x ~17.8k)

def test_c1_shape_full_search_matches_oracle(oracle):
    """BASELINE config C1 is the reference's own gpunit example (82 samples, ~17.8k features, seed, 2
    iterations); its inputs are unreachable URLs, so the same shape is run on synthetic data: EVERY row
    of both iterations against the oracle, winners and seeds exact."""
    from revealer_b200 import synth
    w = synth.make_workload("C1_gpunit_like")
    iters = 2
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"].astype(np.float64))              # the R layout path
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        res = ctx.search(iters)
    ref = oracle.search(w["target"], w["m2"].astype(float), w["seed"].astype(float), iters, hoist=True, nthreads=0)
    assert list(res["top"][0]) == list(ref["top"])
    assert np.array_equal(res["seed_iter"][0], ref["seed_iter"].astype(np.uint8))
    ok, err = score_close(res["cmi"][0], ref["cmi"])
    assert ok, err
    ok, err = score_close(res["cmi_seed"][0], ref["cmi_seed"])
    assert ok, err
    # the full display ranking (R's order()) agrees wherever the oracle's neighbours are separated by
    # more than the tolerance; exact ties keep row order on both sides
    for it in range(iters):
        g_ord = np.argsort(np.where(np.isnan(res["cmi"][0, it]), np.inf, -res["cmi"][0, it]), kind="stable")
        assert g_ord[0] == ref["top"][it]


# ------------------------------------------------------------------ config C4's sample count (n = 10 000)

def test_c4_sample_count_rows_match_oracle(oracle):
    """n = 10 000 (BASELINE config C4): multi-row seed, 157-word bit rows, the n >= 5793 range where the
    original VR_bcv_bin's int product 64*n*n would overflow (evaluated in double here and in the oracle)."""
    from revealer_b200 import synth
    n, F = 10000, 400
    rng = np.random.Generator(np.random.PCG64(synth.BASE_SEED + 3))
    x = synth.make_target(n, rng)
    m2 = synth.make_features(n, F, np.random.Generator(np.random.PCG64(5)))
    seed = (m2[10] | m2[20] | m2[30]).astype(np.uint8)          # three rows OR-ed (LIB:1697)
    with _ctx() as ctx:
        ctx.load_matrix(m2)
        ctx.set_targets(x)
        ctx.set_seed(seed)
        g = ctx.score_once()[0]
        bw = ctx.target_bandwidth(0)
        tab = ctx.binary_bandwidth_table()
        res = ctx.search(2)
    assert abs(bw - oracle.bcv(x)) <= 1e-12 * bw
    rows = [0, 57, 123, 399]
    o = oracle.score_rows(x, m2[rows].astype(float), seed.astype(float))
    ok, err = score_close(g[rows], o)
    assert ok, err
    y = m2[57].astype(float)
    assert abs(tab[int(y.sum())] - oracle.bcv(y)) <= 1e-12 * oracle.bcv(y)
    top = int(res["top"][0, 0])
    assert g[top] == np.nanmax(g) and np.array_equal(res["seed_iter"][0, 0], np.maximum(seed, m2[top]))


# ------------------------------------------------------------------ full-size properties (BASELINE config C3)

@pytest.fixture(scope="module")
def c3():
    from revealer_b200 import synth
    return synth.make_workload("C3_1kx100k")


def test_c3_sampled_rows_match_oracle(oracle, c3):
    with _ctx() as ctx:
        ctx.load_matrix(c3["m2"])
        ctx.set_targets(c3["target"])
        ctx.set_seed(c3["seed"])
        g = ctx.score_once()[0]
    rng = np.random.default_rng(0)
    rows = np.sort(rng.choice(c3["F"], 160, replace=False))
    o = oracle.score_rows(c3["target"], c3["m2"][rows].astype(float), c3["seed"].astype(float))
    ok, err = score_close(g[rows], o)
    assert ok, err
    # exact duplicates -> bit-identical scores; guard rows -> exact zeros
    dup = np.flatnonzero((c3["m2"][1:] == c3["m2"][:-1]).all(axis=1)) + 1
    assert len(dup) > 10000 and np.array_equal(g[dup], g[dup - 1])
    cnt = c3["m2"].sum(axis=1)
    assert np.all(g[(cnt == 0) | (cnt == 1000)] == 0.0)
    assert np.isfinite(g).all()


def test_c3_sample_order_invariance_and_complement_symmetry(c3):
    """Size-independent properties: permuting the samples of target, features and seed together leaves
    every score unchanged (<= 1e-9); with a constant seed, IC(x, 1-y) = -IC(x, y)."""
    rng = np.random.default_rng(1)
    p = rng.permutation(1000)
    sub = c3["m2"][:20000]
    with _ctx() as ctx:
        ctx.load_matrix(sub)
        ctx.set_targets(c3["target"])
        ctx.set_seed(c3["seed"])
        a = ctx.score_once()[0]
        ctx.set_seed(np.zeros(1000, dtype=np.uint8))
        ic = ctx.score_once()[0]
        ctx.load_matrix(1 - sub)
        icc = ctx.score_once()[0]
        ctx.load_matrix(np.ascontiguousarray(sub[:, p]))
        ctx.set_targets(c3["target"][p])
        ctx.set_seed(c3["seed"][p])
        b = ctx.score_once()[0]
    # scores below ~1e-6 are round-off of a vanishing CMI (see the module docstring): score_close's floor
    ok, err = score_close(a, b)
    assert ok, err
    ok, err = score_close(ic, -icc)
    assert ok, err


def test_c3_search_winner_is_consistent(c3):
    """Full-size search: each iteration's winner is the arg-max of that iteration's scores with the
    lowest-index tie-break, the seed is the running OR, and rows are never removed."""
    iters = 3
    with _ctx() as ctx:
        ctx.load_matrix(c3["m2"])
        ctx.set_targets(c3["target"])
        ctx.set_seed(c3["seed"])
        res = ctx.search(iters)
    seed = c3["seed"].copy()
    for it in range(iters):
        col = res["cmi"][0, it]
        top = int(res["top"][0, it])
        assert col[top] == np.nanmax(col) and top == int(np.flatnonzero(col == col[top])[0])
        seed = np.maximum(seed, c3["m2"][top])
        assert np.array_equal(res["seed_iter"][0, it], seed)


def test_programmatic_dependent_launch_changes_nothing():
    """The per-iteration chain k_score -> k_argmax -> k_advance launched as programmatic dependents (default) and as plain
    stream-ordered launches (RVL_PDL=0, read at rvl_create) must give bit-identical searches."""
    w = _workload(300, 900)
    res = {}
    try:
        for pdl in ("0", "1"):
            os.environ["RVL_PDL"] = pdl
            with _ctx() as ctx:
                ctx.load_matrix(w["m2"])
                ctx.set_targets(w["target"])
                ctx.set_seed(w["seed"])
                res[pdl] = ctx.search(6)
    finally:
        os.environ["RVL_PDL"] = "1"
        with _ctx():
            pass                                # the switch is process-wide: leave it on
        del os.environ["RVL_PDL"]
    for k in ("cmi", "top", "top_score", "seed_iter", "cmi_seed"):
        assert np.array_equal(res["0"][k], res["1"][k], equal_nan=True) if res["0"][k].dtype.kind == "f" else np.array_equal(res["0"][k], res["1"][k]), k
