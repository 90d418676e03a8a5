"""Search-vs-oracle parity at the shapes BASELINE.json names (configs C2..C5), through the C ABI.

  C2  500 x 20k, 5 iterations      every score of every iteration, winners, seeds, seed ICs vs oracle.search
  C3  1k x 100k, 10 iterations     per iteration: the GPU's top-32 rows + 64 random rows re-scored by the oracle
                                   with the GPU's seed of that iteration (in-place table updates included)
  C4  n = 10 000, 3-row seed       full 3-iteration search at F = 1200 vs oracle.search
  C5  256 targets x (1k x 100k)    8 of the targets: sampled rows + the winners' neighbourhood vs the oracle
  n.gpus > 1                       rvl_group_connect / rvl_group_search (the path the R wrapper takes) vs one GPU

The oracle restates R + MASS + misc3d (PARITY UNPINNED against real R).  Tolerance: tests/parity_tol.py.
"""
import numpy as np
import pytest

from parity_tol import score_close

pytestmark = pytest.mark.gpu


def _ctx(**kw):
    import revealer_b200 as rb
    return rb.RevealerContext(0, **kw)


def _order(col, negative=False):
    key = np.where(np.isnan(col), np.inf, col if negative else -col)
    return np.argsort(key, kind="stable")


def test_c2_full_search_matches_oracle(oracle):
    from revealer_b200 import synth
    w = synth.make_workload("C2_500x20k")
    iters = w["iters"]
    assert (w["n"], w["F"], iters) == (500, 20000, 5)
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"].astype(np.float64))              # the R layout: column-major doubles
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        res = ctx.search(iters)
    ref = oracle.search(w["target"], w["m2"].astype(float), w["seed"].astype(float), iters, hoist=True)
    assert list(res["top"][0]) == list(ref["top"])
    assert np.array_equal(res["seed_iter"][0], ref["seed_iter"].astype(np.uint8))
    ok, err = score_close(res["cmi"][0], ref["cmi"])
    assert ok, err
    ok, err = score_close(res["cmi_seed"][0], ref["cmi_seed"])
    assert ok, err
    ok, err = score_close(res["ic_seed0"], [ref["ic_seed0"]])
    assert ok, err
    for it in range(iters):   # the display ranking: n.markers = 30 leading rows of R's order()
        g30, o30 = _order(res["cmi"][0, it])[:30], _order(ref["cmi"][it])[:30]
        sep = np.abs(np.diff(ref["cmi"][it][o30])) > 1e-6 * np.abs(ref["cmi"][it][o30][:-1])
        exact_tie = np.diff(ref["cmi"][it][o30]) == 0
        if np.all(sep | exact_tie):
            assert list(g30) == list(o30), it


def test_c3_all_ten_iterations_top32_rescored_by_oracle(oracle):
    from revealer_b200 import synth
    w = synth.make_workload("C3_1kx100k")
    iters = w["iters"]
    assert iters == 10
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        res = ctx.search(iters)
    rng = np.random.default_rng(7)
    seed = w["seed"].copy()
    x = w["target"]
    tops = []
    for it in range(iters):
        col = res["cmi"][0, it]
        order = _order(col)
        top = int(res["top"][0, it])
        assert top == order[0]
        rows = np.unique(np.concatenate([order[:32], rng.choice(w["F"], 64, replace=False)]))
        o = oracle.score_rows(x, w["m2"][rows].astype(float), seed.astype(float))
        ok, err = score_close(col[rows], o)
        assert ok, (it, err)
        # the oracle ranks the GPU's leading rows the same way, and the GPU's winner is the oracle's
        lead = order[:32]
        pos = {r: i for i, r in enumerate(rows)}
        ol = np.array([o[pos[r]] for r in lead])
        o_order = lead[np.lexsort((lead, -ol))]                 # decreasing score, ties -> lower row (R's stable order)
        separated = np.abs(np.diff(np.sort(ol)[::-1])) > 1e-6 * np.abs(np.sort(ol)[::-1][:-1])
        ties = np.diff(np.sort(ol)[::-1]) == 0
        assert o_order[0] == top, it
        if np.all(separated | ties):
            assert list(o_order) == list(lead), it
        assert res["top_score"][0, it] == col[top]
        seed = np.maximum(seed, w["m2"][top])                   # LIB:2168
        assert np.array_equal(res["seed_iter"][0, it], seed), it
        tops.append(top)
    # seed ICs of every iteration (LIB:2170) and the baseline (LIB:1833)
    s = w["seed"].astype(float)
    assert score_close(res["ic_seed0"], [oracle.assoc(x, s)])[0]
    for it in range(iters):
        s = np.maximum(s, w["m2"][tops[it]])
        ok, err = score_close(res["cmi_seed"][0, it], oracle.assoc(x, s))
        assert ok, (it, err)


def test_c4_shape_three_row_seed_full_search_matches_oracle(oracle):
    """n = 10 000 samples (157-word bit rows), a seed that is the OR of three rows (LIB:1697), 3 iterations:
    every score, winner and merged seed against the oracle."""
    from revealer_b200 import synth
    n, F, iters = 10000, 1200, 3
    w = synth.make_workload(n=n, F=F, n_seed_rows=3, seed_offset=synth.CONFIG_INDEX["C4_10kx500k"])
    assert 3 <= int(w["seed"].sum())
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        res = ctx.search(iters)
    ref = oracle.search(w["target"], w["m2"].astype(float), w["seed"].astype(float), iters, hoist=True)
    assert list(res["top"][0]) == list(ref["top"])
    assert np.array_equal(res["seed_iter"][0], ref["seed_iter"].astype(np.uint8))
    ok, err = score_close(res["cmi"][0], ref["cmi"])
    assert ok, err
    ok, err = score_close(res["cmi_seed"][0], ref["cmi_seed"])
    assert ok, err


def test_c5_eight_of_256_targets_match_oracle(oracle):
    """Config C5: 256 targets scored against one 1k x 100k matrix in one pass.  8 targets (the sorted one and 7 of
    the unsorted drug-response-like profiles): 24 random rows + the GPU's 8 leading rows per target vs the
    oracle, and every target's winner is the arg-max of its own score column."""
    from revealer_b200 import synth
    w = synth.make_workload("C5_256targets_1kx100k")
    T, F = w["n_targets"], w["F"]
    assert (T, F) == (256, 100000)
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(w["target"])
        ctx.set_seed(w["seed"])
        res = ctx.search(1)
    rng = np.random.default_rng(11)
    picks = [0] + sorted(rng.choice(np.arange(1, T), 7, replace=False).tolist())
    for t in picks:
        col = res["cmi"][t, 0]
        order = _order(col)
        assert int(res["top"][t, 0]) == order[0], t
        rows = np.unique(np.concatenate([order[:8], rng.choice(F, 24, replace=False)]))
        o = oracle.score_rows(w["target"][t], w["m2"][rows].astype(float), w["seed"].astype(float))
        ok, err = score_close(col[rows], o)
        assert ok, (t, err)
        pos = {r: i for i, r in enumerate(rows)}
        lead = order[:8]
        ol = np.array([o[pos[r]] for r in lead])
        assert lead[np.lexsort((lead, -ol))][0] == order[0], t
        assert np.array_equal(res["seed_iter"][t, 0], np.maximum(w["seed"], w["m2"][order[0]])), t
    # all 256 winners are their column's first maximum
    for t in range(T):
        col = res["cmi"][t, 0]
        assert int(res["top"][t, 0]) == int(np.flatnonzero(col == np.nanmax(col))[0])


def test_group_search_over_two_gpus_equals_one_gpu():
    """rvl_group_connect / rvl_group_search -- one host thread driving one context per GPU, the form an R
    session uses with n.gpus > 1 -- gives the single-GPU result bit for bit."""
    import torch
    import revealer_b200 as rb
    from revealer_b200 import synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    w = synth.make_workload(n=300, F=4001)
    iters = 5                                                    # odd, and repeated: exercises the exchange slots
    one = rb.revealer_search(w["target"], w["m2"], w["seed"], max_n_iter=iters)
    for rep in range(3):
        two = rb.revealer_search_multi_gpu(w["target"], w["m2"], w["seed"], max_n_iter=iters, devices=(0, 1))
        assert np.array_equal(two["top"], one["top"])
        assert np.array_equal(two["seed_iter"], one["seed_iter"])
        assert np.array_equal(two["cmi"], one["cmi"])
        assert np.array_equal(two["cmi_seed"], one["cmi_seed"])


def test_connected_contexts_repeat_searches_with_odd_iteration_counts():
    """The two exchange slots alternate on a running counter, not on the iteration's parity: back-to-back searches with
    an odd iteration count on the SAME connected contexts must neither tear a record nor wait for a stamp that went to
    the other slot -- also when one context ran searches of its own before the group was connected, and with the
    one-launch k_search on both devices."""
    import ctypes as C
    import torch
    import revealer_b200 as rb
    from revealer_b200 import synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    w = synth.make_workload(n=200, F=3001)
    iters = 3
    one = rb.revealer_search(w["target"], w["m2"], w["seed"], max_n_iter=iters)
    F, world = 3001, 2
    ctxs = []
    try:
        for r in range(world):
            lo, cnt = rb.shard_rows(F, world, r)
            c = rb.RevealerContext(r)
            ctxs.append(c)
            c.load_matrix(w["m2"][lo:lo + cnt])
            c.set_targets(w["target"])
            c.set_seed(w["seed"])
            if r == 0:
                c.search(5)                       # a history of its own (world of one) before joining the group
            c.set_shard(r, world, lo, F)
        lib = ctxs[0]._lib
        arr = (C.c_void_p * world)(*[c._h for c in ctxs])
        assert lib.rvl_group_connect(arr, world) == 0
        for fused in (False, True):
            for c in ctxs:
                c.set_fused(fused)
            for rep in range(4):
                rc = lib.rvl_group_search(arr, world, iters)
                assert rc == 0, [lib.rvl_last_error(c._h) for c in ctxs]
                parts = [c.search_fetch(iters) for c in ctxs]
                cmi = np.concatenate([p["cmi"][0] for p in parts], axis=1).T
                assert np.array_equal(cmi, one["cmi"], equal_nan=True), (fused, rep)
                for p in parts:
                    assert np.array_equal(p["top"][0], one["top"]) and np.array_equal(p["seed_iter"][0], one["seed_iter"])
                    assert np.array_equal(p["cmi_seed"][0], one["cmi_seed"])
    finally:
        for c in ctxs:
            c.close()


def test_target_sharded_batch_equals_single_gpu_batch_and_individual_runs():
    """Config C5's multi-GPU form: targets sharded over the devices (whole matrix on each, no exchange).  The batch
    must equal the same batch on one GPU and the targets run one at a time, bit for bit."""
    import torch
    import revealer_b200 as rb
    from revealer_b200 import synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under `gpurun --gpus 2`)")
    w = synth.make_workload(n=200, F=3000)
    rng = np.random.default_rng(5)
    targets = np.stack([w["target"]] + [rng.standard_normal(200) for _ in range(6)])      # 7 targets over 2 devices: 3 + 4
    multi = rb.revealer_search_target_batch_multi_gpu(targets, w["m2"], w["seed"], max_n_iter=2, devices=(0, 1))
    with _ctx() as ctx:
        ctx.load_matrix(w["m2"])
        ctx.set_targets(targets)
        ctx.set_seed(w["seed"])
        one = ctx.search(2)
        for k in ("cmi", "top", "top_score", "seed_iter", "cmi_seed", "ic_seed0"):
            assert np.array_equal(multi[k], one[k], equal_nan=True), k
        for t in (0, 3, 6):
            ctx.set_targets(targets[t])
            ctx.set_seed(w["seed"])
            single = ctx.search(2)
            assert np.array_equal(single["top"][0], multi["top"][t])
            assert np.array_equal(single["cmi"][0], multi["cmi"][t], equal_nan=True)


def test_row_block_of_a_larger_r_matrix_loads_in_place():
    """rvl_load_matrix_f64_colmajor with ld > F: a shard's row block of the whole column-major R matrix, read
    where it lies (what rvlR_search_multi passes).  Only the block's rows may be read and shipped."""
    from revealer_b200 import synth
    w = synth.make_workload(n=333, F=5000)
    full = np.asfortranarray(w["m2"].astype(np.float64))         # F_total x n, column-major like R
    for lo, cnt, threads in ((0, 1700, 0), (1700, 1650, 0), (3350, 1650, 0), (3350, 1650, 3), (4999, 1, 0)):
        with _ctx() as ctx:
            ctx.set_host_threads(threads)
            ctx.load_matrix_block(full, lo, cnt)
            h2d, _ = ctx.upload_stats()
            assert h2d <= cnt * 333 * 8
            ctx.set_targets(w["target"])
            ctx.set_seed(w["seed"])
            a = ctx.score_once()[0]
            cnts = ctx.row_counts()
        with _ctx() as ctx:
            ctx.load_matrix(w["m2"][lo:lo + cnt])
            ctx.set_targets(w["target"])
            ctx.set_seed(w["seed"])
            b = ctx.score_once()[0]
        assert np.array_equal(cnts, w["m2"][lo:lo + cnt].sum(axis=1))
        assert np.array_equal(a, b)
