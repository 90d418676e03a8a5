"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/revealer_b200.h declares; host-side helpers (bit packing, sharding, winner resolution); and the
product path fails loudly without a GPU (no CPU fallback, nothing under oracle/ is imported)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "revealer_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rvl_[a-z0-9_]+)\s*\(", hdr)))


@pytest.fixture(scope="module")
def built_lib():
    from revealer_b200 import build as rb_build
    return rb_build.build()


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 30
    for s in names:
        assert hasattr(lib, s), "missing export: " + s
    lib.rvl_abi_version.restype = ctypes.c_int
    assert lib.rvl_abi_version() == 1


def test_python_binding_covers_the_header(built_lib):
    from revealer_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_means_error_not_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only container")
    import revealer_b200 as rb
    with pytest.raises(rb.RevealerError):
        rb.RevealerContext(0)
    with pytest.raises(rb.RevealerError):
        rb.assoc(np.arange(10.0), np.array([0, 1] * 5))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "revealer_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".R")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower() or f == "synth.py" and "oracle" not in src.lower(), f
    code = "import sys; import revealer_b200; assert not any('oracle' in m for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_pack_bits_layout():
    import revealer_b200 as rb
    rng = np.random.default_rng(0)
    for n in (1, 63, 64, 65, 130, 1000):
        m = (rng.random((5, n)) < 0.4).astype(np.uint8)
        bits, words = rb.pack_bits(m)
        assert words % 2 == 0 and words * 64 >= n and bits.shape == (5, words)
        for k in range(5):
            for s in range(n):
                assert (int(bits[k, s >> 6]) >> (s & 63)) & 1 == m[k, s]
        assert all(int(bits[k, (n - 1) >> 6]) >> ((n - 1) & 63) >> 1 == 0 for k in range(5)) or n % 64 == 0
    with pytest.raises(rb.RevealerError):
        rb.pack_bits(np.array([[0, 2, 1]]))


def test_shard_rows_partition():
    import revealer_b200 as rb
    for F in (0, 1, 7, 100000, 500001):
        for world in (1, 2, 4, 8):
            spans = [rb.shard_rows(F, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == F
            for (lo, c), (lo2, _) in zip(spans, spans[1:]):
                assert lo + c == lo2
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_resolve_winner_matches_r_order(oracle):
    from revealer_b200.api import resolve_winner
    rng = np.random.default_rng(1)
    for trial in range(200):
        k = rng.integers(1, 9)
        rows = rng.choice(1000, k, replace=False)
        sc = rng.choice([0.1, 0.5, 0.5, -0.3, np.nan], k)
        for neg in (False, True):
            # scatter into a full score vector and ask the oracle's order()
            full = np.full(1000, np.nan)
            full[rows] = sc
            want = oracle.top1(full, negative=neg)
            got = rows[resolve_winner(sc, rows, neg)]
            if np.all(np.isnan(sc)):
                assert got == rows.min()
            else:
                assert got == want
    assert resolve_winner([0.2, 0.9], [-1, -1]) == -1
    assert resolve_winner([0.2, np.nan, 0.1], [-1, 5, 9]) == 2


def test_reference_formals_are_preserved():
    """REVEALER.v1's 12 formals (LIB:1504-1516), same names (dots -> underscores) and defaults."""
    import inspect
    import revealer_b200 as rb
    sig = inspect.signature(rb.REVEALER_v1)
    names = list(sig.parameters)[:12]
    assert names == ["ds1", "target_name", "target_match", "ds2", "seed_names", "exclude_features", "max_n_iter",
                     "pdf_output_file", "count_thres_low", "count_thres_high", "n_markers", "locs_table_file"]
    d = {k: v.default for k, v in sig.parameters.items()}
    assert d["target_match"] == "positive" and d["max_n_iter"] == 5 and d["n_markers"] == 30
    assert d["seed_names"] is None and d["count_thres_low"] is None and d["n_grid"] == 25
    assert d["bandwidth_mult"] == 0.25 and d["bandwidth_shrink"] == -0.75
    # the reference's internal constants (LIB:1519, 1531) and the drop-in reader: GCT files go through the host reader unless
    # the caller opts into the device tokeniser (which accepts fewer spellings than read.delim)
    assert d["n_perm"] == 10 and d["r_seed"] == 34578 and d["ingest"] == "host"
    r = open(os.path.join(ROOT, "revealer_b200", "r", "REVEALER_b200.R")).read()
    assert "REVEALER.v1 <- function(" in r and "gct.ingest = FALSE" in r and "n.gpus = 1" in r
    assert 'identical(seed.names, "NULLSEED")' in r and "rvlR_search_multi" in r
    head = r[r.index("REVEALER.v1 <- function("):]
    formals = ["ds1", "target.name", "target.match", "ds2", "seed.names", "exclude.features", "max.n.iter", "pdf.output.file",
               "count.thres.low", "count.thres.high", "n.markers", "locs.table.file", "n.grid", "bandwidth.mult",
               "bandwidth.shrink", "n.gpus"]
    pos = [head.index(f) for f in formals]
    assert pos == sorted(pos)                    # the 12 reference formals first and in order, then the optional four


def test_every_dot_call_in_the_r_wrapper_matches_a_shim_entry_point():
    """Each .Call("rvlR_x", a, b, ...) in REVEALER_b200.R names a function r_shim.c defines, with that many SEXP arguments
    (R would only find a mismatch at run time, and there is no R here to run it)."""
    import re
    r = open(os.path.join(ROOT, "revealer_b200", "r", "REVEALER_b200.R")).read()
    c = open(os.path.join(ROOT, "revealer_b200", "r", "r_shim.c")).read()
    defs = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
            for m in re.finditer(r"^SEXP (rvlR_\w+)\(([^)]*)\)\s*\{", c, flags=re.M | re.S)}
    assert "rvlR_search" in defs and "rvlR_assess" in defs
    calls = 0
    for m in re.finditer(r'\.Call\("(rvlR_\w+)"', r):
        name, i, depth, nargs = m.group(1), m.end(), 1, 0
        while depth:                                   # count the top-level commas up to the closing parenthesis
            ch = r[i]
            depth += ch in "([" 
            depth -= ch in ")]"
            nargs += ch == "," and depth == 1
            i += 1
        assert name in defs, name + " is not defined in r_shim.c"
        assert nargs == defs[name], "%s: called with %d arguments, defined with %d" % (name, nargs, defs[name])
        calls += 1
    assert calls >= 10


def test_prepare_inputs_mirrors_reveler_v1_preprocessing(tmp_path):
    """Host-side preprocessing of REVEALER.v1 (LIB:1575-1720): NA-target samples dropped, samples
    intersected, sorted by target, target row dropped from the features, count filter with seeds kept,
    seed = column max of the seed rows, seed rows removed; GCT files read like MSIG.Gct2Frame."""
    from revealer_b200.api import prepare_inputs, read_gct
    names1 = ["s%d" % i for i in range(8)]
    names2 = ["s7", "s5", "s3", "s1", "s0", "s2", "s9"]
    t = np.array([0.1, np.nan, 2.0, -1.0, 0.5, 3.0, 0.0, 1.5])
    d1 = dict(ds=np.vstack([t, np.arange(8.0)]), row_names=["TGT", "other"], names=names1)
    m = np.array([[1, 1, 1, 1, 1, 1, 1],      # TGT row inside ds2: must be dropped
                  [1, 0, 0, 1, 1, 0, 0],      # A
                  [0, 0, 0, 0, 0, 0, 0],      # B: fails count.thres.low
                  [0, 1, 0, 0, 0, 0, 1],      # SEED1 (kept although only 1 one among the common samples)
                  [1, 1, 1, 0, 1, 1, 0],      # C: 5 ones among common samples -> fails count.thres.high=4
                  [0, 0, 1, 0, 0, 0, 0]], dtype=float)   # SEED2
    d2 = dict(ds=m, row_names=["TGT", "A", "B", "SEED1", "C", "SEED2"], names=names2)
    p = prepare_inputs(d1, "TGT", "positive", d2, seed_names=["SEED1", "SEED2"], count_thres_low=1, count_thres_high=4)
    assert p["sample_names"] == ["s5", "s2", "s7", "s0", "s3"]          # common samples by decreasing target
    assert list(p["target"]) == [3.0, 2.0, 1.5, 0.1, -1.0]
    assert p["feature_names"] == ["A"]
    assert list(p["seed"]) == [1, 0, 0, 0, 1] and list(p["m2"][0]) == [0, 0, 1, 1, 0]
    n = prepare_inputs(d1, "TGT", "negative", d2, seed_names=None)
    assert list(n["target"]) == [-1.0, 0.1, 1.5, 2.0, 3.0] and not n["seed"].any() and n["negative"]
    assert n["feature_names"] == ["A", "B", "SEED1", "C", "SEED2"]
    # the same through GCT files
    def write_gct(path, d):
        with open(path, "w") as fh:
            fh.write("#1.2\n%d\t%d\nName\tDescription\t%s\n" % (len(d["row_names"]), len(d["names"]), "\t".join(d["names"])))
            for r, row in zip(d["row_names"], d["ds"]):
                fh.write("%s\tna\t%s\n" % (r, "\t".join("" if np.isnan(v) else repr(float(v)) for v in row)))
    f1, f2 = str(tmp_path / "a.gct"), str(tmp_path / "b.gct")
    write_gct(f1, d1)
    write_gct(f2, d2)
    g = read_gct(f1)
    assert g["row_names"] == ["TGT", "other"] and np.isnan(g["ds"][0, 1]) and g["names"] == names1
    q = prepare_inputs(f1, "TGT", "positive", f2, seed_names=["SEED1", "SEED2"], count_thres_low=1, count_thres_high=4)
    assert q["sample_names"] == p["sample_names"] and np.array_equal(q["m2"], p["m2"]) and np.array_equal(q["seed"], p["seed"])


def test_r_shim_compiles_against_stub_r_headers():
    """R is absent from this image, so the .Call shim cannot be built for real; it must at least be valid
    C against the R API declarations it uses (tests/rstub/) and the real C-ABI header."""
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    r = subprocess.run([gcc, "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "tests", "rstub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "revealer_b200", "r", "r_shim.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_p_adjust_fdr_is_benjamini_hochberg():
    """p.adjust(p, "fdr") as used at LIB:1870-1871, against the textbook definition."""
    from revealer_b200.api import p_adjust_fdr
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 100):
        p = np.round(rng.random(n), 2)                        # ties included
        o = np.argsort(p, kind="stable")
        s = p[o] * n / np.arange(1, n + 1)
        q = np.minimum(np.minimum.accumulate(s[::-1])[::-1], 1.0)
        want = np.empty(n)
        want[o] = q
        assert np.allclose(p_adjust_fdr(p), want)


def test_synthetic_workload_is_deterministic_and_shardable():
    from revealer_b200 import synth
    a = synth.make_workload(n=120, F=900)
    b = synth.make_workload(n=120, F=900, row_lo=200, row_hi=650)
    assert np.array_equal(a["m2"][200:650], b["m2"]) and np.array_equal(a["seed"], b["seed"])
    assert np.all(np.diff(a["target"]) <= 0)
    dup = ((a["m2"][1:] == a["m2"][:-1]).all(axis=1)).mean()
    assert 0.1 < dup < 0.3
    cnt = a["m2"].sum(axis=1)
    assert (cnt == 0).any() and (cnt == 120).any()


def test_host_packer_threads_produce_the_device_bit_layout(built_lib):
    """hostpack.cpp (the host side of the cooperative upload) needs no GPU: with no DMA side claiming words the
    packer threads take every 64-sample word, and what they leave must be the bit layout pack_bits() defines
    (sample s -> bit s % 64 of word s / 64); entries that are not 0/1 must be reported."""
    from revealer_b200.api import pack_bits
    lib = ctypes.CDLL(built_lib)
    lib.rvl_hostpack_begin.restype = ctypes.c_void_p
    lib.rvl_hostpack_begin.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
    lib.rvl_hostpack_claim_front.restype = ctypes.c_int64
    lib.rvl_hostpack_claim_front.argtypes = [ctypes.c_void_p]
    lib.rvl_hostpack_end.restype = ctypes.c_int
    lib.rvl_hostpack_end.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64)]
    rng = np.random.default_rng(21)
    for F, n, ld_extra, threads in [(1, 2, 0, 1), (37, 64, 0, 2), (5000, 130, 3, 4), (9000, 257, 0, 3)]:
        ld = F + ld_extra
        m = np.zeros((n, ld), dtype=np.float64)                  # C-order n x ld == column-major ld x n
        m[:, :F] = (rng.random((n, F)) < 0.2)
        W = (n + 63) // 64
        packed = np.zeros((W, F), dtype=np.uint64)
        h = lib.rvl_hostpack_begin(m.ctypes.data, F, n, ld, packed.ctypes.data, threads)
        first = ctypes.c_int64(-1)
        bad = lib.rvl_hostpack_end(h, ctypes.byref(first))
        assert bad == 0 and first.value == 0                     # nobody claimed from the front: the packers did all words
        want, words = pack_bits(m[:, :F].T)
        assert np.array_equal(packed.T, want[:, :W])
        # the DMA side takes the first words: the packers must stop where it stands
        h = lib.rvl_hostpack_begin(m.ctypes.data, F, n, ld, packed.ctypes.data, 0)   # no threads: only claims
        got = [lib.rvl_hostpack_claim_front(h) for _ in range(W + 1)]
        assert got == list(range(W)) + [-1]
        assert lib.rvl_hostpack_end(h, ctypes.byref(first)) == 0 and first.value == W
        m[n - 1, F // 2] = 2.0
        h = lib.rvl_hostpack_begin(m.ctypes.data, F, n, ld, packed.ctypes.data, threads)
        assert lib.rvl_hostpack_end(h, ctypes.byref(first)) != 0
        m[n - 1, F // 2] = np.nan
        h = lib.rvl_hostpack_begin(m.ctypes.data, F, n, ld, packed.ctypes.data, threads)
        assert lib.rvl_hostpack_end(h, ctypes.byref(first)) != 0


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the same metric /
    unit / config keys, impl = reference, a cpu_baseline describing the run and an e2e with zero transfer bytes;
    under torchrun only rank 0 prints."""
    import json
    env = dict(os.environ)
    env.pop("RANK", None)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-sample-rows", "48"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "feature-CMI evaluations/sec" and line["unit"] == "evaluations/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"] == "C3_1kx100k" and line["config"]["features"] == 100000
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    env["RANK"] = "1"
    env["WORLD_SIZE"] = "2"
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_bench_score_checksum_is_independent_of_the_sharding():
    """bench.py's `check.cmi_checksum` (position-weighted sum of the score bit patterns, mod 2^64) must be the same
    number whether the rows, or the targets, are split over ranks: that is what lets the driver compare the lines of
    N = 1, 2, 4, 8 -- and it must notice a single flipped bit or two swapped scores."""
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(4)
    T, iters, F = 3, 4, 1000
    cmi = rng.standard_normal((T, iters, F))
    cmi[1, 2, 17] = np.nan
    whole = bench.score_checksum(cmi, 0, F)
    M = (1 << 64) - 1
    for world in (2, 3, 8):
        tot = 0
        for r in range(world):
            lo, hi = r * F // world, (r + 1) * F // world
            tot = (tot + bench.score_checksum(cmi[:, :, lo:hi], lo, F)) & M
        assert tot == whole
        tot = 0
        for r in range(min(world, T)):
            lo, hi = r * T // min(world, T), (r + 1) * T // min(world, T)
            tot = (tot + bench.score_checksum(cmi[lo:hi], 0, F, t_offset=lo)) & M
        assert tot == whole
    flipped = cmi.copy()
    flipped.view(np.uint64)[0, 1, 5] ^= np.uint64(1)
    assert bench.score_checksum(flipped, 0, F) != whole
    swapped = cmi.copy()
    swapped[0, 0, [3, 4]] = swapped[0, 0, [4, 3]]
    assert bench.score_checksum(swapped, 0, F) != whole


def test_score_block_plan_host_logic():
    """rvl_score_plan (host side of k_score's launch): the most worker warps per block whose per-warp scratch -- struct +
    float table of the soft cells + one bit row -- fits the dynamic shared-memory budget.  No GPU needed."""
    import ctypes as C
    from revealer_b200 import _lib
    lib = _lib.load()
    f = lib.rvl_score_plan
    f.restype = C.c_size_t
    f.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_int)]
    budget = (227 - 33) * 1024

    def plan(n, G, max_warps=0):
        words = ((n + 63) // 64 + 1) & ~1
        w = C.c_int(0)
        return f(words, G, budget, max_warps, C.byref(w)), w.value, words

    b3, w3, _ = plan(1000, 25)                       # C3: one 16-warp block per SM
    assert w3 == 16 and 0 < b3 <= budget
    b4, w4, words4 = plan(10000, 25)                 # C4: the longer bit row costs one warp
    assert w4 == 15 and b4 <= budget and b4 % 16 == 0
    assert (b4 // w4) - (b3 // w3) == 8 * (words4 - 16)          # per-warp bytes differ by the bit row only
    _, w32, _ = plan(1000, 32)                       # n_grid 32: float table 2 x 32 x 36 floats
    assert 1 <= w32 < 16
    b, w, _ = plan(1000, 25, max_warps=12)           # tuning override
    assert w == 12 and b == (b3 // 16) * 12
    assert plan(1_600_000, 25)[0] == 0               # one bit row alone exceeds the budget: an error upstream, not a crash
    strides = {G: (((G + 3) >> 2) | 1) * 4 for G in range(2, 33)}
    assert all(s >= G and s % 4 == 0 and (s // 4) % 2 == 1 for G, s in strides.items())
