"""Synthetic workloads of SURVEY 8(d) (bench.py and the parity tests use the same generator).

RNG: numpy PCG64, base seed 20160418 + config index.  target: n i.i.d. N(0,1), sorted decreasing
(mirrors LIB:1626-1633).  features: row density log-uniform on [3/n, 0.05] (count.thres.low=3 /
count.thres.high=50 at n=1000, manifest:147,17), Bernoulli entries, rows with fewer than 3 ones
redrawn; 10 planted mutually-exclusive "driver" rows whose ones fall in the top 20 % of the target;
20 % of rows are exact copies of the preceding row (amplicon blocks: exact score ties are normal on
real data); a few all-zero / all-one rows to hit cond.assoc's `return(0)` guard.
"""
import numpy as np

BASE_SEED = 20160418

CONFIGS = {
    # name: (n samples, F features, iterations, seed rows, targets)
    "C1_gpunit_like": (82, 17800, 2, 1, 1),
    "C2_500x20k": (500, 20000, 5, 1, 1),
    "C3_1kx100k": (1000, 100000, 10, 1, 1),
    "C4_10kx500k": (10000, 500000, 10, 3, 1),
    "C5_256targets_1kx100k": (1000, 100000, 1, 1, 256),
}
CONFIG_INDEX = {k: i for i, k in enumerate(CONFIGS)}


def make_target(n, rng):
    return np.sort(rng.standard_normal(n))[::-1].copy()


def make_features(n, F, rng, n_drivers=10, dup_frac=0.20, n_degenerate=4, row_lo=0, row_hi=None, block=8192):
    """Rows [row_lo, row_hi) of the F x n uint8 matrix.  Every row's content depends only on (seed, row
    block), so shards generated on different ranks are slices of the same global matrix."""
    row_hi = F if row_hi is None else row_hi
    out = np.empty((row_hi - row_lo, n), dtype=np.uint8)
    lo_p, hi_p = 3.0 / n, max(0.05, 6.0 / n)
    top = max(n // 5, n_drivers)
    driver_pool = rng.permutation(top)          # positions among the top 20 % of the (sorted) target
    per = max(1, min(top // n_drivers, max(3, n // 50)))
    b0 = (row_lo // block) * block
    base = rng.integers(0, 2**63 - 1)
    for bs in range(b0, row_hi, block):
        be = min(bs + block, F)
        r = np.random.Generator(np.random.PCG64([base, bs]))
        p = np.exp(r.uniform(np.log(lo_p), np.log(hi_p), size=be - bs))
        m = (r.random((be - bs, n), dtype=np.float32) < p[:, None].astype(np.float32)).astype(np.uint8)
        short = np.flatnonzero(m.sum(axis=1) < 3)
        for i in short:                         # redraw rows with fewer than 3 ones
            idx = r.choice(n, size=3, replace=False)
            m[i] = 0
            m[i, idx] = 1
        dup = r.random(be - bs) < dup_frac
        dup[0] = False
        for i in np.flatnonzero(dup):           # exact copy of the preceding row
            m[i] = m[i - 1]
        lo, hi = max(bs, row_lo), min(be, row_hi)
        out[lo - row_lo: hi - row_lo] = m[lo - bs: hi - bs]
    # planted drivers at fixed global rows, degenerate rows after them
    stride = max(1, F // (n_drivers + 1))
    for d in range(n_drivers):
        g = (d + 1) * stride - 1
        if row_lo <= g < row_hi:
            out[g - row_lo] = 0
            out[g - row_lo, driver_pool[d * per:(d + 1) * per]] = 1
    for d in range(n_degenerate):
        g = d * max(1, F // (n_degenerate + 1)) + 1
        if row_lo <= g < row_hi and F > 4 * n_degenerate:
            out[g - row_lo] = d % 2
    return out


def driver_rows(F, n_drivers=10):
    stride = max(1, F // (n_drivers + 1))
    return [(d + 1) * stride - 1 for d in range(n_drivers)]


def make_workload(name=None, n=None, F=None, n_seed_rows=1, n_targets=1, seed_offset=0, row_lo=0, row_hi=None,
                  null_seed=False):
    """Returns dict(target(s), m2 (uint8 rows [row_lo,row_hi)), seed (uint8[n]), iters)."""
    iters = 5
    if name is not None:
        n, F0, iters, n_seed_rows, n_targets = CONFIGS[name]
        F = F or F0                                   # F given: a row-count override of the named shape
        seed_offset = CONFIG_INDEX[name]
    rng = np.random.Generator(np.random.PCG64(BASE_SEED + seed_offset))
    target = make_target(n, rng)
    frng = np.random.Generator(np.random.PCG64(BASE_SEED + seed_offset + 1000))
    m2 = make_features(n, F, frng, row_lo=row_lo, row_hi=row_hi)
    # the seed = OR of the first n_seed_rows planted drivers (regenerated identically on every rank)
    srng = np.random.Generator(np.random.PCG64(BASE_SEED + seed_offset + 1000))
    top = max(n // 5, 10)
    pool = srng.permutation(top)
    per = max(1, min(top // 10, max(3, n // 50)))
    seed = np.zeros(n, dtype=np.uint8)
    if not null_seed:
        for d in range(n_seed_rows):
            seed[pool[d * per:(d + 1) * per]] = 1
    targets = target
    if n_targets > 1:
        trng = np.random.Generator(np.random.PCG64(BASE_SEED + seed_offset + 2000))
        targets = np.vstack([target[None, :], trng.standard_normal((n_targets - 1, n))])
    return dict(target=targets, m2=m2, seed=seed, iters=iters, n=n, F=F, n_targets=n_targets)
