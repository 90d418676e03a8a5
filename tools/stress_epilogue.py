"""Randomised stress of the production entropy epilogue (closed forms + single-precision soft cells) against the literal
cell-by-cell cube of the validation mode, on the GPU: many shapes, row and seed densities (sparse, dense, complemented),
grids and bandwidth options.  Prints the worst |dCMI| per configuration and fails above 5e-14.
  python tools/stress_epilogue.py [n_configs]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import revealer_b200 as rb  # noqa: E402



def run(ncfg=40, seed=20261020, verbose=True):
  rng = np.random.default_rng(seed)
  worst_all = 0.0
  for c in range(ncfg):
      n = int(rng.choice([40, 64, 82, 129, 200, 333, 500, 1000, 2048, 3000]))
      F = int(rng.integers(200, 1500))
      G = int(rng.choice([5, 8, 13, 17, 25, 25, 25, 32]))
      mult = float(rng.choice([0.1, 0.25, 0.25, 0.5, 1.0]))
      shrink = float(rng.choice([0.0, -0.5, -0.75, -0.75, -0.9]))
      x = rng.standard_normal(n) if rng.random() < 0.7 else rng.exponential(size=n)
      dens = np.exp(rng.uniform(np.log(2.0 / n), np.log(0.98), size=F))
      m2 = (rng.random((F, n)) < dens[:, None]).astype(np.uint8)
      zd = float(np.exp(rng.uniform(np.log(2.0 / n), np.log(0.9))))
      z = (rng.random(n) < zd).astype(np.uint8)
      if z.sum() in (0, n):
          z[0] = 1 - z[0]
      with rb.RevealerContext(0) as ctx:
          ctx.set_options(n_grid=G, bandwidth_mult=mult, bandwidth_shrink=shrink, negative=bool(rng.random() < 0.3))
          ctx.load_matrix(m2)
          ctx.set_targets(x)
          ctx.set_seed(z)
          ctx.set_prune_theta(0.0)
          a = ctx.score_once()[0]
          ctx.set_dense_x(True)
          b = ctx.score_once()[0]
      # a round-off-negative CMI gives NaN (sqrt of a negative number, as in the reference, LIB:2614): where only one of the
      # two modes is NaN the other's CMI must itself be within the rounding band of zero
      info = lambda v: -0.5 * np.log1p(-np.minimum(v ** 2, 1 - 1e-16))
      one = np.isfinite(a) != np.isfinite(b)
      if one.any():
          other = np.where(np.isfinite(a), a, b)[one]
          assert np.all(info(other) <= 5e-14), (c, n, F, G, "NaN where the other mode has CMI", float(info(other).max()))
      fin = np.isfinite(a) & np.isfinite(b)
      ia, ib = (-0.5 * np.log1p(-np.minimum(v[fin] ** 2, 1 - 1e-16)) for v in (a, b))
      w = float(np.max(np.abs(ia - ib))) if fin.any() else 0.0
      big = np.maximum(ia, ib) > 5e-14                       # below the rounding band a score may come out as +-0
      sgn = np.array_equal(np.sign(a[fin])[big], np.sign(b[fin])[big])
      worst_all = max(worst_all, w)
      if verbose:
          print("cfg %2d n=%4d F=%4d G=%2d mult=%.2f shrink=%.2f seed=%4d  worst |dCMI| %.2e  signs %s  nan-mismatch %d" % (c, n, F, G, mult, shrink, int(z.sum()), w, sgn, int(one.sum())))
      assert w <= 5e-14 and sgn, (c, w)
  if verbose:
      print("worst over all configurations: %.3e" % worst_all)
  return worst_all


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 40)
