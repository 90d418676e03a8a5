"""Randomised shapes, densities, grids and bandwidth options: the production entropy epilogue (closed forms + soft cells)
against the literal cell-by-cell cube of the validation mode (tools/stress_epilogue.py)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [20261020, 7])
def test_production_epilogue_random_configurations(seed):
    import stress_epilogue
    worst = stress_epilogue.run(ncfg=30, seed=seed, verbose=False)
    assert worst <= 5e-14, worst
