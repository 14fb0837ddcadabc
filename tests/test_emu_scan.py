"""CPU suite: scan's fused masked inverse + accumulate on the emulated kernels vs the restated reference loop."""
import numpy as np
import pytest

from dspfun_b200 import scan as gscan
from tests import cases
from tests.emu import emu


@pytest.fixture(scope="module")
def lib():
    return emu.load()


@pytest.mark.parametrize("prec", ["f", "d"])
def test_scan_diagonal_and_raster(lib, prec):
    cases.check_scan(lib, prec, 8, 8, 3, "diagonal")
    cases.check_scan(lib, prec, 16, 12, 3, "horizontal", step=16)
    cases.check_scan(lib, prec, 32, 32, 1, "diagonal", step=3)


def test_diagonal_order_matches_reference_readme():
    """scan/README.md:121-129: the 8x8 diagonal order's index serialisation is idx[y][x] = x + y."""
    idx = gscan.order_diagonal(8, 8)
    assert idx[0].tolist() == list(range(8)) and idx[7, 7] == 14 and idx[3, 4] == 7
    assert np.array_equal(idx, idx.T)
