"""CPU suite: spec / ispec fused stages (host emulation of the kernels) vs the restated reference pipelines."""
import pytest

from dspfun_b200 import spec as gspec
from tests import cases
from tests.emu import emu


@pytest.fixture(scope="module")
def lib():
    return emu.load()


@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("shape", [(16, 24, 3), (64, 64, 1), (20, 12, 3), (33, 50, 4)])
def test_presets(lib, prec, shape):
    cases.check_spec_presets(lib, prec, *shape)


@pytest.mark.parametrize("prec", ["f", "d"])
def test_option_overrides(lib, prec):
    cases.check_spec_options(lib, prec)


@pytest.mark.parametrize("prec", ["f", "d"])
def test_c1_roundtrip_small(lib, prec):
    cases.check_spec_c1_roundtrip(lib, prec, 64, 96, 3, "shift")
    cases.check_spec_c1_roundtrip(lib, prec, 64, 96, 3, "flat")


def test_dc_property_wire_format():
    import numpy as np
    dc = np.array([0.25, 0.5, 0.4375])
    s = gspec.base16enc(dc.tobytes())
    assert len(s) == 48 and set(s) <= set("ABCDEFGHIJKLMNOP")
    assert np.array_equal(np.frombuffer(gspec.base16dec(s), dtype=np.float64), dc)
    assert gspec.base16enc(bytes([0x3F])) == "PD"


def test_ispec_needs_dc(lib):
    import numpy as np
    with pytest.raises(ValueError):
        gspec.ispec(np.zeros((8, 8, 3), np.float32), None, "abs", lib=lib)
