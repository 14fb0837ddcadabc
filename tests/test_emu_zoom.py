"""CPU suite: zoom's DCT-domain resampling on the emulated kernels vs the restated reference loops."""
import numpy as np
import pytest

from oracle import pipelines as pl
from tests import cases
from tests.emu import emu


@pytest.fixture(scope="module")
def lib():
    return emu.load()


@pytest.mark.parametrize("prec", ["f", "d"])
def test_interpolated_default_basis(lib, prec):
    path, got, want = cases.check_zoom(lib, prec, 16, 24, scale=2)
    assert path == "shifted-dct"            # integer scaled size: four phase-shifted inverse DCTs
    assert cases.check_zoom(lib, prec, 16, 20, scale=(5, 3))[0] == "dense"      # 26.67 x 33.33: dense contractions
    # SURVEY finding 3: with the default basis the even output samples reproduce the input pixels
    cases.check_zoom(lib, prec, 12, 20, scale=(3, 2))
    cases.check_zoom(lib, prec, 16, 16, scale=0.5)
    cases.check_zoom(lib, prec, 16, 24, scale=2, pos=(3.5, 1.25), view=(20, 12))


@pytest.mark.parametrize("prec", ["f", "d"])
def test_native_basis_is_spectral_zero_pad_or_crop(lib, prec):
    path, _, _ = cases.check_zoom(lib, prec, 16, 24, scale=2, basis="native")
    assert path == "inverse-dct"
    path, _, _ = cases.check_zoom(lib, prec, 16, 24, scale=(1, 2), basis="native")
    assert path == "inverse-dct"
    path, _, _ = cases.check_zoom(lib, prec, 16, 24, scale=2, basis="native", pos=(1.0, 0.0))
    assert path == "dense"
    cases.check_zoom(lib, prec, 10, 14, xscale=(3, 2), yscale=2, basis="native")


@pytest.mark.parametrize("prec", ["f", "d"])
def test_centered_basis(lib, prec):
    _, got, _ = cases.check_zoom(lib, prec, 12, 16, scale=2, basis="centered")


def test_interpolated_even_samples_are_the_input():
    rng = np.random.default_rng(1)
    px = rng.random((8, 12, 3))
    out = pl.zoom_synthesise(px, scale=(2, 1))
    np.testing.assert_allclose(out[::2, ::2], px, atol=1e-12)


@pytest.mark.parametrize("prec", ["f", "d"])
def test_shifted_dct_path_equals_dense_synthesis(lib, prec, monkeypatch):
    """the two evaluations of the interpolated basis agree (offsets, partial views, down-scaling, anisotropic scale)"""
    import numpy as np
    from oracle import dct as od
    for kw in (dict(scale=2), dict(scale=2, pos=(3.5, 1.25), view=(20, 12)), dict(scale=0.5), dict(xscale=(3, 2), yscale=2),
               dict(scale=3, pos=(0.0, 7.0), view=(11, 9))):
        monkeypatch.delenv("DSP_ZOOM_NO_SHIFT", raising=False)
        p1, fast, _ = cases.check_zoom(lib, prec, 16, 24, **kw)
        monkeypatch.setenv("DSP_ZOOM_NO_SHIFT", "1")
        p2, dense, _ = cases.check_zoom(lib, prec, 16, 24, **kw)
        assert (p1, p2) == ("shifted-dct", "dense"), (kw, p1, p2)
        assert od.rel_l2(fast, dense) < cases.OK[prec]
