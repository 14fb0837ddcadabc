"""CPU suite: motion's block pipeline (8-bit load, 3-D DCT, filters, resampling inverse, 8-bit store) on the
emulated kernels vs the restated reference loop."""
import pytest

from tests import cases
from tests.emu import emu


@pytest.fixture(scope="module")
def lib():
    return emu.load()


def test_identity_roundtrip_u8(lib):
    assert cases.check_motion(lib, (4, 8, 8)) == 0.0
    assert cases.check_motion(lib, (1, 16, 24)) == 0.0           # motion's default depth-1 blocks
    assert cases.check_motion(lib, (8, 12, 20), prec="d") == 0.0


def test_filters(lib):
    cases.check_motion(lib, (8, 16, 16), damp=0.0, bandpass=((0, 0, 0), (4, 8, 8)))                # low-pass box
    cases.check_motion(lib, (8, 16, 16), boost=1.5, bandpass=((1, 2, 2), (6, 12, 12)), preserve_dc="dc")
    cases.check_motion(lib, (4, 16, 16), damp=0.25, bandpass=((0, 1, 1), (4, 16, 16)), preserve_dc="grey")
    cases.check_motion(lib, (4, 8, 8), quant=0.02)
    cases.check_motion(lib, (4, 8, 8), threshold=(0.001, 0.5))


def test_resample_by_spectral_pad_and_crop(lib):
    cases.check_motion(lib, (4, 8, 8), scaled=(4, 16, 16))        # 2x spatial upscale: zero-padded spectrum
    cases.check_motion(lib, (8, 16, 16), scaled=(4, 8, 12))       # downscale: cropped spectrum
    cases.check_motion(lib, (4, 10, 12), scaled=(6, 10, 9))       # mixed


def test_float_pixels(lib):
    cases.check_motion(lib, (4, 8, 16), float_pixels=True)
    cases.check_motion(lib, (4, 8, 16), scaled=(4, 12, 16), float_pixels=True, quant=0.01)


def test_all_blocks_of_a_volume_at_once(lib):
    """motion -b 8x8x8 (+ --quant): three per-axis plans over the whole volume instead of a plan per block"""
    cases.check_motion_tiled(lib, (16, 16, 24), (8, 8, 8))
    cases.check_motion_tiled(lib, (16, 16, 24), (8, 8, 8), quant=0.05)
    cases.check_motion_tiled(lib, (8, 32, 16), (4, 16, 8), quant=0.02)
    cases.check_motion_tiled(lib, (2, 24, 40), (1, 8, 8))                   # depth-1 blocks: a 2-D block DCT per frame


@pytest.mark.parametrize("kw", [dict(spec="shift"), dict(spec="flat"), dict(spec="abs"), dict(spec="copy", quant=0.02),
                                dict(ispec="shift"), dict(ispec="flat"), dict(ispec="copy"),
                                dict(spec="shift", ispec="shift"), dict(spec="flat", boost=1.5, bandpass=((0, 1, 1), (4, 6, 6)))])
def test_spectrogram_modes(lib, kw):
    """motion --spec / --ispec (motion.c:627-637, 755-766): the spectrogram stages as pointwise kernels around one transform"""
    cases.check_motion(lib, (4, 8, 8), **kw)
    cases.check_motion(lib, (4, 8, 12), float_pixels=True, **kw)


def test_spec_then_ispec_reproduces_the_pels(lib):
    """a flat spectrogram of a block, fed back through --ispec flat, returns the source pels"""
    import numpy as np
    from dspfun_b200 import motion as gmotion
    rng = np.random.default_rng(3)
    pels = rng.integers(100, 156, (4, 8, 8)).astype(np.float32) / 255.0
    a = gmotion.Motion((4, 8, 8), float_pixels=True, spec="copy", lib=lib)
    s = a.process(pels)
    a.destroy()
    b = gmotion.Motion((4, 8, 8), float_pixels=True, ispec="copy", lib=lib)
    back = b.process(s)
    b.destroy()
    assert np.abs(back - pels).max() < 2e-5


def test_sequential_reference_stages_are_refused(lib):
    from dspfun_b200 import capi, motion as gmotion
    for kw in (dict(coeff_limit=10), dict(expr="c*2"), dict(dither=True), dict(linear=True)):
        with pytest.raises(capi.DspDctError):
            gmotion.Motion((4, 8, 8), lib=lib, **kw)
