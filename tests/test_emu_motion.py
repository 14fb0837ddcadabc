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
