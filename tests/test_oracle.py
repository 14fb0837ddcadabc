"""Pins the oracle: definition-based C restatement and the pocketfft arbiter vs FFTW-generated golden vectors,
Appendix-B known answers, algebraic invariants, and the reference's own speclib (oracle/_ref) where built."""
import ctypes
import os

import numpy as np
import pytest

from oracle import dct as odct
from oracle import pipelines as pl

SIZES = [2, 3, 4, 8, 12, 15, 16, 17, 32, 64, 128, 256, 512, 1024]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("typ,kind", [(2, odct.REDFT10), (3, odct.REDFT01)])
def test_definition_matches_fftw_golden(golden, n, typ, kind):
    x = np.linspace(0, n - 1, n)
    for prec, dt, tol in (("single", np.float32, 2e-6), ("double", np.float64, 1e-14), ("longdouble", np.float64, 1e-15)):
        ref = golden["%s_dct_%d_%d" % (prec, typ, n)].astype(np.float64)
        got = odct.dctn_def(x.astype(dt), [kind]).astype(np.float64)
        assert odct.rel_l2(got, ref) < tol, (prec, n)


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("typ,kind", [(2, odct.REDFT10), (3, odct.REDFT01)])
def test_pocketfft_matches_fftw_golden(golden, n, typ, kind):
    x = np.linspace(0, n - 1, n)
    ref = golden["longdouble_dct_%d_%d" % (typ, n)].astype(np.float64)
    assert odct.rel_l2(odct.dctn_fast(x, [kind]), ref) < 1e-14


def test_appendix_b_known_answers():
    y = odct.dctn_def(np.array([1., 2., 3., 4.]), [odct.REDFT10])
    np.testing.assert_allclose(y, [20, -6.308644059797899, 0, -0.448341529167965], atol=1e-14)
    np.testing.assert_allclose(odct.dctn_def(y, [odct.REDFT01]) / 8, [1, 2, 3, 4], atol=1e-14)
    a = np.arange(12, dtype=np.float64).reshape(3, 4)
    Y = odct.dctn_def(a, [odct.REDFT10] * 2)
    np.testing.assert_allclose(Y[0], [264, -37.85186435879, 0, -2.690049175008], atol=1e-9)
    np.testing.assert_allclose(Y[1, 0], -110.8512516844, atol=1e-9)
    np.testing.assert_allclose(odct.dctn_def(Y, [odct.REDFT01] * 2) / 48, a, atol=1e-13)
    # n = 1 is legal (motion's depth-1 blocks): REDFT10 -> 2x, REDFT01 -> x
    assert odct.dctn_def(np.array([3.0]), [odct.REDFT10])[0] == 6.0
    assert odct.dctn_def(np.array([3.0]), [odct.REDFT01])[0] == 3.0


def test_many_interface_addressing_interleaved_and_embed():
    rng = np.random.default_rng(5)
    h, w, d = 6, 10, 3
    x = rng.random((h, w, d))
    out = np.empty(h * w * d)
    odct.r2r_many(x.reshape(-1).copy(), 2, [h, w], d, None, d, 1, out, None, d, 1, [odct.REDFT10] * 2)
    np.testing.assert_allclose(out.reshape(h, w, d), odct.dctn_fast(x, [odct.REDFT10] * 2, axes=(0, 1)), atol=1e-12)
    # motion-style: logical box inside a larger physical box (motion/motion.c:535-538)
    phys = (5, 7, 9)
    log = (3, 4, 6)
    buf = rng.random(phys)
    ref = buf.copy()
    ref[:3, :4, :6] = odct.dctn_fast(buf[:3, :4, :6], [odct.REDFT10] * 3)
    flat = buf.reshape(-1).copy()
    odct.r2r_many(flat, 3, log, 1, phys, 1, 0, flat, phys, 1, 0, [odct.REDFT10] * 3)
    np.testing.assert_allclose(flat.reshape(phys), ref, atol=1e-12)


@pytest.mark.parametrize("shape", [(16, 16), (12, 20), (7, 9, 5), (1, 8)])
def test_roundtrip_and_fast_vs_definition(shape):
    rng = np.random.default_rng(1)
    x = rng.standard_normal(shape)
    k2 = [odct.REDFT10] * len(shape)
    k3 = [odct.REDFT01] * len(shape)
    Y = odct.dctn_def(x, k2)
    assert odct.rel_l2(odct.dctn_fast(x, k2), Y) < 1e-14
    scale = np.prod([2 * s for s in shape])
    np.testing.assert_allclose(odct.dctn_def(Y, k3) / scale, x, atol=1e-12)


def test_base16_known_answer_and_roundtrip():
    assert pl.base16enc(bytes([0x3F])) == "PD"
    raw = np.array([0.25, 1.5, -3.0]).tobytes()
    assert pl.base16dec(pl.base16enc(raw)) == raw


@pytest.mark.parametrize("preset", ["abs", "shift", "flat", "sign", "copy"])
def test_spec_ispec_pipeline_inverts(preset):
    rng = np.random.default_rng(7)
    px = (rng.integers(0, 256, (16, 24, 3)) / 255.0).astype(np.float64)
    spec, DC = pl.spec_forward(px, preset)
    if preset == "sign":
        assert set(np.unique(spec.reshape(-1)[3:])) <= {0.0, 1.0}
        return
    if preset == "abs":
        sm, _ = pl.spec_forward(px, "sign")
        signmap = np.round(sm * 255).astype(np.uint8)
        signmap.reshape(-1)[:3] = np.round(DC * 255).astype(np.uint8)
        back = pl.ispec_inverse(spec, DC, preset, signmap=None)
        # without the sign map |coefficients| cannot invert; with it they do (DC quantised to 8 bits)
        back = pl.ispec_inverse(spec, DC, preset, signmap=signmap)
        assert np.abs(back - px).max() < 2e-2
        return
    back = pl.ispec_inverse(spec, DC, preset)
    assert np.abs(back - px).max() < 1e-12


def _speclib(tag):
    path = os.path.join(os.path.dirname(odct.__file__), "_ref", "libspeclib_%s.so" % tag)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref not built (reference checkout absent)")
    return ctypes.CDLL(path)


def test_reference_speclib_agrees_with_restated_formulas():
    """The reference's own include/speclib.c (compiled into oracle/_ref) vs the restated spec formulas
    (spec/spec.c:110-139 == include/speclib.c:91-178)."""
    lib = _speclib("dl")

    class P(ctypes.Structure):
        _fields_ = [("scaletype", ctypes.c_int), ("signtype", ctypes.c_int)]
    norm = getattr(lib, "spec_normalization_pcil")
    norm.restype = ctypes.c_longdouble
    norm.argtypes = [ctypes.c_size_t]
    for n, want in enumerate([1, np.sqrt(2), 2, 2 * np.sqrt(2)]):
        assert abs(float(norm(n)) - want) < 1e-15
        assert abs(float(odct._lib().ref_spec_normalization(n)) - want) < 1e-15
    create = getattr(lib, "spec_create_pcil")
    create.restype = ctypes.c_void_p
    create.argtypes = [ctypes.POINTER(P), ctypes.c_double, ctypes.c_double]
    scale = getattr(lib, "spec_scale_pcil")
    scale.restype = ctypes.c_longdouble
    scale.argtypes = [ctypes.c_void_p, ctypes.c_longdouble]
    unscale = lib.spec_unscale
    unscale.restype = ctypes.c_longdouble
    unscale.argtypes = [ctypes.c_void_p, ctypes.c_longdouble]
    # enum values: keyed_enum puts _none at 0; scaletype {linear=1, log=2}; signtype {abs=1, shift=2, saturate=3}
    gain, mx = 127.5 * 64.0, 0.7
    sp = create(ctypes.byref(P(2, 2)), mx, gain)          # log + shift  == spec preset "shift"
    for c in [-0.9, -0.01, 0.0, 0.3, 0.69]:
        ref = float(scale(sp, c))
        v = np.longdouble(c) * gain
        mine = (np.copysign(np.log1p(abs(v)), v) / np.log1p(np.longdouble(gain * mx)) / 2 + 0.5) * 254 / 255
        assert abs(ref - float(mine)) < 1e-15
        assert abs(float(unscale(sp, ref)) - c) < 1e-15
