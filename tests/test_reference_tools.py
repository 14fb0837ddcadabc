"""The reference's UNMODIFIED spec.c / ispec.c, compiled against shim/fftw3.h (oracle/Makefile `reftools`), run as
processes.  The *_emu builds link the host emulation of the kernels (CPU suite); the *_gpu builds link the product
library (GPU suite).  Their output must agree with the restated pipeline of the oracle and with the fused GPU path."""
import os
import subprocess

import numpy as np
import pytest

from dspfun_b200 import spec as gspec
from oracle import dct as od
from oracle import pipelines as pl
from tests import dspraw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def _tool(name):
    path = os.path.join(REF, name)
    if not os.path.exists(path):
        if os.path.exists("/root/reference/spec/spec.c"):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "reftools"], check=True)
        if not os.path.exists(path):
            pytest.skip("oracle/_ref/%s not built (reference checkout absent)" % name)
    return path


def _run_chain(kind, prec, preset, tmp_path):
    rng = np.random.default_rng(12)
    h, w, d = 48, 80, 3
    px = rng.integers(0, 256, (h, w, d)) / 255.0
    src, spc, back = [str(tmp_path / n) for n in ("in.dspraw", "spec.dspraw", "back.dspraw")]
    dspraw.write(src, px)
    subprocess.run([_tool("spec_%s_%s" % (kind, prec)), "-t", preset, src, spc], check=True)
    s_tool, props = dspraw.read(spc)
    dt = np.float32 if prec == "f" else np.float64
    I = np.float64 if prec == "f" else np.longdouble
    s_ref, dc_ref = pl.spec_forward(px.astype(dt), preset, intermediate=I)
    assert od.rel_l2(s_tool, s_ref) < (1e-5 if prec == "f" else 1e-12)
    dc_tool = np.frombuffer(gspec.base16dec(props["DC"]), dtype=np.float64)
    np.testing.assert_allclose(dc_tool, dc_ref, rtol=0, atol=1e-6 if prec == "f" else 1e-13)
    subprocess.run([_tool("ispec_%s_%s" % (kind, prec)), "-t", preset, spc, back], check=True)
    b_tool, _ = dspraw.read(back)
    assert np.abs(b_tool - px).max() < (2e-5 if prec == "f" else 1e-12)
    assert np.array_equal(pl.quantize_unorm(b_tool, 8), np.round(px * 255).astype(np.uint8))
    return s_tool


@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("preset", ["shift", "flat", "copy"])
def test_unmodified_reference_tools_on_emulated_kernels(prec, preset, tmp_path):
    _run_chain("emu", prec, preset, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("preset", ["shift", "flat", "copy"])
def test_unmodified_reference_tools_on_gpu(prec, preset, tmp_path):
    _run_chain("gpu", prec, preset, tmp_path)


def _run_draw(kind, prec, tmp_path):
    """applybasis/draw.c (call site P8, fftw_plan_r2r_2d REDFT01 x REDFT01 on a planar canvas): the unmodified tool's
    canvas == DC/2 + sum of the requested cosine components, by the definition of REDFT01"""
    out = str(tmp_path / "canvas.dspraw")
    w, h = 96, 64
    comps = [(3, 5, 0.25), (0, 7, 0.125), (10, 0, 0.2)]
    args = [_tool("draw_%s_%s" % (kind, prec)), "-b", "%dx%d" % (w, h)]
    for x, y, s in comps:
        args += ["-f", "%dx%d:%g" % (x, y, s)]
    subprocess.run(args + [out], check=True)
    img, _ = dspraw.read(out)
    coefs = np.zeros((h, w))
    for x, y, s in comps:
        coefs[y, x] = s / 4                                       # draw.c:69
    coefs[0, 0] += 0.5                                             # draw.c:70
    ref = od.dctn_def(coefs, [od.REDFT01, od.REDFT01])
    assert od.rel_l2(img[:, :, 0], ref) < (1e-5 if prec == "f" else 1e-12)


@pytest.mark.parametrize("prec", ["f", "d"])
def test_unmodified_draw_on_emulated_kernels(prec, tmp_path):
    _run_draw("emu", prec, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["f", "d"])
def test_unmodified_draw_on_gpu(prec, tmp_path):
    _run_draw("gpu", prec, tmp_path)
