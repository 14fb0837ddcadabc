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
    if not os.path.exists(path) and "_emu_" in name:
        from tests.emu import emu
        emu.load()                                  # the *_emu_* tools link tests/emu/libdspdct_emu.so: make sure it is there first
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


# ======================================================================================================================
# scan.c, motion.c, zoom.c: the reference's own tool code (unmodified) over tools/ffstub + tools/wandstub.  These runs pin
# oracle/pipelines.py: scan_frames, motion_block, zoom_synthesise -- the restatements every GPU parity test of those
# pipelines is compared with -- to what the reference's code computes.
# ======================================================================================================================
def _dims(d, h, w):
    return "%dx%dx%d" % (w, h, d)


def _tie_equal(got, want, pel, what):
    """8-bit pels equal, except where the restatement's unrounded value sits on a rounding tie"""
    diff = got != want
    if diff.any():
        assert np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1, what
        frac = np.abs(np.asarray(pel, dtype=np.float64))
        assert (np.abs(frac - np.floor(frac) - 0.5)[diff] < 2e-3).all(), what


MOTION_CASES = {   # name -> (tool arguments, restatement keyword arguments); volume 8 x 16 x 24 (d, h, w)
    "identity": ([], {}),
    "lowpass": (["-p", "0x0x0-8x6x4", "-D", "0"], dict(damp=0.0, bandpass=((0, 0, 0), (4, 6, 8)))),
    "boost_damp_grey": (["-p", "2x1x1-12x10x6", "-B", "1.5", "-D", "0.25", "--preserve-dc=grey"],
                        dict(boost=1.5, damp=0.25, bandpass=((1, 1, 2), (6, 10, 12)), preserve_dc="grey")),
    "preserve_dc": (["-p", "2x1x1-12x10x6", "-D", "0", "--preserve-dc"], dict(damp=0.0, bandpass=((1, 1, 2), (6, 10, 12)), preserve_dc="dc")),
    "quant": (["-q", "0.02"], dict(quant=0.02)),
    "threshold": (["--threshold", "0.001-0.5"], dict(threshold=(0.001, 0.5))),
    "upscale": (["-s", "48x32x8"], dict(scaled=(8, 32, 48))),
    "downscale": (["-s", "12x12x4"], dict(scaled=(4, 12, 12))),
    "spec_abs": (["--spectrogram"], dict(spec="abs")),
    "spec_shift": (["--spectrogram=shift"], dict(spec="shift")),
    "spec_flat": (["--spectrogram=flat"], dict(spec="flat")),
    "spec_copy_quant": (["--spectrogram=copy", "-q", "0.02"], dict(spec="copy", quant=0.02)),
    "ispec_shift": (["--ispectrogram"], dict(ispec="shift")),
    "ispec_flat": (["--ispectrogram=flat"], dict(ispec="flat")),
}


def _run_motion(kind, prec, case, tmp_path, fmt="gray"):
    args, kw = MOTION_CASES[case]
    D, H, W = 8, 16, 24
    rng = np.random.default_rng(31)
    C = np.float32 if prec == "f" else np.float64
    src, dst = str(tmp_path / "in.dspv"), str(tmp_path / "out.dspv")
    if fmt == "grayf32le":
        comps = [rng.random((D, H, W)).astype(np.float32)]
    elif fmt == "yuv420p":
        comps = [rng.integers(0, 256, (D, H, W)).astype(np.uint8)] + [rng.integers(0, 256, (D, H // 2, W // 2)).astype(np.uint8) for _ in range(2)]
    else:
        comps = [rng.integers(0, 256, (D, H, W)).astype(np.uint8)]
    dspraw.write_video(src, fmt, W, H, [[c[z] for c in comps] for z in range(D)])
    subprocess.run([_tool("motion_%s_%s" % (kind, prec)), "-Q", "-b", _dims(D, H, W)] + args + [src, dst], check=True)
    ofmt, ow, oh, frames = dspraw.read_video(dst)
    assert ofmt == fmt
    for ci, vol in enumerate(comps):
        d, h, w = vol.shape
        kw_c = dict(kw)
        if ci and "scaled" in kw_c:
            kw_c["scaled"] = (kw_c["scaled"][0], kw_c["scaled"][1] // 2, kw_c["scaled"][2] // 2)
        scaled = kw_c.get("scaled", (d, h, w))
        minbuf = tuple(max(a, b) for a, b in zip((d, h, w), scaled))
        stage = np.zeros(minbuf, dtype=vol.dtype)
        stage[:d, :h, :w] = vol
        want, _, pel = pl.motion_block(stage, (d, h, w), coeff=C, intermediate=np.longdouble, **kw_c)
        got = np.stack([f[ci] for f in frames])
        assert got.shape == tuple(scaled), (got.shape, scaled)
        want = want[:scaled[0], :scaled[1], :scaled[2]]
        if fmt == "grayf32le":
            assert np.abs(got - want).max() < (2e-5 if prec == "f" else 1e-6) * max(1.0, float(np.abs(want).max())), case
        else:
            _tie_equal(got, want, pel, case)


@pytest.mark.parametrize("case", sorted(MOTION_CASES))
def test_reference_motion_tool_pins_the_restatement(case, tmp_path):
    """motion.c's block loop (motion/motion.c:591-811) == oracle/pipelines.py: motion_block, option by option"""
    _run_motion("emu", "f", case, tmp_path, "grayf32le" if case in ("spec_flat", "spec_copy_quant") else "gray")      # motion.c:313-315


@pytest.mark.parametrize("fmt,prec,case", [("yuv420p", "f", "identity"), ("yuv420p", "f", "quant"), ("yuv420p", "f", "upscale"),
                                           ("grayf32le", "f", "lowpass"), ("grayf32le", "d", "quant"), ("gray", "d", "boost_damp_grey")])
def test_reference_motion_tool_formats_and_precisions(fmt, prec, case, tmp_path):
    _run_motion("emu", prec, case, tmp_path, fmt)


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,prec,case", [("gray", "f", "identity"), ("gray", "f", "lowpass"), ("yuv420p", "f", "quant"), ("gray", "f", "upscale"),
                                           ("gray", "f", "spec_shift"), ("gray", "f", "ispec_flat"), ("grayf32le", "d", "quant")])
def test_reference_motion_tool_on_gpu(fmt, prec, case, tmp_path):
    """the same unmodified tool linked to the product library: the reference's loops around GPU transforms"""
    _run_motion("gpu", prec, case, tmp_path, fmt)


def test_reference_motion_tool_tiled_blocks(tmp_path):
    """`motion -b 8x8x4` walks the blocks of the volume one by one (motion.c:591-615): each == motion_block of that block"""
    D, H, W, bd, bh, bw = 8, 16, 24, 4, 8, 8
    vol = np.random.default_rng(8).integers(0, 256, (D, H, W)).astype(np.uint8)
    src, dst = str(tmp_path / "in.dspv"), str(tmp_path / "out.dspv")
    dspraw.write_video(src, "gray", W, H, [[vol[z]] for z in range(D)])
    subprocess.run([_tool("motion_emu_f"), "-Q", "-b", _dims(bd, bh, bw), "-q", "0.05", src, dst], check=True)
    _, _, _, frames = dspraw.read_video(dst)
    got = np.stack([f[0] for f in frames])
    for z in range(0, D, bd):
        for y in range(0, H, bh):
            for x in range(0, W, bw):
                sl = (slice(z, z + bd), slice(y, y + bh), slice(x, x + bw))
                want, _, pel = pl.motion_block(vol[sl].copy(), (bd, bh, bw), quant=0.05, intermediate=np.longdouble)
                _tie_equal(got[sl], want, pel, (z, y, x))


def _run_scan(kind, tmp_path, method="diag", step=1, shape=(12, 20)):
    h, w = shape
    px = np.random.default_rng(5).integers(0, 256, (h, w, 3)) / 255.0
    img, idx, vid = str(tmp_path / "img.dspraw"), str(tmp_path / "scan.idx"), str(tmp_path / "scan.dspv")
    dspraw.write(img, px)
    subprocess.run([_tool("scan_%s_f" % kind), "-q", "-m", method, "-S", str(step), "-p", "false", "-f", idx, "-t", "index", img, vid], check=True)
    index_map = np.loadtxt(idx, dtype=np.int64).reshape(h, w)
    fmt, ow, oh, frames = dspraw.read_video(vid)
    assert (fmt, ow, oh) == ("gbrpf32le", w, h)
    want, _ = pl.scan_frames(px.astype(np.float32), index_map, step=step)
    assert len(frames) == len(want)
    for f, wt in zip(frames, want):
        got = np.stack(f, axis=-1)                                   # components R, G, B
        assert np.abs(got - wt).max() < 2e-5
    assert np.abs(np.stack(frames[-1], axis=-1) - px).max() < 2e-5   # the last frame is the image (scan/scan.c:508-526)


@pytest.mark.parametrize("method,step", [("diag", 1), ("diag", 3), ("horizontal", 2), ("radial", 1), ("zigzag", 7), ("random", 16)])
def test_reference_scan_tool_pins_the_restatement(method, step, tmp_path):
    """scan.c's frame loop (scan/scan.c:421-459) on the scan order scan_methods.c generates == oracle scan_frames"""
    _run_scan("emu", tmp_path, method, step)


@pytest.mark.gpu
def test_reference_scan_tool_on_gpu(tmp_path):
    _run_scan("gpu", tmp_path, "diag", 2, shape=(24, 40))


ZOOM_CASES = {
    "up_3_2": (["-s", "3/2"], dict(scale=(3, 2))),
    "down_2_3": (["-s", "2/3"], dict(scale=(2, 3))),
    "native_x2": (["-s", "2", "--basis", "native"], dict(scale=(2, 1), basis="native")),
    "centered_view": (["-s", "5/2", "--basis", "centered", "-p", "4.5x2.25", "-v", "16x10"], dict(scale=(5, 2), basis="centered", pos=(4.5, 2.25), view=(16, 10))),
    "anisotropic": (["-s", "2x3/2"], dict(xscale=(2, 1), yscale=(3, 2))),
}


def _run_zoom(kind, prec, case, tmp_path):
    args, kw = ZOOM_CASES[case]
    h, w = 12, 20
    px = np.random.default_rng(9).integers(0, 256, (h, w, 3)) / 255.0
    img, vid = str(tmp_path / "img.dspraw"), str(tmp_path / "zoom.dspv")
    dspraw.write(img, px)
    subprocess.run([_tool("zoom_%s_%s" % (kind, prec)), "-q"] + args + [img, vid], check=True)
    _, ow, oh, frames = dspraw.read_video(vid)
    want = pl.zoom_synthesise(px.astype(np.float32 if prec == "f" else np.float64), **kw)
    got = np.stack(frames[0], axis=-1)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert np.abs(got - want).max() < (3e-5 if prec == "f" else 2e-6)          # the video frame is float32


@pytest.mark.parametrize("prec", ["d", "f"])
@pytest.mark.parametrize("case", sorted(ZOOM_CASES))
def test_reference_zoom_tool_pins_the_restatement(prec, case, tmp_path):
    """zoom.c's basis generation and synthesis loops (zoom/zoom.c:36-68, 263-375) == oracle zoom_synthesise"""
    _run_zoom("emu", prec, case, tmp_path)


@pytest.mark.gpu
def test_reference_zoom_tool_on_gpu(tmp_path):
    _run_zoom("gpu", "d", "up_3_2", tmp_path)
    _run_zoom("gpu", "f", "centered_view", tmp_path)
