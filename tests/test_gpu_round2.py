"""GPU suite (-m gpu), round-2 additions: the persistent bulk-copy-fed row kernel (dct_ring.cuh), the chunked
(L2-resident) schedule, and the fused stages on the wide (column-pass) path -- same cases as tests/test_emu_round2.py,
through the product library's C ABI."""
import os

import numpy as np
import pytest

from dspfun_b200 import REDFT01, REDFT10, Plan, capi
from dspfun_b200 import spec as gspec
from oracle import dct as od
from oracle import pipelines as pl
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return capi.load()


@pytest.fixture
def chunk_env():
    old = os.environ.get("DSP_DCT_L2_CHUNK_MB")

    def set_mb(v):
        os.environ["DSP_DCT_L2_CHUNK_MB"] = str(v)
    yield set_mb
    if old is None:
        os.environ.pop("DSP_DCT_L2_CHUNK_MB", None)
    else:
        os.environ["DSP_DCT_L2_CHUNK_MB"] = old


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape", [(2, 256), (70, 256), (6, 512), (36, 512), (4, 1024), (22, 1024), (10, 2048), (6, 4096),
                                   (2, 8192), (10, 8192), (3000, 256), (2048, 1024), (1000, 4096), (700, 8192)])
def test_ring_row_kernel_planar(lib, kind, shape):
    """whole buffers, ragged last buffers, fewer iterations than SMs, many iterations per CTA (ring wrap-around, both
    mbarrier phases)"""
    h, w = shape
    cases.check_interleaved_2d(lib, "f", h, w, 1, kind)


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
def test_ring_row_kernel_strided_lines_and_scale(lib, kind):
    n, lines, dist = 1024, 1200, 1024 + 64
    rng = np.random.default_rng(31)
    buf = rng.random(lines * dist).astype(np.float32)
    ref = buf.astype(np.float64).copy()
    ok = cases.ORK[kind]
    for b in range(lines):
        ref[b * dist:b * dist + n] = od.dctn_fast(buf[b * dist:b * dist + n].astype(np.float64) * 0.5, [ok]) * 3.0
    p = Plan("f", [n], [kind], lines, None, 1, dist, None, 1, dist, lib=lib).fuse_scale(0.5, 3.0)
    y = p.execute_host(buf.copy())
    p.destroy()
    sel = np.zeros(lines * dist, bool)
    for b in range(lines):
        sel[b * dist:b * dist + n] = True
    assert od.rel_l2(y[sel], ref[sel]) < 1e-5
    assert np.array_equal(y[~sel], buf[~sel])


def test_ring_row_kernel_repeatable(lib):
    """the ring hands buffers between two thread groups: the same input twice must give the same bits"""
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand((4096, 8192), device="cuda", dtype=torch.float32, generator=g)
    p = Plan("f", [8192], [REDFT10], 4096, None, 1, 8192, None, 1, 8192, lib=lib)
    a, b = x.clone(), x.clone()
    p.execute_dev(a.data_ptr(), a.data_ptr(), None)
    p.execute_dev(b.data_ptr(), b.data_ptr(), None)
    torch.cuda.synchronize()
    p.destroy()
    assert torch.equal(a, b)
    ref = od.dctn_fast(x[1234].cpu().numpy().astype(np.float64), [od.REDFT10])
    assert od.rel_l2(a[1234].cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("mb", [0, 1, 4])
def test_chunked_batches_match_oracle(lib, chunk_env, prec, mb):
    chunk_env(mb)
    cases.check_batched_images(lib, prec, 7, 128, 256, 3)
    cases.check_batched_images(lib, prec, 16, 256, 256, 1)


def test_chunked_frames_of_a_rank3_plan(lib, chunk_env):
    dims, embed = (12, 64, 96), (12, 80, 112)
    for kind in (REDFT10, REDFT01):
        chunk_env(0)
        a = cases.run_planar_3d_embed(lib, "f", dims, embed, kind)
        chunk_env(0.1)
        b = cases.run_planar_3d_embed(lib, "f", dims, embed, kind)
        assert np.array_equal(a, b)
        cases.check_planar_3d_embed(lib, "f", dims, embed, kind)


def test_chunked_motion_keeps_fused_coordinates(lib, chunk_env):
    chunk_env(0.002)
    cases.check_motion(lib, (8, 16, 16), boost=1.5, bandpass=((1, 2, 2), (6, 12, 12)), preserve_dc="dc")
    cases.check_motion(lib, (8, 16, 16), scaled=(4, 8, 12))


@pytest.mark.parametrize("prec,shape", [("d", (2, 8192, 3)), ("f", (3, 8192, 4))])
def test_fused_scan_on_the_wide_path(lib, prec, shape):
    cases.check_scan(lib, prec, *shape, order="horizontal", step=shape[0] * shape[1] // 3 + 1)


@pytest.mark.parametrize("prec,shape", [("d", (2, 8192, 3)), ("f", (2, 16384, 2))])
@pytest.mark.parametrize("rng_", ["dc", "dcs"])
def test_fused_spec_ranges_on_the_wide_path(lib, prec, shape, rng_):
    rs = np.random.default_rng(11)
    px = (rs.integers(0, 256, shape) / 255.0).astype(cases.DT[prec])
    params = ("log", "shift", "reference", rng_)
    s1, dc1 = gspec.spec(px, None, lib=lib, scale="log", sign="shift", range_=rng_, gain="reference")
    s0, dc0 = pl.spec_forward(px, params=params, custom_gain=1.0, intermediate=cases.INTERMEDIATE[prec])
    assert np.all(np.isfinite(s1))
    assert od.rel_l2(s1, s0) < cases.OK[prec] * 4, od.rel_l2(s1, s0)
    assert np.allclose(dc1, dc0, rtol=1e-5 if prec == "f" else 1e-12)


# ---------------------------------------------------------------------------------------------- ring column sub-passes
@pytest.fixture
def small_panels():
    old = os.environ.get("DSP_DCT_SPLIT_PANEL_MB")
    os.environ["DSP_DCT_SPLIT_PANEL_MB"] = "1"
    yield
    if old is None:
        os.environ.pop("DSP_DCT_SPLIT_PANEL_MB", None)
    else:
        os.environ["DSP_DCT_SPLIT_PANEL_MB"] = old


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape", [(4096, 64, 1), (4096, 160, 1), (8192, 96, 1), (8192, 32, 3), (4096, 32, 2)])
def test_ring_column_subpasses_small_panels(lib, small_panels, kind, shape):
    cases.check_interleaved_2d(lib, "f", *shape, kind)


@pytest.mark.parametrize("shape", [(4096, 160, 1), (8192, 96, 1), (8192, 1024, 1)])
def test_ring_column_inverse_opt_in(shape):
    """the DCT-III ring sub-passes (DSP_DCT_COLRING_INV=1: correct, measured slower than the default two-kernel split, so
    opt-in): a fresh process, because the switch is read once"""
    import subprocess, sys, os
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from dspfun_b200 import capi, REDFT01\nfrom tests import cases\n"
            "print(cases.check_interleaved_2d(capi.load(), 'f', %d, %d, %d, REDFT01))" % (os.getcwd(), *shape))
    env = dict(os.environ, DSP_DCT_COLRING_INV="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert float(r.stdout.strip().splitlines()[-1]) < 1e-5


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape", [(4096, 2048, 1), (8192, 1024, 1), (8192, 512, 3), (4096, 4096, 1)])
def test_ring_column_subpasses(lib, kind, shape):
    """default 32 MB panels: hundreds of iterations per CTA (ring wrap-around, deferred refills after tensor stores)"""
    cases.check_interleaved_2d(lib, "f", *shape, kind)


def test_ring_column_subpasses_batched_and_scaled(lib, small_panels):
    rng = np.random.default_rng(41)
    x = rng.random((2, 4096, 64)).astype(np.float32)
    p = Plan("f", [4096, 64], [REDFT10, REDFT10], 1, None, 1, 0, None, 1, 0, 2, 4096 * 64, 4096 * 64, lib=lib).fuse_scale(1.0, 0.25)
    y = p.execute_host(x.copy())
    p.destroy()
    ref = od.dctn_fast(x.astype(np.float64), [od.REDFT10] * 2, axes=(1, 2)) * 0.25
    assert od.rel_l2(y, ref) < 1e-5


# ---------------------------------------------------------------------------------------------- motion spectrograms
@pytest.mark.parametrize("kw", [dict(spec="shift"), dict(spec="flat"), dict(spec="abs"), dict(spec="copy", quant=0.02),
                                dict(ispec="shift"), dict(ispec="flat"), dict(ispec="copy"),
                                dict(spec="shift", ispec="shift"), dict(spec="flat", boost=1.5, bandpass=((0, 1, 1), (4, 6, 6)))])
def test_motion_spectrogram_modes(lib, kw):
    cases.check_motion(lib, (4, 8, 8), **kw)
    cases.check_motion(lib, (8, 30, 40), **kw)
    cases.check_motion(lib, (4, 8, 12), float_pixels=True, **kw)


# ---------------------------------------------------------------------------------------------- block DCT on the tensor cores
# float tolerance of the 3 x TF32 GEMM path: products carry ~2^-21 relative error (the FFT passes: ~1e-7); measured
# 6e-7 (B = 8) .. 1e-6 (B = 64) relative L2 on normal data, asserted at 3e-6
BLOCK_MM_TOL = 3e-6


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("B", [8, 16, 32, 64])
def test_block_dct2d_tensor_cores_vs_oracle(lib, kind, B):
    """whole tiles, several tiles per CTA (more tiles than SMs), planes that are not multiples of the 128 x 128 tile"""
    cases.check_block_dct2d(lib, (2, 256, 384), B, kind, BLOCK_MM_TOL, device="cuda")
    cases.check_block_dct2d(lib, (3, 192, 320), B, kind, BLOCK_MM_TOL, device="cuda", in_place=True, scale=0.5)
    cases.check_block_dct2d(lib, (1, 64, 64), B, kind, BLOCK_MM_TOL, device="cuda")


def test_block_dct2d_ragged_and_many_tiles(lib):
    cases.check_block_dct2d(lib, (1, 1080, 1920), 8, REDFT10, BLOCK_MM_TOL, device="cuda")        # C5's frame: 8.4 tiles high
    cases.check_block_dct2d(lib, (1, 144, 16), 16, REDFT01, BLOCK_MM_TOL, device="cuda")
    cases.check_block_dct2d(lib, (40, 256, 256), 32, REDFT10, BLOCK_MM_TOL, device="cuda")        # 160 tiles > 148 SMs: ring wrap
    cases.check_block_dct2d(lib, (5, 1024, 640), 64, REDFT01, BLOCK_MM_TOL, device="cuda", in_place=True)


def test_block_dct2d_equals_the_plan_path_and_round_trips(lib):
    """the same blocks through two per-axis plans (FFT passes); forward then inverse / (4 B^2) returns the planes"""
    import torch
    P, H, W, B = 4, 512, 768, 16
    x = torch.randn(P, H, W, device="cuda")
    y = torch.empty_like(x)
    assert lib.dsp_block_dct2d(b"f", x.data_ptr(), y.data_ptr(), P, H, W, B, REDFT10, 1.0, None) == 0
    pw = Plan("f", [B], [REDFT10], (W // B) * H * P, None, 1, B, None, 1, B, lib=lib)
    ph = Plan("f", [B], [REDFT10], W, None, W, 1, None, W, 1, (H // B) * P, B * W, B * W, lib=lib)
    z = x.clone()
    pw.execute_dev(z.data_ptr(), z.data_ptr(), None)
    ph.execute_dev(z.data_ptr(), z.data_ptr(), None)
    torch.cuda.synchronize()
    assert od.rel_l2(y.cpu().numpy(), z.cpu().numpy().astype(np.float64)) < BLOCK_MM_TOL
    assert lib.dsp_block_dct2d(b"f", y.data_ptr(), y.data_ptr(), P, H, W, B, REDFT01, 1.0 / (4.0 * B * B), None) == 0
    torch.cuda.synchronize()
    assert od.rel_l2(y.cpu().numpy(), x.cpu().numpy().astype(np.float64)) < 2 * BLOCK_MM_TOL
    pw.destroy()
    ph.destroy()


def test_block_dct2d_full_size_properties(lib):
    """64 planes of 2048 x 2048 (1 GiB): linearity and the round trip, with sampled blocks against the oracle"""
    import torch
    P, H, W, B = 64, 2048, 2048, 8
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(P, H, W, device="cuda", generator=g)
    y = torch.empty_like(x)
    assert lib.dsp_block_dct2d(b"f", x.data_ptr(), y.data_ptr(), P, H, W, B, REDFT10, 1.0, None) == 0
    torch.cuda.synchronize()
    rng = np.random.default_rng(4)
    for _ in range(32):
        p, by, bx = int(rng.integers(P)), int(rng.integers(H // B)) * B, int(rng.integers(W // B)) * B
        blk = x[p, by:by + B, bx:bx + B].cpu().numpy().astype(np.float64)
        assert od.rel_l2(y[p, by:by + B, bx:bx + B].cpu().numpy(), od.dctn_fast(blk, [od.REDFT10, od.REDFT10])) < 2 * BLOCK_MM_TOL
    # Parseval per plane for the orthogonalised transform: sum y^2 w = (2B)^2 sum x^2 with w = 1/2 on each zero index
    w = torch.ones(B, device="cuda", dtype=torch.float64)
    w[0] = 0.5
    wy = (w.repeat(H // B)[:, None] * w.repeat(W // B)[None, :])
    e_y = float((y[0].double() ** 2 * wy).sum())
    e_x = float((x[0].double() ** 2).sum()) * (2.0 * B) ** 2
    assert abs(e_y / e_x - 1.0) < 1e-5
    assert lib.dsp_block_dct2d(b"f", y.data_ptr(), y.data_ptr(), P, H, W, B, REDFT01, 1.0 / (4.0 * B * B), None) == 0
    torch.cuda.synchronize()
    assert float((y - x).double().norm() / x.double().norm()) < 2 * BLOCK_MM_TOL


def test_motion_tiled_on_the_gemm_path(lib):
    cases.check_motion_tiled(lib, (16, 64, 96), (8, 8, 8), device="cuda", gemm=False)
    cases.check_motion_tiled(lib, (4, 128, 256), (2, 32, 32), quant=0.02, device="cuda")
    cases.check_motion_tiled(lib, (2, 128, 128), (1, 64, 64), device="cuda")


def test_motion_tiled_c_session_on_gpu(lib):
    cases.check_motion_tiled_c_session(lib, (16, 128, 256), (8, 8, 8), 0.05, device="cuda")
    cases.check_motion_tiled_c_session(lib, (8, 64, 96), (4, 16, 8), 0.02, device="cuda")
    cases.check_motion_tiled_c_session(lib, (4, 128, 128), (1, 64, 64), 0.0, device="cuda")


# ---------------------------------------------------------------------------------------------- zoom's dense path on the tensor cores
@pytest.mark.parametrize("kw", [dict(scale=(3, 2)), dict(scale=(5, 3), basis="centered"), dict(scale=(7, 4), pos=(10.5, 3.25), view=(200, 150)),
                                dict(xscale=(2, 1), yscale=(3, 2), basis="centered"), dict(scale=(2, 3))])
def test_zoom_dense_path_tensor_core_gemm(lib, kw):
    """k tails (cw, ch not multiples of 32), M / N tails (views not multiples of 128), several k blocks and tiles; against the
    restated synthesis and against the SIMT GEMM of round 1 (DSP_ZOOM_NO_TC is read once per process, so that comparison
    runs in the fullsize test below through a subprocess-free route: the double-precision session)"""
    path, got, want = cases.check_zoom(lib, "f", 301, 421, **kw)
    assert path == "dense-tensor-core"
    path_d, got_d, _ = cases.check_zoom(lib, "d", 301, 421, **kw)
    assert path_d == "dense"
    assert od.rel_l2(got, got_d) < 1e-5                      # float bases and float coefficients against the double session


# ---------------------------------------------------------------------------------------------- sweeps added at the end of round 2
@pytest.mark.parametrize("kw", [dict(), dict(damp=0.25, boost=1.5, bandpass=((0, 1, 1), (12, 60, 100))), dict(quant=0.05),
                                dict(threshold=(0.001, 0.4), preserve_dc=1, damp=0.0, bandpass=((1, 0, 2), (16, 64, 128)))])
def test_motion_coeff_stage_sweep_equals_the_fused_pass_on_gpu(lib, kw):
    import ctypes
    import torch
    from dspfun_b200.dist3d import motion_params
    D, H, W = 16, 72, 160
    mp = motion_params((D, H, W), **kw)
    c = (torch.randn(D, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)) * 50).contiguous()
    fused = Plan("f", [D, H, W], [REDFT01] * 3, lib=lib).fuse_motion_coeff(mp)
    want = c.clone()
    fused.execute_dev(want.data_ptr(), want.data_ptr(), None)
    a = c.clone()
    assert lib.dsp_motion_coeff_stage(b"f", ctypes.byref(mp), a.data_ptr(), None, None) == 0, capi.last_error(lib)
    b = c.clone()
    assert lib.dsp_motion_coeff_stage_flat(b"f", ctypes.byref(mp), b.data_ptr(), D, H * W, W, 0, None, None) == 0, capi.last_error(lib)
    assert torch.equal(a, b)
    plain = Plan("f", [D, H, W], [REDFT01] * 3, lib=lib)
    plain.execute_dev(a.data_ptr(), a.data_ptr(), None)
    torch.cuda.synchronize()
    assert torch.equal(a, want)
    fused.destroy(); plain.destroy()


@pytest.mark.parametrize("bd", [2, 4, 8, 16])
def test_block_dquant_equals_the_three_sweeps_on_gpu(lib, bd):
    import torch
    D, H, W, bh, bw = 4 * bd, 64, 128, 8, 8
    c = (torch.randn(D, H, W, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 200).contiguous()
    q = float(np.float32(0.05 * 8.0 * np.sqrt(bd * bh * bw)))
    a, b = c.clone(), c.clone()
    ca = torch.zeros(1, dtype=torch.int64, device="cuda")
    cb = torch.zeros(1, dtype=torch.int64, device="cuda")
    assert lib.dsp_block_dquant(a.data_ptr(), D, H, W, bd, bh, bw, q, ca.data_ptr(), None) == 0, capi.last_error(lib)
    fwd = Plan("f", [bd], [REDFT10], H * W, None, H * W, 1, None, H * W, 1, D // bd, bd * H * W, bd * H * W, lib=lib)
    inv = Plan("f", [bd], [REDFT01], H * W, None, H * W, 1, None, H * W, 1, D // bd, bd * H * W, bd * H * W, lib=lib)
    fwd.execute_dev(b.data_ptr(), b.data_ptr(), None)
    assert lib.dsp_block_quant(b"f", b.data_ptr(), D, H, W, bd, bh, bw, q, cb.data_ptr(), None) == 0
    inv.execute_dev(b.data_ptr(), b.data_ptr(), None)
    torch.cuda.synchronize()
    # a coefficient that sits on a quantiser tie may round the other way (the direct sums and the FFT passes differ in the last
    # bits): a handful of full quantiser steps in 10^5 coefficients, hence the looser bound with the quantiser on
    assert od.rel_l2(a.cpu().numpy(), b.cpu().numpy().astype(np.float64)) < 2e-4
    assert abs(int(ca.item()) - int(cb.item())) <= max(2, int(cb.item()) // 1000)
    a, b = c.clone(), c.clone()                                           # quantiser off: float accuracy
    assert lib.dsp_block_dquant(a.data_ptr(), D, H, W, bd, bh, bw, 0.0, None, None) == 0
    fwd.execute_dev(b.data_ptr(), b.data_ptr(), None)
    assert lib.dsp_block_quant(b"f", b.data_ptr(), D, H, W, bd, bh, bw, 0.0, None, None) == 0
    inv.execute_dev(b.data_ptr(), b.data_ptr(), None)
    torch.cuda.synchronize()
    assert od.rel_l2(a.cpu().numpy(), b.cpu().numpy().astype(np.float64)) < 2e-6
    fwd.destroy(); inv.destroy()
