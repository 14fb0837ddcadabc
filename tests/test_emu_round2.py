"""CPU suite, round-2 additions on the emulated kernels:
  * the chunked (L2-resident) schedule: consecutive passes that share their outermost loop level (batch elements,
    frames) run chunk by chunk -- results must not depend on the chunk size, fused coordinates included;
  * fused scan / spec stages when the contiguous axis runs as a strided (column) pass because the interleaved line
    does not fit on chip (the channel coordinate was overwritten there)."""
import os

import numpy as np
import pytest

from dspfun_b200 import REDFT01, REDFT10, Plan, capi
from dspfun_b200 import spec as gspec
from oracle import dct as od
from oracle import pipelines as pl
from tests import cases
from tests.emu import emu


@pytest.fixture(scope="module")
def lib():
    return emu.load()


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


@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("mb", [0, 0.05, 0.2])
def test_chunked_batches_match_oracle(lib, chunk_env, prec, mb):
    """7 images of 32x48x3: 18 KB (f) each -> chunks of 1..2 (0.05 MB), 5 (0.2 MB, ragged tail), or one launch (0 = off)"""
    chunk_env(mb)
    cases.check_batched_images(lib, prec, 7, 32, 48, 3)
    cases.check_batched_images(lib, prec, 5, 64, 64, 1)


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
def test_chunked_frames_of_a_rank3_plan(lib, chunk_env, kind):
    """rank 3: the w (row) and h (column) passes share the frame level and run per chunk of frames; the temporal pass
    runs whole.  Same result as the unchunked plan, bit for bit, and within tolerance of the oracle."""
    dims, embed = (12, 16, 24), (12, 20, 28)
    chunk_env(0)
    a = cases.run_planar_3d_embed(lib, "f", dims, embed, kind)
    chunk_env(0.005)           # 20*28*4 B = 2.2 KB per frame -> 2 frames per chunk
    b = cases.run_planar_3d_embed(lib, "f", dims, embed, kind)
    assert np.array_equal(a, b)
    cases.check_planar_3d_embed(lib, "f", dims, embed, kind)


def test_chunked_motion_keeps_fused_coordinates(lib, chunk_env):
    """motion's fused band-pass / preserve-dc stages read the frame coordinate: chunked launches must hand it on"""
    chunk_env(0.002)
    cases.check_motion(lib, (8, 16, 16), boost=1.5, bandpass=((1, 2, 2), (6, 12, 12)), preserve_dc="dc")
    cases.check_motion(lib, (8, 16, 16), scaled=(4, 8, 12))


@pytest.mark.parametrize("prec,shape", [("d", (2, 8192, 3)), ("f", (3, 8192, 4))])
def test_fused_scan_on_the_wide_path(lib, prec, shape):
    """contiguous axis as a column pass (line too long for the chip): channel-aware scan accumulate"""
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


# ---------------------------------------------------------------------------------------------- ring row kernel
@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape", [(2, 256), (70, 256), (6, 512), (36, 512), (4, 1024), (22, 1024), (10, 2048), (6, 4096), (2, 8192), (10, 8192)])
def test_ring_row_kernel_planar(lib, kind, shape):
    """dct_ring.cuh: planar power-of-two lines through the persistent bulk-copy-fed kernel (emulated: memcpy + the same
    index arithmetic); line counts that fill whole buffers, leave a ragged last buffer, and spread over several CTAs"""
    h, w = shape
    cases.check_interleaved_2d(lib, "f", h, w, 1, kind)


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
def test_ring_row_kernel_strided_lines_and_scale(lib, kind):
    """lines at a stride larger than n (embedded boxes: one bulk copy per line) and the fused load / store scale"""
    n, lines, dist = 1024, 12, 1024 + 64
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


def test_ring_row_kernel_matches_one_shot_kernel(lib):
    """same lines through the ring kernel and the one-shot row kernel (DSP_DCT_NO_RING is read once per process, so the
    comparison goes through a shape the ring kernel does not take: an odd line count)"""
    rng = np.random.default_rng(32)
    x = rng.random((5, 2048)).astype(np.float32)
    p = Plan("f", [2048], [REDFT10], 4, None, 1, 2048, None, 1, 2048, lib=lib)       # even: ring
    q = Plan("f", [2048], [REDFT10], 5, None, 1, 2048, None, 1, 2048, lib=lib)       # odd: one-shot kernel
    a = p.execute_host(x[:4].copy().reshape(-1)).reshape(4, 2048)
    b = q.execute_host(x.copy().reshape(-1)).reshape(5, 2048)
    p.destroy(); q.destroy()
    assert od.rel_l2(a, b[:4]) < 2e-7


# ---------------------------------------------------------------------------------------------- ring column sub-passes
@pytest.fixture
def small_panels():
    old = os.environ.get("DSP_DCT_SPLIT_PANEL_MB")
    os.environ["DSP_DCT_SPLIT_PANEL_MB"] = "1"          # 64-column panels at n = 4096 / 8192: several panels per plane
    yield
    if old is None:
        os.environ.pop("DSP_DCT_SPLIT_PANEL_MB", None)
    else:
        os.environ["DSP_DCT_SPLIT_PANEL_MB"] = old


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape", [(4096, 64, 1), (4096, 160, 1), (8192, 96, 1), (8192, 32, 3), (4096, 32, 2)])
def test_ring_column_subpasses(lib, small_panels, kind, shape):
    """dct_colring.cuh: long power-of-two columns through the tensor-copy-fed sub-pass kernels (emulated boxes):
    several tiles per panel, several panels, interleaved channels as plain columns, a last panel narrower than the rest"""
    cases.check_interleaved_2d(lib, "f", *shape, kind)


@pytest.mark.parametrize("shape", [(4096, 160, 1), (8192, 96, 1), (4096, 32, 2)])
def test_ring_column_inverse_opt_in(shape):
    """the DCT-III ring sub-passes (DSP_DCT_COLRING_INV=1, opt-in): pre-twiddle pairs across the two sub-sequence boxes,
    single sub-sequences 0 and 8, outer DIT butterflies with the un-permuting store -- emulated boxes, fresh process"""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from dspfun_b200 import REDFT01\nfrom tests import cases\nfrom tests.emu import emu\n"
            "print(cases.check_interleaved_2d(emu.load(), 'f', %d, %d, %d, REDFT01))" % (os.getcwd(), *shape))
    env = dict(os.environ, DSP_DCT_COLRING_INV="1", DSP_DCT_SPLIT_PANEL_MB="1", DSP_DCT_TRACE="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "ring sub-pass kernel" in r.stderr
    assert float(r.stdout.strip().splitlines()[-1]) < 1e-5


def test_ring_column_subpasses_batched_and_scaled(lib, small_panels):
    """outer (batch) offsets enter through the tensor maps' base pointers; fused store scale"""
    rng = np.random.default_rng(41)
    x = rng.random((2, 4096, 64)).astype(np.float32)
    p = Plan("f", [4096, 64], [REDFT10, REDFT10], 1, None, 1, 0, None, 1, 0, 2, 4096 * 64, 4096 * 64, lib=lib).fuse_scale(1.0, 0.25)
    y = p.execute_host(x.copy())
    p.destroy()
    ref = od.dctn_fast(x.astype(np.float64), [od.REDFT10] * 2, axes=(1, 2)) * 0.25
    assert od.rel_l2(y, ref) < 1e-5


# ---------------------------------------------------------------------------------------------- block DCT by GEMM (host logic)
@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape,B", [((2, 32, 48), 8), ((1, 64, 32), 16), ((3, 32, 96), 32), ((1, 128, 64), 64), ((1, 200, 328), 8)])
def test_block_dct2d_matches_oracle_blocks(lib, kind, shape, B):
    """the emulation build runs the same contraction as loops (the tcgen05 path itself is covered by tests/test_gpu_round2.py)"""
    cases.check_block_dct2d(lib, shape, B, kind, 2e-6)
    cases.check_block_dct2d(lib, shape, B, kind, 2e-6, in_place=True, scale=0.25)


def test_block_dct2d_refuses_what_it_cannot_do(lib):
    x = np.zeros((1, 48, 48), np.float32)
    for args in ((b"f", 1, 48, 48, 12), (b"f", 1, 48, 40, 16), (b"d", 1, 48, 48, 8), (b"f", 0, 48, 48, 8)):
        prec, P, H, W, B = args
        assert lib.dsp_block_dct2d(prec, x.ctypes.data, x.ctypes.data, P, H, W, B, REDFT10, 1.0, None) != 0
        assert capi.last_error(lib)
    assert lib.dsp_block_dct2d(b"f", x.ctypes.data, x.ctypes.data, 1, 48, 48, 8, 3, 1.0, None) != 0      # not a DCT kind


def test_motion_tiled_gemm_and_plan_paths_agree(lib):
    """MotionTiled with the spatial axes on dsp_block_dct2d == the three-plan path == the reference block loop"""
    cases.check_motion_tiled(lib, (16, 16, 24), (8, 8, 8), gemm=False)
    cases.check_motion_tiled(lib, (16, 16, 24), (8, 8, 8), quant=0.05, gemm=False)
    cases.check_motion_tiled(lib, (2, 32, 64), (1, 16, 16), quant=0.02)
    cases.check_motion_tiled(lib, (2, 32, 64), (2, 32, 32))


def test_motion_tiled_c_session(lib):
    """dsp_motion_tiled_create / process_dev / destroy: the block-tiled pipeline behind the C ABI (missing 5)"""
    cases.check_motion_tiled_c_session(lib, (16, 16, 24), (8, 8, 8), 0.05)          # GEMM spatial axes + d plan
    cases.check_motion_tiled_c_session(lib, (8, 32, 16), (4, 16, 8), 0.02)          # three per-axis plans
    cases.check_motion_tiled_c_session(lib, (2, 24, 40), (1, 8, 8), 0.0)            # depth-1 blocks, no quantiser
    assert not lib.dsp_motion_tiled_create(16, 16, 24, 8, 8, 7, 0.0) and capi.last_error(lib)


# ---------------------------------------------------------------------------------------------- motion coefficient stage as a sweep
@pytest.mark.parametrize("kw", [dict(), dict(damp=0.25, boost=1.5, bandpass=((0, 1, 1), (3, 6, 9))), dict(quant=0.05),
                                dict(threshold=(0.001, 0.4), preserve_dc="dc", damp=0.0, bandpass=((1, 0, 2), (4, 8, 12)))])
def test_motion_coeff_stage_sweep_equals_the_fused_pass(lib, kw):
    """dsp_motion_coeff_stage (one sweep over the volume) and dsp_motion_coeff_stage_flat (the [D][h*w] layout of the sharded
    path) apply exactly the map dsp_dct_fuse_motion_coeff carries in a pass: same coefficients in, same bits out"""
    import ctypes
    from dspfun_b200.dist3d import motion_params
    D, H, W = 4, 8, 12
    kw = dict(kw)
    if "preserve_dc" in kw:
        kw["preserve_dc"] = {"dc": 1, "grey": 2}[kw["preserve_dc"]]
    mp = motion_params((D, H, W), **kw)
    c = (np.random.default_rng(2).standard_normal((D, H, W)) * 50).astype(np.float32)
    fused = Plan("f", [D, H, W], [REDFT01] * 3, lib=lib).fuse_motion_coeff(mp)
    want = fused.execute_host(c.copy())
    fused.destroy()
    plain = Plan("f", [D, H, W], [REDFT01] * 3, lib=lib)
    a = c.copy()
    assert lib.dsp_motion_coeff_stage(b"f", ctypes.byref(mp), a.ctypes.data, None, None) == 0, capi.last_error(lib)
    b = c.copy().reshape(D, H * W)
    assert lib.dsp_motion_coeff_stage_flat(b"f", ctypes.byref(mp), b.ctypes.data, D, H * W, W, 0, None, None) == 0, capi.last_error(lib)
    assert np.array_equal(a.reshape(D, H * W), b)
    # a slice of the columns with its base, as a rank of the sharded path sees it
    half = np.ascontiguousarray(c.reshape(D, H * W)[:, H * W // 2:])
    assert lib.dsp_motion_coeff_stage_flat(b"f", ctypes.byref(mp), half.ctypes.data, D, H * W // 2, W, H * W // 2, None, None) == 0
    assert np.array_equal(half, b[:, H * W // 2:])
    got = plain.execute_host(a)
    plain.destroy()
    assert np.array_equal(got, want)


@pytest.mark.parametrize("bd", [2, 4, 8, 16])
def test_block_dquant_equals_the_three_sweeps(lib, bd):
    """dsp_block_dquant (d forward + quantiser stage + d inverse per pixel, in registers) against the d plan, dsp_block_quant
    and the inverse d plan; the direct 8-point sums and the FFT passes round differently, so the comparison is to float accuracy"""
    import ctypes
    D, H, W, bh, bw = 2 * bd, 8, 16, 8, 8
    c = (np.random.default_rng(3).standard_normal((D, H, W)) * 200).astype(np.float32)
    q = float(np.float32(0.05 * 8.0 * np.sqrt(bd * bh * bw)))
    a = c.copy()
    cnt_a = ctypes.c_ulonglong(0)
    assert lib.dsp_block_dquant(a.ctypes.data, D, H, W, bd, bh, bw, q, ctypes.byref(cnt_a), None) == 0, capi.last_error(lib)
    b = c.copy()
    cnt_b = ctypes.c_ulonglong(0)
    fwd = Plan("f", [bd], [REDFT10], H * W, None, H * W, 1, None, H * W, 1, D // bd, bd * H * W, bd * H * W, lib=lib)
    inv = Plan("f", [bd], [REDFT01], H * W, None, H * W, 1, None, H * W, 1, D // bd, bd * H * W, bd * H * W, lib=lib)
    fwd.execute_host(b)
    assert lib.dsp_block_quant(b"f", b.ctypes.data, D, H, W, bd, bh, bw, q, ctypes.byref(cnt_b), None) == 0
    inv.execute_host(b)
    assert od.rel_l2(a, b.astype(np.float64)) < 2e-4                                             # a coefficient on a quantiser tie may flip
    assert abs(int(cnt_a.value) - int(cnt_b.value)) <= max(2, int(cnt_b.value) // 1000)
    a, b = c.copy(), c.copy()                                                                    # quantiser off: float accuracy
    assert lib.dsp_block_dquant(a.ctypes.data, D, H, W, bd, bh, bw, 0.0, None, None) == 0
    fwd.execute_host(b)
    assert lib.dsp_block_quant(b"f", b.ctypes.data, D, H, W, bd, bh, bw, 0.0, None, None) == 0
    inv.execute_host(b)
    fwd.destroy(); inv.destroy()
    assert od.rel_l2(a, b.astype(np.float64)) < 2e-6
    assert lib.dsp_block_dquant(a.ctypes.data, D, H, W, 3, bh, bw, q, None, None) != 0           # unsupported depth: refused


def test_gpu_count_knob_is_inert_without_gpus(lib):
    """dsp_dct_plan_with_ngpus on the emulation build: accepted, plans stay on one (emulated) device"""
    lib.dsp_dct_plan_with_ngpus(4)
    p = Plan("f", [4, 8, 8], [REDFT10] * 3, lib=lib)
    x = np.random.default_rng(1).random((4, 8, 8)).astype(np.float32)
    y = p.execute_host(x.copy())
    assert lib.dsp_dct_plan_ngpus(p._h) == 1
    assert od.rel_l2(y, od.dctn_fast(x.astype(np.float64), [od.REDFT10] * 3)) < 1e-5
    p.destroy()
    lib.dsp_dct_plan_with_ngpus(1)
