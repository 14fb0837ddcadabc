"""CPU suite: the kernel sources compiled as a host SIMT emulation (tests/emu) vs the oracle.  Checks the host
logic (planner, tables, layouts) and the kernels' index arithmetic; the GPU suite repeats these on the device."""
import numpy as np
import pytest

from dspfun_b200 import REDFT01, REDFT10, Plan, capi
from tests import cases
from tests.emu import emu


@pytest.fixture(scope="module")
def lib():
    return emu.load()


@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape", cases.SHAPES_2D)
def test_interleaved_2d(lib, prec, kind, shape):
    cases.check_interleaved_2d(lib, prec, *shape, kind, definition=max(shape[:2]) <= 64)


@pytest.mark.parametrize("prec", ["f", "d"])
def test_out_of_place_preserves_input(lib, prec):
    cases.check_interleaved_2d(lib, prec, 24, 40, 3, REDFT01, inplace=False)
    cases.check_interleaved_2d(lib, prec, 24, 40, 1, REDFT10, inplace=False)


@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
def test_rank1_batches(lib, prec, kind):
    cases.check_rank1_batch(lib, prec, 64, 5, kind)
    cases.check_rank1_batch(lib, prec, 60, 7, kind, dist=67)
    cases.check_rank1_batch(lib, prec, 2048, 3, kind)
    cases.check_rank1_batch(lib, prec, 1, 4, kind)


@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
def test_planar_3d_embed(lib, prec, kind):
    cases.check_planar_3d_embed(lib, prec, (8, 8, 8), (8, 8, 8), kind)
    cases.check_planar_3d_embed(lib, prec, (4, 6, 10), (7, 9, 12), kind)
    cases.check_planar_3d_embed(lib, prec, (1, 16, 24), (1, 16, 24), kind)      # motion's default depth-1 blocks
    cases.check_planar_3d_embed(lib, prec, (16, 15, 20), (16, 27, 36), kind)


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape", [(4096, 64, 1), (4096, 40, 1), (8192, 32, 1), (4096, 24, 3)])
def test_split_column_pass(lib, kind, shape):
    """long power-of-two columns: the two-sub-pass split path (dct_split.cuh), full fixed-length tiles and ragged ones"""
    cases.check_interleaved_2d(lib, "f", *shape, kind)
    if shape[0] == 4096:
        cases.check_interleaved_2d(lib, "d", *shape, kind)     # double: forward split + DIF-style inverse split


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
@pytest.mark.parametrize("shape", [(8, 1080, 1), (6, 1920, 1), (540, 16, 1), (20, 360, 1), (1080, 36, 1)])
def test_mixed_radix_video_sizes(lib, kind, shape):
    """motion's frame sizes: merged radices 15 / 9 and the lean planar moves of the generic engine"""
    cases.check_interleaved_2d(lib, "f", *shape, kind)
    cases.check_interleaved_2d(lib, "d", *shape, kind)


def test_fixed_length_column_tiles(lib):
    """batches large enough for 32- and 16-column tiles: the compile-time-length column moves (col_tile_fixed)"""
    cases.check_batched_images(lib, "f", 10, 256, 1024, 1)      # n = 256, 32-column tiles
    cases.check_batched_images(lib, "f", 25, 1024, 64, 3)       # n = 1024, 16-column tiles, RGB rows


@pytest.mark.parametrize("prec", ["f", "d"])
def test_batched_images_roundtrip(lib, prec):
    cases.check_batched_images(lib, prec, 5, 16, 24, 3)
    cases.check_batched_images(lib, prec, 3, 32, 32, 1)


@pytest.mark.parametrize("prec", ["f", "d"])
@pytest.mark.parametrize("typ", [2, 3])
@pytest.mark.parametrize("n", [2, 3, 4, 8, 12, 15, 16, 17, 32, 64, 128, 256, 512, 1024])
def test_fftw_golden_vectors(lib, golden, prec, typ, n):
    cases.check_golden_1d(lib, golden, prec, n, typ)


def test_plan_2d_matches_plan_many(lib):
    rng = np.random.default_rng(4)
    x = rng.random((12, 16))
    h = lib.dsp_dct_plan_2d(b"d", 12, 16, None, None, REDFT01, REDFT01, 0)
    assert h
    y = x.copy()
    assert lib.dsp_dct_execute_host(h, y.ctypes.data, y.ctypes.data) == 0
    lib.dsp_dct_destroy(h)
    from oracle import dct as od
    np.testing.assert_allclose(y, od.dctn_def(x, [od.REDFT01] * 2), rtol=0, atol=1e-11)


def test_rejects_loudly(lib):
    for args in [dict(prec="l", n=[8], kinds=[REDFT10]), dict(prec="f", n=[8], kinds=[3]),
                 dict(prec="f", n=[0], kinds=[REDFT10]), dict(prec="f", n=[8, 8], kinds=[REDFT10] * 2, howmany=2, istride=5, idist=1, ostride=5, odist=1)]:
        with pytest.raises(capi.DspDctError) as e:
            Plan(lib=lib, **args)
        assert str(e.value)


@pytest.mark.parametrize("kind", [REDFT10, REDFT01])
def test_segmented_output_of_last_pass(lib, kind):
    """dsp_dct_set_output_segments: the last pass scatters runs of axis positions to separate buffers (the peer
    buffers of the slab-sharded 3-D transform), for the batched 2-D plan and for the wide-interleave temporal plan"""
    from oracle import dct as od
    ork = od.REDFT10 if kind == REDFT10 else od.REDFT01
    rng = np.random.default_rng(4)
    # batched 2-D plan, last pass along h (forward) -- rows h split in two, frame stride (h/2) w in the destinations
    B, h, w = 3, 16, 32
    x = rng.random((B, h, w)).astype(np.float32)
    ref = od.dctn_fast(x.astype(np.float64), [ork] * 2, axes=(1, 2))
    if kind == REDFT10:
        p = Plan.interleaved_2d("f", h, w, 1, kind, nbatch=B, lib=lib)
        segs = [np.full((B, h // 2, w), -1.0, np.float32) for _ in range(2)]
        p.set_output_segments(h // 2, [a.ctypes.data for a in segs], outer_stride=(h // 2) * w)
        buf = x.copy()
        p.execute_dev(buf.ctypes.data, buf.ctypes.data, None)
        p.destroy()
        got = np.concatenate(segs, axis=1)
        assert od.rel_l2(got, ref) < cases.OK["f"]
    # temporal plan over [D][P] (stride P): frames d split in two runs
    D, P = 32, 64
    v = rng.random((D, P)).astype(np.float32)
    refv = od.dctn_fast(v.astype(np.float64), [ork], axes=(0,))
    q = Plan("f", [D], [kind], P, None, P, 1, None, P, 1, lib=lib)
    segs = [np.full((D // 2, P), -1.0, np.float32) for _ in range(2)]
    q.set_output_segments(D // 2, [a.ctypes.data for a in segs])
    buf = v.copy()
    q.execute_dev(buf.ctypes.data, buf.ctypes.data, None)
    q.destroy()
    assert od.rel_l2(np.concatenate(segs, axis=0), refv) < cases.OK["f"]
    # a plan whose last pass is a row pass refuses
    r = Plan.interleaved_2d("f", h, w, 1, REDFT01, lib=lib)
    with pytest.raises(capi.DspDctError):
        r.set_output_segments(h // 2, [segs[0].ctypes.data, segs[1].ctypes.data])
    r.destroy()


def test_size_limits_and_argument_errors(lib):
    """the loud failures at the edges: lengths beyond the on-chip limit, bad segment geometry, bad block-stage arguments"""
    import ctypes
    for n in (70000, 65536):                                   # table index type / shared-memory capacity
        with pytest.raises(capi.DspDctError) as e:
            Plan("f", [n], [REDFT10], lib=lib)
        assert "too large" in str(e.value) or "does not fit" in str(e.value)
    p = Plan.interleaved_2d("f", 16, 32, 1, REDFT10, nbatch=2, lib=lib)
    a = np.zeros((2, 8, 32), np.float32)
    with pytest.raises(capi.DspDctError):                      # 3 x 8 rows do not tile the 16-point axis
        p.set_output_segments(8, [a.ctypes.data] * 3)
    with pytest.raises(capi.DspDctError):                      # more than 8 segments
        p.set_output_segments(1, [a.ctypes.data] * 16)
    p.set_output_segments(8, [a.ctypes.data, a.ctypes.data], outer_stride=8 * 32)
    lib.dsp_dct_set_output_segments(p._h, 0, 0, None, 0, 0)    # switching it off again is legal
    p.destroy()
    c = np.zeros(64, np.float32)
    assert lib.dsp_block_quant(b"x", c.ctypes.data, 4, 4, 4, 2, 2, 2, 0.0, None, None) != 0
    assert lib.dsp_block_quant(b"f", None, 4, 4, 4, 2, 2, 2, 0.0, None, None) != 0
    assert lib.dsp_block_store_u8(b"f", c.ctypes.data, None, 64, 1.0, None) != 0
    assert capi.last_error(lib)
