"""CPU suite, round-2 additions on the emulated kernels:
  * the chunked (L2-resident) schedule: consecutive passes that share their outermost loop level (batch elements,
    frames) run chunk by chunk -- results must not depend on the chunk size, fused coordinates included;
  * fused scan / spec stages when the contiguous axis runs as a strided (column) pass because the interleaved line
    does not fit on chip (the channel coordinate was overwritten there)."""
import os

import numpy as np
import pytest

from dspfun_b200 import REDFT01, REDFT10, Plan
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
