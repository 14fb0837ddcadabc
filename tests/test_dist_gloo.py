"""CPU suite, world_size 2 over gloo: the slab-sharded 3-D transform's host logic (partitioning, all-to-all order,
plan geometry) with the emulated kernels, and the batched-images sharding used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import dct as od


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, D, H, W, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dspfun_b200.dist3d import Dist3D
        from tests.emu import emu
        lib = emu.load()
        rng = np.random.default_rng(7)
        vol = rng.random((D, H, W)).astype(np.float64)
        d3 = Dist3D(D, H, W, prec="d", lib=lib)
        Dl, Pl = D // world, H * W // world
        slab = torch.from_numpy(vol[rank * Dl:(rank + 1) * Dl].copy())
        cols = d3.forward(slab)
        ref = od.dctn_fast(vol, [od.REDFT10] * 3).reshape(D, H * W)[:, rank * Pl:(rank + 1) * Pl]
        e_fwd = od.rel_l2(cols.numpy(), ref)
        back = d3.inverse(cols.clone()) / (8.0 * D * H * W)
        e_inv = od.rel_l2(back.numpy(), vol[rank * Dl:(rank + 1) * Dl])
        d3.destroy()
        q.put((rank, e_fwd, e_inv, d3.a2a_bytes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(8, 6, 10), (16, 9, 16), (4, 27, 30)])
def test_slab_sharded_3d_world2(shape):
    D, H, W = shape
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, D, H, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e_fwd, e_inv, nbytes in res:
        assert e_fwd < 1e-12 and e_inv < 1e-12, (rank, e_fwd, e_inv)
        assert nbytes == 2 * (D // 2) * (H * W) * 8 // 2          # two exchanges, half the local slab leaves each time


def test_single_rank_uses_one_rank3_plan():
    from dspfun_b200.dist3d import Dist3D
    from tests.emu import emu
    rng = np.random.default_rng(3)
    vol = rng.random((4, 6, 8))
    d3 = Dist3D(4, 6, 8, prec="d", lib=emu.load())
    out = d3.forward(torch.from_numpy(vol.copy()))
    assert od.rel_l2(out.numpy(), od.dctn_fast(vol, [od.REDFT10] * 3)) < 1e-12
    d3.destroy()


def _scan_worker(rank, world, port, h, w, d, step, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dspfun_b200 import scan as gscan
        from tests.emu import emu
        rng = np.random.default_rng(5)
        px = rng.random((h, w, d))
        idx = gscan.order_diagonal(h, w)
        f0, frames = gscan.scan_frames_sharded(px, idx, step=step, lib=emu.load())     # rank / world from the group
        q.put((rank, f0, [f.copy() for f in frames]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cfg", [(12, 10, 3, 1), (16, 16, 1, 3), (9, 14, 2, 2)])
def test_scan_sharded_world2(cfg):
    """scan's frames over 2 ranks with the linear-prefix start == the sequential oracle"""
    from oracle import pipelines as op
    from dspfun_b200 import scan as gscan
    h, w, d, step = cfg
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_scan_worker, args=(r, 2, port, h, w, d, step, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    px = np.random.default_rng(5).random((h, w, d))
    ref, _ = op.scan_frames(px, gscan.order_diagonal(h, w), step=step)
    got = [f for _, _, fr in res for f in fr]
    assert res[0][1] == 0 and res[1][1] == len(res[0][2]) and len(got) == len(ref)
    for g, r in zip(got, ref):
        assert od.rel_l2(g, r) < 1e-12
    assert od.rel_l2(got[-1], px) < 1e-12                       # the last frame is the image itself


def test_scan_shard_ranges_cover():
    from dspfun_b200.scan import shard_range
    for n in (1, 5, 16, 17):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))


def _pel_diff_ok(got, want, pel):
    """8-bit pels: identical except where the reference's own unrounded value sits on a rounding tie"""
    diff = got.astype(np.int64) != want.astype(np.int64)
    if diff.any():
        assert np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1
        frac = np.abs(pel.astype(np.float64))
        assert (np.abs(frac - np.floor(frac) - 0.5)[diff] < 2e-3).all()
    return float(diff.mean())


def _motion_worker(rank, world, port, dims, filt, q, fuse=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if fuse:
        os.environ["DSP_DIST_FUSE_COEFF"] = "1"        # the stage carried in the temporal pass's store instead of its own sweep
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dspfun_b200.dist3d import Dist3D, motion_params
        from tests.emu import emu
        D, H, W = dims
        vol = np.random.default_rng(21).integers(16, 236, dims).astype(np.uint8)
        kw = dict(filt)
        kw["preserve_dc"] = {None: 0, "dc": 1, "grey": 2}[kw.get("preserve_dc")]
        d3 = Dist3D(D, H, W, prec="f", lib=emu.load(), motion=motion_params(dims, **kw))
        Dl = D // world
        out = d3.process(torch.from_numpy(vol[rank * Dl:(rank + 1) * Dl].copy()))
        d3.destroy()
        q.put((rank, out.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("filt", [dict(), dict(damp=0.0, bandpass=((0, 0, 0), (4, 6, 8))),
                                  dict(boost=1.5, bandpass=((1, 2, 2), (6, 10, 12)), preserve_dc="dc"), dict(quant=0.02)])
def test_motion_volume_u8_world2(filt):
    """motion -b 0x0x0 over two ranks: 8-bit pels in, fused pel load / coefficient stages (flat coordinates on the
    temporal pass) / pel store, 8-bit pels out == the restated reference block loop on the whole volume"""
    from oracle import pipelines as op
    dims = (8, 12, 16)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_motion_worker, args=(r, 2, port, dims, filt, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([r[1] for r in res])
    vol = np.random.default_rng(21).integers(16, 236, dims).astype(np.uint8)
    want, _, pel = op.motion_block(vol, dims, **filt)
    _pel_diff_ok(got, want, pel)
    if not filt:
        assert np.array_equal(got, vol)          # no filter: the round trip reproduces the 8-bit source


def test_motion_volume_u8_world2_stage_carried_in_the_pass():
    """DSP_DIST_FUSE_COEFF=1: the coefficient stage rides in the store of the forward temporal pass (flat coordinates) -- the
    variant the sweep replaced; same pels"""
    from oracle import pipelines as op
    dims = (8, 12, 16)
    filt = dict(boost=1.5, bandpass=((1, 2, 2), (6, 10, 12)), preserve_dc="dc")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_motion_worker, args=(r, 2, port, dims, filt, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = np.concatenate([r[1] for r in res])
    vol = np.random.default_rng(21).integers(16, 236, dims).astype(np.uint8)
    want, _, pel = op.motion_block(vol, dims, **filt)
    _pel_diff_ok(got, want, pel)


def test_motion_volume_single_rank_matches_session():
    from dspfun_b200.dist3d import Dist3D, motion_params
    from oracle import pipelines as op
    from tests.emu import emu
    dims = (4, 10, 12)
    vol = np.random.default_rng(22).integers(16, 236, dims).astype(np.uint8)
    d3 = Dist3D(*dims, prec="f", lib=emu.load(), motion=motion_params(dims, damp=0.25, bandpass=((0, 1, 1), (4, 8, 9))))
    out = d3.process(torch.from_numpy(vol.copy())).numpy()
    d3.destroy()
    want, _, pel = op.motion_block(vol, dims, damp=0.25, bandpass=((0, 1, 1), (4, 8, 9)))
    _pel_diff_ok(out, want, pel)
