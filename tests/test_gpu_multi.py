"""GPU suite, needs >= 2 GPUs (skipped otherwise): the slab-sharded 3-D transform with the exchange fused into the
preceding pass (peer-mapped symmetric buffers over NVLink) against the NCCL all-to-all path and the oracle."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, D, H, W, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dspfun_b200.dist3d import Dist3D
        from oracle import dct as od
        vol = np.random.default_rng(7).random((D, H, W)).astype(np.float32)
        Dl, Pl = D // world, H * W // world
        ref = od.dctn_fast(vol.astype(np.float64), [od.REDFT10] * 3).reshape(D, H * W)[:, rank * Pl:(rank + 1) * Pl]
        res = {}
        for mode in ("peer", "nccl"):
            d3 = Dist3D(D, H, W, prec="f", exchange=mode)
            assert d3.mode == mode
            slab = torch.from_numpy(vol[rank * Dl:(rank + 1) * Dl].copy()).cuda()
            cols = d3.forward(slab)
            e_fwd = od.rel_l2(cols.cpu().numpy(), ref)
            back = d3.inverse(cols) / (8.0 * D * H * W)
            e_inv = od.rel_l2(back.cpu().numpy(), vol[rank * Dl:(rank + 1) * Dl])
            # a second round trip reuses the symmetric buffers
            cols2 = d3.forward(back.clone())
            e_fwd2 = od.rel_l2(cols2.cpu().numpy(), ref)
            res[mode] = (e_fwd, e_inv, e_fwd2, cols2.cpu().numpy().copy())
            torch.cuda.synchronize()
            dist.barrier()
            d3.destroy()
        same = np.array_equal(res["peer"][3], res["nccl"][3])
        q.put((rank, res["peer"][:3], res["nccl"][:3], same))
    except Exception as e:                                     # report at once instead of leaving the parent to time out
        q.put((rank, "error", repr(e), False))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(16, 24, 32), (32, 540, 960)])
def test_peer_fused_exchange_world2(shape):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    D, H, W = shape
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, D, H, W, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert not any(r[1] == "error" for r in res), res
    assert all(p.exitcode == 0 for p in procs)
    for rank, peer, nccl, same in res:
        assert max(peer) < 1e-5 and max(nccl) < 1e-5, (rank, peer, nccl)
        assert same, "peer-fused and NCCL exchanges must give identical coefficients"


def _motion_worker(rank, world, port, dims, filt, mode, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from dspfun_b200.dist3d import Dist3D, motion_params
        D, H, W = dims
        vol = np.random.default_rng(21).integers(96, 160, dims).astype(np.uint8)
        d3 = Dist3D(D, H, W, prec="f", exchange=mode, motion=motion_params(dims, **filt))
        Dl = D // world
        pels = torch.from_numpy(vol[rank * Dl:(rank + 1) * Dl].copy()).cuda()
        out = d3.process(pels)
        out2 = d3.process(pels)                                   # symmetric buffers reused
        torch.cuda.synchronize()
        dist.barrier()
        q.put((rank, d3.mode, out.cpu().numpy().copy(), bool(torch.equal(out, out2))))
        d3.destroy()
    except Exception as e:
        q.put((rank, "error", repr(e), False))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["peer", "nccl"])
@pytest.mark.parametrize("filt", [dict(), dict(damp=0.0, bandpass=((0, 0, 0), (8, 60, 120)))])
def test_motion_volume_u8_world2(mode, filt):
    """motion -b 0x0x0 on two GPUs, 8-bit pels in and out (fused pel load, coefficient stage on the temporal pass with flat
    coordinates, pel store): no filter reproduces the source exactly; the low-pass box == the oracle's block loop"""
    import torch
    import torch.multiprocessing as mp
    from oracle import pipelines as op
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    dims = (16, 120, 240)             # 120 = 8 * 15: a radix path (a prime factor > 13 would take the dense axis, which cannot store segments)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_motion_worker, args=(r, 2, port, dims, filt, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert not any(r[1] == "error" for r in res), [r[:3] for r in res]
    assert all(p.exitcode == 0 for p in procs)
    res.sort(key=lambda r: r[0])
    assert all(r[1] == mode and r[3] for r in res)
    got = np.concatenate([r[2] for r in res])
    vol = np.random.default_rng(21).integers(96, 160, dims).astype(np.uint8)
    if not filt:
        assert np.array_equal(got, vol)
        return
    want, _, pel = op.motion_block(vol, dims, **filt)
    diff = got.astype(np.int64) != want.astype(np.int64)
    if diff.any():
        assert np.abs(got.astype(np.int64) - want.astype(np.int64)).max() <= 1
        frac = np.abs(pel.astype(np.float64))
        assert (np.abs(frac - np.floor(frac) - 0.5)[diff] < 2e-3).all()
