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


# ---------------------------------------------------------------------------------------------- several GPUs behind the C ABI
def _single_process_multi_gpu(q):
    """runs in a fresh process: dsp_dct_plan_with_ngpus is process-wide"""
    try:
        import ctypes
        from dspfun_b200 import Plan, REDFT01, REDFT10, capi
        from oracle import dct as od
        lib = capi.load()
        out = []
        for (D, H, W) in ((16, 24, 32), (32, 120, 240), (8, 1080, 1920)):
            x = np.random.default_rng(4).random((D, H, W)).astype(np.float32)
            res = {}
            for ng in (1, 2):
                lib.dsp_dct_plan_with_ngpus(ng)
                fwd = Plan("f", [D, H, W], [REDFT10] * 3)
                inv = Plan("f", [D, H, W], [REDFT01] * 3)
                y = fwd.execute_host(x.copy())
                z = inv.execute_host(y.copy())
                res[ng] = (y, z, lib.dsp_dct_plan_ngpus(fwd._h), lib.dsp_dct_plan_ngpus(inv._h))
                fwd.destroy(); inv.destroy()
            ref = od.dctn_fast(x.astype(np.float64), [od.REDFT10] * 3) if D * H * W < 1 << 21 else None
            out.append(dict(shape=(D, H, W), ngpus=res[2][2:], single=res[1][2:],
                            same_fwd=bool(np.array_equal(res[1][0], res[2][0])), same_inv=bool(np.array_equal(res[1][1], res[2][1])),
                            err_fwd=float(od.rel_l2(res[2][0], ref)) if ref is not None else float(od.rel_l2(res[2][0], res[1][0].astype(np.float64))),
                            err_rt=float(od.rel_l2(res[2][1] / (8.0 * D * H * W), x.astype(np.float64)))))
        # not eligible: an odd frame count stays on one GPU and still works
        lib.dsp_dct_plan_with_ngpus(2)
        p = Plan("f", [5, 8, 16], [REDFT10] * 3)
        x = np.random.default_rng(5).random((5, 8, 16)).astype(np.float32)
        y = p.execute_host(x.copy())
        out.append(dict(shape=(5, 8, 16), ngpus=(lib.dsp_dct_plan_ngpus(p._h),), err_fwd=float(od.rel_l2(y, od.dctn_fast(x.astype(np.float64), [od.REDFT10] * 3)))))
        p.destroy()
        q.put(out)
    except Exception as e:
        q.put(repr(e))
        raise


def test_rank3_host_plans_over_two_gpus_in_one_process():
    """fftw_plan_with_nthreads(2) -> dsp_dct_plan_with_ngpus(2): motion's whole-clip rank-3 plan, host buffers, slab-sharded
    over two GPUs of ONE process with the exchange fused into the transform (peer stores): same coefficients as one GPU"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_single_process_multi_gpu, args=(q,))
    p.start()
    res = q.get(timeout=300)
    p.join(timeout=60)
    assert not isinstance(res, str), res
    for r in res[:-1]:
        assert r["ngpus"] == (2, 2) and r["single"] == (1, 1), r
        assert r["err_fwd"] < 1e-5 and r["err_rt"] < 1e-5, r
        assert r["same_fwd"] and r["same_inv"], r                 # the same kernels on the same lines: identical bits
    assert res[-1]["ngpus"] == (1,) and res[-1]["err_fwd"] < 1e-5


def test_unmodified_motion_tool_over_two_gpus(tmp_path):
    """the reference's motion.c, unmodified, `--fftw-threads 2 -b 0x0x0`: its rank-3 FFTW plans run on two GPUs through the shim"""
    import subprocess
    import torch
    from tests import dspraw
    from tests.test_reference_tools import _tool
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    D, H, W = 8, 64, 96
    vol = np.random.default_rng(6).integers(0, 256, (D, H, W)).astype(np.uint8)
    src = str(tmp_path / "in.dspv")
    dspraw.write_video(src, "gray", W, H, [[vol[z]] for z in range(D)])
    outs = {}
    for ng in (1, 2):
        dst = str(tmp_path / ("out%d.dspv" % ng))
        env = dict(os.environ, DSP_DCT_TRACE="1")
        r = subprocess.run([_tool("motion_gpu_f"), "-Q", "--fftw-threads", str(ng), "-b", "%dx%dx%d" % (W, H, D), "-q", "0.02", src, dst],
                           check=True, env=env, capture_output=True, text=True)
        assert ("multi-GPU plan: 2 devices" in r.stderr) == (ng == 2), r.stderr[-2000:]
        outs[ng] = np.stack([f[0] for f in dspraw.read_video(dst)[3]])
    assert np.array_equal(outs[1], outs[2])
