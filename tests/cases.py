"""Shared parity cases: the same functions run against the host SIMT emulation (CPU suite) and against the
product CUDA library through its C ABI (GPU suite).  Every case compares with the oracle on seeded inputs."""
import numpy as np

from dspfun_b200 import REDFT01, REDFT10, Plan
from oracle import dct as od

OK = {"f": 1e-5, "d": 1e-12}          # BASELINE.json north_star tolerances (relative L2 on coefficients)
DT = {"f": np.float32, "d": np.float64}
ORK = {REDFT10: od.REDFT10, REDFT01: od.REDFT01}


def check_interleaved_2d(lib, prec, h, w, d, kind, seed=0, inplace=True, definition=False):
    rng = np.random.default_rng(seed)
    x = rng.random((h, w, d)).astype(DT[prec])
    p = Plan.interleaved_2d(prec, h, w, d, kind, lib=lib)
    if inplace:
        y = p.execute_host(x.copy())
    else:
        src = x.copy()
        y = p.execute_host(src, np.full_like(x, 7.0))
        assert np.array_equal(src, x), "out-of-place execute must not modify its input"
    p.destroy()
    x64 = x.astype(np.float64)
    ref = od.dctn_def(x64, [ORK[kind]] * 2, axes=(0, 1)) if definition else od.dctn_fast(x64, [ORK[kind]] * 2, axes=(0, 1))
    err = od.rel_l2(y, ref)
    assert err < OK[prec], (prec, h, w, d, kind, err)
    return err


def check_rank1_batch(lib, prec, n, howmany, kind, dist=None, seed=1):
    """plan_many_r2r(1,{n},howmany, x,NULL,1,dist, ...) planar lines `dist` apart."""
    rng = np.random.default_rng(seed)
    dist = n if dist is None else dist
    buf = rng.random(howmany * dist).astype(DT[prec])
    ref = buf.astype(np.float64).copy()
    for b in range(howmany):
        ref[b * dist:b * dist + n] = od.dctn_fast(buf[b * dist:b * dist + n].astype(np.float64), [ORK[kind]])
    p = Plan(prec, [n], [kind], howmany, None, 1, dist, None, 1, dist, lib=lib)
    y = p.execute_host(buf.copy())
    p.destroy()
    err = od.rel_l2(y, ref)
    assert err < OK[prec], (prec, n, howmany, kind, err)
    # gaps between lines must be untouched
    if dist > n:
        gaps = np.ones(howmany * dist, bool)
        for b in range(howmany):
            gaps[b * dist:b * dist + n] = False
        assert np.array_equal(y[gaps], buf[gaps])
    return err


def check_planar_3d_embed(lib, prec, dims, embed, kind, seed=2):
    """motion's rank-3 plan over a sub-box of a larger zero-initialised buffer (motion/motion.c:535-552)."""
    rng = np.random.default_rng(seed)
    buf = rng.random(embed).astype(DT[prec])
    ref = buf.astype(np.float64).copy()
    sl = tuple(slice(0, s) for s in dims)
    ref[sl] = od.dctn_fast(buf[sl].astype(np.float64), [ORK[kind]] * 3)
    p = Plan.planar_3d(prec, dims, embed, kind, lib=lib)
    y = p.execute_host(buf.copy())
    p.destroy()
    err = od.rel_l2(y[sl], ref[sl])
    assert err < OK[prec], (prec, dims, embed, kind, err)
    mask = np.ones(embed, bool)
    mask[sl] = False
    assert np.array_equal(y[mask], buf[mask]), "elements outside the logical box must be untouched"
    return err


def check_batched_images(lib, prec, nb, h, w, d, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.random((nb, h, w, d)).astype(DT[prec])
    p = Plan.interleaved_2d(prec, h, w, d, REDFT10, nbatch=nb, lib=lib)
    y = p.execute_host(x.copy())
    p.destroy()
    ref = od.dctn_fast(x.astype(np.float64), [od.REDFT10] * 2, axes=(1, 2))
    err = od.rel_l2(y, ref)
    assert err < OK[prec], err
    q = Plan.interleaved_2d(prec, h, w, d, REDFT01, nbatch=nb, lib=lib)
    z = q.execute_host(y.copy()) / (4.0 * h * w)
    q.destroy()
    assert od.rel_l2(z, x) < OK[prec]
    return err


def check_golden_1d(lib, golden, prec, n, typ):
    """FFTW-generated known answers (tests/golden/fftw_dct_ref.npz) through a rank-1 plan."""
    kind = REDFT10 if typ == 2 else REDFT01
    x = np.linspace(0, n - 1, n).astype(DT[prec])
    p = Plan(prec, [n], [kind], lib=lib)
    y = p.execute_host(x.copy())
    p.destroy()
    ref = golden["%s_dct_%d_%d" % ("single" if prec == "f" else "double", typ, n)].astype(np.float64)
    err = od.rel_l2(y, ref)
    assert err < (2e-6 if prec == "f" else 1e-14), (prec, n, typ, err)
    return err


SHAPES_2D = [
    (8, 8, 1), (16, 32, 3), (12, 20, 3), (30, 14, 1), (64, 48, 3), (1, 16, 1), (16, 1, 3), (1, 1, 3), (5, 7, 2),
    (2, 3, 4), (256, 256, 1), (100, 135, 3), (33, 77, 1), (26, 22, 3), (128, 512, 3), (1024, 16, 1), (49, 81, 2),
]
