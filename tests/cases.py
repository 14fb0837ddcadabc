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


def run_planar_3d_embed(lib, prec, dims, embed, kind, seed=2):
    """The raw result of the same plan (for bit-for-bit comparisons between schedules)."""
    buf = np.random.default_rng(seed).random(embed).astype(DT[prec])
    p = Plan.planar_3d(prec, dims, embed, kind, lib=lib)
    y = p.execute_host(buf.copy())
    p.destroy()
    return y


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
    (17, 17, 1), (34, 19, 3), (509, 3, 1), (6, 211, 2),          # prime factors > 13: dense fallback
]


# ---------------------------------------------------------------------------------------------- spec / ispec
from dspfun_b200 import spec as gspec          # noqa: E402
from oracle import pipelines as pl             # noqa: E402

INTERMEDIATE = {"f": np.float64, "d": np.longdouble}   # reference default: INTERMEDIATE_PRECISION = COEFF << 1


def _quantised_matches(ours, ref, bits, tie_tol):
    """Quantised pixels must be identical except where the reference's own unquantised value sits within
    `tie_tol` LSB of a rounding tie (BASELINE.json north_star); even there the difference is at most one LSB."""
    m = (1 << bits) - 1
    qo, qr = pl.quantize_unorm(ours, bits).astype(np.int64), pl.quantize_unorm(ref, bits).astype(np.int64)
    diff = qo != qr
    if diff.any():
        assert np.abs(qo - qr).max() <= 1
        frac = np.clip(np.asarray(ref, dtype=np.float64), 0, 1) * m
        dist = np.abs(frac - np.floor(frac) - 0.5)
        assert (dist[diff] < tie_tol).all(), "quantised mismatch away from a rounding tie: max dist %.3g" % dist[diff].max()
    return float(diff.mean())


def check_spec_presets(lib, prec, h, w, d, seed=0, fast=False):
    """Fused spec epilogue / ispec prologue vs the restated reference loops (spec/spec.c:66-139, ispec.c:84-163)."""
    rng = np.random.default_rng(seed)
    px = (rng.integers(0, 256, (h, w, d)) / 255.0).astype(DT[prec])
    I = INTERMEDIATE[prec]
    for preset in ("abs", "shift", "flat", "sign", "copy"):
        s1, dc1 = gspec.spec(px, preset, lib=lib)
        s0, dc0 = pl.spec_forward(px, preset, intermediate=I, fast=fast)
        np.testing.assert_allclose(dc1, dc0, rtol=0, atol=2e-6 if prec == "f" else 1e-13)
        if preset == "sign":
            # a sign map: every coefficient whose magnitude is above the float noise floor must get the same bit
            raw = od.dctn_fast(px.astype(np.float64), [od.REDFT10] * 2, axes=(0, 1))
            solid = np.abs(raw) > (1e-4 if prec == "f" else 1e-10) * np.abs(raw).max()
            solid.reshape(-1)[:d] = False
            assert np.array_equal(s1[solid], s0[solid])
            continue
        assert od.rel_l2(s1, s0) < OK[prec], (preset, od.rel_l2(s1, s0))
        sm = None
        if preset == "abs":
            smf, _ = pl.spec_forward(px, "sign", intermediate=I, fast=fast)
            sm = np.round(smf * 255).astype(np.uint8)
            sm.reshape(-1)[:d] = np.round(dc0 * 255).astype(np.uint8)
        b1 = gspec.ispec(s0, dc0, preset, signmap=sm, lib=lib)
        b0 = pl.ispec_inverse(s0, dc0, preset, intermediate=I, signmap=sm, fast=fast)
        assert od.rel_l2(b1, b0) < OK[prec], (preset, od.rel_l2(b1, b0))


def check_spec_c1_roundtrip(lib, prec, h=512, w=512, d=3, preset="shift", seed=0):
    """BASELINE config 0: spec -t <preset> -> 16-bit image -> ispec -t <preset> -> 8- and 16-bit pixels, against the
    same chain through the oracle.  Returns the fractions of differing 16-bit spectrogram / 8-bit pixel values."""
    rng = np.random.default_rng(seed)
    px = (rng.integers(0, 256, (h, w, d)) / 255.0).astype(DT[prec])
    I = INTERMEDIATE[prec]
    s1, dc1 = gspec.spec(px, preset, lib=lib)
    s0, dc0 = pl.spec_forward(px, preset, intermediate=I, fast=True)
    # The log scale divides the transform's absolute noise (~1e-7 of the DC term) by (1+|v|): for near-zero
    # coefficients that is several 16-bit steps in single precision -- for FFTW's float path as much as for ours --
    # so the 16-bit spectrogram itself is held to the coefficient tolerance in float and to tie-exactness in double.
    assert od.rel_l2(s1, s0) < OK[prec]
    f16 = _quantised_matches(s1, s0, 16, 1e-6) if prec == "d" else float((pl.quantize_unorm(s1, 16) != pl.quantize_unorm(s0, 16)).mean())
    img = (pl.quantize_unorm(s0, 16) / 65535.0).astype(DT[prec])          # what ispec reads back from the PNG
    b1 = gspec.ispec(img, dc0, preset, lib=lib)
    b0 = pl.ispec_inverse(img, dc0, preset, intermediate=I, fast=True)
    f8 = _quantised_matches(b1, b0, 8, 0.01 if prec == "f" else 1e-6)
    _quantised_matches(b1, b0, 16, 0.5)                                    # at most one LSB anywhere
    # and with the log-scaled template the round trip really reproduces the 8-bit source (spec/README.md:64-70)
    if preset == "shift":
        assert np.array_equal(pl.quantize_unorm(b1, 8), np.round(px.astype(np.float64) * 255).astype(np.uint8))
    return f16, f8


def check_spec_options(lib, prec):
    """Individual -R/-T/-S/-G overrides (spec/spec.h:112-155) incl. rangetype dcs and the reference gain."""
    rng = np.random.default_rng(3)
    px = (rng.integers(0, 256, (24, 40, 3)) / 255.0).astype(DT[prec])
    I = INTERMEDIATE[prec]
    for params in [("log", "shift", "reference", "dcs"), ("linear", "abs", "native", "dc"), ("log", "retain", "custom", "one")]:
        sc, sg, gt, rg = params
        kw = dict(scale=sc, sign=sg, range_=rg, gain=(gt if gt != "custom" else 37.5))
        s1, dc1 = gspec.spec(px, None, lib=lib, **kw)
        s0, dc0 = pl.spec_forward(px, params=params, custom_gain=37.5, intermediate=I)
        assert od.rel_l2(s1, s0) < OK[prec], (params, od.rel_l2(s1, s0))
        if sg != "abs":
            b1 = gspec.ispec(s0, dc0, None, lib=lib, **kw)
            b0 = pl.ispec_inverse(s0, dc0, params=params, custom_gain=37.5, intermediate=I)
            assert od.rel_l2(b1, b0) < OK[prec], (params, od.rel_l2(b1, b0))
            b2 = gspec.ispec(s0, dc0, None, preserve_dc=True, lib=lib, **kw)
            assert od.rel_l2(b2, pl.ispec_inverse(s0, dc0, params=params, custom_gain=37.5, intermediate=I, preserve_dc=True)) < OK[prec]


# ---------------------------------------------------------------------------------------------- scan
from dspfun_b200 import scan as gscan           # noqa: E402


def check_scan(lib, prec, h, w, d, order="diagonal", step=1, seed=4):
    """Fused masked inverse + accumulate (scan/scan.c:421-459) vs the restated loop; the last frame must also pass
    the reference's own --measure-parity self-check (scan/scan.c:508-526) at the source bit depth."""
    rng = np.random.default_rng(seed)
    px8 = rng.integers(0, 256, (h, w, d))
    px = (px8 / 255.0).astype(DT[prec])
    idx = gscan.order_diagonal(h, w) if order == "diagonal" else gscan.order_horizontal(h, w)
    ref_frames, ref_coeffs = pl.scan_frames(px, idx, step=step, fast=max(h, w) > 64)
    s = gscan.Scan(px, idx, lib=lib)
    assert od.rel_l2(s.coeffs(), ref_coeffs) < OK[prec]
    limit = int(idx.max()) + 1
    nframes = (limit + step - 1) // step
    assert nframes == len(ref_frames)
    worst = 0.0
    for i in range(nframes):
        want = i % max(1, nframes // 6) == 0 or i == nframes - 1          # D2H only a few frames; all are computed
        f = s.frame(i * step, min((i + 1) * step, limit), want=want)
        if want:
            worst = max(worst, od.rel_l2(f, ref_frames[i]))
    s.destroy()
    assert worst < OK[prec] * 4, worst
    # scan.c:508-526: lround(orig * (2^depth - 1)) == lround(sum * (2^depth - 1))
    assert np.array_equal(np.round(f.astype(np.float64) * 255).astype(np.int64), px8)
    return worst


# ---------------------------------------------------------------------------------------------- motion
from dspfun_b200 import motion as gmotion       # noqa: E402


def check_motion(lib, block, scaled=None, prec="f", float_pixels=False, seed=6, **filt):
    """motion's per-block body on the GPU vs the restated loop (motion/motion.c:617-788).  8-bit output must be
    bit-exact except where the reference's own unrounded pel sits on a rounding tie."""
    rng = np.random.default_rng(seed)
    scaled = tuple(scaled) if scaled is not None else tuple(block)
    minbuf = tuple(max(a, b) for a, b in zip(block, scaled))
    if float_pixels:
        pels = (rng.integers(16, 236, minbuf) / 255.0).astype(np.float32)
    else:
        pels = rng.integers(16, 236, minbuf).astype(np.uint8)
    # smooth the volume a little so that quantisation / thresholds leave something behind
    m = gmotion.Motion(block, scaled, float_pixels=float_pixels, prec=prec, lib=lib, **filt)
    got = m.process(pels)
    coded = m.coeffs_coded
    m.destroy()
    want, coded_ref, pel = pl.motion_block(pels, block, scaled, coeff=DT[prec],
                                           intermediate=np.float64 if prec == "f" else np.longdouble, **filt)
    ssl = tuple(slice(0, v) for v in scaled)
    outside = np.ones(minbuf, bool)
    outside[ssl] = False
    assert np.array_equal(got[outside], pels[outside]), "pels outside the scaled box must be untouched"
    if float_pixels:
        assert od.rel_l2(got[ssl], want[ssl]) < 1e-5
        return 0.0
    diff = got[ssl].astype(np.int64) != want[ssl].astype(np.int64)
    if diff.any():
        assert np.abs(got[ssl].astype(np.int64) - want[ssl].astype(np.int64)).max() <= 1
        frac = np.abs(pel.astype(np.float64))
        dist = np.abs(frac - np.floor(frac) - 0.5)
        # a tie is "the reference's own unrounded pel within the float error of x.5": 2e-3 of a level for pels in 0..255,
        # proportionally more when the unclamped values are far larger (a random block read as a spectrogram: --ispec)
        tie = max(2e-3, 4e-7 * float(np.abs(pel).max())) if prec == "f" else 1e-8
        assert (dist[diff] < tie).all(), (dist[diff].max(), tie)
    if filt.get("quant"):
        assert abs(coded - coded_ref) <= max(2, coded_ref // 500), (coded, coded_ref)
    return float(diff.mean())


# ---------------------------------------------------------------------------------------------- zoom
from dspfun_b200 import zoom as gzoom           # noqa: E402


def check_zoom(lib, prec, h, w, seed=8, **kw):
    """zoom's scaled-basis synthesis (zoom/zoom.c:36-68, 361-375) on the GPU vs the restated loops."""
    rng = np.random.default_rng(seed)
    px = (rng.integers(0, 256, (h, w, 3)) / 255.0).astype(DT[prec])
    z = gzoom.Zoom(px, lib=lib)
    got = z.frame(**kw)
    path = z.last_path
    z.destroy()
    okw = dict(kw)
    for k in ("scale", "xscale", "yscale"):
        if k in okw and not isinstance(okw[k], (tuple, list)):
            okw[k] = (okw[k], 1)
    want = pl.zoom_synthesise(px, intermediate=np.longdouble, **okw)
    assert got.shape == want.shape, (got.shape, want.shape)
    err = od.rel_l2(got, want)
    assert err < OK[prec], (kw, err)
    return path, got, want


def block_dct2d_reference(x, B, kind):
    """every B x B block of [P][H][W] through the oracle's 2-D transform, in double"""
    P, H, W = x.shape
    want = np.empty(x.shape, dtype=np.float64)
    for p in range(P):
        for y in range(0, H, B):
            for xx in range(0, W, B):
                want[p, y:y + B, xx:xx + B] = od.dctn_fast(x[p, y:y + B, xx:xx + B].astype(np.float64), [ORK[kind], ORK[kind]])
    return want


def check_block_dct2d(lib, shape, B, kind, tol, device="cpu", seed=17, in_place=False, scale=1.0):
    """dsp_block_dct2d (tensor-core GEMM on the GPU, plain loops in the emulation build) == the oracle per block"""
    import torch
    P, H, W = shape
    x = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    d_in = torch.from_numpy(x.copy()).to(device)
    d_out = d_in if in_place else torch.full_like(d_in, np.nan)
    stream = torch.cuda.current_stream().cuda_stream if device != "cpu" else None
    rc = lib.dsp_block_dct2d(b"f", d_in.data_ptr(), d_out.data_ptr(), P, H, W, B, kind, scale, stream)
    assert rc == 0, capi_last_error(lib)
    if device != "cpu":
        torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    if not in_place:
        assert np.array_equal(d_in.cpu().numpy(), x), "out-of-place call must leave the input alone"
    err = od.rel_l2(got, block_dct2d_reference(x, B, kind) * scale)
    assert err < tol, (shape, B, kind, err)
    return err


def check_motion_tiled_c_session(lib, dims, block, quant, device="cpu", seed=12):
    """dsp_motion_tiled_* (the C-ABI session) == MotionTiled (the Python orchestration of the same calls), pels and counts"""
    import ctypes
    import torch
    from dspfun_b200.motion import MotionTiled
    v = torch.from_numpy(np.random.default_rng(seed).integers(16, 236, dims).astype(np.uint8)).to(device)
    t = lib.dsp_motion_tiled_create(*dims, *block, float(quant))
    assert t, capi_last_error(lib)
    out = torch.empty_like(v)
    cnt = ctypes.c_ulonglong(0)
    stream = torch.cuda.current_stream().cuda_stream if device != "cpu" else None
    assert lib.dsp_motion_tiled_process_dev(t, v.data_ptr(), out.data_ptr(), ctypes.byref(cnt), stream) == 0, capi_last_error(lib)
    again = v.clone()                                             # in place, no count requested
    assert lib.dsp_motion_tiled_process_dev(t, again.data_ptr(), again.data_ptr(), None, stream) == 0
    if device != "cpu":
        torch.cuda.synchronize()
    lib.dsp_motion_tiled_destroy(t)
    mt = MotionTiled(dims, block, quant=quant, lib=lib)
    want = mt.process(v.clone())
    assert torch.equal(out, want) and torch.equal(again, want)
    assert cnt.value == mt.coeffs_coded
    mt.destroy()


def capi_last_error(lib):
    from dspfun_b200 import capi
    return capi.last_error(lib)


def check_motion_tiled(lib, dims, block, quant=0.0, seed=12, device="cpu", gemm=True):
    """MotionTiled (all blocks of a volume through three per-axis plans) == the reference block loop, block by block:
    8-bit output identical except where the reference's unrounded pel sits on a rounding tie; coded counts equal."""
    import torch
    from dspfun_b200.motion import MotionTiled
    D, H, W = dims
    bd, bh, bw = block
    v = np.random.default_rng(seed).integers(16, 236, dims).astype(np.uint8)
    mt = MotionTiled(dims, block, quant=quant, lib=lib, gemm=gemm)
    out = mt.process(torch.from_numpy(v.copy()).to(device)).cpu().numpy()
    # sharding along d in whole blocks needs no exchange: the two halves processed separately give the same pels
    if (D // bd) % 2 == 0:
        half = MotionTiled((D // 2, H, W), block, quant=quant, lib=lib, gemm=gemm)
        lo = half.process(torch.from_numpy(v[:D // 2].copy()).to(device)).cpu().numpy()
        hi = half.process(torch.from_numpy(v[D // 2:].copy()).to(device)).cpu().numpy()
        half.destroy()
        assert np.array_equal(np.concatenate([lo, hi]), out)
    coded = 0
    for z in range(0, D, bd):
        for y in range(0, H, bh):
            for x in range(0, W, bw):
                sl = (slice(z, z + bd), slice(y, y + bh), slice(x, x + bw))
                o, c, pel = pl.motion_block(v[sl].copy(), block, quant=quant)
                coded += c
                diff = out[sl] != o
                if diff.any():
                    frac = np.abs(pel - np.floor(pel))[diff]
                    assert np.all(np.abs(frac - 0.5) < 1e-3), "8-bit mismatch away from a rounding tie"
    if quant:
        assert abs(mt.coeffs_coded - coded) <= max(2, coded // 10000)       # a coefficient on a quantiser tie may flip
    mt.destroy()
