"""GPU suite (-m gpu): the BASELINE.json configurations at FULL size ("tested at reduced size only").

The oracle cannot transform these sizes whole in seconds, so every case checks the CUDA result against the oracle on
sampled lines, using separability: a 2-D / 3-D DCT restricted to one output line is a 1-D oracle transform of a weighted
sum of the input lines (the weights are the cosine basis of the other axes).  The weighted sums are float64 reductions
(torch on the device holds the operands; no transform code of the product is involved), the 1-D transforms are the
oracle's.  Tolerances: BASELINE.json north_star (relative L2 1e-5 float / 1e-12 double; 8-bit pels exact off ties).
"""
import numpy as np
import pytest

from dspfun_b200 import REDFT01, REDFT10, Plan, capi
from oracle import dct as od

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return capi.load()


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "the gpu suite needs a CUDA device"
    return torch


def _cos2(torch, n, k, dtype):
    """REDFT10 weights of output index k along an axis of length n: 2 cos(pi (j + 1/2) k / n)"""
    j = torch.arange(n, device="cuda", dtype=torch.float64)
    return 2 * torch.cos(torch.pi * (j + 0.5) * float(k) / n)


@pytest.mark.parametrize("prec,tol", [("f", 1e-5), ("d", 1e-12)])
def test_c3_plane8192_forward_64_rows_64_columns(lib, torch, prec, tol):
    """C3: spec's forward transform of an 8192 x 8192 single-plane image, float and double: 64 output rows and 64 output
    columns against the separable definition; then the inverse must give 4 w h x back."""
    n = 8192
    tdt = torch.float32 if prec == "f" else torch.float64
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand((n, n), device="cuda", dtype=tdt, generator=g)
    y = x.clone()
    fwd = Plan.interleaved_2d(prec, n, n, 1, REDFT10)
    fwd.execute_dev(y.data_ptr(), y.data_ptr(), None)
    torch.cuda.synchronize()
    xh = x.double()
    rng = np.random.default_rng(0)
    ks = sorted(set([0, 1, n // 2, n - 1] + [int(v) for v in rng.integers(0, n, 60)]))[:64]
    # The tolerance is BASELINE's relative L2 over the coefficients, here over the sampled lines together (they include row
    # 0 and column 0, which carry the DC terms).  A single line is held to 10x that: every FFT-based transform -- FFTW's
    # too -- leaves an absolute error proportional to the largest term of each 1-D pass, and column 0 holds Y[0,0] ~ 1e8
    # next to terms of ~ 1e4.
    worst, e2, n2 = 0.0, 0.0, 0.0
    for k in ks:
        # row k of the result: Y[k, :] = REDFT10_x( sum_y 2 cos(pi (y + 1/2) k / n) x[y, :] )
        row = (_cos2(torch, n, k, tdt)[:, None] * xh).sum(0).cpu().numpy()
        # column k: Y[:, k] = REDFT10_y( sum_x 2 cos(pi (x + 1/2) k / n) x[:, x] )
        col = (xh * _cos2(torch, n, k, tdt)[None, :]).sum(1).cpu().numpy()
        for got, src in ((y[k], row), (y[:, k], col)):
            ref = od.dctn_fast(src, [od.REDFT10])
            g = got.cpu().numpy().astype(np.float64)
            worst = max(worst, od.rel_l2(g, ref))
            e2 += float(np.sum((g - ref) ** 2)); n2 += float(np.sum(ref ** 2))
    assert np.sqrt(e2 / n2) < tol, np.sqrt(e2 / n2)
    assert worst < 10 * tol, worst
    inv = Plan.interleaved_2d(prec, n, n, 1, REDFT01).fuse_scale(1.0, 1.0 / (4.0 * n * n))
    inv.execute_dev(y.data_ptr(), y.data_ptr(), None)
    torch.cuda.synchronize()
    err = (torch.linalg.norm((y - x).double()) / torch.linalg.norm(xh)).item()
    assert err < tol, err
    fwd.destroy(); inv.destroy()


def test_c4_batch_of_1024_rgb_images(lib, torch):
    """C4 shape: a batch of 1024 x 1024 RGB float images through one batched plan pair (384 images, 4.8 GB: the batch level
    is what the ranks shard); 6 sampled images against the oracle, every image against the round-trip identity."""
    nb, h, w, d = 384, 1024, 1024, 3
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand((nb, h, w, d), device="cuda", dtype=torch.float32, generator=g)
    y = x.clone()
    fwd = Plan.interleaved_2d("f", h, w, d, REDFT10, nbatch=nb)
    inv = Plan.interleaved_2d("f", h, w, d, REDFT01, nbatch=nb).fuse_scale(1.0, 1.0 / (4.0 * h * w))
    fwd.execute_dev(y.data_ptr(), y.data_ptr(), None)
    torch.cuda.synchronize()
    for b in (0, 1, 127, 200, 382, 383):
        ref = od.dctn_fast(x[b].cpu().numpy().astype(np.float64), [od.REDFT10] * 2, axes=(0, 1))
        assert od.rel_l2(y[b].cpu().numpy(), ref) < 1e-5
    inv.execute_dev(y.data_ptr(), y.data_ptr(), None)
    torch.cuda.synchronize()
    per_image = (torch.linalg.vector_norm((y - x).double().flatten(1), dim=1) / torch.linalg.vector_norm(x.double().flatten(1), dim=1))
    assert per_image.max().item() < 1e-5
    fwd.destroy(); inv.destroy()


def _projector(n, keep):
    """P = IDCT . diag(mask) . DCT / (2n) as a dense [n][n] float64 matrix (oracle transforms of the unit vectors)"""
    eye = np.eye(n)
    C = od.dctn_fast(eye, [od.REDFT10], axes=(0,))              # C[k, j]: coefficient k of unit vector j
    C[keep:, :] = 0.0
    return od.dctn_fast(C, [od.REDFT01], axes=(0,)) / (2.0 * n)  # [out sample][in sample]


@pytest.mark.parametrize("plane", ["Y", "U"])
def test_c5_motion_volume_u8_full_size(lib, torch, plane):
    """C5: `motion -b 0x0x0` on a 256-frame plane of a 1920x1080 yuv420p volume (Y 1080x1920, chroma 540x960), 8-bit pels
    in and out through the fused pel load / coefficient stage / pel store.
      * no filter: the output must equal the 8-bit source exactly;
      * -p 0x0x0-(w/2)x(h/2)x128 -D 0 (spectral low-pass box): sampled lines along x, y and t against the oracle -- the
        box mask is a product of per-axis masks, so the pipeline is P_t (x) P_y (x) P_x with P = IDCT . mask . DCT."""
    from dspfun_b200.dist3d import Dist3D, motion_params
    D, H, W = (256, 1080, 1920) if plane == "Y" else (256, 540, 960)
    g = torch.Generator(device="cuda").manual_seed(9)
    # smooth-ish content so that the low-pass keeps pels inside [0, 255]
    pels = torch.randint(96, 160, (D, H, W), device="cuda", dtype=torch.uint8, generator=g)
    d3 = Dist3D(D, H, W, "f", motion=motion_params((D, H, W)))
    out = d3.process(pels)
    torch.cuda.synchronize()
    assert torch.equal(out, pels)
    d3.destroy()

    box = (D // 2, H // 2, W // 2)
    d3 = Dist3D(D, H, W, "f", motion=motion_params((D, H, W), damp=0.0, bandpass=((0, 0, 0), box)))
    out = d3.process(pels)
    torch.cuda.synchronize()
    d3.destroy()
    Pt, Py, Px = (torch.from_numpy(_projector(n, k)).cuda() for n, k in zip((D, H, W), box))
    xd = pels.double()
    rng = np.random.default_rng(5)
    checked = off = 0
    for _ in range(4):
        z0, y0, x0 = int(rng.integers(0, D)), int(rng.integers(0, H)), int(rng.integers(0, W))
        lines = [
            ((z0, y0, slice(None)), Px @ torch.einsum("z,y,zyx->x", Pt[z0], Py[y0], xd)),
            ((z0, slice(None), x0), Py @ torch.einsum("z,x,zyx->y", Pt[z0], Px[x0], xd)),
            ((slice(None), y0, x0), Pt @ torch.einsum("y,x,zyx->z", Py[y0], Px[x0], xd)),
        ]
        for sl, want in lines:
            got = out[sl].cpu().numpy().astype(np.int64)
            want = want.cpu().numpy()
            ref = np.clip(np.floor(np.abs(want) + 0.5) * np.sign(want), 0, 255).astype(np.int64)      # motion.c:776 lround + clamp
            diff = got != ref
            checked += got.size; off += int(diff.sum())
            if diff.any():
                # only where the reference's own unrounded pel sits on a rounding tie (float coefficients: 2e-3 of a level)
                assert np.abs(got - ref).max() <= 1
                frac = np.abs(want - np.floor(want) - 0.5)
                assert (frac[diff] < 2e-3).all(), frac[diff].max()
    assert off <= max(1, checked // 200), (off, checked)


def test_c2_zoom_2x_of_4096_rgb(lib):
    """C2: zoom -s 2 of a 4096 x 4096 RGB float image (default interpolated basis) -> 8192 x 8192: the even output samples
    are the input samples; 48 random outputs against the synthesis sum evaluated directly in float64 (zoom.c:361-375)."""
    from dspfun_b200 import zoom as gzoom
    rng = np.random.default_rng(1)
    n = 4096
    px = (rng.integers(0, 256, (n, n, 3)) / 255.0).astype(np.float32)
    z = gzoom.Zoom(px)
    out = z.frame(scale=2)
    assert z.last_path == "shifted-dct"
    z.destroy()
    assert out.shape == (2 * n, 2 * n, 3)
    assert np.abs(out[::2, ::2] - px).max() < 2e-5
    C = od.dctn_fast(px.astype(np.float64), [od.REDFT10] * 2, axes=(0, 1))
    u = np.arange(n)
    for _ in range(48):
        j, i, c = int(rng.integers(0, 2 * n)), int(rng.integers(0, 2 * n)), int(rng.integers(0, 3))
        xb = np.cos(np.pi * (i / 2 + 0.5) * u / n); xb[0] = 0.5
        yb = np.cos(np.pi * (j / 2 + 0.5) * u / n); yb[0] = 0.5
        want = yb @ C[:, :, c] @ xb / (n * n)
        assert abs(out[j, i, c] - want) < 2e-5, (j, i, c, out[j, i, c], want)
