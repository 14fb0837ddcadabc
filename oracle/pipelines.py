"""oracle.pipelines -- TEST INFRASTRUCTURE ONLY.

numpy restatements of the per-tool pointwise pipelines that surround the FFTW calls in the reference, each
following the cited lines loop by loop, with the reference's two-level precision system
(/root/reference/include/precision.h:59-122): `coeff` = storage dtype, `intermediate` = arithmetic dtype
(np.longdouble reproduces the reference's default INTERMEDIATE_PRECISION=L on x86-64).  After every reference
loop the buffer is rounded back to `coeff`, exactly where the C code stores into `coeff* f`.
"""
import numpy as np

from . import dct as odct

SQRT2_L = np.sqrt(np.longdouble(2))

# spec/spec.h:71-77 presets: (scaletype, signtype, gaintype, rangetype)
SPEC_PRESETS = {
    "abs":   ("log",    "abs",      "native", "dc"),
    "shift": ("log",    "shift",    "native", "one"),
    "flat":  ("linear", "shift",    "custom", "one"),
    "sign":  ("linear", "saturate", "custom", "one"),
    "copy":  ("linear", "retain",   "custom", "one"),
}


def _transform(f, kinds, fast):
    """2-D separable r2r over axes (0,1) of an [h][w][d] interleaved buffer (plan_many_r2r howmany=d,stride=d,dist=1)."""
    if fast:
        return odct.dctn_fast(f, kinds, axes=(0, 1)).astype(f.dtype)
    return odct.dctn_def(f, kinds, axes=(0, 1))


def spec_gain(gaintype, w, h, custom_gain, I):
    # spec/spec.c:81-87, spec/ispec.c:111-117
    if gaintype in ("none", "native"):
        return I(127.5) * np.sqrt(I(w * h * 4))
    if gaintype == "reference":
        return I(127.5) * I(1024)
    return I(custom_gain)


def spec_forward(pixels, preset="abs", custom_gain=1.0, intermediate=np.longdouble, fast=False, params=None):
    """spec/spec.c:59-139.  pixels: [h][w][d] coeff-dtype array in [0,1].  Returns (spectrogram, DC[d] float64)."""
    C = pixels.dtype.type
    I = intermediate
    scaletype, signtype, gaintype, rangetype = params if params else SPEC_PRESETS[preset]
    h, w, d = pixels.shape
    f = _transform(np.ascontiguousarray(pixels), [odct.REDFT10, odct.REDFT10], fast)      # :63-65
    DC = np.array([np.float64(f[0, 0, z]) / (w * h * 4) for z in range(d)])                 # :66-68
    f[0, :, :] = (f[0, :, :].astype(I) / SQRT2_L.astype(I)).astype(C)                       # :70-71
    f[:, 0, :] = (f[:, 0, :].astype(I) / SQRT2_L.astype(I)).astype(C)                       # :72-74
    norm = I(w * h * 2)
    f = (f.astype(I) / norm).astype(C)                                                      # :76-78
    gain = spec_gain(gaintype, w, h, custom_gain, I)
    f = (f.astype(I) * gain).astype(C)                                                      # :89-90
    mx = np.empty(d, dtype=C)
    if rangetype == "one":                                                                  # :92-108
        mx[:] = C(gain)
    elif rangetype in ("none", "dc"):
        mx[:] = f[0, 0, :].max()
    elif rangetype == "dcs":
        mx[:] = f[0, 0, :]
    if scaletype in ("none", "log"):                                                        # :110-117
        mx = np.log1p(mx)                                                                   # mc(log1p): coeff precision
        fi = f.astype(I)
        f = (np.copysign(np.log1p(np.abs(f).astype(I)), fi) / mx.astype(I)).astype(C)
    else:                                                                                   # :118-121
        f = (f / mx).astype(C)
    if signtype in ("none", "abs"):                                                         # :124-139
        f = np.abs(f)
    elif signtype == "shift":
        f = ((f.astype(I) / I(2) + I(0.5)) * 254 / 255).astype(C)
    elif signtype == "saturate":
        flat = f.reshape(-1)
        flat[d:] = (~np.signbit(flat[d:])).astype(C)
    return f, DC


def ispec_inverse(spec_img, DC, preset="abs", custom_gain=1.0, intermediate=np.longdouble, fast=False,
                  preserve_dc=False, signmap=None, params=None):
    """spec/ispec.c:79-167.  spec_img: [h][w][d] coeff-dtype spectrogram; DC: float64[d] (the image property).
    signmap: optional uint8 [h][w][d] sign image (spec/ispec.c:87-98)."""
    C = spec_img.dtype.type
    I = intermediate
    scaletype, signtype, gaintype, rangetype = params if params else SPEC_PRESETS[preset]
    h, w, d = spec_img.shape
    f = np.array(spec_img, copy=True)
    DC = np.array(DC, dtype=np.float64, copy=True)
    if signtype in ("none", "abs"):                                                         # :84-99
        if signmap is not None:
            tmp = signmap.reshape(-1)
            DC = (tmp[:d].astype(I) / I(255.)).astype(np.float64)
            flat = f.reshape(-1)
            flat[d:] = np.copysign(flat[d:], (tmp[d:].astype(np.int32) - 128).astype(C))
    elif signtype == "shift":                                                               # :100-103
        f = ((f.astype(I) * I(255.) / 254 - I(0.5)) * 2).astype(C)
    elif signtype == "saturate":                                                            # :104-107
        flat = f.reshape(-1)
        flat[d:] = flat[d:] * 2 - 1
    gain = spec_gain(gaintype, w, h, custom_gain, I)
    mx = np.empty(d, dtype=C)
    if rangetype == "one":                                                                  # :119-134
        mx[:] = C(gain)
    elif rangetype in ("none", "dc"):
        mx[:] = C((DC.astype(I) * gain).max())
    elif rangetype == "dcs":
        mx[:] = (DC.astype(I) * gain).astype(C)
    if scaletype in ("none", "log"):                                                        # :136-143
        mx = np.log1p(mx.astype(np.float64)).astype(C)                                      # plain log1p (double)
        prod = f * mx                                                                       # coeff*coeff
        f = np.copysign(np.expm1(np.abs(prod).astype(I)), f.astype(I)).astype(C)
    else:
        f = (f * mx).astype(C)                                                              # :144-147
    f = (f.astype(I) / gain).astype(C)                                                      # :150-151
    f[0, :, :] = (f[0, :, :].astype(I) * SQRT2_L.astype(I)).astype(C)                       # :153-154
    f[:, 0, :] = (f[:, 0, :].astype(I) * SQRT2_L.astype(I)).astype(C)                       # :155-157
    f = (f / C(2)).astype(C)                                                                # :158-159
    if preserve_dc:                                                                         # :161-163
        f[0, 0, :] = DC.astype(C)
    return _transform(np.ascontiguousarray(f), [odct.REDFT01, odct.REDFT01], fast)          # :165-167


def quantize_unorm(x, bits):
    """[0,1] float -> unsigned integer of `bits` bits, clamp then round half away from zero.  This is what
    ImageMagick's ClampToQuantum (Q16: floor(v*65535+0.5) after clamping) and swscale do when the reference's
    MagickConstituteImage/MagickWriteImage quantise the float buffer (spec/spec.c:141-150, spec/ispec.c:170-182).
    Neither library is in this image; ties are excluded by BASELINE.json's north_star."""
    m = (1 << bits) - 1
    v = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0) * m
    return np.floor(v + 0.5).astype(np.uint16 if bits > 8 else np.uint8)


def base16enc(raw: bytes) -> str:
    """spec/spec.h:157-163: low nibble first, alphabet 'A'+nibble."""
    return "".join(chr((b & 15) + 65) + chr((b >> 4) + 65) for b in raw)


def base16dec(s: str) -> bytes:
    """spec/spec.h:164-168."""
    return bytes(((ord(s[i]) - 65) | ((ord(s[i + 1]) - 65) << 4)) & 0xFF for i in range(0, len(s), 2))


def scan_frames(pixels, index_map, step=1, nframes=None, fast=False):
    """scan/scan.c:292-298 (forward, /= 4wh), :377-383 (sum = DC), :421-459 (per frame: zero reconstruction, copy the
    interval's coefficients, clear DC, REDFT01 x REDFT01, sum += image, emit sum).  Default options."""
    C = pixels.dtype.type
    h, w, d = pixels.shape
    coeffs = _transform(np.ascontiguousarray(pixels), [odct.REDFT10, odct.REDFT10], fast)   # :292-294
    coeffs = (coeffs / C(w * h * 4)).astype(C)                                                # :297-298
    total = np.empty_like(coeffs)
    total[:, :, :] = coeffs[0, 0, :]                                                          # :381-383
    limit = int(index_map.max()) + 1
    if not nframes or nframes > limit // step:
        nframes = (limit + step - 1) // step                                                  # :346-347
    frames = []
    for i in range(nframes):
        lo, hi = i * step, min(i * step + step, limit)
        recon = np.zeros_like(coeffs)                                                         # :429
        sel = (index_map >= lo) & (index_map < hi)
        recon[sel] = coeffs[sel]                                                              # :430-432
        recon[0, 0, :] = 0                                                                    # :445
        image = _transform(recon, [odct.REDFT01, odct.REDFT01], fast)                         # :447
        total = (total + image).astype(C)                                                     # :454
        frames.append(total.copy())
    return frames, coeffs


def motion_block(pels, block, scaled=None, damp=1.0, boost=1.0, bandpass=None, threshold=(0.0, 0.0), quant=0.0,
                 preserve_dc=None, coeff=np.float32, intermediate=np.float64, fast=True, spec=None, ispec=None):
    """motion/motion.c:617-788 for one plane block with linear/dither/coeff-limit/eval off.
    spec / ispec: None | "abs" (spec only) | "shift" | "flat" | "copy" (motion.c:627-637, 755-766).
    pels: staging block [minbuf.d][minbuf.h][minbuf.w], uint8 or float32.  Sizes are (d, h, w).
    Returns (processed block, coefficients coded)."""
    C, I = coeff, intermediate
    scaled = tuple(scaled) if scaled is not None else tuple(block)
    block = tuple(block)
    minbuf = tuple(max(a, b) for a, b in zip(block, scaled))                                  # :491-493
    active = tuple(min(a, b) for a, b in zip(block, scaled))                                  # :494-496
    assert pels.shape == minbuf
    float_pixels = pels.dtype != np.uint8
    sf = I(scaled[0] * scaled[1] * scaled[2]) / I(block[0] * block[1] * block[2])             # :566
    norm = 1 / np.sqrt(I(scaled[0] * scaled[1] * scaled[2] * 8))                              # :567
    quantizer = C(quant * 8 * np.sqrt(I(scaled[0] * scaled[1] * scaled[2])))                  # :570
    thr = (C(threshold[0] * 255 / norm / norm), C(threshold[1] * 255 / norm / norm))          # :571-572
    bsl = tuple(slice(0, v) for v in block)
    asl = tuple(slice(0, v) for v in active)
    ssl = tuple(slice(0, v) for v in scaled)
    coeffs = np.zeros(minbuf, dtype=C)                                                        # :617
    src = pels[bsl].astype(I)
    pel_in = src * 255 if float_pixels else src                                               # :618-624
    cshift = I(127.5) / np.log1p(I(scaled[0] * scaled[1] * scaled[2]) * norm * 255 * 8)       # :568-569 (c[i] and ic[i])
    if ispec == "shift":
        pel_in = np.copysign(np.expm1(np.abs((pel_in - I(127.5)) / cshift)), pel_in - I(127.5)) / norm      # :628
    elif ispec == "flat":
        pel_in = (pel_in - I(127.5)) * 2 / norm / norm                                        # :629
    elif ispec == "copy":
        pel_in = pel_in / norm / norm                                                         # :630
    coeffs[bsl] = pel_in.astype(C)                                                            # :637
    z, y, x = np.meshgrid(*[np.arange(v) for v in active], indexing="ij")
    s2 = np.sqrt(I(2))
    nf = (2 * s2) / (np.where(x > 0, I(1), s2) * np.where(y > 0, I(1), s2) * np.where(z > 0, I(1), s2))
    if not ispec:
        coeffs[bsl] = (odct.dctn_fast(coeffs[bsl], [odct.REDFT10] * 3) if fast else odct.dctn_def(coeffs[bsl], [odct.REDFT10] * 3)).astype(C)   # :641
        a = (coeffs[asl].astype(I) * nf).astype(C)                                            # :644-647
    else:
        a = coeffs[asl].copy()
    dc = a[0, 0, 0]                                                                           # :649
    bb, be = bandpass if bandpass is not None else ((0, 0, 0), active)
    inside = ((z >= bb[0]) & (z < be[0]) & (y >= bb[1]) & (y < be[1]) & (x >= bb[2]) & (x < be[2]))
    if damp != 1:
        a = np.where(inside, a, (a * C(damp)).astype(C))                                      # :683-713
    if boost != 1:
        a = np.where(inside, (a * C(boost)).astype(C), a)                                     # :714-719
    if threshold[1]:
        aa = np.abs(a)
        a = np.where((aa < thr[0]) | (aa > thr[1]), C(0), a)                                  # :721-728
    if preserve_dc:
        dcstop = bool(bb[0] or bb[1] or bb[2])
        if dcstop or boost != 1 or threshold[1]:                                              # :730-738
            if preserve_dc == "dc":
                a[0, 0, 0] = dc
            else:
                a[0, 0, 0] = C(I(a[0, 0, 0]) + (1 - (damp if dcstop else boost)) * I(127.5) / (norm * norm * sf))
    coded = 0
    if quant:
        a = (np.round(a.astype(I) / quantizer) * quantizer).astype(C)                         # :740-744
        coded = int(np.count_nonzero(a))
    if not spec:
        coeffs[asl] = (a.astype(I) / nf).astype(C)                                            # :748-751
        coeffs[ssl] = (odct.dctn_fast(coeffs[ssl], [odct.REDFT01] * 3) if fast else odct.dctn_def(coeffs[ssl], [odct.REDFT01] * 3)).astype(C)   # :753
        pel = coeffs[ssl].astype(I) * sf * norm * norm                                        # :757,767
    else:
        coeffs[asl] = a
        if active != minbuf:                                                                  # (outside the active box the buffer is zero: :617)
            keep = np.zeros(minbuf, dtype=bool); keep[asl] = True
            coeffs[~keep] = 0
        pel = coeffs[ssl].astype(I) * sf * norm                                               # :757
        if spec == "abs":
            pel = (255 / np.log1p(np.abs(I(dc) * sf * norm))) * np.log1p(np.abs(pel))         # :754, :760
        elif spec == "shift":
            pel = cshift * np.copysign(np.log1p(np.abs(pel)), pel) + I(127.5)                 # :761
        elif spec == "flat":
            pel = pel * norm / 2 + I(127.5)                                                   # :762
        else:
            pel = pel * norm                                                                  # :764-765
    out = pels.copy()
    if float_pixels:
        out[ssl] = (pel / 255).astype(np.float32)                                             # :773
    else:
        out[ssl] = np.where(pel > 255, 255, np.where(pel < 0, 0, np.floor(np.abs(pel) + 0.5) * np.sign(pel))).astype(np.uint8)   # :776 lround
    return out, coded, pel


def zoom_basis(scaling_type, num, den, offset, nvectors, sampling_len, coeff=np.float64, intermediate=np.longdouble):
    """zoom/zoom.c:36-68 generate_scaled_basis.  Returns (basis[nvectors][ncomponents-1] in coeff precision, ncomponents)."""
    I = intermediate
    num, den, offset = I(num), I(den), I(offset)
    if sampling_len * num / den < 1:                                                          # :37-40
        num, den = I(1), I(sampling_len)
    ncomp = int(min(sampling_len, np.round(sampling_len * num / den)))                        # :41
    b = np.arange(nvectors, dtype=I)[:, None]
    n = np.arange(1, ncomp, dtype=I)[None, :]
    if scaling_type == "native":                                                              # :50-53
        k, N = b + offset, sampling_len * num / den
    elif scaling_type == "interpolated":                                                      # :54-57
        k, N = (b + offset) * den / num, I(sampling_len)
    else:                                                                                     # :58-61 centered
        k, N = (b + offset) * (sampling_len - 1) * den / (sampling_len * num - den), I(sampling_len)
    pi = I(np.pi) if I is not np.longdouble else np.longdouble("3.14159265358979323846264338327950288")
    return np.cos(pi * (k + I(0.5)) * n / N).astype(coeff), ncomp                             # :63


def zoom_synthesise(pixels, scale=(1, 1), basis="interpolated", pos=(0.0, 0.0), view=(0, 0), xscale=None, yscale=None,
                    intermediate=np.longdouble, fast=True):
    """zoom/zoom.c:263-266 (forward), :268-289 (scale / view), :347-358 (bases), :361-375 (synthesis), one frame."""
    C, I = pixels.dtype.type, intermediate
    H, W, _ = pixels.shape
    coeffs = _transform(np.ascontiguousarray(pixels), [odct.REDFT10, odct.REDFT10], fast)     # :263-265
    xn, xd = (xscale if xscale is not None else scale)
    yn, yd = (yscale if yscale is not None else scale)
    xn, xd, yn, yd = I(xn), I(xd), I(yn), I(yd)
    if W * xn / xd < 1:
        xn, xd = I(1), I(W)                                                                   # :277-280
    if H * yn / yd < 1:
        yn, yd = I(1), I(H)                                                                   # :281-284
    vw = view[0] or int(W * xn / xd)                                                          # :286-289
    vh = view[1] or int(H * yn / yd)
    xb, cw = zoom_basis(basis, xn, xd, pos[0], vw, W, C, I)                                   # :348
    yb, ch = zoom_basis(basis, yn, yd, pos[1], vh, H, C, I)                                   # :356
    out = np.empty((vh, vw, 3), dtype=C)
    for z in range(3):                                                                        # :361-375
        cz = coeffs[:ch, :cw, z].astype(I)
        tmp = cz[:, 0:1] / 2 + cz[:, 1:] @ xb.astype(I).T                                     # tmp[row] per output column i
        s = tmp[0:1, :] / 2 + yb.astype(I) @ tmp[1:, :]
        out[:, :, z] = (s / (W * H)).astype(C)
    return out
