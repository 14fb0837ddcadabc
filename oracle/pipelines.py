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
