"""spec / ispec -- host-side mirror of the reference tools' numeric flow (everything between the image decode and
the image encode), running on the GPU through fused plans.

  spec  : /root/reference/spec/spec.c:59-148   pixels -> REDFT10 x REDFT10 -> DC capture, normalise, gain, range,
          log|linear scale, sign map -> spectrogram (+ the DC property)
  ispec : /root/reference/spec/ispec.c:66-167  spectrogram (+ DC, optional sign map) -> un-sign, un-scale, un-gain,
          un-normalise -> REDFT01 x REDFT01 -> pixels

Option names and presets follow spec/spec.h:71-77 (`-t` templates) and spec/spec.h:112-155 (`-R -T -S -G`).  The
reference's image I/O (MagickWand) is outside the hot path; callers pass and receive [h][w][d] float arrays in
[0,1], the buffers MagickExportImagePixels / MagickConstituteImage exchange (spec/spec.c:60,142).
"""
import math

import numpy as np

from . import capi
from .plan import Plan

# spec/spec.h:71-77: template -> (scaletype, signtype, gaintype, rangetype)
PRESETS = {
    "abs":   ("log",    "abs",      "native", "dc"),
    "shift": ("log",    "shift",    "native", "one"),
    "flat":  ("linear", "shift",    "custom", "one"),
    "sign":  ("linear", "saturate", "custom", "one"),
    "copy":  ("linear", "retain",   "custom", "one"),
}
_SCALE = {"none": capi.SCALE_LOG, "log": capi.SCALE_LOG, "linear": capi.SCALE_LINEAR}
_SIGN = {"none": capi.SIGN_ABS, "abs": capi.SIGN_ABS, "shift": capi.SIGN_SHIFT, "saturate": capi.SIGN_SATURATE,
         "retain": capi.SIGN_RETAIN}
_RANGE = {"none": capi.RANGE_DC, "one": capi.RANGE_ONE, "dc": capi.RANGE_DC, "dcs": capi.RANGE_DCS}


def resolve_params(preset="abs", scale=None, sign=None, gain=None, range_=None):
    """spec_opt_proc (spec/spec.h:112-155): a template, then individual overrides.  `gain` may be 'native',
    'reference' or a number (custom)."""
    sc, sg, gt, rg = PRESETS[preset] if preset else ("none", "none", "none", "none")
    custom = 1.0                                      # spec_opt_defaults.gain (spec/spec.h:81)
    if scale is not None:
        sc = scale
    if sign is not None:
        sg = sign
    if range_ is not None:
        rg = range_
    if gain is not None:
        if isinstance(gain, str):
            gt = gain
        else:
            gt, custom = "custom", float(gain)
    return sc, sg, gt, rg, custom


def resolve_gain(gaintype, w, h, custom):
    """spec/spec.c:81-87 == spec/ispec.c:111-117."""
    if gaintype in ("none", "native"):
        return 127.5 * math.sqrt(w * h * 4)
    if gaintype == "reference":
        return 127.5 * 1024
    return float(custom)


def spec(pixels, preset="abs", lib=None, **overrides):
    """spec/spec.c:59-148 on an [h][w][d] float32/float64 array.  Returns (spectrogram, DC[d] float64)."""
    x = np.ascontiguousarray(pixels)
    h, w, d = x.shape
    prec = "f" if x.dtype == np.float32 else "d"
    sc, sg, gt, rg, custom = resolve_params(preset, **overrides)
    gain = resolve_gain(gt, w, h, custom)
    p = Plan.interleaved_2d(prec, h, w, d, capi.REDFT10, lib=lib)
    p.fuse_spec(_SCALE[sc], _SIGN[sg], _RANGE[rg], gain)
    out = p.execute_host(x.copy())
    dc = p.spec_dc(d)
    p.destroy()
    return out, dc


def ispec(spectrogram, dc=None, preset="abs", preserve_dc=False, signmap=None, lib=None, **overrides):
    """spec/ispec.c:66-167.  `dc` is the decoded DC property (float64[d]); `signmap` an optional uint8 [h][w][d]
    sign image (spec/ispec.c:87-98), whose first pixel carries DC/255 as in the reference."""
    x = np.ascontiguousarray(spectrogram)
    h, w, d = x.shape
    prec = "f" if x.dtype == np.float32 else "d"
    dt = x.dtype.type
    sc, sg, gt, rg, custom = resolve_params(preset, **overrides)
    gain = resolve_gain(gt, w, h, custom)
    if dc is None:
        if signmap is None and (preserve_dc or rg in ("dc", "dcs", "none")):
            raise ValueError("DC not found in header")            # spec/ispec.c:73-77
        dc = np.zeros(d)
    dc = np.array(dc, dtype=np.float64, copy=True)
    if sg in ("none", "abs") and signmap is not None:
        signmap = np.ascontiguousarray(signmap, dtype=np.uint8)
        dc = signmap.reshape(-1)[:d].astype(np.float64) / 255.0   # spec/ispec.c:92-93
    # spec/ispec.c:119-134: range in coeff precision
    if rg == "one":
        mx = np.full(d, dt(gain), dtype=np.float64)
    elif rg in ("none", "dc"):
        mx = np.full(d, dt((dc * gain).max()), dtype=np.float64)
    else:
        mx = (dc * gain).astype(dt).astype(np.float64)
    p = Plan.interleaved_2d(prec, h, w, d, capi.REDFT01, lib=lib)
    p.fuse_ispec(_SCALE[sc], _SIGN[sg], gain, mx, preserve_dc=preserve_dc, dc=dc,
                 signmap=signmap if sg in ("none", "abs") else None)
    out = p.execute_host(x.copy())
    p.destroy()
    return out


def quantize_unorm(x, bits):
    """[0,1] float -> 8/16-bit unsigned, clamp then round half up: what MagickWriteImage does to the float buffer
    handed to MagickConstituteImage (spec/spec.c:141-150)."""
    m = (1 << bits) - 1
    v = np.clip(np.asarray(x, dtype=np.float64), 0.0, 1.0) * m
    return np.floor(v + 0.5).astype(np.uint16 if bits > 8 else np.uint8)


def base16enc(raw: bytes) -> str:
    """spec/spec.h:157-163 (the DC image property: raw double[d], low nibble first, 'A' + nibble)."""
    return "".join(chr((b & 15) + 65) + chr((b >> 4) + 65) for b in raw)


def base16dec(s: str) -> bytes:
    """spec/spec.h:164-168."""
    return bytes(((ord(s[i]) - 65) | ((ord(s[i + 1]) - 65) << 4)) & 0xFF for i in range(0, len(s), 2))
