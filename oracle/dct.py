"""oracle.dct -- TEST INFRASTRUCTURE ONLY.

Two arbiters for the FFTW r2r transforms the reference plans (REDFT10 = unnormalised DCT-II, REDFT01 =
unnormalised DCT-III; call sites /root/reference/spec/spec.c:63, spec/ispec.c:165, zoom/zoom.c:263,
scan/scan.c:292,359, motion/motion.c:535-552, applybasis/draw.c:74):

* ``r2r_many``  -- ctypes binding of oracle/ref_dct.c: the FFTW-manual definitions evaluated in long double,
  O(n^2) per line, with FFTW's advanced-interface howmany/stride/dist/embed addressing.  Small sizes.
* ``dctn_fast`` -- scipy.fft (pocketfft) ``dctn(type=2|3, norm=None)``: an independent O(n log n)
  implementation of the same definitions, used at sizes the definition-based arbiter cannot finish in seconds,
  itself checked against ``r2r_many`` and against the FFTW-generated golden vectors in tests/test_oracle.py.

FFTW itself (the third-party module that holds the reference's arithmetic; un-vendored, no version pinned) is not
installed in this image, so parity against the reference binary is unpinned; see ref_dct.c's header.
"""
import ctypes
import os
import subprocess

import numpy as np

REDFT10 = 10
REDFT01 = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_dct.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.ref_spec_normalization.restype = ctypes.c_longdouble
        _LIB.ref_spec_normalization.argtypes = [ctypes.c_size_t]
        _LIB.ref_quant_u8.restype = ctypes.c_ubyte
        _LIB.ref_quant_u8.argtypes = [ctypes.c_longdouble]
    return _LIB


def _iarr(v):
    return (ctypes.c_int * len(v))(*[int(x) for x in v])


def r2r_many(buf_in, rank, n, howmany, inembed, istride, idist, buf_out, onembed, ostride, odist, kinds):
    """fftw{f,,l}_plan_many_r2r + execute in one call on flat numpy buffers (definition-based, long double)."""
    assert buf_in.dtype == buf_out.dtype and buf_in.flags.c_contiguous and buf_out.flags.c_contiguous
    fn = {np.dtype(np.float32): "ref_r2r_many_f", np.dtype(np.float64): "ref_r2r_many_d",
          np.dtype(np.longdouble): "ref_r2r_many_l"}[buf_in.dtype]
    f = getattr(_lib(), fn)
    f.restype = ctypes.c_int
    rc = f(ctypes.c_int(rank), _iarr(n), ctypes.c_int(howmany),
           buf_in.ctypes.data_as(ctypes.c_void_p), _iarr(inembed) if inembed is not None else None,
           ctypes.c_int(istride), ctypes.c_int(idist),
           buf_out.ctypes.data_as(ctypes.c_void_p), _iarr(onembed) if onembed is not None else None,
           ctypes.c_int(ostride), ctypes.c_int(odist), _iarr(kinds))
    if rc != 0:
        raise ValueError("ref_r2r_many rejected its arguments (rc=%d)" % rc)
    return buf_out


def dctn_def(x, kinds, axes=None):
    """Definition-based N-d transform of a dense array over `axes` (default: all), one kind per axis."""
    x = np.ascontiguousarray(x)
    axes = list(range(x.ndim)) if axes is None else [a % x.ndim for a in axes]
    y = x
    for ax, kind in zip(axes, kinds):
        moved = np.ascontiguousarray(np.moveaxis(y, ax, -1))
        n = moved.shape[-1]
        flat = moved.reshape(-1)
        out = np.empty_like(flat)
        r2r_many(flat, 1, [n], flat.size // n, None, 1, n, out, None, 1, n, [kind])
        y = np.moveaxis(out.reshape(moved.shape), -1, ax)
    return np.ascontiguousarray(y)


def dctn_fast(x, kinds, axes=None, workers=None, compute_dtype=None):
    """pocketfft arbiter: same definitions as FFTW (norm=None).  compute_dtype=np.float64 gives the 'true'
    answer for a float32 input."""
    import scipy.fft as sf
    axes = list(range(x.ndim)) if axes is None else [a % x.ndim for a in axes]
    y = x if compute_dtype is None else x.astype(compute_dtype)
    w = workers if workers is not None else (os.cpu_count() or 1)
    # group consecutive equal kinds so pocketfft can thread over the whole array
    i = 0
    while i < len(axes):
        j = i
        while j + 1 < len(axes) and kinds[j + 1] == kinds[i]:
            j += 1
        y = sf.dctn(y, type=2 if kinds[i] == REDFT10 else 3, axes=axes[i:j + 1], norm=None, workers=w)
        i = j + 1
    return y


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))
