"""zoom -- host-side mirror of the reference tool's numeric core (/root/reference/zoom/zoom.c:263-266 forward DCT,
:36-68 scaled cosine basis, :361-375 separable synthesis): resample an RGB image in the DCT domain at any scale,
offset and view.  Expression evaluation, animation and FFmpeg output (zoom.c:306-345, 393-415) are outside the hot
path; callers ask for one view at a time."""
import ctypes

import numpy as np

from . import capi

BASIS = {"interpolated": 0, "centered": 1, "native": 2}


class Zoom:
    def __init__(self, pixels, lib=None):
        self.lib = lib if lib is not None else capi.load()
        x = np.ascontiguousarray(pixels)
        assert x.ndim == 3 and x.shape[2] == 3 and x.dtype in (np.float32, np.float64), "zoom works on [h][w][3] RGB"
        self.dtype = x.dtype
        self.h, self.w = x.shape[:2]
        self._h = self.lib.dsp_zoom_create(b"f" if x.dtype == np.float32 else b"d", self.h, self.w, x.ctypes.data)
        if not self._h:
            raise capi.DspDctError(capi.last_error(self.lib))

    def frame(self, scale=1.0, basis="interpolated", pos=(0.0, 0.0), view=(0, 0), xscale=None, yscale=None, pinned=False):
        """One output view [vh][vw][3].  scale / xscale / yscale: a number or a (num, den) pair (zoom -s / -x / -y).
        pinned=True returns a view of a page-locked buffer owned by this object (dsp_dct_alloc, what the tool's
        fftw_alloc_real gives it through the shim): valid until the next pinned frame or destroy(), and copied out of the
        GPU at PCIe speed instead of through pageable memory."""
        def frac(v):
            return (float(v[0]), float(v[1])) if isinstance(v, (tuple, list)) else (float(v), 1.0)
        xs, ys = frac(xscale if xscale is not None else scale), frac(yscale if yscale is not None else scale)
        zp = capi.ZoomParams(BASIS[basis], xs[0], xs[1], ys[0], ys[1], float(pos[0]), float(pos[1]), int(view[0]), int(view[1]))
        vw, vh = ctypes.c_int(0), ctypes.c_int(0)
        self.lib.dsp_zoom_view_size(self._h, ctypes.byref(zp), ctypes.byref(vw), ctypes.byref(vh))
        if pinned:
            nbytes = vh.value * vw.value * 3 * np.dtype(self.dtype).itemsize
            if getattr(self, "_pin_bytes", 0) < nbytes:
                if getattr(self, "_pin", None):
                    self.lib.dsp_dct_free(self._pin)
                self._pin = self.lib.dsp_dct_alloc(nbytes)
                if not self._pin:
                    raise capi.DspDctError(capi.last_error(self.lib))
                self._pin_bytes = nbytes
            buf = (ctypes.c_char * nbytes).from_address(self._pin)
            out = np.frombuffer(buf, dtype=self.dtype).reshape(vh.value, vw.value, 3)
        else:
            out = np.empty((vh.value, vw.value, 3), self.dtype)
        if self.lib.dsp_zoom_frame(self._h, ctypes.byref(zp), out.ctypes.data) != 0:
            raise capi.DspDctError(capi.last_error(self.lib))
        self.last_path = {1: "inverse-dct", 2: "shifted-dct", 3: "dense-tensor-core"}.get(self.lib.dsp_zoom_last_path(self._h), "dense")
        return out

    def destroy(self):
        if getattr(self, "_h", None):
            self.lib.dsp_zoom_destroy(self._h)
            self._h = None
        if getattr(self, "_pin", None):
            self.lib.dsp_dct_free(self._pin)
            self._pin, self._pin_bytes = None, 0

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
