"""motion -- host-side mirror of the per-block numeric body of the reference tool
(/root/reference/motion/motion.c:525-573 plans and constants, :617-788 block loop): 3-D DCT of a block of pels,
coefficient-space filters, 3-D inverse DCT at a possibly different size (spectral zero-pad / crop), 8-bit store.

Frame decoding / encoding and the scatter of frames into per-block staging buffers (motion.c:591-612, 791-811) are
FFmpeg I/O outside the hot path: callers hand over the staging block [minbuf.d][minbuf.h][minbuf.w] directly.
Sizes are (d, h, w) triples, like the reference's coords struct.
"""
import ctypes

import numpy as np

from . import capi

PRESERVE_DC = {None: 0, "none": 0, "dc": 1, "grey": 2}


class Motion:
    def __init__(self, block, scaled=None, float_pixels=False, damp=1.0, boost=1.0, bandpass=None, threshold=(0.0, 0.0),
                 quant=0.0, preserve_dc=None, prec="f", lib=None):
        self.lib = lib if lib is not None else capi.load()
        self.block = tuple(int(v) for v in block)
        self.scaled = tuple(int(v) for v in (scaled if scaled is not None else block))
        self.minbuf = tuple(max(a, b) for a, b in zip(self.block, self.scaled))
        active = tuple(min(a, b) for a, b in zip(self.block, self.scaled))
        bp = bandpass if bandpass is not None else ((0, 0, 0), active)
        mp = capi.MotionParams()
        for i in range(3):
            mp.block[i], mp.scaled[i] = self.block[i], self.scaled[i]
            mp.bp_begin[i], mp.bp_end[i] = int(bp[0][i]), int(bp[1][i])
        mp.float_pixels = int(bool(float_pixels))
        mp.damp, mp.boost = float(damp), float(boost)
        mp.threshold_min, mp.threshold_max = float(threshold[0]), float(threshold[1])
        mp.quant = float(quant)
        mp.preserve_dc = PRESERVE_DC[preserve_dc]
        self.float_pixels = bool(float_pixels)
        self.coeffs_coded = 0
        self._h = self.lib.dsp_motion_create(prec.encode(), ctypes.byref(mp))
        if not self._h:
            raise capi.DspDctError(capi.last_error(self.lib))

    def process(self, pels):
        """One staging block [minbuf] (uint8, or float32 in [0,1]) -> processed block, same layout; pels outside the
        scaled box are returned unchanged, as in the reference's in-place staging buffers."""
        x = np.ascontiguousarray(pels, dtype=np.float32 if self.float_pixels else np.uint8)
        assert x.shape == self.minbuf, (x.shape, self.minbuf)
        out = np.empty_like(x)
        cnt = ctypes.c_ulonglong(0)
        if self.lib.dsp_motion_block(self._h, x.ctypes.data, out.ctypes.data, ctypes.byref(cnt)) != 0:
            raise capi.DspDctError(capi.last_error(self.lib))
        self.coeffs_coded += cnt.value
        return out

    def process_dev(self, d_in, d_out, stream=None):
        if self.lib.dsp_motion_block_dev(self._h, d_in, d_out, stream) != 0:
            raise capi.DspDctError(capi.last_error(self.lib))

    def destroy(self):
        if getattr(self, "_h", None):
            self.lib.dsp_motion_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
