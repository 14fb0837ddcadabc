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
SPEC = {None: 0, "none": 0, "abs": 1, "shift": 2, "flat": 3, "copy": 4}          # motion --spec / --ispec TYPE


class Motion:
    def __init__(self, block, scaled=None, float_pixels=False, damp=1.0, boost=1.0, bandpass=None, threshold=(0.0, 0.0),
                 quant=0.0, preserve_dc=None, spec=None, ispec=None, prec="f", lib=None, coeff_limit=0, expr=None, dither=False,
                 linear=False):
        if coeff_limit or expr or dither or linear:
            # sequential / host-side stages of the reference (repeated qsort, FFmpeg's expression VM, Floyd-Steinberg error
            # diffusion, libavutil transfer curves): not part of the GPU path, refused rather than silently ignored
            raise capi.DspDctError("motion: --coeff-limit, --eval, --dither and --linear are not supported on the GPU path")
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
        mp.spec, mp.ispec = SPEC[spec], SPEC[ispec]
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


class MotionTiled:
    """Every block of a plane volume in one go: `motion -b WxHxD` with block == scaled (no resampling), the reference
    README's 8x8x8 + --quant example (motion/README.md:75-77, block loop motion/motion.c:591-615).

    The reference walks the blocks one by one and runs an 8x8x8 FFTW plan on each; here the whole [D][H][W] volume is
    transformed block-wise by three plans, one per axis, each a single launch over the volume:
        w: rank-1 length bw, howmany = (W/bw) H D contiguous segments            (dist = bw)
        h: rank-1 length bh, stride W, W adjacent columns, batch over (H/bh) D bands (band stride bh W)
        d: rank-1 length bd, stride H W, H W adjacent columns, batch over D/bd slabs
    The blocks are independent, so a multi-GPU run shards the volume along d in whole blocks with no collective
    (SURVEY 8e).  The per-coefficient stages between and after the transforms (normalise, quantise, de-normalise;
    clamp / round / 8-bit store: motion.c:644-647, 740-751, 757-776) are one pass each (dsp_block_quant,
    dsp_block_store_u8), evaluated in double like the fused per-block path; filters other than --quant stay with
    `Motion` (block at a time).
    Square spatial blocks of 8, 16, 32 or 64 (the README's 8x8x8 among them) skip the w and h plans: dsp_block_dct2d
    contracts both spatial axes of every block in one pass over the volume on the tensor cores (tcgen05 MMAs, TF32
    operands split hi + lo for float accuracy; csrc/kern_block_mm.cu), and only the d axis remains a plan.
    `gemm=False` forces the three-plan path.
    Works on torch tensors: CUDA with the product library, CPU with the emulation library (tests).
    """

    def __init__(self, dims, block, quant=0.0, float_pixels=False, lib=None, gemm=True):
        import torch
        self.torch = torch
        self.lib = lib if lib is not None else capi.load()
        self.dims = D, H, W = tuple(int(v) for v in dims)
        self.block = bd, bh, bw = tuple(int(v) for v in block)
        if D % bd or H % bh or W % bw:
            raise ValueError("the volume must be a whole number of blocks (the reference pads the last block with zeros)")
        from .plan import Plan
        self.quant, self.float_pixels = float(quant), bool(float_pixels)
        self.coeffs_coded = 0

        def plans(kind):
            return (Plan("f", [bw], [kind], (W // bw) * H * D, None, 1, bw, None, 1, bw, lib=self.lib),
                    Plan("f", [bh], [kind], W, None, W, 1, None, W, 1, (H // bh) * D, bh * W, bh * W, lib=self.lib),
                    Plan("f", [bd], [kind], H * W, None, H * W, 1, None, H * W, 1, D // bd, bd * H * W, bd * H * W, lib=self.lib))
        self.gemm = bool(gemm) and bh == bw and bw in (8, 16, 32, 64)
        self.dquant = self.gemm and bd in (2, 4, 8, 16)       # d forward + coefficient stage + d inverse in one pass (dsp_block_dquant)
        if self.gemm:
            def plans(kind):                                  # the d axis only (none for depth-1 blocks or with dsp_block_dquant)
                if bd == 1 or self.dquant:
                    return ()
                return (Plan("f", [bd], [kind], H * W, None, H * W, 1, None, H * W, 1, D // bd, bd * H * W, bd * H * W, lib=self.lib),)
        self.fwd = plans(capi.REDFT10)
        self.inv = plans(capi.REDFT01)[::-1]

    def _spatial(self, c, kind, stream):
        D, H, W = self.dims
        # depth-1 blocks have no d plan: FFTW's REDFT10 of length 1 doubles its sample (REDFT01 of length 1 copies it)
        scale = 2.0 if (self.block[0] == 1 and kind == capi.REDFT10) else 1.0
        if self.lib.dsp_block_dct2d(b"f", c.data_ptr(), c.data_ptr(), D, H, W, self.block[2], kind, scale, stream) != 0:
            raise capi.DspDctError(capi.last_error(self.lib))

    def process(self, pels):
        """pels: [D][H][W] torch tensor, uint8 (or float32 in [0,1] with float_pixels).  Returns the processed volume."""
        t = self.torch
        assert tuple(pels.shape) == self.dims and pels.is_contiguous()
        D, H, W = self.dims
        bd, bh, bw = self.block
        stream = t.cuda.current_stream().cuda_stream if pels.is_cuda else None
        if self.float_pixels:
            c = (pels.to(t.float64) * 255.0).to(t.float32).contiguous()                         # motion.c:618-637
        else:
            c = pels.to(t.float32).contiguous()                                                 # (8-bit values are exact)
        if self.gemm:
            self._spatial(c, capi.REDFT10, stream)
        for p in self.fwd:
            p.execute_dev(c.data_ptr(), c.data_ptr(), stream)                                   # :641 for every block
        # normalise, quantise + count, de-normalise: one pass (dsp_block_quant, :644-647, :740-751)
        q = float(np.float32(self.quant * 8.0 * np.sqrt(np.float64(bd * bh * bw)))) if self.quant else 0.0   # :570
        cnt = t.zeros(1, dtype=t.int64, device=c.device)
        if self.dquant:
            rc = self.lib.dsp_block_dquant(c.data_ptr(), D, H, W, bd, bh, bw, q, cnt.data_ptr(), stream)
        else:
            rc = self.lib.dsp_block_quant(b"f", c.data_ptr(), D, H, W, bd, bh, bw, q, cnt.data_ptr(), stream)
        if rc != 0:
            raise capi.DspDctError(capi.last_error(self.lib))
        for p in self.inv:
            p.execute_dev(c.data_ptr(), c.data_ptr(), stream)                                   # :753
        if self.gemm:
            self._spatial(c, capi.REDFT01, stream)
        scale = float((1.0 / np.sqrt(np.float64(bd * bh * bw * 8))) ** 2)                       # :757,767 (sf = 1)
        if self.quant:
            self.coeffs_coded += int(cnt.item())
        if self.float_pixels:
            return (c.to(t.float64) * scale / 255.0).to(t.float32)                              # :773
        out = t.empty(self.dims, dtype=t.uint8, device=c.device)
        if self.lib.dsp_block_store_u8(b"f", c.data_ptr(), out.data_ptr(), D * H * W, scale, stream) != 0:   # :776
            raise capi.DspDctError(capi.last_error(self.lib))
        return out

    def destroy(self):
        for p in tuple(getattr(self, "fwd", ())) + tuple(getattr(self, "inv", ())):
            p.destroy()
        self.fwd = self.inv = ()
