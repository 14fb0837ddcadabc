"""Plan -- Python handle on a dsp_dct_plan (include/dsp_dct.h), mirroring FFTW's plan/execute/destroy cycle
as dspfun's tools use it (/root/reference/spec/spec.c:63-65 and the other call sites listed in dsp_dct.h)."""
import ctypes

import numpy as np

from . import capi


def _iarr(v):
    return None if v is None else (ctypes.c_int * len(v))(*[int(x) for x in v])


class Plan:
    def __init__(self, prec, n, kinds, howmany=1, inembed=None, istride=1, idist=0, onembed=None, ostride=1, odist=0,
                 nbatch=1, ibdist=0, obdist=0, in_ptr=None, out_ptr=None, flags=0, lib=None):
        self.lib = lib if lib is not None else capi.load()
        self.prec = prec
        self.dtype = np.float32 if prec == "f" else np.float64
        self.n = list(n)
        self._h = self.lib.dsp_dct_plan_many_batched(
            prec.encode(), len(n), _iarr(n), int(howmany), in_ptr, _iarr(inembed), int(istride), int(idist),
            out_ptr, _iarr(onembed), int(ostride), int(odist), _iarr(kinds), int(flags),
            int(nbatch), int(ibdist), int(obdist))
        if not self._h:
            raise capi.DspDctError(capi.last_error(self.lib))

    # -- layout helpers ------------------------------------------------------------------------------------
    @classmethod
    def interleaved_2d(cls, prec, h, w, d, kind, nbatch=1, **kw):
        """plan_many_r2r(2,{h,w},d,f,NULL,d,1,f,NULL,d,1,{kind,kind}) -- the image tools' layout (spec/spec.c:63)."""
        if d == 1:
            return cls(prec, [h, w], [kind, kind], 1, None, 1, 0, None, 1, 0, nbatch, h * w, h * w, **kw)
        return cls(prec, [h, w], [kind, kind], d, None, d, 1, None, d, 1, nbatch, h * w * d, h * w * d, **kw)

    @classmethod
    def planar_3d(cls, prec, dims, embed, kind, **kw):
        """motion's rank-3 plan over a sub-box of a padded buffer (motion/motion.c:535-538, 549-552)."""
        return cls(prec, list(dims), [kind] * 3, 1, list(embed), 1, 0, list(embed), 1, 0, **kw)

    # -- fused stages ----------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise capi.DspDctError(capi.last_error(self.lib))

    def fuse_scale(self, load_scale=1.0, store_scale=1.0):
        self._check(self.lib.dsp_dct_fuse_scale(self._h, float(load_scale), float(store_scale)))
        return self

    def set_output_segments(self, seg_rows, bases, outer_stride=0, row_stride=0):
        """Last pass stores run g of `seg_rows` axis positions at device pointer bases[g] (include/dsp_dct.h)."""
        arr = (ctypes.c_void_p * len(bases))(*[int(b) for b in bases])
        self._check(self.lib.dsp_dct_set_output_segments(self._h, len(bases), int(seg_rows), arr, int(outer_stride),
                                                          int(row_stride)))
        return self

    def clear_output_segments(self):
        """Back to the plan's own output addressing (nseg = 0 also restores overridden strides)."""
        self._check(self.lib.dsp_dct_set_output_segments(self._h, 0, 0, None, 0, 0))
        return self

    # motion's stages on a caller-owned plan (include/dsp_dct.h; motion/motion.c:618-624, 617-751, 757-776)
    def fuse_pel_load(self, float_pixels=False):
        self._check(self.lib.dsp_dct_fuse_pel_load(self._h, int(bool(float_pixels))))
        return self

    def fuse_motion_coeff(self, mp, d_counter=None, flat_w=0, flat_base=0):
        self._mp = mp
        self._check(self.lib.dsp_dct_fuse_motion_coeff(self._h, ctypes.byref(mp), d_counter, int(flat_w), int(flat_base)))
        return self

    def fuse_pel_store(self, mp):
        self._check(self.lib.dsp_dct_fuse_pel_store(self._h, ctypes.byref(mp)))
        return self

    def fuse_spec(self, scaletype, signtype, rangetype, gain):
        sp = capi.SpecParams(int(scaletype), int(signtype), int(rangetype), float(gain))
        self._check(self.lib.dsp_dct_fuse_spec(self._h, ctypes.byref(sp)))
        return self

    def spec_dc(self, d):
        out = (ctypes.c_double * d)()
        self._check(self.lib.dsp_dct_spec_dc(self._h, out, d))
        return np.array(list(out), dtype=np.float64)

    def fuse_ispec(self, scaletype, signtype, gain, maxv, preserve_dc=False, dc=None, signmap=None):
        ip = capi.IspecParams()
        ip.scaletype, ip.signtype, ip.gain = int(scaletype), int(signtype), float(gain)
        for z in range(4):
            ip.max[z] = float(maxv[z]) if z < len(maxv) else 0.0
            ip.dc[z] = float(dc[z]) if dc is not None and z < len(dc) else 0.0
        ip.preserve_dc = int(bool(preserve_dc))
        self._signmap = None
        if signmap is not None:
            self._signmap = np.ascontiguousarray(signmap, dtype=np.uint8)
            ip.signmap = self._signmap.ctypes.data
        self._check(self.lib.dsp_dct_fuse_ispec(self._h, ctypes.byref(ip)))
        return self

    # -- execution ---------------------------------------------------------------------------------------------
    def execute_host(self, src, dst=None):
        """Host buffers (numpy, flat or shaped, C-contiguous) staged through the GPU: H2D, passes, D2H."""
        assert src.dtype == self.dtype and src.flags.c_contiguous
        if dst is None:
            dst = src
        assert dst.dtype == self.dtype and dst.flags.c_contiguous
        self._check(self.lib.dsp_dct_execute_host(self._h, src.ctypes.data, dst.ctypes.data))
        return dst

    def execute_dev(self, d_in, d_out=None, stream=None):
        """Device pointers (ints); enqueues on `stream` (cudaStream_t as int, None = default) and returns."""
        self._check(self.lib.dsp_dct_execute_dev(self._h, d_in, d_out if d_out is not None else d_in, stream))

    def profile(self, enable=True):
        self.lib.dsp_dct_profile(self._h, int(bool(enable)))
        return self

    def pass_stats(self):
        """Per-pass device timings accumulated since the last call (see dsp_dct_pass_stat_get)."""
        out = []
        for i in range(self.lib.dsp_dct_num_passes(self._h)):
            st = capi.PassStat()
            self._check(self.lib.dsp_dct_pass_stat_get(self._h, i, ctypes.byref(st)))
            out.append(dict(kernel="row" if st.is_row else ("col-split" if st.split_panels else "col"), axis=st.axis, n=st.n, grid=st.grid, block=st.block,
                            smem_bytes=int(st.smem_bytes), launches=st.launches, ms_total=st.ms_total,
                            samples=st.samples, split_panels=st.split_panels,
                            kernel_launches=st.kernel_launches))
        return out

    def destroy(self):
        if getattr(self, "_h", None):
            self.lib.dsp_dct_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
