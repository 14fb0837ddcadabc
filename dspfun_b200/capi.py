"""ctypes binding of include/dsp_dct.h.

``load()`` binds the product library (dspfun_b200/libdspdct.so) and raises loudly if it has not been built.
``bind(path)`` attaches the same prototypes to any build of the library sources; the test-suite uses it for the
host SIMT emulation under tests/emu (a test harness, never used by this package).
"""
import ctypes
import os

REDFT01 = 4
REDFT10 = 5

SCALE_LOG, SCALE_LINEAR = 0, 1
SIGN_ABS, SIGN_SHIFT, SIGN_SATURATE, SIGN_RETAIN = 0, 1, 2, 3
RANGE_ONE, RANGE_DC, RANGE_DCS = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
# DSPFUN_B200_LIB points at another build of the same CUDA library (kernel tuning experiments)
LIB_PATH = os.environ.get("DSPFUN_B200_LIB") or os.path.join(_HERE, "libdspdct.so")

# every symbol include/dsp_dct.h declares (tests check the built library exports all of them)
SYMBOLS = [
    "dsp_dct_plan_many", "dsp_dct_plan_many_batched", "dsp_dct_plan_2d", "dsp_dct_execute", "dsp_dct_execute_host",
    "dsp_dct_execute_dev", "dsp_dct_destroy", "dsp_dct_alloc", "dsp_dct_free", "dsp_dct_cleanup",
    "dsp_dct_last_error", "dsp_dct_launch_count", "dsp_dct_fuse_scale", "dsp_dct_set_output_segments", "dsp_dct_fuse_spec",
    "dsp_dct_spec_dc",
    "dsp_dct_fuse_ispec", "dsp_dct_profile", "dsp_dct_num_passes", "dsp_dct_pass_stat_get",
    "dsp_scan_create", "dsp_scan_frame", "dsp_scan_coeffs", "dsp_scan_sum", "dsp_scan_destroy",
    "dsp_motion_create", "dsp_motion_block", "dsp_motion_block_dev", "dsp_motion_destroy", "dsp_block_quant",
    "dsp_block_store_u8",
    "dsp_motion_coeff_stage", "dsp_motion_coeff_stage_flat", "dsp_block_dquant", "dsp_block_dct2d", "dsp_dct_plan_with_ngpus", "dsp_dct_plan_ngpus",
    "dsp_motion_tiled_create", "dsp_motion_tiled_process_dev", "dsp_motion_tiled_destroy",
    "dsp_block_dct2d_debug",
    "dsp_zoom_create", "dsp_zoom_view_size", "dsp_zoom_frame", "dsp_zoom_last_path", "dsp_zoom_destroy",
    "dsp_dct_fuse_pel_load", "dsp_dct_fuse_motion_coeff", "dsp_dct_fuse_pel_store", "dsp_dct_is_emulation",
]


class DspDctError(RuntimeError):
    pass


class SpecParams(ctypes.Structure):
    _fields_ = [("scaletype", ctypes.c_int), ("signtype", ctypes.c_int), ("rangetype", ctypes.c_int),
                ("gain", ctypes.c_double)]


class PassStat(ctypes.Structure):
    _fields_ = [("is_row", ctypes.c_int), ("axis", ctypes.c_int), ("n", ctypes.c_int), ("grid", ctypes.c_int),
                ("block", ctypes.c_int), ("smem_bytes", ctypes.c_size_t), ("launches", ctypes.c_int),
                ("ms_total", ctypes.c_double), ("samples", ctypes.c_double), ("split_panels", ctypes.c_int),
                ("kernel_launches", ctypes.c_int)]


class MotionParams(ctypes.Structure):
    _fields_ = [("block", ctypes.c_int * 3), ("scaled", ctypes.c_int * 3), ("float_pixels", ctypes.c_int),
                ("damp", ctypes.c_double), ("boost", ctypes.c_double), ("bp_begin", ctypes.c_int * 3),
                ("bp_end", ctypes.c_int * 3), ("threshold_min", ctypes.c_double), ("threshold_max", ctypes.c_double),
                ("quant", ctypes.c_double), ("preserve_dc", ctypes.c_int), ("spec", ctypes.c_int), ("ispec", ctypes.c_int)]


class ZoomParams(ctypes.Structure):
    _fields_ = [("basis", ctypes.c_int), ("xscale_num", ctypes.c_double), ("xscale_den", ctypes.c_double),
                ("yscale_num", ctypes.c_double), ("yscale_den", ctypes.c_double), ("vx", ctypes.c_double),
                ("vy", ctypes.c_double), ("vw", ctypes.c_int), ("vh", ctypes.c_int)]


class IspecParams(ctypes.Structure):
    _fields_ = [("scaletype", ctypes.c_int), ("signtype", ctypes.c_int), ("gain", ctypes.c_double),
                ("max", ctypes.c_double * 4), ("preserve_dc", ctypes.c_int), ("dc", ctypes.c_double * 4),
                ("signmap", ctypes.c_void_p)]


def bind(path):
    lib = ctypes.CDLL(path)
    vp, ci, cu, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint, ctypes.c_double
    ip = ctypes.POINTER(ctypes.c_int)
    lib.dsp_dct_plan_many.restype = vp
    lib.dsp_dct_plan_many.argtypes = [ctypes.c_char, ci, ip, ci, vp, ip, ci, ci, vp, ip, ci, ci, ip, cu]
    lib.dsp_dct_plan_many_batched.restype = vp
    lib.dsp_dct_plan_many_batched.argtypes = [ctypes.c_char, ci, ip, ci, vp, ip, ci, ci, vp, ip, ci, ci, ip, cu,
                                              ci, ctypes.c_ssize_t, ctypes.c_ssize_t]
    lib.dsp_dct_plan_2d.restype = vp
    lib.dsp_dct_plan_2d.argtypes = [ctypes.c_char, ci, ci, vp, vp, ci, ci, cu]
    lib.dsp_dct_execute.restype = None
    lib.dsp_dct_execute.argtypes = [vp]
    lib.dsp_dct_execute_host.restype = ci
    lib.dsp_dct_execute_host.argtypes = [vp, vp, vp]
    lib.dsp_dct_execute_dev.restype = ci
    lib.dsp_dct_execute_dev.argtypes = [vp, vp, vp, vp]
    lib.dsp_dct_destroy.restype = None
    lib.dsp_dct_destroy.argtypes = [vp]
    lib.dsp_dct_plan_with_ngpus.restype = None
    lib.dsp_dct_plan_with_ngpus.argtypes = [ci]
    lib.dsp_dct_plan_ngpus.restype = ci
    lib.dsp_dct_plan_ngpus.argtypes = [vp]
    lib.dsp_dct_alloc.restype = vp
    lib.dsp_dct_alloc.argtypes = [ctypes.c_size_t]
    lib.dsp_dct_free.restype = None
    lib.dsp_dct_free.argtypes = [vp]
    lib.dsp_dct_cleanup.restype = None
    lib.dsp_dct_last_error.restype = ctypes.c_char_p
    lib.dsp_dct_launch_count.restype = ctypes.c_ulonglong
    lib.dsp_dct_fuse_scale.restype = ci
    lib.dsp_dct_fuse_scale.argtypes = [vp, cd, cd]
    lib.dsp_block_quant.restype = ci
    lib.dsp_block_quant.argtypes = [ctypes.c_char, vp, ci, ci, ci, ci, ci, ci, cd, vp, vp]
    lib.dsp_block_store_u8.restype = ci
    lib.dsp_block_store_u8.argtypes = [ctypes.c_char, vp, vp, ctypes.c_longlong, cd, vp]
    lib.dsp_motion_tiled_create.restype = vp
    lib.dsp_motion_tiled_create.argtypes = [ci, ci, ci, ci, ci, ci, cd]
    lib.dsp_motion_tiled_process_dev.restype = ci
    lib.dsp_motion_tiled_process_dev.argtypes = [vp, vp, vp, ctypes.POINTER(ctypes.c_ulonglong), vp]
    lib.dsp_motion_tiled_destroy.restype = None
    lib.dsp_motion_tiled_destroy.argtypes = [vp]
    lib.dsp_motion_coeff_stage.restype = ci
    lib.dsp_motion_coeff_stage.argtypes = [ctypes.c_char, ctypes.POINTER(MotionParams), vp, vp, vp]
    lib.dsp_motion_coeff_stage_flat.restype = ci
    lib.dsp_motion_coeff_stage_flat.argtypes = [ctypes.c_char, ctypes.POINTER(MotionParams), vp, ci, ctypes.c_longlong, ci, ctypes.c_longlong, vp, vp]
    lib.dsp_block_dquant.restype = ci
    lib.dsp_block_dquant.argtypes = [vp, ci, ci, ci, ci, ci, ci, cd, vp, vp]
    lib.dsp_block_dct2d.restype = ci
    lib.dsp_block_dct2d.argtypes = [ctypes.c_char, vp, vp, ctypes.c_longlong, ci, ci, ci, ci, cd, vp]
    lib.dsp_block_dct2d_debug.restype = ci
    lib.dsp_block_dct2d_debug.argtypes = [vp, vp, ctypes.c_longlong, ci, ci, ci, ci, cd, vp, vp]
    lib.dsp_dct_set_output_segments.restype = ci
    lib.dsp_dct_set_output_segments.argtypes = [vp, ci, ci, ctypes.POINTER(vp), ctypes.c_longlong, ctypes.c_longlong]
    lib.dsp_dct_fuse_spec.restype = ci
    lib.dsp_dct_fuse_spec.argtypes = [vp, ctypes.POINTER(SpecParams)]
    lib.dsp_dct_spec_dc.restype = ci
    lib.dsp_dct_spec_dc.argtypes = [vp, ctypes.POINTER(cd), ci]
    lib.dsp_dct_fuse_ispec.restype = ci
    lib.dsp_dct_fuse_ispec.argtypes = [vp, ctypes.POINTER(IspecParams)]
    lib.dsp_scan_create.restype = vp
    lib.dsp_scan_create.argtypes = [ctypes.c_char, ci, ci, ci, vp, vp]
    lib.dsp_scan_frame.restype = ci
    lib.dsp_scan_frame.argtypes = [vp, ci, ci, vp]
    lib.dsp_scan_coeffs.restype = ci
    lib.dsp_scan_coeffs.argtypes = [vp, vp]
    lib.dsp_scan_sum.restype = ci
    lib.dsp_scan_sum.argtypes = [vp, vp]
    lib.dsp_scan_destroy.restype = None
    lib.dsp_scan_destroy.argtypes = [vp]
    lib.dsp_motion_create.restype = vp
    lib.dsp_motion_create.argtypes = [ctypes.c_char, ctypes.POINTER(MotionParams)]
    lib.dsp_motion_block.restype = ci
    lib.dsp_motion_block.argtypes = [vp, vp, vp, ctypes.POINTER(ctypes.c_ulonglong)]
    lib.dsp_motion_block_dev.restype = ci
    lib.dsp_motion_block_dev.argtypes = [vp, vp, vp, vp]
    lib.dsp_motion_destroy.restype = None
    lib.dsp_motion_destroy.argtypes = [vp]
    lib.dsp_zoom_create.restype = vp
    lib.dsp_zoom_create.argtypes = [ctypes.c_char, ci, ci, vp]
    lib.dsp_zoom_view_size.restype = ci
    lib.dsp_zoom_view_size.argtypes = [vp, ctypes.POINTER(ZoomParams), ip, ip]
    lib.dsp_zoom_frame.restype = ci
    lib.dsp_zoom_frame.argtypes = [vp, ctypes.POINTER(ZoomParams), vp]
    lib.dsp_zoom_last_path.restype = ci
    lib.dsp_zoom_last_path.argtypes = [vp]
    lib.dsp_zoom_destroy.restype = None
    lib.dsp_zoom_destroy.argtypes = [vp]
    lib.dsp_dct_profile.restype = ci
    lib.dsp_dct_profile.argtypes = [vp, ci]
    lib.dsp_dct_num_passes.restype = ci
    lib.dsp_dct_num_passes.argtypes = [vp]
    lib.dsp_dct_pass_stat_get.restype = ci
    lib.dsp_dct_pass_stat_get.argtypes = [vp, ci, ctypes.POINTER(PassStat)]
    lib.dsp_dct_fuse_pel_load.restype = ci
    lib.dsp_dct_fuse_pel_load.argtypes = [vp, ci]
    lib.dsp_dct_fuse_motion_coeff.restype = ci
    lib.dsp_dct_fuse_motion_coeff.argtypes = [vp, ctypes.POINTER(MotionParams), vp, ci, ctypes.c_longlong]
    lib.dsp_dct_fuse_pel_store.restype = ci
    lib.dsp_dct_fuse_pel_store.argtypes = [vp, ctypes.POINTER(MotionParams)]
    lib.dsp_dct_is_emulation.restype = ci
    lib.dsp_dct_is_emulation.argtypes = []
    return lib


_LIB = None


def load():
    """The product library.  Raises if the CUDA extension has not been built -- there is no fallback."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise DspDctError(
                "dspfun_b200/libdspdct.so is missing: build it with `make -C dspfun_b200/csrc` "
                "(or python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
        lib = bind(LIB_PATH)
        if lib.dsp_dct_is_emulation():
            raise DspDctError("%s is the host emulation of the kernels (a test harness): the product loader only takes "
                              "the CUDA build.  There is no CPU path." % LIB_PATH)
        _LIB = lib
    return _LIB


def last_error(lib):
    return lib.dsp_dct_last_error().decode("utf-8", "replace")
