"""scan -- host-side mirror of the reference tool's numeric loop (/root/reference/scan/scan.c:292-298, 352-459):
progressive reconstruction of an image from its DCT coefficients in a given scan order.

The scan order arrives as an index map (index[y][x] = scan step that delivers coefficient (y, x)), the "index"
serialisation of scan/scan_precomputed.c:133-153 -- generating orders (scan/scan_methods.c) is host-side work the
reference does once per run and is outside the hot path.  Two orders are provided for tests and examples.
"""
import numpy as np

from . import capi


def order_horizontal(h, w):
    """raster order: coefficient (y, x) arrives at step y*w + x."""
    return np.arange(h * w, dtype=np.int32).reshape(h, w)


def order_diagonal(h, w):
    """scan/README.md:121-129: index[y][x] = x + y (each anti-diagonal is one interval)."""
    return (np.arange(h, dtype=np.int32)[:, None] + np.arange(w, dtype=np.int32)[None, :]).astype(np.int32)


class Scan:
    """dsp_scan session (include/dsp_dct.h): coefficients, index map and running sum stay on the GPU."""

    def __init__(self, pixels, index_map, lib=None):
        self.lib = lib if lib is not None else capi.load()
        x = np.ascontiguousarray(pixels)
        assert x.ndim == 3 and x.dtype in (np.float32, np.float64)
        self.shape, self.dtype = x.shape, x.dtype
        h, w, d = x.shape
        idx = np.ascontiguousarray(index_map, dtype=np.int32)
        assert idx.shape == (h, w)
        self._h = self.lib.dsp_scan_create(b"f" if x.dtype == np.float32 else b"d", h, w, d, x.ctypes.data, idx.ctypes.data)
        if not self._h:
            raise capi.DspDctError(capi.last_error(self.lib))

    def frame(self, lo, hi, want=True):
        """One output frame: coefficients with lo <= index < hi join the reconstruction (scan.c:421-459)."""
        out = np.empty(self.shape, self.dtype) if want else None
        if self.lib.dsp_scan_frame(self._h, int(lo), int(hi), out.ctypes.data if want else None) != 0:
            raise capi.DspDctError(capi.last_error(self.lib))
        return out

    def coeffs(self):
        out = np.empty(self.shape, self.dtype)
        if self.lib.dsp_scan_coeffs(self._h, out.ctypes.data) != 0:
            raise capi.DspDctError(capi.last_error(self.lib))
        return out

    def destroy(self):
        if getattr(self, "_h", None):
            self.lib.dsp_scan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def scan_frames(pixels, index_map, step=1, nframes=None, lib=None):
    """scan.c main loop with default options (no --offset/--invert/--visualize): yields every output frame."""
    s = Scan(pixels, index_map, lib=lib)
    limit = int(np.max(index_map)) + 1
    if not nframes or nframes > limit // step:
        nframes = (limit + step - 1) // step                       # scan.c:346-347
    frames = [s.frame(i * step, min((i + 1) * step, limit)) for i in range(nframes)]
    s.destroy()
    return frames


def shard_range(nframes, rank, world):
    """Contiguous frame range [f0, f1) of `rank` (the first nframes % world ranks take one frame more)."""
    base, rem = divmod(nframes, world)
    f0 = rank * base + min(rank, rem)
    return f0, f0 + base + (1 if rank < rem else 0)


def scan_frames_sharded(pixels, index_map, step=1, nframes=None, rank=None, world=None, lib=None):
    """The scan main loop split over `world` GPUs, one process per GPU (frames are independent given the
    coefficient plane and the index map; only the running sum is a prefix dependency, scan.c:454).

    The prefix is linear: sum before frame f0 = DC + IDCT(all coefficients with index < f0*step), so a rank builds
    its starting sum with ONE extra masked inverse instead of replaying f0 frames -- no data-path collective.
    Every rank holds the pixels and the index map (a broadcast at start in a real run) and returns
    (f0, frames of its range).  The prefix inverse adds the f0 partial images in one rounding instead of f0, so a
    sharded frame agrees with the sequential one to the coefficient tolerance, not bit for bit.
    """
    if rank is None or world is None:
        import torch.distributed as dist
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    limit = int(np.max(index_map)) + 1
    if not nframes or nframes > limit // step:
        nframes = (limit + step - 1) // step                       # scan.c:346-347
    f0, f1 = shard_range(nframes, rank, world)
    s = Scan(pixels, index_map, lib=lib)
    if f0 > 0:
        s.frame(0, min(f0 * step, limit), want=False)              # the prefix: one inverse, frame not fetched
    frames = [s.frame(i * step, min((i + 1) * step, limit)) for i in range(f0, f1)]
    s.destroy()
    return f0, frames
