"""Reader / writer of the .dspraw files understood by tools/wandstub (the raw-file MagickWand stand-in used to
run the reference's unmodified spec / ispec against libdspdct)."""
import numpy as np


def write(path, px, props=None):
    h, w, d = px.shape
    with open(path, "wb") as f:
        f.write(("DSPRAW %d %d %d\n" % (w, h, d)).encode())
        for k, v in (props or {}).items():
            f.write(("PROP %s %s\n" % (k, v)).encode())
        f.write(b"DATA\n")
        f.write(np.ascontiguousarray(px, dtype=np.float64).tobytes())


def read(path):
    with open(path, "rb") as f:
        w, h, d = [int(x) for x in f.readline().split()[1:4]]
        props = {}
        while True:
            line = f.readline().decode().rstrip("\n")
            if line.startswith("DATA"):
                break
            _, k, v = line.split(" ", 2)
            props[k] = v
        px = np.frombuffer(f.read(h * w * d * 8), dtype=np.float64).reshape(h, w, d)
    return px, props


# ---- DSPV: the raw video files of tools/ffstub (the FFmpeg stand-in behind the reference's ffapi.h) -----------------
_FMT = {  # name -> (dtype, [(plane, log2 sub w, log2 sub h) per component], planes)
    "gray": (np.uint8, [(0, 0, 0)]),
    "yuv420p": (np.uint8, [(0, 0, 0), (1, 1, 1), (2, 1, 1)]),
    "yuv444p": (np.uint8, [(0, 0, 0), (1, 0, 0), (2, 0, 0)]),
    "gbrp": (np.uint8, [(2, 0, 0), (0, 0, 0), (1, 0, 0)]),
    "grayf32le": (np.float32, [(0, 0, 0)]),
    "gbrpf32le": (np.float32, [(2, 0, 0), (0, 0, 0), (1, 0, 0)]),
}


def _plane_shapes(fmt, w, h):
    dt, comps = _FMT[fmt]
    shapes = {}
    for plane, sw, sh in comps:
        shapes[plane] = (-(-h >> sh), -(-w >> sw))
    return dt, comps, [shapes[p] for p in sorted(shapes)]


def write_video(path, fmt, w, h, frames, rate=(25, 1)):
    """frames: list of per-frame lists of COMPONENT arrays (component order of the pixel format: Y,U,V / R,G,B)"""
    dt, comps, shapes = _plane_shapes(fmt, w, h)
    with open(path, "wb") as f:
        f.write(("DSPV1 %s %d %d %020d %d %d\n" % (fmt, w, h, len(frames), rate[0], rate[1])).encode())
        for fr in frames:
            planes = [None] * len(shapes)
            for c, (plane, _, _) in enumerate(comps):
                planes[plane] = np.ascontiguousarray(fr[c], dtype=dt)
            for p, shp in zip(planes, shapes):
                assert p.shape == shp, (p.shape, shp)
                f.write(p.tobytes())


def read_video(path):
    """-> (fmt, w, h, frames) with frames[i][c] the component arrays"""
    with open(path, "rb") as f:
        hdr = f.readline().split()
        fmt, w, h, n = hdr[1].decode(), int(hdr[2]), int(hdr[3]), int(hdr[4])
        dt, comps, shapes = _plane_shapes(fmt, w, h)
        frames = []
        for _ in range(n):
            planes = [np.frombuffer(f.read(int(np.prod(s)) * np.dtype(dt).itemsize), dtype=dt).reshape(s) for s in shapes]
            frames.append([planes[plane] for plane, _, _ in comps])
    return fmt, w, h, frames
