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
