import sys; sys.path.insert(0, '.')
import numpy as np
from dspfun_b200 import capi, REDFT10, REDFT01
from tests import cases
lib = capi.load()
for prec in ("f", "d"):
    for shape in ((512, 512, 3), (256, 1024, 1), (4096, 128, 1), (8192, 64, 1), (1080, 64, 1), (64, 1920, 1), (100, 135, 3), (4096, 48, 3)):
        for kind in (REDFT10, REDFT01):
            e = cases.check_interleaved_2d(lib, prec, *shape, kind)
    cases.check_batched_images(lib, prec, 40, 256, 256, 1)
    cases.check_batched_images(lib, prec, 25, 1024, 64, 3)
print("done")
