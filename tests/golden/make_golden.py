"""Copy the FFTW-generated DCT-II/III known-answer vectors that ship with scipy into tests/golden/.

scipy/fftpack/tests/fftw_{single,double,longdouble}_ref.npz were produced by scipy's maintainers by running
FFTW itself (REDFT10 / REDFT01, i.e. exactly the two kinds dspfun plans) on x = linspace(0, n-1, n) for
n in {2,3,4,8,12,15,16,17,32,64,128,256,512,1024}; see scipy/fftpack/tests/test_real_transforms.py:20-28.
They are the only FFTW outputs available in this image (FFTW is not installed), so they pin the oracle.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import numpy as np
import scipy.fftpack.tests as t

src = os.path.dirname(t.__file__)
out = {}
for prec in ("single", "double", "longdouble"):
    z = np.load(os.path.join(src, "fftw_%s_ref.npz" % prec))
    for k in z.files:
        if k.startswith("dct_2_") or k.startswith("dct_3_"):
            # longdouble is stored as float64 pairs? keep whatever dtype scipy stored, as float64 at most
            out["%s_%s" % (prec, k)] = np.asarray(z[k], dtype=np.float64 if prec != "single" else np.float32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fftw_dct_ref.npz"), **out)
print("wrote", len(out), "vectors")
