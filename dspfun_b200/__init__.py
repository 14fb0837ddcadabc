"""dspfun_b200 -- B200-native DCT-II/III hot path behind dspfun's FFTW call sites.

The product is the C-ABI CUDA library ``dspfun_b200/libdspdct.so`` (include/dsp_dct.h) built from the
hand-written sm_100a kernels in ``dspfun_b200/csrc``.  The Python modules here are the host-side mirror of the
reference tools' numeric flow (spec / ispec / scan / motion / zoom) used by the tests and the benchmark; they
only ever call the CUDA library.  There is no CPU path: without the built extension or without a GPU every
entry point raises.
"""
from .capi import REDFT01, REDFT10, DspDctError, load  # noqa: F401
from .plan import Plan  # noqa: F401
