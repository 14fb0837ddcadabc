"""tests/emu -- TEST HARNESS ONLY.

Builds dspfun_b200/csrc with g++ -DDSP_EMULATE: the very same kernel sources (dct_core.cuh, dct_ops.cuh) and
planner (dsp_dct.cu), with every CTA executed as a sequential loop over its threads on the host.  It exists so
the index arithmetic of the kernels (digit reversal, bank-skew padding, Makhoul permutation, tile edges) can be
checked against the oracle in the `-m "not gpu"` suite on machines without a GPU.  It is not a CPU fallback:
nothing under dspfun_b200/ can load it, and the `-m gpu` suite never touches it.
"""
import os
import subprocess

from dspfun_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EMU_PATH = os.path.join(ROOT, "tests", "emu", "libdspdct_emu.so")
_LIB = None


def load():
    global _LIB
    if _LIB is None:
        subprocess.run(["make", "-s", "-j%d" % max(2, min(16, os.cpu_count() or 2)), "-C",
                        os.path.join(ROOT, "dspfun_b200", "csrc"), "emu"], check=True)
        _LIB = capi.bind(EMU_PATH)
    return _LIB
