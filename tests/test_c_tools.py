"""tools/dct_cli.c: C host code calling the FFTW names (shim/fftw3.h) with the argument lists of every reference call
site (image / scan out-of-place / motion embed / draw).  Built here with gcc; linked to the emulation (CPU suite)
or to the product library (GPU suite)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import dct as od

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(kind, tmp_path):
    exe = str(tmp_path / ("dct_cli_" + kind))
    if kind == "emu":
        from tests.emu import emu
        emu.load()
        libdir, lib, extra = os.path.join(ROOT, "tests", "emu"), "dspdct_emu", ["-lstdc++"]
    else:
        libdir, lib, extra = os.path.join(ROOT, "dspfun_b200"), "dspdct", []
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-I" + os.path.join(ROOT, "shim"), "-o", exe,
                    os.path.join(ROOT, "tools", "dct_cli.c"), "-L" + libdir, "-l" + lib, "-Wl,-rpath," + libdir] + extra,
                   check=True)
    return exe


def _run(exe, tmp_path, site, prec, kind, n, d, x, embed=None):
    fi, fo = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    x.tofile(fi)
    n3 = list(n) + [1] * (3 - len(n))
    cmd = [exe, site, prec, kind] + [str(v) for v in n3] + [str(d), fi, fo]
    if embed:
        cmd += [str(v) for v in embed]
    subprocess.run(cmd, check=True)
    return np.fromfile(fo, dtype=x.dtype).reshape(x.shape)


def _check_all(exe, tmp_path):
    rng = np.random.default_rng(21)
    for prec, dt, tol in (("f", np.float32, 1e-5), ("d", np.float64, 1e-12)):
        for kind, ok in (("10", od.REDFT10), ("01", od.REDFT01)):
            x = rng.random((20, 36, 3)).astype(dt)
            ref = od.dctn_fast(x.astype(np.float64), [ok] * 2, axes=(0, 1))
            assert od.rel_l2(_run(exe, tmp_path, "image", prec, kind, (20, 36), 3, x), ref) < tol
            assert od.rel_l2(_run(exe, tmp_path, "scan", prec, kind, (20, 36), 3, x), ref) < tol
            g = rng.random((16, 24, 1)).astype(dt)
            assert od.rel_l2(_run(exe, tmp_path, "draw", prec, kind, (16, 24), 1, g),
                             od.dctn_fast(g.astype(np.float64), [ok] * 2, axes=(0, 1))) < tol
            vol = rng.random((5, 9, 12)).astype(dt)
            out = _run(exe, tmp_path, "motion", prec, kind, (4, 6, 10), 1, vol, embed=(5, 9, 12))
            want = vol.astype(np.float64).copy()
            want[:4, :6, :10] = od.dctn_fast(vol[:4, :6, :10].astype(np.float64), [ok] * 3)
            assert od.rel_l2(out, want) < tol


def test_c_driver_on_emulated_kernels(tmp_path):
    _check_all(_build("emu", tmp_path), tmp_path)


@pytest.mark.gpu
def test_c_driver_on_gpu(tmp_path):
    _check_all(_build("gpu", tmp_path), tmp_path)


def test_shim_rejects_long_double_at_compile_time(tmp_path):
    src = tmp_path / "l.c"
    src.write_text("#include <fftw3.h>\nint main(void){ long double *x = fftwl_alloc_real(4); return x != 0; }\n")
    r = subprocess.run(["gcc", "-I" + os.path.join(ROOT, "shim"), "-c", str(src), "-o", str(tmp_path / "l.o")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "not supported on the GPU" in r.stderr
