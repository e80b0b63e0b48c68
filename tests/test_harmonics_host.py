"""
CPU tests of the C_l row's product code that can run without a GPU: the scaled Legendre recursion of csrc/sht_kernels.cu is
compiled for the host as well (bfg_test_sht_lambda_host) and must equal the oracle's recursion (oracle/anafast_rings.py) through
the underflow regime; the Python front end refuses to run without a CUDA device.
"""
import numpy as np
import pytest


def test_device_legendre_recursion_compiled_for_the_host_equals_the_oracle():
    from baryonforge_b200 import _lib
    from oracle.anafast_rings import RingSHT
    L = _lib.lib()
    r = RingSHT(256)                                  # lmax = 767: sin^m(theta) down to 1e-1900 near the poles
    for m in (0, 1, 2, 17, 300, 511, 700, 767):
        want = {l: lam.copy() for l, lam in r._lambdas(m)}
        for ring in (0, 1, 5, 100, 511, 600, 1022):
            out = np.full(r.lmax - m + 1, np.nan)
            rc = L.bfg_test_sht_lambda_host(m, r.lmax, float(r._ln_mm[m]), float(r.x[ring]), float(r.sin2[ring]), out.ctypes.data)
            assert rc == 0
            ref = np.array([want[l][ring] for l in range(m, r.lmax + 1)])
            assert np.all(np.isfinite(out))
            # same operations in the same order: equal to round-off; values below 1e-280 may be flushed to zero
            assert np.max(np.abs(out - ref)) <= 1e-12 * np.max(np.abs(ref)) + 1e-280, (m, ring)
    out = np.zeros(3)
    assert L.bfg_test_sht_lambda_host(5, 7, 0.0, 1.0, 0.0, out.ctypes.data) == 0 and np.all(out == 0)   # exactly on a pole
    assert L.bfg_test_sht_lambda_host(3, 2, 0.0, 0.5, 0.75, out.ctypes.data) == -1                       # lmax < m


def test_shell_harmonics_front_end_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import baryonforge_b200 as b
    from oracle.anafast_rings import RingSHT
    sh = b.harmonics.ShellHarmonics(8)
    assert sh.lmax == 23 and sh.n_alm == 300 and sh.npix == 768
    assert np.allclose(sh.ln_mm, RingSHT(8)._ln_mm, rtol=0, atol=0)
    assert b._lib.lib().bfg_sht_workspace_elems(8, 23) == 24 * 31
    with pytest.raises(b._lib.BFGError):
        sh.anafast(np.zeros(768))
    with pytest.raises(ValueError):
        b.harmonics.anafast(np.zeros(100))
