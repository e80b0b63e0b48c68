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


@pytest.mark.parametrize("nside,ring", [(4, 1), (4, 3), (4, 4), (4, 5), (4, 8), (4, 13), (4, 15), (64, 1), (64, 40), (64, 64),
                                        (64, 127), (64, 128), (64, 200), (512, 300), (512, 1024)])
def test_device_ring_sums_compiled_for_the_host_follow_the_oracle_phase_conventions(nside, ring):
    """k_sht_ring_analysis / k_sht_ring_synthesis bodies on the host: cap rings (half-pixel phase), belt rings of both parities,
    the equator, south-cap rings, rings longer than the recurrence's re-seed interval and than one shared-memory chunk."""
    from baryonforge_b200 import _lib
    from oracle import hpo
    from oracle.anafast_rings import RingSHT
    L = _lib.lib()
    r = RingSHT(nside)
    i = ring - 1
    n, start, lmax = int(r.n_ring[i]), int(r.start[i]), r.lmax
    theta, phi = hpo.pix2ang(nside, start + np.arange(n))
    odd = int(round(phi[0] / np.pi * n))                          # phi_0 / pi = odd / n
    assert odd in (0, 1) and np.allclose(phi, (2 * np.arange(n) + odd) * np.pi / n, rtol=0, atol=1e-12)
    rng = np.random.default_rng(ring)
    f = rng.normal(size=n)
    b = rng.normal(size=(lmax + 1, 2))
    b[0, 1] = 0.0
    F = np.zeros((lmax + 1, 2))
    out = np.zeros(n)
    assert L.bfg_test_sht_ring_host(n, odd, lmax, f.ctypes.data, F.ctypes.data, b.ctypes.data, out.ctypes.data) == 0
    m = np.arange(lmax + 1)
    # analysis: the oracle's F_m(r) = exp(-i m phi_0) FFT[m mod n]
    want_F = np.fft.fft(f)[m % n] * np.exp(-1j * m * phi[0])
    got_F = F[:, 0] + 1j * F[:, 1]
    assert np.max(np.abs(got_F - want_F)) < 1e-11 * np.sqrt(n)
    # synthesis: f_j = sum_m c_m Re(b_m exp(i m phi_j))
    bm = (b[:, 0] + 1j * b[:, 1]) * np.where(m == 0, 1.0, 2.0)
    want = (np.exp(1j * np.outer(phi, m)) @ bm).real
    assert np.max(np.abs(out - want)) < 1e-11 * np.sqrt(lmax + 1.0) * 4


@pytest.mark.parametrize("nside,ms", [(4, (0, 1, 2, 5, 11)), (16, (0, 3, 20, 47)), (64, (0, 1, 100, 191))])
def test_device_legendre_stage_compiled_for_the_host_equals_the_oracle(nside, ms):
    """The lane functions of k_sht_leg_analysis / k_sht_leg_synthesis over all ring pairs of one m: north/south pairing through
    lambda_lm(-x) = (-1)^(l+m) lambda_lm(x), the equator on its own, healpy's packing, the 4 pi / npix weight."""
    from baryonforge_b200 import _lib
    from oracle.anafast_port import alm_index
    from oracle.anafast_rings import RingSHT
    L = _lib.lib()
    r = RingSHT(nside)
    lmax, nr = r.lmax, r.n_ring.size
    rng = np.random.default_rng(nside)
    f = rng.normal(size=r.npix)
    Fm = r.ring_ffts(f)                                                   # [ring, m]
    alm_want = r.analysis(f)
    ln_mm = np.ascontiguousarray(r._ln_mm)
    for m in ms:
        F_m = np.ascontiguousarray(np.stack([Fm[:, m].real, Fm[:, m].imag], axis=1))
        a_in = rng.normal(size=(lmax + 1, 2))
        a_in[:m] = 0.0
        a_out, B = np.zeros((lmax + 1, 2)), np.full((nr, 2), np.nan)
        rc = L.bfg_test_sht_legendre_host(nside, lmax, m, ln_mm.ctypes.data, F_m.ctypes.data, a_out.ctypes.data, a_in.ctypes.data,
                                          B.ctypes.data)
        assert rc == 0
        ls = np.arange(m, lmax + 1)
        want = alm_want[alm_index(lmax, ls, m)]
        got = a_out[m:, 0] + 1j * a_out[m:, 1]
        assert np.max(np.abs(got - want)) < 1e-12 * np.max(np.abs(alm_want)), m
        b_want = np.zeros(nr, dtype=np.complex128)
        for l, lam in r._lambdas(m):
            b_want += (a_in[l, 0] + 1j * a_in[l, 1]) * lam
        b_got = B[:, 0] + 1j * B[:, 1]
        assert np.all(np.isfinite(b_got))                                  # every ring of the map was written
        assert np.max(np.abs(b_got - b_want)) < 1e-11 * np.max(np.abs(b_want)), m
