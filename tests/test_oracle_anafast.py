"""
CPU tests of oracle/anafast_port.py (what `hp.anafast(map)` computes with healpy's defaults; PARITY UNPINNED -- healpy is absent,
so the pins are analytic identities of the spherical-harmonic transform on the HEALPix grid).
"""
import numpy as np
import pytest

from oracle.anafast_port import DenseSHT, alm_index, alm_size


@pytest.fixture(scope="module")
def sht8():
    return DenseSHT(8)


def test_alm_packing_is_healpys():
    lmax = 5
    assert alm_size(lmax) == 21
    assert [alm_index(lmax, l, 0) for l in range(6)] == [0, 1, 2, 3, 4, 5]
    assert alm_index(lmax, 1, 1) == 6 and alm_index(lmax, 5, 5) == 20 and alm_index(lmax, 2, 2) == 11


def test_constant_map_has_only_a_monopole(sht8):
    cl = sht8.anafast(np.full(sht8.npix, 3.0))
    # one quadrature pass gives a_00 = sqrt(4 pi) c exactly (the pixel areas add up to 4 pi) ...
    cl_once = sht8.alm2cl(sht8.analysis(np.full(sht8.npix, 3.0)))
    assert np.isclose(cl_once[0], 4 * np.pi * 9.0, rtol=1e-13)
    # ... the Jacobi iterations then trade that for a smaller residual over ALL l <= 3 nside - 1 (the harmonics are not
    # orthogonal on the pixel grid), which moves C_0 by ~1e-4 and leaves ~1e-8 of it in the even multipoles
    assert np.isclose(cl[0], 4 * np.pi * 9.0, rtol=2e-4)
    assert np.max(cl[1:]) < 1e-7 * cl[0] and np.max(cl[1::2]) < 1e-25 * cl[0]  # odd l vanish by north-south symmetry


@pytest.mark.parametrize("l,m", [(1, 0), (2, 1), (5, 2), (10, 7), (16, 16)])
def test_pure_harmonic_maps(sht8, l, m):
    """f = Re Y_lm  =>  a_lm = 1/2 (1 for m = 0), so C_l = 1 / (2 (2l+1)) (1 / (2l+1) for m = 0); the three Jacobi iterations take
    the plain quadrature's 1e-3 error down by orders of magnitude for l well below the band limit."""
    f = sht8.Y[:, alm_index(sht8.lmax, l, m)].real
    want = (1.0 if m == 0 else 0.5) / (2 * l + 1)
    cl0 = sht8.alm2cl(sht8.analysis(f))
    cl3 = sht8.anafast(f)
    assert abs(cl3[l] - want) <= abs(cl0[l] - want) + 1e-15                    # iterating never hurts here
    assert np.isclose(cl3[l], want, rtol=1e-3 if l <= 10 else 2e-2)
    others = np.delete(cl3, l)
    assert np.max(others) < 1e-3 * want


def test_parseval_scaling_and_linearity(sht8):
    rng = np.random.default_rng(0)
    lmax_in = 8                                                                # band-limited well below 3 nside - 1
    alm = np.zeros(alm_size(sht8.lmax), dtype=np.complex128)
    for l in range(lmax_in + 1):
        for m in range(l + 1):
            alm[alm_index(sht8.lmax, l, m)] = rng.normal() + (0 if m == 0 else 1j * rng.normal())
    f = sht8.synthesis(alm)
    cl_true = sht8.alm2cl(alm)
    cl = sht8.anafast(f)
    assert np.allclose(cl[:lmax_in + 1], cl_true[:lmax_in + 1], rtol=5e-3)
    assert np.max(cl[lmax_in + 1:]) < 1e-4 * np.max(cl_true)
    ells = np.arange(sht8.lmax + 1)
    assert np.isclose(np.sum((2 * ells + 1) * cl), sht8.weight * np.sum(f * f), rtol=2e-3)     # Parseval
    assert np.allclose(sht8.anafast(-2.5 * f), 6.25 * cl, rtol=1e-12)                          # C_l is quadratic in the map
    g = rng.normal(size=sht8.npix)
    assert np.allclose(sht8.map2alm(f + g), sht8.map2alm(f) + sht8.map2alm(g), rtol=1e-10, atol=1e-12)   # map2alm is linear


@pytest.mark.parametrize("nside", [4, 8])
def test_ring_route_equals_the_dense_definition(nside):
    """oracle/anafast_rings.py (ring FFTs + scaled Legendre recursion, the route the device kernels are planned to follow) against
    the dense Y_lm matrix: analysis, synthesis and the iterated anafast agree to round-off."""
    from oracle.anafast_rings import RingSHT
    d, r = DenseSHT(nside), RingSHT(nside)
    f = np.random.default_rng(nside).normal(size=d.npix)
    a_d, a_r = d.analysis(f), r.analysis(f)
    assert np.max(np.abs(a_d - a_r)) < 1e-13 * np.max(np.abs(a_d))
    s_d, s_r = d.synthesis(a_d), r.synthesis(a_d)
    assert np.max(np.abs(s_d - s_r)) < 1e-12 * np.max(np.abs(s_d))
    c_d, c_r = d.anafast(f), r.anafast(f)
    assert np.allclose(c_r, c_d, rtol=1e-11, atol=0)


def test_scaled_legendre_recursion_survives_underflow():
    """m in the hundreds near the poles: sin^m(theta) is far below the fp64 range (1e-1900 for NSIDE = 256), the recursion carries a
    power-of-two exponent instead; values agree with scipy wherever scipy itself is representable and accurate (l <= ~600)."""
    from scipy.special import sph_harm_y
    from oracle.anafast_rings import RingSHT
    r = RingSHT(256)
    theta = np.arccos(r.x)
    for m, picks in ((0, (0, 1, 120)), (17, (17, 67)), (300, (300, 301, 420)), (511, (511, 561)), (600, (600, 601))):
        for l, lam in r._lambdas(m):
            if l > max(picks):
                break
            if l in picks:
                ref = sph_harm_y(l, m, theta, 0.0).real
                ok = np.abs(ref) > 1e-150
                assert ok.sum() > 600
                assert np.max(np.abs(lam[ok] - ref[ok]) / np.abs(ref[ok])) < 1e-9
                assert np.all(np.isfinite(lam)) and np.max(np.abs(lam[~ok])) < 1e-140 if (~ok).any() else True


def test_band_limited_round_trip_at_nside_64():
    from oracle.anafast_rings import RingSHT
    r = RingSHT(64)
    rng = np.random.default_rng(3)
    lmax_in = 40
    alm = np.zeros(alm_size(r.lmax), dtype=np.complex128)
    for l in range(2, lmax_in + 1):
        for m in range(l + 1):
            alm[alm_index(r.lmax, l, m)] = (rng.normal() + (0 if m == 0 else 1j * rng.normal())) / (l + 1.0)
    f = r.synthesis(alm)
    cl_true = r.alm2cl(alm)
    cl = r.anafast(f)
    assert np.allclose(cl[2:lmax_in + 1], cl_true[2:lmax_in + 1], rtol=1e-4)
    assert np.max(cl[lmax_in + 1:]) < 1e-7 * np.max(cl_true)
    ells = np.arange(r.lmax + 1)
    assert np.isclose(np.sum((2 * ells + 1) * cl), r.weight * np.sum(f * f), rtol=1e-4)
