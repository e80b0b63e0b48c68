"""
CPU: the HEALPix RING device functions of csrc/bfg_common.cuh -- disc_rings / disc_ring_span (query_disc), pix2vec, get_interpol,
ang2pix_ring, ring2nest / nest2ring: the source every shell kernel inlines -- compiled for the HOST (bfg_test_healpix_host) and
held against the oracle's C restatement of healpy's T_Healpix_Base (oracle/healpix_ring.c) over thousands of discs and 10^5
directions per NSIDE, powers of two or not.  Index sets and pixel numbers bit-exact, floating-point outputs to 1e-15 (the same libm
on both sides; on the GPU the compiler contracts a*b+c into FMAs, which tests/test_gpu_parity.py covers on hardware).
tools/sass_fingerprint.py shows that making these functions host-compilable changed none of the library's kernels.
"""
import numpy as np
import pytest

from oracle import hpo


def host(what, nside, idx=None, a=None, b=None, cap=0, n=None, n_i=0, n_d=0):
    from baryonforge_b200 import _lib
    idx = None if idx is None else np.ascontiguousarray(idx, dtype=np.int64)
    a = None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
    out_i = np.zeros(max(1, n_i), dtype=np.int64)
    out_d = np.zeros(max(1, n_d), dtype=np.float64)
    _lib.check(_lib.lib().bfg_test_healpix_host(what, int(nside), int(n or 0), _lib.ptr(idx), _lib.ptr(a), _lib.ptr(b), int(cap),
                                                out_i.ctypes.data, out_d.ctypes.data))
    return out_i, out_d


def device_source_query_disc(nside, theta, phi, radius):
    cap = 64
    while True:
        out, _ = host(0, nside, a=[theta, phi, radius], cap=cap, n_i=cap + 1)
        if out[cap] <= cap:
            return out[:out[cap]]
        cap = int(out[cap])


@pytest.mark.parametrize("nside", [1, 3, 32, 1000, 4096])
def test_query_disc_index_sets_are_bit_exact(nside):
    rng = np.random.default_rng(70 + nside)
    pixsize = np.sqrt(4 * np.pi / (12 * nside * nside))
    n_disc = 1500 if nside >= 1000 else 600
    theta = np.arccos(rng.uniform(-1, 1, n_disc))
    phi = rng.uniform(0, 2 * np.pi, n_disc)
    radius = pixsize * 10 ** rng.uniform(-0.5, 2.0 if nside > 3 else 0.7, n_disc)
    # the special places: both poles inside / on the edge of the disc, the phi = 0 seam, tiny and empty discs, the whole sphere
    theta[:40] = rng.uniform(0, 3 * pixsize, 40)
    theta[40:80] = np.pi - rng.uniform(0, 3 * pixsize, 40)
    phi[80:120] = rng.choice([0.0, 1e-12, 2 * np.pi - 1e-12], 40)
    radius[120:140] = pixsize * 1e-3
    radius = np.minimum(radius, np.pi)
    radius[140:142] = [np.pi, 3.0]
    n_pix_total = 0
    for t, p, r in zip(theta, phi, radius):
        got = device_source_query_disc(nside, t, p, r)
        want = hpo.query_disc(nside, t, p, r)
        assert np.array_equal(np.sort(got), np.sort(want)), (nside, t, p, r, got.size, want.size)
        assert np.unique(got).size == got.size
        n_pix_total += got.size
    assert n_pix_total > n_disc                                          # the comparison was not vacuous


@pytest.mark.parametrize("nside", [1, 2, 3, 64, 1000, 4096])
def test_pix2vec_interp_weights_ang2pix_match_the_oracle(nside):
    rng = np.random.default_rng(170 + nside)
    npix = 12 * nside * nside
    pix = np.arange(npix) if npix <= 50000 else np.concatenate([rng.integers(0, npix, 100000), np.arange(2000),
                                                                npix - 1 - np.arange(2000)])
    n = pix.size
    _, xyz = host(1, nside, idx=pix, n=n, n_d=3 * n)
    xyz = xyz.reshape(n, 3)
    ox, oy, oz = hpo.pix2vec(nside, pix)
    assert np.abs(xyz - np.stack([ox, oy, oz], axis=1)).max() < 1e-15
    th = np.arccos(rng.uniform(-1, 1, 100000))
    ph = rng.uniform(0, 2 * np.pi, 100000)
    th[:200] = rng.uniform(0, 3.0 / nside, 200)                           # inside the first ring: the pole branches of get_interpol
    th[200:400] = np.pi - rng.uniform(0, 3.0 / nside, 200)
    ph[400:600] = rng.choice([0.0, 1e-13, 2 * np.pi - 1e-13], 200)         # the azimuth seam
    m = th.size
    gi, gw = host(2, nside, a=th, b=ph, n=m, n_i=4 * m, n_d=4 * m)
    wi, ww = hpo.get_interpol(nside, th, ph)
    assert np.array_equal(gi.reshape(m, 4), wi.T)
    assert np.abs(gw.reshape(m, 4) - ww.T).max() < 1e-15
    gp, _ = host(3, nside, a=th, b=ph, n=m, n_i=m)
    assert np.array_equal(gp, hpo.ang2pix(nside, th, ph))
    # pixel centres map back to their pixel
    t_c, p_c = hpo.pix2ang(nside, pix[:20000])
    back, _ = host(3, nside, a=t_c, b=p_c, n=t_c.size, n_i=t_c.size)
    assert np.array_equal(back, pix[:20000])
    if nside & (nside - 1) == 0:                                           # NESTED needs a power of two
        nest, _ = host(4, nside, idx=pix, n=n, n_i=n)
        assert np.array_equal(nest, hpo.ring2nest(nside, pix))
        ring, _ = host(5, nside, idx=nest, n=n, n_i=n)
        assert np.array_equal(ring, pix)
