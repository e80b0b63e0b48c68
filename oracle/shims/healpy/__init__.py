"""
Stand-in for the third-party `healpy` package (absent from this image, no network) --
TEST INFRASTRUCTURE.  Exposes exactly the functions the reference calls
(/root/reference/BaryonForge/Runners/HealpixRunner.py:327-361,426; utils/io.py:346-353;
utils/Pixel.py nside2resol) on top of oracle/healpix_ring.c.  The pure-numpy wrappers
(ang2vec, vec2ang, lonlat conversion) restate healpy/rotator.py + pixelfunc.py.
"""
import numpy as np
from oracle import hpo as _hpo


def nside2npix(nside):
    return 12 * int(nside) * int(nside)


def npix2nside(npix):
    nside = int(round(np.sqrt(npix / 12.0)))
    if 12 * nside * nside != npix:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")
    return nside


def nside2pixarea(nside, degrees=False):
    pixarea = 4 * np.pi / nside2npix(nside)
    if degrees:
        pixarea = np.rad2deg(np.rad2deg(pixarea))
    return pixarea


def nside2resol(nside, arcmin=False):
    resol = np.sqrt(nside2pixarea(nside))
    if arcmin:
        resol = np.rad2deg(resol) * 60
    return resol


def _lonlat2thetaphi(lon, lat):
    return np.pi / 2.0 - np.radians(lat), np.radians(lon)


def _thetaphi2lonlat(theta, phi):
    return np.degrees(phi), 90.0 - np.degrees(theta)


def ang2vec(theta, phi, lonlat=False):
    if lonlat:
        theta, phi = _lonlat2thetaphi(theta, phi)
    sintheta = np.sin(theta)
    return np.array([sintheta * np.cos(phi), sintheta * np.sin(phi), np.cos(theta)]).T


def vec2ang(vectors, lonlat=False):
    vectors = np.asarray(vectors).reshape(-1, 3)
    dnorm = np.sqrt(np.sum(np.square(vectors), axis=1))
    theta = np.arccos(vectors[:, 2] / dnorm)
    phi = np.arctan2(vectors[:, 1], vectors[:, 0])
    phi[phi < 0] += 2 * np.pi
    if lonlat:
        return _thetaphi2lonlat(theta, phi)
    return theta, phi


def pix2vec(nside, ipix, nest=False):
    assert not nest, "shim: RING only (the reference never uses NEST)"
    return _hpo.pix2vec(nside, ipix)


def pix2ang(nside, ipix, nest=False, lonlat=False):
    assert not nest
    t, p = _hpo.pix2ang(nside, ipix)
    if lonlat:
        return _thetaphi2lonlat(t, p)
    return t, p


def ang2pix(nside, theta, phi, nest=False, lonlat=False):
    assert not nest
    if lonlat:
        theta, phi = _lonlat2thetaphi(theta, phi)
    return _hpo.ang2pix(nside, theta, phi)


def query_disc(nside, vec, radius, inclusive=False, fact=4, nest=False, buff=None):
    assert not inclusive and not nest, "shim: only the reference's mode (inclusive=False, RING)"
    theta, phi = _hpo.vec2pointing(np.asarray(vec, dtype=np.float64))
    return _hpo.query_disc(nside, theta, phi, radius)


def get_interp_weights(nside, theta, phi=None, nest=False, lonlat=False):
    assert not nest
    if phi is None:
        theta, phi = pix2ang(nside, theta)
    elif lonlat:
        theta, phi = _lonlat2thetaphi(theta, phi)
    scalar = np.ndim(theta) == 0 and np.ndim(phi) == 0
    theta, phi = np.broadcast_arrays(np.asarray(theta, dtype=np.float64), np.asarray(phi, dtype=np.float64))
    pix, wgt = _hpo.get_interpol(nside, theta.ravel(), phi.ravel())
    if scalar:
        return pix[:, 0], wgt[:, 0]
    return pix, wgt


def read_map(*a, **k):
    raise NotImplementedError("healpy shim: FITS I/O is outside the hot path")
