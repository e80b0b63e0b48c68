"""
Seeded synthetic inputs (catalogues, maps, tables) shared by tests/, bench.py and the golden-fixture generator.
Distributions follow the only ones the reference's tests define (tests/test_healpix.py:29-55, tests/defaults.py:5)
as laid out in SURVEY.md §8(d).
"""
import numpy as np

COSMO = dict(Omega_m=0.30, Omega_b=0.04, h=0.7, sigma8=0.8, n_s=0.96, w0=-1.0)   # tests/defaults.py:5


def sky_halos(n, seed=42, logM=(12.0, 15.5), z=(0.4, 0.5), mass_function=False):
    """ra, dec [deg], M [Msun], z.  The intended sampler of tests/test_healpix.py:10-24,31-32 (uniform on the sphere)."""
    rng = np.random.default_rng(seed)
    ra = rng.uniform(0, 360, n)
    dec = np.degrees(np.arcsin(rng.uniform(-1, 1, n)))
    if mass_function:   # dn/dlogM ~ M^-0.9 between the same limits
        u = rng.uniform(0, 1, n)
        lo, hi = 10 ** (-0.9 * logM[0]), 10 ** (-0.9 * logM[1])
        M = (lo + u * (hi - lo)) ** (-1 / 0.9)
    else:
        M = 10 ** rng.uniform(logM[0], logM[1], n)
    zz = rng.uniform(z[0], z[1], n)
    return ra, dec, M, zz


def box_halos(n, L, seed=42, logM=(12.0, 15.5), ndim=3):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0, L, (ndim, n))
    M = 10 ** rng.uniform(logM[0], logM[1], n)
    return pos, M


def table_axes(nz=10, nM=10, nr=500, z_min=0.01, z_max=1.0, M_min=1e12, M_max=10 ** 15.5, r_min=1e-3, r_max=3e2,
               z_linear=False):
    z = np.linspace(z_min, z_max, nz) if z_linear else np.geomspace(z_min, z_max, nz)
    M = np.geomspace(M_min, M_max, nM)
    r = np.geomspace(r_min, r_max, nr)
    return np.log(1 + z), np.log(M), np.log(r)


def _R_of_M(M):
    return (M / 1e14) ** (1. / 3.)   # Mpc, order of R200c


def displacement_values(axes, inject_nan=False):
    """Smooth, sign-changing displacement d(z, M, r) [comoving Mpc] that -> 0 at large r."""
    lnz, lnM, lnr = axes
    zp1 = np.exp(lnz)[:, None, None]
    R = _R_of_M(np.exp(lnM))[None, :, None]
    x = np.exp(lnr)[None, None, :] / R
    d = -0.08 * R * x * (1.0 - x / 2.5) * np.exp(-x / 1.5) / zp1
    if inject_nan:
        d = d.copy()
        d[1, 2, 40:60] = np.nan
        d[3, 5, 200] = np.inf
    return d


def profile_values(axes, with_zeros=True):
    """Positive projected/real profile with a hard cut (exact zeros -> -inf in log space) at r > 8 R."""
    lnz, lnM, lnr = axes
    zp1 = np.exp(lnz)[:, None, None]
    M = np.exp(lnM)[None, :, None]
    R = _R_of_M(M)
    x = np.exp(lnr)[None, None, :] / R
    p = 1e-6 * (M / 1e14) ** (5. / 3.) * zp1 ** 2 / (np.sqrt(x) * (1 + x) ** 3.5)
    if with_zeros:
        p = np.where(x < 8.0, p, 0.0)
    return p


def shell_map(nside, seed=7, lo=0.0, hi=10.0):
    return np.random.default_rng(seed).uniform(lo, hi, 12 * nside * nside)   # tests/test_healpix.py:49,53
