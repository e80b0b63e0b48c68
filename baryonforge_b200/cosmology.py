"""
Per-halo host scalars (R_200c, D_A) -- SURVEY §8a row 11: the two pyccl calls the reference makes per loop
iteration (Runners/HealpixRunner.py:320-321, Map2DRunner.py:491,734, SnapshotRunner.py:226,
Profiles/BaryonCorrection.py:399), vectorised once per process().

If `pyccl` is importable it is used (so numbers are CCL's own); otherwise `Background`, a flat wCDM +
photons + massless-neutrino background written from CCL's published formulas, stands in.  Either way the
device path only ever sees the resulting arrays.
"""
import numpy as np
from scipy import interpolate

try:  # the reference's dependency; optional here
    import pyccl as _ccl
except Exception:  # pragma: no cover - depends on the environment
    _ccl = None

# CCL's constants (ccl_constants.h)
_RHO_CRIT = 2.7753662724570803e11     # h^2 Msun / Mpc^3
_CLIGHT_HMPC = 2997.92458
_STBOLTZ, _CLIGHT, _SOLAR_MASS, _MPC = 5.670367e-8, 299792458.0, 1.988409870698051e30, 3.085677581491367e22
_T_CMB, _T_NCDM, _NEFF = 2.7255, 0.71611, 3.044


def have_pyccl():
    return _ccl is not None


class Background(object):
    """Flat wCDM background used when pyccl is absent."""

    def __init__(self, Omega_m, Omega_b, h, w0=-1.0, **unused):
        self.Omega_m, self.Omega_b, self.h, self.w0 = Omega_m, Omega_b, h, w0
        rho_g = 4 * _STBOLTZ / _CLIGHT ** 3 * _T_CMB ** 4
        rho_c = _RHO_CRIT * h * h * _SOLAR_MASS / _MPC ** 3
        self.Omega_g = rho_g / rho_c
        self.Omega_r = self.Omega_g * (1 + _NEFF * 7.0 / 8.0 * _T_NCDM ** 4)
        self.Omega_l = 1.0 - Omega_m - self.Omega_r

    def E2(self, a):
        a = np.asarray(a, dtype=np.float64)
        ia = 1.0 / a
        ia3 = ia * ia * ia
        de = 1.0 if self.w0 == -1.0 else a ** (-3 * (1 + self.w0))
        return self.Omega_m * ia3 + self.Omega_r * ia3 * ia + self.Omega_l * de

    def rho_crit(self, a):
        return _RHO_CRIT * self.h ** 2 * self.E2(a)

    def rho_matter(self, a):
        return _RHO_CRIT * self.h ** 2 * self.Omega_m * np.asarray(a, dtype=np.float64) ** -3

    def comoving_distance(self, a):
        """int_a^1 da' / (a'^2 E(a')) * c / H0, for any array of scale factors <= 1.  One cumulative Gauss-Legendre pass over
        panels no wider than 0.01 in a (12 nodes each: relative error ~1e-16), so that the 1000-node table behind the D_A
        spline costs ~1 ms instead of the ~25 ms of integrating every node from scratch."""
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        af = a.ravel()
        a_min = float(min(np.min(af), 1.0)) if af.size else 1.0
        grid = a_min + 0.01 * np.arange(int(np.ceil((1.0 - a_min) / 0.01)) + 1)
        edges = np.unique(np.concatenate([af[af <= 1.0], grid[grid < 1.0], [1.0]]))
        lo, hi = edges[:-1, None], edges[1:, None]
        x, w = np.polynomial.legendre.leggauss(12)
        aa = 0.5 * (hi - lo) * x + 0.5 * (hi + lo)
        f = 1.0 / (aa * aa * np.sqrt(self.E2(aa)))
        seg = 0.5 * (hi - lo)[:, 0] * np.sum(w * f, axis=1)
        to_one = np.concatenate([np.cumsum(seg[::-1])[::-1], [0.0]])          # to_one[i] = integral from edges[i] to 1
        out = to_one[np.searchsorted(edges, np.minimum(af, 1.0))]
        return out.reshape(a.shape) * _CLIGHT_HMPC / self.h

    def angular_diameter_distance(self, a):
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        return self.comoving_distance(a) * a


def runner_cosmology(cosmo_dict, with_w0):
    """The cosmology object a runner builds at the top of process() (HealpixRunner.py:280-284 passes w0;
    Map2DRunner.py:462-465 and SnapshotRunner.py:198-201 do not)."""
    if _ccl is not None:
        kw = dict(Omega_c=cosmo_dict['Omega_m'] - cosmo_dict['Omega_b'], Omega_b=cosmo_dict['Omega_b'],
                  h=cosmo_dict['h'], sigma8=cosmo_dict['sigma8'], n_s=cosmo_dict['n_s'], matter_power_spectrum='linear')
        if with_w0:
            kw['w0'] = cosmo_dict['w0']
        return _ccl.Cosmology(**kw)
    return Background(cosmo_dict['Omega_m'], cosmo_dict['Omega_b'], cosmo_dict['h'],
                      cosmo_dict['w0'] if with_w0 else -1.0)


def _rho(cosmo, a, rho_type):
    if isinstance(cosmo, Background):
        return cosmo.rho_crit(a) if rho_type == 'critical' else cosmo.rho_matter(a)
    return _ccl.rho_x(cosmo, a, rho_type) if hasattr(_ccl, 'rho_x') else cosmo.rho_x(a, rho_type)


def rho_matter(cosmo, a, is_comoving=False):
    """cosmo.rho_x(a, species='matter', is_comoving=...) (HealpixRunner.py:580, Map2DRunner.py:886), Msun / Mpc^3."""
    if isinstance(cosmo, Background):
        r = cosmo.rho_matter(a)
        return r * np.asarray(a, dtype=np.float64) ** 3 if is_comoving else r
    if hasattr(cosmo, 'rho_x'):
        return cosmo.rho_x(a, 'matter', is_comoving=is_comoving)
    return _ccl.rho_x(cosmo, a, 'matter', is_comoving=is_comoving)


def radius_of_mass(cosmo, M, a, mass_def=None):
    """mass_def.get_radius(cosmo, M, a), physical Mpc, vectorised over (M, a).  mass_def=None means 200c."""
    M = np.asarray(M)
    a = np.asarray(a, dtype=np.float64)
    Delta, rho_type = 200, 'critical'
    if mass_def is not None:
        Delta, rho_type = mass_def.Delta, mass_def.rho_type
        if not isinstance(Delta, (int, float)):
            # 'vir' / 'fof': defer to the object, one call per distinct scale factor
            out = np.empty(M.shape, dtype=np.float64)
            ab = np.broadcast_to(a, M.shape)
            for au in np.unique(ab):
                sel = ab == au
                out[sel] = mass_def.get_radius(cosmo, M[sel], float(au))
            return out
    return (M / (4.18879020479 * Delta * _rho(cosmo, a, rho_type))) ** (1. / 3.)


def angular_diameter_distance(cosmo, a):
    if isinstance(cosmo, Background):
        return cosmo.angular_diameter_distance(a)
    return _ccl.angular_diameter_distance(cosmo, a)


def D_A_spline_to(cosmo, z_m):
    """CubicSpline of D_A over 1000 nodes in z in [0, z_m + 0.1] (HealpixRunner.py:296-299)."""
    z_t = np.linspace(0, z_m + 0.1, 1000)
    return interpolate.CubicSpline(z_t, angular_diameter_distance(cosmo, 1 / (1 + z_t)))


def D_A_spline(cosmo, z):
    """The CubicSpline object itself (HealpixRunner.py:297-299), for chunked evaluation."""
    z_m = np.max(z)
    assert z_m <= 30, f"We assume max(z) = 30, but your catalog has max(z) = {z_m}"   # HealpixRunner.py:301
    return D_A_spline_to(cosmo, z_m)


RADIUS_SPLINE_NODES = 2048


def radius_factor_spline(cosmo, mass_def, z_m):
    """
    g(u), u = ln(1+z), with mass_def.get_radius(cosmo, M, a) = cbrt(M) * g(u) (physical Mpc): every spherical-overdensity
    definition has R = (M / (4 pi/3 Delta(a) rho(a)))^(1/3).  Tabulated from the cosmology object in use on 2048 nodes
    up to z_m + 0.1 and splined; in u the integrand is close to exp(-u), so the spline error is ~ du^4/384 < 1e-13.
    The device record kernel (bfg_shell_records) evaluates it per halo.
    """
    u = np.linspace(0, np.log(1 + z_m + 0.1), RADIUS_SPLINE_NODES)
    a = np.exp(-u)
    a[0] = 1.0
    g = radius_of_mass(cosmo, np.ones_like(a), a, mass_def)
    return interpolate.CubicSpline(u, g)


def D_A_of_z(cosmo, z):
    """The reference's D_a: CubicSpline over 1000 nodes in z in [0, zmax+0.1] (HealpixRunner.py:297-299)."""
    z = np.asarray(z, dtype=np.float64)
    z_m = np.max(z)
    assert z_m <= 30, f"We assume max(z) = 30, but your catalog has max(z) = {z_m}"   # HealpixRunner.py:301
    z_t = np.linspace(0, z_m + 0.1, 1000)
    D_a = interpolate.CubicSpline(z_t, angular_diameter_distance(cosmo, 1 / (1 + z_t)))
    return D_a(z)
