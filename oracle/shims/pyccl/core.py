import numpy as np
from . import physical_constants as const


class Cosmology(object):
    """Flat wCDM background: enough for R_200c and D_A."""

    def __init__(self, Omega_c=None, Omega_b=None, h=None, n_s=None, sigma8=None, A_s=None,
                 Omega_k=0.0, Omega_g=None, Neff=3.044, m_nu=0.0, w0=-1.0, wa=0.0, T_CMB=const.T_CMB,
                 matter_power_spectrum='halofit', transfer_function='boltzmann_camb', **kw):
        self._params = dict(Omega_c=Omega_c, Omega_b=Omega_b, h=h, n_s=n_s, sigma8=sigma8, w0=w0, wa=wa,
                            Omega_k=Omega_k, Neff=Neff, T_CMB=T_CMB, m_nu=m_nu)
        self.Omega_m = Omega_c + Omega_b
        rho_g = 4 * const.STBOLTZ / const.CLIGHT ** 3 * T_CMB ** 4          # kg / m^3
        rho_crit = const.RHO_CRITICAL * const.SOLAR_MASS / const.MPC_TO_METER ** 3 * h * h
        self.Omega_g = rho_g / rho_crit
        # massless neutrinos: 7/8 * N * (T_nu/T_g)^4 with CCL's T_ncdm ratio
        self.Omega_nu_rel = Neff * 7.0 / 8.0 * const.T_NCDM ** 4 * self.Omega_g
        self.Omega_r = self.Omega_g + self.Omega_nu_rel
        self.Omega_l = 1.0 - self.Omega_m - self.Omega_r - Omega_k
        self.h = h
        self.w0, self.wa = w0, wa
        self._chi_tab = None

    def __getitem__(self, key):
        if key in self._params:
            return self._params[key]
        return getattr(self, key)

    def compute_sigma(self):
        return None

    def compute_distances(self):
        return None

    # ---- background ----
    def _E(self, a):
        a = np.asarray(a, dtype=np.float64)
        de = a ** (-3 * (1 + self.w0 + self.wa)) * np.exp(3 * self.wa * (a - 1))
        return np.sqrt(self.Omega_m * a ** -3 + self.Omega_r * a ** -4 + self.Omega_l * de)

    def _chi(self, a):
        """comoving radial distance [Mpc] by 64-pt Gauss-Legendre on each of 256 log-a panels."""
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        x, w = np.polynomial.legendre.leggauss(48)
        out = np.empty_like(a)
        for i, ai in enumerate(a.ravel()):
            if ai >= 1.0:
                out.ravel()[i] = 0.0
                continue
            edges = np.linspace(ai, 1.0, 17)
            lo, hi = edges[:-1, None], edges[1:, None]
            aa = 0.5 * (hi - lo) * x[None, :] + 0.5 * (hi + lo)
            f = 1.0 / (aa * aa * self._E(aa))
            out.ravel()[i] = np.sum(0.5 * (hi - lo) * w[None, :] * f)
        return out * const.CLIGHT_HMPC / self.h


def h_over_h0(cosmo, a):
    return cosmo._E(a)


def comoving_radial_distance(cosmo, a):
    r = cosmo._chi(a)
    return r if np.ndim(a) else float(r[0])


def angular_diameter_distance(cosmo, a1, a2=None):
    assert a2 is None, "pyccl shim: single-argument form only"
    r = cosmo._chi(a1) * np.atleast_1d(a1)
    return r if np.ndim(a1) else float(r[0])


def rho_x(cosmo, a, species, is_comoving=False):
    rho_c0 = const.RHO_CRITICAL * cosmo.h ** 2
    a = np.asarray(a, dtype=np.float64)
    if species == 'critical':
        rho = rho_c0 * cosmo._E(a) ** 2
        return rho * a ** 3 if is_comoving else rho
    if species == 'matter':
        rho = rho_c0 * cosmo.Omega_m * a ** -3
        return rho * a ** 3 if is_comoving else rho
    raise NotImplementedError(species)


Cosmology.rho_x = lambda self, a, species, is_comoving=False: rho_x(self, a, species, is_comoving)
Cosmology.angular_diameter_distance = lambda self, a: angular_diameter_distance(self, a)
