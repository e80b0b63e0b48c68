"""
oracle/anafast_port.py -- TEST INFRASTRUCTURE (never imported by the product).  PARITY UNPINNED.

Restatement of what `hp.anafast(map)` computes with healpy's defaults -- the C_l measurement that follows BaryonifyShell in the
reference's workflow (/root/reference/examples/04_Baryonify_Density_Shell.ipynb cell 18).  healpy
(third-party, unpinned in pyproject.toml, absent here) implements it in healpix_cxx: `map2alm_iter` with lmax = 3 nside - 1,
mmax = lmax, iter = 3 Jacobi iterations and unit ring weights, then `alm2cl`.  Restated from the published algorithm with a dense
spherical-harmonic matrix (scipy.special.sph_harm_y), i.e. the DEFINITION rather than the ring-FFT + Legendre-recursion route a fast
implementation takes -- small NSIDE only (the matrix is npix x (lmax+1)(lmax+2)/2 complex128).

"Parity unpinned": neither healpy nor any golden vector of it exists in this container or in the reference's tests; the pins in
tests/test_oracle_anafast.py are analytic identities (constant map, pure Y_lm maps, Parseval, quadratic scaling).  This file is the
first step (oracle before kernels) of the C_l row; no product code uses it yet.
"""
import numpy as np

from . import hpo


def alm_index(lmax, l, m):
    """healpy.Alm.getidx: m-major packing of the m >= 0 coefficients."""
    return m * (2 * lmax + 1 - m) // 2 + l


def alm_size(lmax):
    return (lmax + 1) * (lmax + 2) // 2


class DenseSHT(object):
    """Y[p, idx(l, m)] = Y_lm(theta_p, phi_p) at the RING pixel centres of `nside`, for 0 <= m <= l <= lmax."""

    def __init__(self, nside, lmax=None):
        from scipy.special import sph_harm_y
        self.nside = int(nside)
        self.lmax = 3 * self.nside - 1 if lmax is None else int(lmax)
        self.npix = 12 * self.nside * self.nside
        theta, phi = hpo.pix2ang(self.nside, np.arange(self.npix))
        self.Y = np.empty((self.npix, alm_size(self.lmax)), dtype=np.complex128)
        self.l_of = np.empty(alm_size(self.lmax), dtype=np.int64)
        self.m_of = np.empty(alm_size(self.lmax), dtype=np.int64)
        for m in range(self.lmax + 1):
            ls = np.arange(m, self.lmax + 1)
            idx = alm_index(self.lmax, ls, m)
            self.Y[:, idx] = sph_harm_y(ls[None, :], m, theta[:, None], phi[:, None])
            self.l_of[idx], self.m_of[idx] = ls, m
        self.weight = 4 * np.pi / self.npix              # unit ring weights: every pixel carries its area

    def analysis(self, f):
        """One quadrature pass: a_lm = sum_p w f_p conj(Y_lm(p))  (healpix_cxx map2alm, weights = 1)."""
        return self.weight * (self.Y.conj().T @ np.asarray(f, dtype=np.float64))

    def synthesis(self, alm):
        """Real map from the m >= 0 coefficients: f = sum_l a_l0 Y_l0 + 2 Re sum_{m>0} a_lm Y_lm  (alm2map)."""
        fac = np.where(self.m_of == 0, 1.0, 2.0)
        return (self.Y @ (fac * alm)).real

    def map2alm(self, f, iter=3):
        """healpix_cxx map2alm_iter: alm = A f, then `iter` times alm += A (f - S alm)."""
        f = np.asarray(f, dtype=np.float64)
        alm = self.analysis(f)
        for _ in range(int(iter)):
            alm = alm + self.analysis(f - self.synthesis(alm))
        return alm

    def alm2cl(self, alm):
        """C_l = (|a_l0|^2 + 2 sum_{m=1..l} |a_lm|^2) / (2 l + 1)  (healpix_cxx extract_powspec / hp.alm2cl)."""
        p = np.abs(alm) ** 2 * np.where(self.m_of == 0, 1.0, 2.0)
        return np.bincount(self.l_of, weights=p, minlength=self.lmax + 1) / (2 * np.arange(self.lmax + 1) + 1)

    def anafast(self, f, iter=3):
        """hp.anafast(f) with healpy's defaults (lmax = 3 nside - 1, iter = 3, use_weights = False)."""
        return self.alm2cl(self.map2alm(f, iter=iter))
