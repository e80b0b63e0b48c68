"""
oracle/anafast_rings.py -- TEST INFRASTRUCTURE (never imported by the product).  PARITY UNPINNED (see anafast_port.py).

The same transform as oracle/anafast_port.py (what `hp.anafast(map)` computes), but by the route a fast implementation takes and
the device kernels of the C_l row are planned to follow: per-ring FFTs + a Legendre recursion in l for
every (m, ring) with a power-of-two rescaling exponent, so that it also runs where the dense matrix cannot (NSIDE in the
hundreds) and exercises the underflow handling (sin^m theta for m in the hundreds near the poles).  Checked against the dense
definition at small NSIDE in tests/test_oracle_anafast.py.

    F_m(r)  = exp(-i m phi0_r) * FFT_{n_r}[f(r, .)][m mod n_r]                      (ring r: n_r pixels, first one at phi0_r)
    a_lm    = (4 pi / npix) * sum_r lambda_lm(cos theta_r) F_m(r)                    (analysis, unit ring weights)
    f(r, j) = Re sum_k G_k exp(2 pi i j k / n_r),  G_k = sum_{m = k mod n_r} c_m b_m(r) exp(i m phi0_r),  c_0 = 1, c_{m>0} = 2,
    b_m(r)  = sum_l a_lm lambda_lm(cos theta_r)                                      (synthesis)
    lambda_mm = (-1)^m sqrt((2m+1)/(4 pi) prod_{k<=m} (2k-1)/(2k)) sin^m theta,
    lambda_lm = sqrt((4l^2-1)/(l^2-m^2)) [x lambda_{l-1,m} - sqrt(((l-1)^2-m^2)/(4(l-1)^2-1)) lambda_{l-2,m}]
"""
import numpy as np

from . import hpo
from .anafast_port import alm_index, alm_size

SCALE_BITS = 256          # mantissas are kept below 2^SCALE_BITS; the true value is mantissa * 2^exponent, exponent <= 0


class RingSHT(object):
    def __init__(self, nside, lmax=None):
        self.nside = int(nside)
        self.lmax = 3 * self.nside - 1 if lmax is None else int(lmax)
        self.npix = 12 * self.nside * self.nside
        self.weight = 4 * np.pi / self.npix
        nr = 4 * self.nside - 1
        ir = np.arange(1, nr + 1)
        n_north = np.where(ir < self.nside, 4 * ir, 4 * self.nside)
        self.n_ring = np.where(ir > 3 * self.nside, 4 * (4 * self.nside - ir), n_north).astype(np.int64)
        self.start = np.concatenate([[0], np.cumsum(self.n_ring)[:-1]]).astype(np.int64)
        theta, phi = hpo.pix2ang(self.nside, self.start)
        self.x = np.cos(theta)
        self.sin2 = np.sin(theta) ** 2
        self.phi0 = phi
        self.l_of = np.empty(alm_size(self.lmax), dtype=np.int64)
        self.m_of = np.empty(alm_size(self.lmax), dtype=np.int64)
        for m in range(self.lmax + 1):
            ls = np.arange(m, self.lmax + 1)
            self.l_of[alm_index(self.lmax, ls, m)], self.m_of[alm_index(self.lmax, ls, m)] = ls, m
        k = np.arange(1, self.lmax + 1)
        # ln of sqrt((2m+1)/(4 pi) prod_{k<=m} (2k-1)/(2k)) for m = 0 .. lmax
        self._ln_mm = 0.5 * (np.log(2 * np.arange(self.lmax + 1) + 1.0) - np.log(4 * np.pi)
                             + np.concatenate([[0.0], np.cumsum(np.log((2 * k - 1.0) / (2 * k)))]))

    # ---- Legendre recursion for one m over all rings: yields (l, lambda_lm(x_r)) for l = m .. lmax
    def _lambdas(self, m):
        with np.errstate(divide='ignore'):
            ln = self._ln_mm[m] + 0.5 * m * np.log(self.sin2)                 # ln |lambda_mm|, -inf at an exact pole
        ln = np.where(m == 0, self._ln_mm[0], ln)
        log2v = ln / np.log(2.0)
        expo = np.minimum(0, np.floor(log2v / SCALE_BITS) * SCALE_BITS)
        expo = np.where(np.isfinite(expo), expo, -1e9).astype(np.int64)
        with np.errstate(over='ignore', invalid='ignore'):
            cur = np.where(expo > -10 ** 8, np.exp2(log2v - expo), 0.0) * (-1.0 if m % 2 else 1.0)
        prev = np.zeros_like(cur)
        big = 2.0 ** SCALE_BITS
        coef_prev = 0.0
        for l in range(m, self.lmax + 1):
            if l > m:
                a = np.sqrt((4.0 * l * l - 1.0) / (l * l - m * m))
                nxt = a * (self.x * cur - coef_prev * prev)
                prev, cur = cur, nxt
                grow = np.abs(cur) > big
                if grow.any():
                    cur = np.where(grow, cur / big, cur)
                    prev = np.where(grow, prev / big, prev)
                    expo = np.where(grow, expo + SCALE_BITS, expo)
            coef_prev = np.sqrt((l * l - m * m) / (4.0 * l * l - 1.0)) if l >= m else 0.0
            yield l, np.ldexp(cur, np.maximum(expo, -100000).astype(np.int32))

    def ring_ffts(self, f):
        """F_m(r) for m = 0 .. lmax, all rings: [n_rings, lmax + 1] complex."""
        f = np.asarray(f, dtype=np.float64)
        m = np.arange(self.lmax + 1)
        out = np.empty((self.n_ring.size, self.lmax + 1), dtype=np.complex128)
        for r, (s, n) in enumerate(zip(self.start, self.n_ring)):
            F = np.fft.fft(f[s:s + n])
            out[r] = F[m % n] * np.exp(-1j * m * self.phi0[r])
        return out

    def analysis(self, f):
        Fm = self.ring_ffts(f)
        alm = np.zeros(alm_size(self.lmax), dtype=np.complex128)
        for m in range(self.lmax + 1):
            for l, lam in self._lambdas(m):
                alm[alm_index(self.lmax, l, m)] = self.weight * np.dot(lam, Fm[:, m])
        return alm

    def synthesis(self, alm):
        b = np.zeros((self.n_ring.size, self.lmax + 1), dtype=np.complex128)
        for m in range(self.lmax + 1):
            for l, lam in self._lambdas(m):
                b[:, m] += alm[alm_index(self.lmax, l, m)] * lam
        f = np.empty(self.npix, dtype=np.float64)
        m = np.arange(self.lmax + 1)
        fac = np.where(m == 0, 1.0, 2.0)
        for r, (s, n) in enumerate(zip(self.start, self.n_ring)):
            G = np.zeros(n, dtype=np.complex128)
            np.add.at(G, m % n, fac * b[r] * np.exp(1j * m * self.phi0[r]))
            f[s:s + n] = (np.fft.ifft(G) * n).real
        return f

    def map2alm(self, f, iter=3):
        f = np.asarray(f, dtype=np.float64)
        alm = self.analysis(f)
        for _ in range(int(iter)):
            alm = alm + self.analysis(f - self.synthesis(alm))
        return alm

    def alm2cl(self, alm):
        p = np.abs(alm) ** 2 * np.where(self.m_of == 0, 1.0, 2.0)
        return np.bincount(self.l_of, weights=p, minlength=self.lmax + 1) / (2 * np.arange(self.lmax + 1) + 1)

    def anafast(self, f, iter=3):
        return self.alm2cl(self.map2alm(f, iter=iter))
