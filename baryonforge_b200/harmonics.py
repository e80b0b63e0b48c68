"""
C_l of a HEALPix shell on the device -- `hp.anafast(map)`, the measurement that follows BaryonifyShell.process() in the
reference's workflow (/root/reference/examples/04_Baryonify_Density_Shell.ipynb cell 18).

The algorithm is the ring route of oracle/anafast_rings.py (per-ring DFTs + a rescaled Legendre recursion in l), with healpy's
defaults: lmax = 3 nside - 1, iter = 3, unit ring weights.  First run on a B200 in round 2: tests/test_gpu_harmonics.py (the
dense definition at NSIDE <= 8, the ring-route oracle up to NSIDE = 128, underflow at NSIDE = 512).  PARITY UNPINNED at the
healpy boundary (healpy is absent here and the reference's tests hold no C_l vector).  No CPU fallback.
"""
import numpy as np

from . import _lib

__all__ = ['ShellHarmonics', 'anafast']


def _torch():
    import torch
    return torch


class ShellHarmonics(object):
    """map2alm / alm2map / anafast for RING maps of one NSIDE; a_lm in healpy's packing (m-major, m >= 0)."""

    def __init__(self, nside, lmax=None, device=None):
        self.nside = int(nside)
        self.lmax = 3 * self.nside - 1 if lmax is None else int(lmax)
        self.npix = 12 * self.nside * self.nside
        self.device = device
        k = np.arange(1, self.lmax + 1)
        self.ln_mm = 0.5 * (np.log(2 * np.arange(self.lmax + 1) + 1.0) - np.log(4 * np.pi)
                            + np.concatenate([[0.0], np.cumsum(np.log((2 * k - 1.0) / (2 * k)))]))
        self.n_alm = (self.lmax + 1) * (self.lmax + 2) // 2
        self._buf = None

    def _dev(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.BFGError("baryonforge_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        return torch.device('cuda', torch.cuda.current_device() if self.device is None else int(self.device))

    def _buffers(self, dev):
        torch = _torch()
        if self._buf is None or self._buf[0] != dev:
            n_work = int(_lib.lib().bfg_sht_workspace_elems(self.nside, self.lmax))
            self._buf = (dev, torch.from_numpy(self.ln_mm).to(dev), torch.empty(2 * n_work, dtype=torch.float64, device=dev))
        return self._buf[1], self._buf[2]

    def _to_map(self, m, dev):
        torch = _torch()
        d = m.to(device=dev, dtype=torch.float64) if torch.is_tensor(m) else \
            torch.from_numpy(np.ascontiguousarray(m, dtype=np.float64)).to(dev)
        if d.numel() != self.npix:
            raise ValueError("Wrong pixel number (it is not 12*nside**2)")
        return d.contiguous()

    def map2alm_on_device(self, m, iter=3):
        """healpix_cxx map2alm_iter: alm = A f, then `iter` times alm += A (f - S alm).  Returns a float64 device tensor
        [n_alm, 2] (re, im)."""
        torch = _torch()
        dev = self._dev()
        L = _lib.lib()
        with torch.cuda.device(dev):
            d_ln, d_work = self._buffers(dev)
            d_map = self._to_map(m, dev)
            d_alm = torch.zeros((self.n_alm, 2), dtype=torch.float64, device=dev)
            st = _lib.current_stream()
            _lib.check(L.bfg_sht_map2alm_pass(self.nside, self.lmax, d_map.data_ptr(), d_ln.data_ptr(), d_work.data_ptr(),
                                              d_alm.data_ptr(), st))
            d_syn = torch.empty_like(d_map) if iter > 0 else None
            for _ in range(int(iter)):
                _lib.check(L.bfg_sht_alm2map(self.nside, self.lmax, d_alm.data_ptr(), d_ln.data_ptr(), d_work.data_ptr(),
                                             d_syn.data_ptr(), st))
                torch.sub(d_map, d_syn, out=d_syn)                                  # residual f - S alm
                _lib.check(L.bfg_sht_map2alm_pass(self.nside, self.lmax, d_syn.data_ptr(), d_ln.data_ptr(), d_work.data_ptr(),
                                                  d_alm.data_ptr(), st))
        return d_alm

    def map2alm(self, m, iter=3):
        a = self.map2alm_on_device(m, iter).cpu().numpy()
        return a[:, 0] + 1j * a[:, 1]

    def alm2map(self, alm):
        torch = _torch()
        dev = self._dev()
        alm = np.ascontiguousarray(alm, dtype=np.complex128)
        if alm.size != self.n_alm:
            raise ValueError("alm has the wrong size for this lmax")
        with torch.cuda.device(dev):
            d_ln, d_work = self._buffers(dev)
            d_alm = torch.from_numpy(alm.view(np.float64).reshape(-1, 2)).to(dev)
            d_map = torch.empty(self.npix, dtype=torch.float64, device=dev)
            _lib.check(_lib.lib().bfg_sht_alm2map(self.nside, self.lmax, d_alm.data_ptr(), d_ln.data_ptr(), d_work.data_ptr(),
                                                  d_map.data_ptr(), _lib.current_stream()))
            return d_map.cpu().numpy()

    def alm2cl_on_device(self, d_alm):
        torch = _torch()
        d_cl = torch.empty(self.lmax + 1, dtype=torch.float64, device=d_alm.device)
        with torch.cuda.device(d_alm.device):
            _lib.check(_lib.lib().bfg_sht_alm2cl(self.lmax, d_alm.data_ptr(), d_cl.data_ptr(), _lib.current_stream()))
        return d_cl

    def anafast(self, m, iter=3):
        """hp.anafast(m) with healpy's defaults: C_l for l = 0 .. lmax as a numpy array."""
        return self.alm2cl_on_device(self.map2alm_on_device(m, iter)).cpu().numpy()


def anafast(m, lmax=None, iter=3, device=None):
    """Drop-in for `hp.anafast(map)` (one map, no polarisation): NSIDE from the map size."""
    n = m.numel() if hasattr(m, 'numel') else np.asarray(m).size
    nside = int(round(np.sqrt(n / 12.0)))
    if 12 * nside * nside != n:
        raise ValueError("Wrong pixel number (it is not 12*nside**2)")
    return ShellHarmonics(nside, lmax, device).anafast(m, iter)
