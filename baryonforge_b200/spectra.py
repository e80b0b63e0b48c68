"""
P(k) of a particle set on the device -- the measurement that follows BaryonifySnapshot.process() in the reference's
workflow.

The reference has no library function for this step; the algorithm is the cell code of
/root/reference/examples/10_Reproduce_Schneider_deltaPk.ipynb, cited below as nb10:cell.  `ShellPowerSpectrum` keeps the
variable names of those cells (Ngrd, Nk, Lbox, kbins, klin, k_c, k_cen) as attributes and `measure()` is one pass of the
`for factor in [1, 8]` loop of nb10:15: fold the box, count particles per cell, FFT, |F|^2, mean per k-shell.  The
deposit, the FFT (cuFFT, a plain library transform) and the shell sums (bfg_power_bin_spectrum) run on the device; with
`BaryonifySnapshot.process_on_device()` the displaced particles never leave HBM.  No CPU fallback.
"""
import numpy as np

from . import _lib

__all__ = ['ShellPowerSpectrum']


def _torch():
    import torch
    return torch


class ShellPowerSpectrum(object):
    """
    Ngrd, Nk, Lbox as in nb10:12 (the notebook uses Ngrd = 256, Nk = 180, Lbox = Snap.L).

    Attributes (numpy, same expressions as nb10:12): kbins, klin; k_c (modes per shell, int64) and k_cen (mean |k| per
    shell; NaN for an empty shell, like the notebook's 0/0) come from one device pass over the mode grid.
    """

    def __init__(self, Ngrd=256, Nk=180, Lbox=1.0, device=None):
        self.Ngrd, self.Nk, self.Lbox = int(Ngrd), int(Nk), float(Lbox)
        if self.Ngrd < 2 or self.Nk < 1 or not self.Lbox > 0:
            raise ValueError("ShellPowerSpectrum needs Ngrd >= 2, Nk >= 1, Lbox > 0")
        self.device = device
        Lbox, Ngrd = self.Lbox, self.Ngrd
        self.kbins = np.linspace(2 * np.pi / Lbox, 2 * np.pi / Lbox * Ngrd / 2, self.Nk + 1)      # nb10:12
        self.klin = np.fft.fftfreq(Ngrd, 1 / (2 * np.pi / (Lbox)) / Ngrd)                          # nb10:12
        self._d_klin = None
        self._shells = None

    # ---- plumbing
    def _dev(self):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.BFGError("baryonforge_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        return torch.device('cuda', torch.cuda.current_device() if self.device is None else int(self.device))

    def _klin_on(self, dev):
        torch = _torch()
        if self._d_klin is None or self._d_klin.device != dev:
            self._d_klin = torch.from_numpy(np.ascontiguousarray(self.klin)).to(dev)
        return self._d_klin

    def _bin(self, d_grid, dev, spectrum=None):
        """(sum |F|^2, sum |k|, count) per shell as numpy arrays; d_grid None -> mode counting only."""
        torch = _torch()
        L = _lib.lib()
        with torch.cuda.device(dev):
            out = torch.empty((2, self.Nk), dtype=torch.float64, device=dev)
            cnt = torch.empty(self.Nk, dtype=torch.int64, device=dev)
            k0, dk = float(self.kbins[0]), float(self.kbins[1] - self.kbins[0])
            d_klin = self._klin_on(dev)
            if d_grid is None:
                _lib.check(L.bfg_power_bin_spectrum(self.Ngrd, _lib.ptr(spectrum), d_klin.data_ptr(), k0, dk, self.Nk,
                                                    out[0].data_ptr(), out[1].data_ptr(), cnt.data_ptr(),
                                                    _lib.current_stream()))
            else:
                _lib.check(L.bfg_grid_power_spectrum(self.Ngrd, d_grid.data_ptr(), d_klin.data_ptr(), k0, dk, self.Nk,
                                                     out[0].data_ptr(), out[1].data_ptr(), cnt.data_ptr(),
                                                     _lib.current_stream()))
            o = out.cpu().numpy()
            return o[0], o[1], cnt.cpu().numpy()

    def _shell_table(self):
        if self._shells is None:
            _, ksum, cnt = self._bin(None, self._dev())
            with np.errstate(invalid='ignore', divide='ignore'):
                self._shells = (cnt, ksum / cnt)
        return self._shells

    @property
    def k_c(self):
        """np.bincount(kinds[kmsk], minlength = Nk)  (nb10:12)"""
        return self._shell_table()[0]

    @property
    def k_cen(self):
        """np.bincount(kinds[kmsk], minlength = Nk, weights = k[kmsk]) / k_c  (nb10:12)"""
        return self._shell_table()[1]

    # ---- the measurement
    def deposit_on_device(self, coords, factor=1):
        """
        `numba_histogram3d(Part % Lbox, bins = Ngrd, min_vals = 0.0, max_vals = Lbox)` with Lbox = self.Lbox / factor
        (nb10:1, nb10:15) as a float64 device tensor (Ngrd, Ngrd, Ngrd).  coords: an (n, 3) array like the notebook's Part_B,
        or three 1-D arrays / float64 CUDA tensors (x, y, z).  Sets self.last_dropped = number of non-finite particles.
        """
        torch = _torch()
        dev = self._dev()
        if isinstance(coords, np.ndarray) and coords.ndim == 2:
            if coords.shape[1] != 3:
                raise ValueError("coords must be (n, 3)")
            coords = [coords[:, 0], coords[:, 1], coords[:, 2]]
        if len(coords) != 3:
            raise ValueError("P(k) is measured on 3-D particle sets (x, y, z)")
        with torch.cuda.device(dev):
            d_p = []
            for c in coords:
                if torch.is_tensor(c):
                    d_p.append(c.to(device=dev, dtype=torch.float64).contiguous())
                else:
                    d_p.append(torch.from_numpy(np.ascontiguousarray(c, dtype=np.float64)).to(dev, non_blocking=True))
            n = d_p[0].numel()
            if d_p[1].numel() != n or d_p[2].numel() != n:
                raise ValueError("x, y, z must have the same length")
            d_grid = torch.zeros((self.Ngrd,) * 3, dtype=torch.float64, device=dev)
            d_drop = torch.zeros(1, dtype=torch.int64, device=dev)
            _lib.check(_lib.lib().bfg_snap_deposit_folded(n, _lib.ptr(d_p[0]), _lib.ptr(d_p[1]), _lib.ptr(d_p[2]),
                                                         self.Lbox / factor, self.Ngrd, d_grid.data_ptr(), d_drop.data_ptr(),
                                                         _lib.current_stream()))
            self.last_dropped = int(d_drop.cpu()[0])
        return d_grid

    def measure_runner(self, runner, factors=(1,)):
        """
        `Runner.process()` followed by one `measure()` per folding factor (the body of the parameter loop of nb10:15), with the
        deposits taken straight from the cell-ordered particles of the runner's cell list (bfg_snap_apply_deposit_folded): the
        displaced particles are never scattered back to the caller's order.  Returns [P(k) for factor in factors].
        Grids are bit-identical to `measure(runner.process_on_device(), factor)` (tests/test_gpu_spectrum.py).
        """
        torch = _torch()
        S = runner._displace_sorted()
        if S['ndim'] != 3:
            raise ValueError("P(k) is measured on 3-D particle sets (x, y, z)")
        d_s, out = S['d_s'], []
        with torch.cuda.device(S['dev']):
            d_drop = torch.zeros(1, dtype=torch.int64, device=S['dev'])
            for factor in factors:
                d_grid = torch.zeros((self.Ngrd,) * 3, dtype=torch.float64, device=S['dev'])
                _lib.check(_lib.lib().bfg_snap_apply_deposit_folded(
                    S['n_part'], d_s[0].data_ptr(), d_s[1].data_ptr(), d_s[2].data_ptr(), S['d_tot'].data_ptr(), S['L'],
                    self.Lbox / factor, self.Ngrd, d_grid.data_ptr(), d_drop.data_ptr(), _lib.current_stream()))
                out.append(self.measure_grid(d_grid))
            self.last_dropped = int(d_drop.cpu()[0])
            runner._finish_stats(S)
        return out

    def measure_grid(self, grid):
        """nb10:15 from the FFT on: np.bincount(kinds[kmsk], weights = |fftn(grid)|^2) / k_c for an (Ngrd,)*3 grid (numpy array
        or float64 CUDA tensor)."""
        torch = _torch()
        dev = self._dev()
        if torch.is_tensor(grid):
            d_grid = grid.to(device=dev, dtype=torch.float64).contiguous()
        else:
            d_grid = torch.from_numpy(np.ascontiguousarray(grid, dtype=np.float64)).to(dev)
        if tuple(d_grid.shape) != (self.Ngrd,) * 3:
            raise ValueError("grid must have shape (Ngrd, Ngrd, Ngrd)")
        pk_sum, _, _ = self._bin(d_grid, dev)
        with np.errstate(invalid='ignore', divide='ignore'):
            return pk_sum / self.k_c

    def measure(self, coords, factor=1):
        """One pass of the `for factor in [1, 8]` loop of nb10:15: P(k) (un-normalised, like the notebook's PkB) of the
        particle set folded `factor` times per axis; the matching wavenumbers are self.k_cen * factor."""
        return self.measure_grid(self.deposit_on_device(coords, factor))
