"""
oracle/pk_port.py -- TEST INFRASTRUCTURE (never imported by the product).

numpy restatement of the P(k) measurement that FOLLOWS BaryonifySnapshot.process() in the reference's workflow
(SURVEY.md section 8(f) item 4).  The reference has no library function for it: the algorithm lives in the cells of
/root/reference/examples/10_Reproduce_Schneider_deltaPk.ipynb, cited below as nb10:cell.  Pinned by
tests/golden/pk_nb10_*.npz, which oracle/make_golden.py produces by executing those very cells (source text read from the
notebook file, only Ngrd / Nk replaced by test sizes) -- see tests/test_oracle_pk.py.
"""
import numpy as np


def fold_positions(points, Lfold):
    """`Part_B % Lbox` (nb10:15, Lbox = Snap.L / factor): numpy's floored float remainder."""
    return np.asarray(points, dtype=np.float64) % Lfold


def histogram3d(points, bins, min_val, max_val):
    """
    numba_histogram3d (nb10:1): counts[int((p - min) / width)] += 1 per particle, int64 counts; `points` is (n, 3).
    int() truncates towards zero; the notebook only ever passes 0 <= p < max_val.
    """
    points = np.asarray(points, dtype=np.float64)
    width = (max_val - min_val) / bins
    idx = ((points - min_val) / width).astype(np.int64)          # truncation, like int() of a non-negative float
    flat = (idx[:, 0] * bins + idx[:, 1]) * bins + idx[:, 2]
    return np.bincount(flat, minlength=bins ** 3).reshape(bins, bins, bins).astype(np.int64)


class KBinning(object):
    """The k-shell set-up of nb10:12, line by line (note the axis order of the sum: axis 0, axis 2, axis 1)."""

    def __init__(self, Ngrd, Nk, Lbox):
        self.Ngrd, self.Nk, self.Lbox = int(Ngrd), int(Nk), float(Lbox)
        self.kbins = np.linspace(2 * np.pi / Lbox, 2 * np.pi / Lbox * Ngrd / 2, Nk + 1)
        self.klin = np.fft.fftfreq(Ngrd, 1 / (2 * np.pi / (Lbox)) / Ngrd)
        klin = self.klin
        k = np.sqrt(klin[:, None, None] ** 2 + klin[None, None, :] ** 2 + klin[None, :, None] ** 2).flatten()
        self.kinds = np.floor((k - self.kbins[0]) / (self.kbins[1] - self.kbins[0])).astype(int)
        self.kmsk = (self.kinds >= 0) & (self.kinds < Nk)
        self.k_c = np.bincount(self.kinds[self.kmsk], minlength=Nk)
        with np.errstate(invalid='ignore', divide='ignore'):
            self.k_cen = np.bincount(self.kinds[self.kmsk], minlength=Nk, weights=k[self.kmsk]) / self.k_c

    def power(self, grid):
        """nb10:15: FFT of the (count) grid, |F|^2, mean per k-shell."""
        F = np.fft.fftn(grid)
        P = (np.conjugate(F) * F).real.flatten()
        with np.errstate(invalid='ignore', divide='ignore'):
            return np.bincount(self.kinds[self.kmsk], minlength=self.Nk, weights=P[self.kmsk]) / self.k_c


def folded_power(points, kb, factor):
    """One pass of the `for factor in [1, 8]` loop of nb10:15 for the particle set `points` (n, 3): P(k) of the box folded
    `factor` times per axis; the matching wavenumbers are kb.k_cen * factor."""
    Lfold = kb.Lbox / factor
    grid = histogram3d(fold_positions(points, Lfold), kb.Ngrd, 0.0, Lfold)
    return kb.power(grid), grid
