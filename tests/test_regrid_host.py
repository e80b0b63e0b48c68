"""
CPU: the re-binning target of the shell re-binning kernels (regrid_target_fast in csrc/bfg_common.cuh -- small-angle azimuth and
colatitude against a per-nside ring table, no acos / atan2 / degrees round trip, no 64-bit division) compiled for the HOST
(bfg_test_regrid_target_host runs the same source the kernels run) against the oracle's literal chain
(HealpixRunner.py:357-361: pix2vec + offset -> hp.vec2ang(lonlat=True) -> hp.get_interp_weights(lonlat=True)).
tools/sass_fingerprint.py shows that making the function host-compilable left every kernel's SASS unchanged.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import hpo
from oracle import runners_port as rp


def fast_targets(nside, pix, off):
    from baryonforge_b200 import _lib
    n = pix.size
    pix = np.ascontiguousarray(pix, dtype=np.int64)
    off = np.ascontiguousarray(off, dtype=np.float64)
    out_pix = np.empty((n, 4), dtype=np.int64)
    out_w = np.empty((n, 4), dtype=np.float64)
    fast = np.empty(n, dtype=np.int32)
    _lib.check(_lib.lib().bfg_test_regrid_target_host(int(nside), n, pix.ctypes.data, off.ctypes.data, out_pix.ctypes.data,
                                                      out_w.ctypes.data, fast.ctypes.data))
    return out_pix, out_w, fast.astype(bool)


def literal_targets(nside, pix, off):
    """The oracle port's chain (oracle/runners_port.shell_regrid), per direction."""
    vec = np.stack(hpo.pix2vec(nside, pix), axis=1) + off.T
    dnorm = np.sqrt(np.sum(np.square(vec), axis=1))
    theta = np.arccos(vec[:, 2] / dnorm)
    phi = np.arctan2(vec[:, 1], vec[:, 0])
    phi[phi < 0] += 2 * np.pi
    lon, lat = np.degrees(phi), 90.0 - np.degrees(theta)
    c_pix, c_w = rp._interp_weights_lonlat(nside, lon, lat)
    return np.ascontiguousarray(c_pix.T), np.ascontiguousarray(c_w.T)


@pytest.mark.parametrize("nside", [64, 1000, 4096])          # 1000: not a power of two (RING maps may have any nside)
def test_small_angle_regrid_target_matches_the_literal_chain(nside):
    npix = 12 * nside * nside
    n = 200000
    rng = np.random.default_rng(900 + nside)
    pix = rng.integers(0, npix, n)
    pix[:2000] = rng.integers(0, min(npix, 40000), 2000)                   # north polar cap, first rings included
    pix[2000:4000] = npix - 1 - rng.integers(0, min(npix, 40000), 2000)    # south polar cap
    pixsize = np.sqrt(4 * np.pi / npix)
    off = rng.normal(0, 1.0, (3, n)) * pixsize * rng.choice([0.0, 0.01, 0.3, 2.5], n)[None, :]
    vx, vy, _ = hpo.pix2vec(nside, pix[4000:4050])
    off[:, 4000:4050] = 0.3 * np.stack([-vy, vx, np.zeros(50)]) / np.hypot(vx, vy)   # 0.3 rad sideways: far beyond the guards
    f_pix, f_w, fast = fast_targets(nside, pix, off)
    l_pix, l_w = literal_targets(nside, pix, off)
    # the function declines only where it must: next to the poles and for the huge displacements
    assert fast.mean() > (0.97 if nside >= 1000 else 0.85), fast.mean()     # NSIDE = 64: a 2.5-pixel offset is 0.04 rad
    assert not fast[4000:4050].any()
    declined = ~fast
    declined[4000:4050] = False
    z_src = hpo.pix2vec(nside, pix)[2]
    z_cut = 1.0 - (min(400, nside) / nside) ** 2 / 3.0                     # |z| of ring 400 (or of the polar-cap boundary)
    if nside >= 1000:      # pixel-sized offsets are far inside the guards there: only the polar neighbourhood is declined
        assert not declined.any() or np.abs(z_src[declined]).min() >= z_cut - 1e-12, np.abs(z_src[declined]).min()
    fp, fw, lp, lw = f_pix[fast], f_w[fast], l_pix[fast], l_w[fast]
    assert np.all((fp >= 0) & (fp < npix))
    assert np.allclose(fw.sum(axis=1), 1.0, rtol=0, atol=1e-12)
    assert fw.min() > -1e-8 and fw.max() < 1 + 1e-8
    # (1) as deposits: what each path adds to the map, robust against ring-boundary ties (the bilinear weights are continuous)
    m_f, m_l = np.zeros(npix), np.zeros(npix)
    np.add.at(m_f, fp.ravel(), fw.ravel())
    np.add.at(m_l, lp.ravel(), lw.ravel())
    assert np.abs(m_f - m_l).max() < 2e-8, np.abs(m_f - m_l).max()
    # (2) direction by direction: same four pixels for all but round-off ties, and then the same weights.  Next to the poles the
    # LITERAL arccos(z / |v|) loses 1 / sin(theta) in precision, hence 1e-8 rather than 1e-11.
    of, ol = np.argsort(fp, axis=1), np.argsort(lp, axis=1)
    sfp, slp = np.take_along_axis(fp, of, 1), np.take_along_axis(lp, ol, 1)
    same = np.all(sfp == slp, axis=1)
    # (an undisplaced pixel centre sits exactly on its ring and exactly between two azimuthal cells: every choice there is a tie
    # with weight 0 on the pixels that differ -- deposits (1) cover those; the per-direction comparison takes the displaced ones)
    moved = np.any(off[:, fast] != 0.0, axis=0)
    assert same[moved].mean() > 0.999, same[moved].mean()
    # weights of equal pixels may be split differently when a pixel appears twice (degenerate pairs): compare per-pixel sums
    for k in range(4):
        wf = np.where(sfp == sfp[:, [k]], np.take_along_axis(fw, of, 1), 0.0).sum(axis=1)
        wl = np.where(slp == slp[:, [k]], np.take_along_axis(lw, ol, 1), 0.0).sum(axis=1)
        assert np.abs(wf - wl)[same].max() < 1e-8
    if nside == 4096:      # the equatorial belt, where both chains are clean: agreement at the level of phi / dphi round-off
        belt = same & (np.abs(hpo.pix2vec(nside, pix[fast])[2]) < 0.6)
        wf = np.take_along_axis(fw, of, 1)
        wl = np.take_along_axis(lw, ol, 1)
        distinct = np.all(np.diff(sfp, axis=1) != 0, axis=1)
        assert np.abs(wf - wl)[belt & distinct].max() < 5e-11
