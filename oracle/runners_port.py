"""
oracle/runners_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Self-contained CPU restatement (numpy + scipy + oracle/*.c) of the five `process()` loops of the
reference's Runners and of the two table read-outs, written so that it can run on a machine where
/root/reference does not exist (the GPU box).  Each function cites the reference lines it follows.
The third-party scalars the reference obtains from pyccl per halo (R_200c in the runner cosmology,
R_200c in the model cosmology, D_A(z)) are INPUTS here, exactly as they are inputs to the CUDA path.

Pinned: tests/test_oracle_port.py compares every function below with the reference's own runner
code executed in the build container (fixtures in tests/golden/, generator oracle/make_golden.py),
and -- when /root/reference is importable -- with the reference live.

Per-halo Python loops are kept on purpose: that is what the reference executes, so timing this
file is an honest stand-in (`cpu_baseline.kind = "port"`) for the reference's CPU path.
"""
import ctypes as C
import os
import warnings

import numpy as np
from scipy.interpolate import RegularGridInterpolator

from . import hpo

_HERE = os.path.dirname(os.path.abspath(__file__))
_GRID = None


def _gridlib():
    global _GRID
    if _GRID is None:
        hpo.build()
        L = C.CDLL(os.path.join(_HERE, "_build", "libgrid.so"))
        pdbl = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.grido_regrid_2d.argtypes = [pdbl, C.c_int64, C.c_int64, pdbl, pdbl]
        L.grido_regrid_3d.argtypes = [pdbl, C.c_int64, C.c_int64, pdbl, pdbl]
        _GRID = L
    return _GRID


# ------------------------------------------------------------------------------------------------
# table read-outs
# ------------------------------------------------------------------------------------------------
class DisplacementTable(object):
    """
    What BaryonificationClass.setup_interpolator leaves behind
    (/root/reference/BaryonForge/Profiles/BaryonCorrection.py:307-323): the axes
    (ln(1+z), ln M, ln r | ln r/R, *extras), the values, Rdelta_sampling and the model's epsilon_max.
    """

    def __init__(self, axes, values, epsilon_max, Rdelta_sampling=False, p_keys=()):
        self.axes = tuple(np.asarray(a, dtype=np.float64) for a in axes)
        self.values = np.asarray(values, dtype=np.float64)
        self.epsilon_max = epsilon_max
        self.Rdelta_sampling = Rdelta_sampling
        self.p_keys = list(p_keys)
        # BaryonCorrection.py:322
        self.interp_d = RegularGridInterpolator(self.axes, self.values, bounds_error=False, fill_value=np.nan)

    def displacement(self, r, M, a, R_com, warn=True, **kwargs):
        """
        BaryonificationClass._readout for scalar M, a (BaryonCorrection.py:331-419).  `R_com` replaces
        `self.mass_def.get_radius(self.cosmo, M, a)/a` (:399).  The per-call range checks (:378-394) are
        kept because they are part of what the reference spends per halo.
        """
        r_use = np.atleast_1d(r)
        empty = np.ones_like(r_use)
        z_in = np.log(1 / a) * empty                       # :371
        r_in = np.log(r_use)                               # :372
        k_in = [kwargs[k] * empty for k in self.p_keys]    # :373
        z_use = 1 / np.atleast_1d(a) - 1
        z_tab = np.exp(self.axes[0]) - 1                   # :378-380
        M_tab = np.exp(self.axes[1])
        r_tab = np.exp(self.axes[2])
        M_use = np.atleast_1d(M)
        if warn:
            if (np.min(z_use) < np.min(z_tab)) | (np.max(z_use) > np.max(z_tab)):
                warnings.warn("Requested redshift range outside table's range", UserWarning)
            if (np.min(M_use) < np.min(M_tab)) | (np.max(M_use) > np.max(M_tab)):
                warnings.warn("Requested log_Mass range outside table's range", UserWarning)
            if not self.Rdelta_sampling:
                if (np.min(r_use) < np.min(r_tab)) | (np.max(r_use) > np.max(r_tab)):
                    warnings.warn("Requested Radius range outside table's range", UserWarning)
        M_in = np.log(M) * empty                           # :398 (float32 M -> float32 log, as numpy does)
        if not self.Rdelta_sampling:
            p_in = tuple([z_in, M_in, r_in] + k_in)        # :404-405
        else:
            p_in = tuple([z_in, M_in, r_in - np.log(R_com)] + k_in)   # :407-408
        displ = self.interp_d(p_in)
        inside = (r_use < self.epsilon_max * R_com)        # :410
        return np.where(inside, displ, 0)                  # :411


class ProfileTable(object):
    """
    TabulatedProfile after setup_interpolator (/root/reference/BaryonForge/utils/Tabulate.py:261-271):
    RegularGridInterpolator over np.log(table), NaN outside, exp() on read-out (:279-327).
    `raw3D` / `raw2D` are the un-logged tables (either may be None).
    """

    def __init__(self, axes, raw3D=None, raw2D=None, p_keys=()):
        self.axes = tuple(np.asarray(a, dtype=np.float64) for a in axes)
        self.p_keys = list(p_keys)
        with np.errstate(divide='ignore', invalid='ignore'):
            self.log3D = None if raw3D is None else np.log(np.asarray(raw3D, dtype=np.float64))
            self.log2D = None if raw2D is None else np.log(np.asarray(raw2D, dtype=np.float64))
        self.interp3D = None if raw3D is None else RegularGridInterpolator(self.axes, self.log3D, bounds_error=False)
        self.interp2D = None if raw2D is None else RegularGridInterpolator(self.axes, self.log2D, bounds_error=False)

    def _readout(self, r, M, a, table, **kwargs):
        r_use = np.atleast_1d(r)
        empty = np.ones_like(r_use)
        z_in = np.log(1 / a) * empty       # Tabulate.py:312
        r_in = np.log(r_use)               # :313
        M_in = np.log(M) * empty           # :317
        k_in = [kwargs[k] * empty for k in self.p_keys]
        with np.errstate(over='ignore', invalid='ignore'):
            return np.exp(table(tuple([z_in, M_in, r_in] + k_in)))  # :318-319

    def projected(self, r, M, a, **kw):
        return self._readout(r, M, a, self.interp2D, **kw)   # Tabulate.py:362-391

    def real(self, r, M, a, **kw):
        return self._readout(r, M, a, self.interp3D, **kw)   # Tabulate.py:330-359


# ------------------------------------------------------------------------------------------------
# HEALPix shells
# ------------------------------------------------------------------------------------------------
def _ang2vec_lonlat(ra, dec):
    # healpy.ang2vec(lonlat=True): theta = pi/2 - radians(lat), phi = radians(lon)
    theta, phi = np.pi / 2.0 - np.radians(dec), np.radians(ra)
    st = np.sin(theta)
    return np.array([st * np.cos(phi), st * np.sin(phi), np.cos(theta)]).T


def _query_disc_vec(nside, vec, radius):
    theta, phi = hpo.vec2pointing(vec)
    return hpo.query_disc(nside, theta, phi, radius)


def _interp_weights_lonlat(nside, lon, lat):
    theta, phi = np.pi / 2.0 - np.radians(lat), np.radians(lon)
    return hpo.get_interpol(nside, theta, phi)


def shell_offsets(nside, cat, R_run, D_A, R_model_com, eps_runner, table, extras=None, fallback=True,
                  count_only=False, warn=True):
    """
    The halo loop of BaryonifyShell.process (/root/reference/BaryonForge/Runners/HealpixRunner.py:313-355).
    cat: dict/structured array with 'M','z','ra','dec'.  Returns (pix_offsets[npix,3], n_updates).
    """
    npix = 12 * nside * nside
    pix_offsets = np.zeros([npix, 3])
    n_updates = 0
    keys = table.p_keys
    for j in range(len(cat['M'])):
        M_j = cat['M'][j]
        z_j = cat['z'][j]
        a_j = 1 / (1 + z_j)
        R_j = R_run[j]
        D_j = D_A[j]
        o_j = {key: extras[key][j] for key in keys}
        ra_j, dec_j = cat['ra'][j], cat['dec'][j]
        vec_j = _ang2vec_lonlat(ra_j, dec_j)
        radius = R_j * eps_runner / D_j
        pixind = _query_disc_vec(nside, vec_j, radius)
        if fallback and pixind.size < 4:                                  # :333-334
            pixind = _interp_weights_lonlat(nside, ra_j, dec_j)[0][:, 0]
        n_updates += pixind.size
        if count_only:
            continue
        vec = np.stack(hpo.pix2vec(nside, pixind), axis=1)                # :336
        pos_j = vec_j * D_j
        pos = vec * D_j
        diff = pos - pos_j
        r_sep = np.sqrt(np.sum(diff ** 2, axis=1))
        with np.errstate(divide='ignore', invalid='ignore'):
            offset = table.displacement(r_sep / a_j, M_j, a_j, R_model_com[j], warn=warn, **o_j) * a_j   # :345
            offset = offset[:, None] * (diff / r_sep[:, None])
            offset = np.where(np.isfinite(offset), offset, 0)             # :347
            nw_pos = pos + offset
            nw_vec = nw_pos / np.sqrt(np.sum(nw_pos ** 2, axis=1))[:, None]
        offset = nw_vec - vec
        pix_offsets[pixind, :] += offset                                  # :355
    return pix_offsets, n_updates


def shell_regrid(nside, orig_map, pix_offsets):
    """BaryonifyShell.process after the loop (HealpixRunner.py:357-370)."""
    npix = orig_map.size
    new_vec = np.stack(hpo.pix2vec_range(nside, 0, npix), axis=1) + pix_offsets      # :357
    # hp.vec2ang(lonlat=True) then get_interp_weights(lonlat=True): radians -> degrees -> radians
    dnorm = np.sqrt(np.sum(np.square(new_vec), axis=1))
    theta = np.arccos(new_vec[:, 2] / dnorm)
    phi = np.arctan2(new_vec[:, 1], new_vec[:, 0])
    phi[phi < 0] += 2 * np.pi
    lon, lat = np.degrees(phi), 90.0 - np.degrees(theta)                             # :358
    p_pix = np.where(orig_map != 0)[0]                                               # :359
    c_pix, c_weight = _interp_weights_lonlat(nside, lon[p_pix], lat[p_pix])          # :361
    c_pix, c_weight = np.ascontiguousarray(c_pix.T), np.ascontiguousarray(c_weight.T)
    new_map = np.zeros(orig_map.size, dtype=float)
    new_map = hpo.regrid_scatter(new_map, np.ascontiguousarray(orig_map[p_pix], dtype=np.float64), c_pix, c_weight)
    new_sum, old_sum = np.sum(new_map), np.sum(orig_map)
    assert np.isclose(new_sum, old_sum), "ERROR in pixel regridding"                 # :368-370
    return new_map


def baryonify_shell(nside, orig_map, cat, R_run, D_A, R_model_com, eps_runner, table, extras=None, warn=True):
    """BaryonifyShell.process (HealpixRunner.py:252-373)."""
    if np.allclose(orig_map, 0):       # :293-294
        return orig_map
    off, _ = shell_offsets(nside, cat, R_run, D_A, R_model_com, eps_runner, table, extras, warn=warn)
    return shell_regrid(nside, orig_map, off)


def paint_shell(nside, cat, R_run, D_A, eps_runner, table, include_pixel_size=False, extras=None):
    """PaintProfilesShell.process (HealpixRunner.py:390-483).  Returns (new_map, n_updates)."""
    npix = 12 * nside * nside
    new_map = np.zeros(npix, dtype=np.float64)
    pixarea = 4 * np.pi / npix
    n_updates = 0
    keys = table.p_keys
    for j in range(len(cat['M'])):
        M_j = cat['M'][j]
        z_j = cat['z'][j]
        a_j = 1 / (1 + z_j)
        R_j, D_j = R_run[j], D_A[j]
        o_j = {key: extras[key][j] for key in keys}
        vec_j = _ang2vec_lonlat(cat['ra'][j], cat['dec'][j])
        radius = R_j * eps_runner / D_j
        pixind = _query_disc_vec(nside, vec_j, radius)
        n_updates += pixind.size
        vec = np.stack(hpo.pix2vec(nside, pixind), axis=1)
        pos_j = vec_j * D_j
        pos = vec * D_j
        diff = pos - pos_j
        r_sep = np.sqrt(np.sum(diff ** 2, axis=1))
        with np.errstate(divide='ignore', invalid='ignore'):
            Paint = table.projected(r_sep / a_j, M_j, a_j, **o_j)         # :472
        Paint = np.where(np.isfinite(Paint), Paint, 0)                    # :473
        if include_pixel_size:
            Paint = Paint * (pixarea * D_j ** 2)                          # :478
        new_map[pixind] += Paint                                          # :481
    return new_map, n_updates


def paint_anis_shell(nside, orig_map, cat, R_run, D_A, eps_runner, table, tracer_table, mtot_table, dD, rho_m,
                     proj_cutoff, background_val, global_tracer_fraction, include_pixel_size=False, extras=None,
                     mtot_extras=None):
    """
    PaintProfilesAnisShell.process (/root/reference/BaryonForge/Runners/HealpixRunner.py:510-640).
    Third-party scalars are inputs: dD = D_a(LightconeShell.redshift) (:574), rho_m = cosmo.rho_x(a, 'matter',
    is_comoving=False) (:580).  Returns (new_map, n_updates, Mtot_map incl. background).
    """
    npix = 12 * nside * nside
    new_map = np.zeros(npix, dtype=np.float64)
    pixarea = 4 * np.pi / npix
    Mtot_map, _ = paint_shell(nside, cat, R_run, D_A, eps_runner, mtot_table, True, extras=mtot_extras)   # :565-570
    dL = 2 * proj_cutoff                                                   # :573
    dV = pixarea * ((dD + dL) ** 3 - dD ** 3)                              # :575
    rho_halos = np.sum(Mtot_map) / (dV * Mtot_map.size)                    # :576
    drho_m = np.clip(rho_m - rho_halos, 0, None)                           # :581
    Mtot_map += dV * drho_m                                                # :582
    n_updates = 0
    keys = table.p_keys
    for j in range(len(cat['M'])):
        M_j = cat['M'][j]
        z_j = cat['z'][j]
        a_j = 1 / (1 + z_j)
        R_j, D_j = R_run[j], D_A[j]
        o_j = {key: extras[key][j] for key in keys}
        vec_j = _ang2vec_lonlat(cat['ra'][j], cat['dec'][j])
        radius = R_j * eps_runner / D_j
        pixind = _query_disc_vec(nside, vec_j, radius)
        n_updates += pixind.size
        vec = np.stack(hpo.pix2vec(nside, pixind), axis=1)
        pos_j = vec_j * D_j
        pos = vec * D_j
        diff = pos - pos_j
        r_sep = np.sqrt(np.sum(diff ** 2, axis=1))
        with np.errstate(divide='ignore', invalid='ignore'):
            Painting = table.projected(r_sep / a_j, M_j, a_j, **o_j)       # :610
            Painting = np.where(np.isfinite(Painting), Painting, 0)        # :611
            Canvas = tracer_table.projected(r_sep / a_j, M_j, a_j, **o_j)  # :612
        Canvas = np.where(np.isfinite(Canvas) & np.invert(np.isnan(Canvas)), Canvas, 0)   # :613
        Mfrac = np.divide(Canvas, Mtot_map[pixind], out=np.zeros_like(Canvas), where=Mtot_map[pixind] > 0)   # :614
        Mfrac *= orig_map[pixind]                                          # :615
        if include_pixel_size:
            Painting = Painting * (pixarea * D_j ** 2)                     # :620
        new_map[pixind] += Painting * Mfrac                                # :623
    Mfrac = np.divide(dV * drho_m, Mtot_map, out=np.zeros_like(Mtot_map), where=Mtot_map > 0)   # :626
    Mfrac *= orig_map
    new_map += background_val * global_tracer_fraction * Mfrac             # :628
    return new_map.reshape(orig_map.shape), n_updates, Mtot_map


# ------------------------------------------------------------------------------------------------
# periodic grids
# ------------------------------------------------------------------------------------------------
def _pick_indices(center, width, Npix):
    # Map2DRunner.py:400-429
    inds = np.arange(center - width, center + width)
    inds = np.where(inds < 0, inds + Npix, inds)
    inds = np.where(inds >= Npix, inds - Npix, inds)
    return inds


def _flat_inds(N, ndim, *axis_inds):
    """GriddedMap.inds[x_inds,...][:, y_inds,...][..., z_inds].flatten() without the N^d index cube (io.py:470)."""
    if ndim == 2:
        return (axis_inds[0][:, None] * N + axis_inds[1][None, :]).ravel()
    return ((axis_inds[0][:, None, None] * N + axis_inds[1][None, :, None]) * N + axis_inds[2][None, None, :]).ravel()


def build_Rmat(A, q):
    """DefaultRunnerGrid.build_Rmat, 2-D branch (Map2DRunner.py:281-350), literal."""
    A = A / np.linalg.norm(A)            # the reference does this in place (:310)
    ref = np.array([1., 0.])
    beta = np.arccos(np.dot(A, ref))
    eta = -np.log(q)
    if eta > 1e-4:
        eta2g = np.tanh(0.5 * eta) / eta
    else:
        etasq = eta * eta
        eta2g = 0.5 + etasq * ((-1 / 24) + etasq * (1 / 240))
    g = eta2g * eta * np.exp(2j * beta)
    g1, g2 = g.real, g.imag
    det = np.sqrt(1 - np.abs(g) ** 2)
    return np.array([[1 + g1, g2], [g2, 1 - g1]]) / det


def _ell_radius(grids, d, A_j, q_j):
    """Map2DRunner.py:531-536 / :769-774: r of the sheared coordinates."""
    A_j = A_j / np.sqrt(np.sum(A_j ** 2))                               # :497
    Rmat = build_Rmat(A_j, q_j)
    xy = np.vstack([(grids[0] + d[0]).flatten(), (grids[1] + d[1]).flatten()]).T   # coord_array
    xe, ye = (xy @ Rmat).T
    return np.sqrt(xe ** 2 + ye ** 2)


def _cutout(bins, res, Nfloat, pos):
    """Map2DRunner.py:500-528 / :548-566 -- shared cutout construction.  pos = (x_j, y_j[, z_j])."""
    Nsize = int(Nfloat // 2) * 2
    Nsize = int(np.clip(Nsize, 2, bins.size // 2))
    x = np.linspace(-Nsize / 2, Nsize / 2, Nsize) * res
    cw = Nsize // 2
    cens = [int(np.argmin(np.abs(bins - p))) for p in pos]
    axis_inds = [_pick_indices(c, cw, bins.size) for c in cens]
    d = [bins[c] - p for c, p in zip(cens, pos)]
    grids = np.meshgrid(*([x] * len(pos)), indexing='xy')
    return Nsize, axis_inds, d, grids


def grid_offsets(shape, bins, cat, a, R_phys, R_model_com, eps_runner, table, extras=None, warn=True,
                 count_only=False, ell=None):
    """
    Halo loop of BaryonifyGrid.process (/root/reference/BaryonForge/Runners/Map2DRunner.py:474-586).
    cat has float32 fields 'M','x','y','z' (HaloNDCatalog).  R_phys = get_radius(cosmo, M, a) (physical).
    Returns (pix_offsets[N^d, d] BEFORE the isfinite clean, n_updates).
    """
    ndim = len(shape)
    N = shape[0]
    res = bins[1] - bins[0]
    pix_offsets = np.zeros([N ** ndim, ndim])
    n_updates = 0
    keys = table.p_keys
    for j in range(len(cat['M'])):
        M_j = cat['M'][j]
        pos = [cat['x'][j], cat['y'][j]] + ([cat['z'][j]] if ndim == 3 else [])
        o_j = {key: extras[key][j] for key in keys}
        R_q = eps_runner * R_phys[j] / a                                   # :492
        R_q = np.clip(R_q, 0, np.max(bins) / 2)                            # :493
        Nsize, axis_inds, d, grids = _cutout(bins, res, 2 * R_q / res, pos)
        n_updates += Nsize ** ndim
        if count_only:
            continue
        inds = _flat_inds(N, ndim, *axis_inds)
        with np.errstate(divide='ignore', invalid='ignore'):
            r_grid = np.sqrt(sum((g + dd) ** 2 for g, dd in zip(grids, d)))
            hats = [(g + dd) / r_grid for g, dd in zip(grids, d)]
            if ell is not None:                                                # (q_ell, A_ell) columns, 2-D only
                r_grid = _ell_radius(grids, d, ell[1][j], ell[0][j])
            offset = table.displacement(r_grid.flatten(), M_j, a, R_model_com[j], warn=warn, **o_j) / res   # :540/:583
            for k in range(ndim):
                pix_offsets[inds, k] += offset * hats[k].flatten()
    return pix_offsets, n_updates


def grid_regrid(orig_map, pix_offsets):
    """BaryonifyGrid.process after the loop (Map2DRunner.py:589-619)."""
    ndim = orig_map.ndim
    N = orig_map.shape[0]
    x = np.arange(N)
    grids = np.meshgrid(*([x] * ndim), indexing='xy')
    pix_offsets = np.where(np.isfinite(pix_offsets), pix_offsets, 0)       # :597/:607
    for k in range(ndim):
        pix_offsets[:, k] += grids[k].flatten()
    new_map = np.zeros(orig_map.shape, dtype=np.float64)
    flat = np.ascontiguousarray(orig_map.flatten(), dtype=np.float64)
    pos = np.ascontiguousarray(pix_offsets)
    if ndim == 2:
        _gridlib().grido_regrid_2d(new_map, N, flat.size, pos, flat)
    else:
        _gridlib().grido_regrid_3d(new_map, N, flat.size, pos, flat)
    assert np.isclose(np.sum(new_map), np.sum(flat)), "ERROR in pixel regridding"
    return new_map


def baryonify_grid(orig_map, bins, cat, a, R_phys, R_model_com, eps_runner, table, extras=None, warn=True, ell=None):
    off, _ = grid_offsets(orig_map.shape, bins, cat, a, R_phys, R_model_com, eps_runner, table, extras, warn, ell=ell)
    return grid_regrid(orig_map, off)


def paint_grid(shape, bins, cat, a, R_com, eps_runner, table, include_pixel_size=True, extras=None, ell=None):
    """PaintProfilesGrid.process (Map2DRunner.py:676-829).  R_com = get_radius(cosmo, M, a)/a.  Returns (map, n_updates)."""
    ndim = len(shape)
    N = shape[0]
    res = bins[1] - bins[0]
    new_map = np.zeros(N ** ndim, dtype=np.float64)
    dV = np.power(res, ndim)
    n_updates = 0
    keys = table.p_keys
    profile = table.projected if ndim == 2 else table.real                # :763 / :792
    for j in range(len(cat['M'])):
        M_j = cat['M'][j]
        pos = [cat['x'][j], cat['y'][j]] + ([cat['z'][j]] if ndim == 3 else [])
        o_j = {key: extras[key][j] for key in keys}
        R_j = R_com[j]
        Nsize, axis_inds, d, grids = _cutout(bins, res, 2 * eps_runner * R_j / res, pos)   # :740-746
        n_updates += Nsize ** ndim
        inds = _flat_inds(N, ndim, *axis_inds)
        r_grid = np.sqrt(sum((g + dd) ** 2 for g, dd in zip(grids, d)))
        if ell is not None:
            r_grid = _ell_radius(grids, d, ell[1][j], ell[0][j])
        with np.errstate(divide='ignore', invalid='ignore'):
            Painting = profile(r_grid.flatten(), M_j, a, **o_j)           # :812
        mask = np.isfinite(Painting)
        mask = mask & (r_grid.flatten() < R_j * eps_runner)               # :814-815
        if mask.sum() == 0:
            continue
        Painting = np.where(mask, Painting, 0)
        new_map[inds] += Painting                                         # :821
    if include_pixel_size:
        new_map *= dV                                                     # :825
    return new_map.reshape(shape), n_updates


def paint_anis_grid(orig_map, bins, cat, a, R_com, eps_runner, table, tracer_table, mtot_table, rho_m, proj_cutoff,
                    background_val, global_tracer_fraction, include_pixel_size=True, extras=None, mtot_extras=None,
                    ell=None):
    """
    PaintProfilesAnisGrid.process (/root/reference/BaryonForge/Runners/Map2DRunner.py:845-1017), 2-D maps only (:847).
    rho_m = cosmo.rho_x(a, 'matter', is_comoving=True) (:886) is an input.  Returns (new_map, n_updates).
    """
    assert orig_map.ndim == 2, "Can only paint tSZ on 2D maps. You have passed a 3D Map"
    shape = orig_map.shape
    N = shape[0]
    res = bins[1] - bins[0]
    new_map = np.zeros(orig_map.size, dtype=np.float64)
    flat = orig_map.flatten()
    Mtot_map, _ = paint_grid(shape, bins, cat, a, R_com, eps_runner, mtot_table, include_pixel_size=False,
                             extras=mtot_extras, ell=ell)                  # :866-871
    Mtot_map = Mtot_map.flatten()
    dL = 2 * proj_cutoff                                                   # :877
    dV = np.power(res, 2) * dL
    rho_halos = np.average(Mtot_map) / dL
    drho_m = np.clip(rho_m - rho_halos, 0, None)                           # :887
    Mtot_map += dV * drho_m
    n_updates = 0
    keys = table.p_keys
    for j in range(len(cat['M'])):
        M_j = cat['M'][j]
        pos = [cat['x'][j], cat['y'][j]]
        o_j = {key: extras[key][j] for key in keys}
        R_j = R_com[j]
        Nsize, axis_inds, d, grids = _cutout(bins, res, 2 * eps_runner * R_j / res, pos)   # :907-915
        n_updates += Nsize ** 2
        inds = _flat_inds(N, 2, *axis_inds)
        r_grid = np.sqrt(sum((g + dd) ** 2 for g, dd in zip(grids, d)))
        if ell is not None:
            r_grid = _ell_radius(grids, d, ell[1][j], ell[0][j])
        with np.errstate(divide='ignore', invalid='ignore'):
            Painting = table.projected(r_grid.flatten(), M_j, a, **o_j)    # :981
            Canvas = tracer_table.projected(r_grid.flatten(), M_j, a, **o_j)
        Canvas = np.where(np.isfinite(Canvas) & np.invert(np.isnan(Canvas)), Canvas, 0)        # :983
        Mfrac = np.divide(Canvas, Mtot_map[inds], out=np.zeros_like(Canvas), where=Mtot_map[inds] > 0)
        Mfrac *= flat[inds]
        mask = np.isfinite(Painting) & np.invert(np.isnan(Painting))       # :987
        mask = mask & (r_grid.flatten() < R_j * eps_runner)
        if mask.sum() == 0:
            continue
        Painting = np.where(mask, Painting, 0)
        new_map[inds] += Painting * Mfrac                                  # :996
    Mfrac = np.divide(dV * drho_m, Mtot_map, out=np.zeros_like(Mtot_map), where=Mtot_map > 0)   # :1004
    Mfrac *= flat
    new_map += background_val * global_tracer_fraction * Mfrac
    new_map = new_map.reshape(shape)
    if include_pixel_size:
        new_map *= np.power(res, 2)                                        # :1012-1015
    return new_map, n_updates


# ------------------------------------------------------------------------------------------------
# particle snapshots
# ------------------------------------------------------------------------------------------------
def _periodic(dx, L):
    # SnapshotRunner.py:135-158
    dx = np.where(dx > L / 2, dx - L, dx)
    dx = np.where(dx < -L / 2, dx + L, dx)
    return dx


def baryonify_snapshot(px, L, cat, a, R_phys, R_model_com, eps_runner, table, tree=None, extras=None, warn=True,
                       KDTree_kwargs=None):
    """
    BaryonifySnapshot.process (/root/reference/BaryonForge/Runners/SnapshotRunner.py:83-100,176-274).
    px: list of 2 or 3 coordinate arrays (f64).  Returns (list of displaced coordinate arrays, n_pairs, tree).
    """
    from scipy.spatial import KDTree
    ndim = len(px)
    if tree is None:
        tree = KDTree(np.vstack(px).T, boxsize=L, **(KDTree_kwargs or {}))   # :100
    tot = np.zeros([px[0].size, ndim])
    names = ['x', 'y', 'z'][:ndim]
    n_pairs = 0
    keys = table.p_keys
    for j in range(len(cat['M'])):
        M_j = cat['M'][j]
        cen = [cat[n][j] for n in names]
        o_j = {key: extras[key][j] for key in keys}
        R_q = eps_runner * R_phys[j] / a           # :227
        R_q = np.clip(R_q, 0, L / 2)               # :228
        inds = tree.query_ball_point(cen, R_q)     # :232/:247
        n_pairs += len(inds)
        dxs = [p[inds] - c for p, c in zip(px, cen)]
        d = np.sqrt(sum(_periodic(dx, L) ** 2 for dx in dxs))      # compute_distance :103-132
        with np.errstate(divide='ignore', invalid='ignore'):
            hats = [_periodic(dx, L) / d for dx in dxs]
            offset = table.displacement(d, M_j, a, R_model_com[j], warn=warn, **o_j)
        offset = np.where(np.isfinite(offset), offset, 0)          # :259
        tot[inds] += np.vstack([offset * h for h in hats]).T       # :260
    out = []
    for k in range(ndim):
        q = px[k] + tot[:, k]                      # :264-266
        q = np.where(q > L, q - L, q)              # :272
        q = np.where(q < 0, q + L, q)              # :273
        out.append(q)
    return out, n_pairs, tree


def make_map_ngp(px, M, L, N_grid):
    """ParticleSnapshot.make_map (/root/reference/BaryonForge/utils/io.py:629-677)."""
    bins = np.linspace(0, L, N_grid + 1)
    coords = np.vstack(px).T
    return np.histogramdd(coords, bins=tuple([bins] * len(px)), weights=M)[0]
