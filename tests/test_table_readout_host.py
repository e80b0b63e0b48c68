"""
CPU: the table read-out of the halo-loop kernels (RowBlender + row_lookup in csrc/bfg_common.cuh: the 2^(ndim-1) corner rows of the
non-radial axes blended into one radial row, then interpolation along ln r) compiled for the HOST (bfg_test_table_readout_host runs
the kernels' own source), against the real scipy.interpolate.RegularGridInterpolator the reference calls
(BaryonCorrection.py:331-419, utils/Tabulate.py:279-327): inside, on and beyond every edge, NaN / -inf table entries, NaN and infinite
radii, uniform and non-uniform radial axes, 3-D to 6-D tables (p_keys), log-valued profile tables.
tools/sass_fingerprint.py shows no kernel changed when these functions became host-compilable.
"""
import ctypes as C

import numpy as np
from scipy.interpolate import RegularGridInterpolator as RGI

from baryonforge_b200 import _lib, synth
from helpers import assert_close


def host_readout(axes, values, lnz, lnM, x, extras=None, flags=0, force_search=False):
    axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
    values = np.ascontiguousarray(values, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    shape = (C.c_int64 * len(axes))(*[a.size for a in axes])
    ptrs = (C.c_void_p * len(axes))(*[a.ctypes.data for a in axes])
    ex = None if extras is None else np.ascontiguousarray(extras, dtype=np.float64)
    _lib.check(_lib.lib().bfg_test_table_readout_host(len(axes), shape, ptrs, values.ctypes.data, int(flags),
                                                      1 if force_search else 0, float(lnz), float(lnM), _lib.ptr(ex), x.size,
                                                      x.ctypes.data, out.ctypes.data))
    return out


def test_readout_source_on_host_matches_scipy_regular_grid_interpolator():
    axes = synth.table_axes(nz=7, nM=9, nr=300)
    vals = synth.displacement_values(axes, inject_nan=True)
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(axes[2][0] - 0.5, axes[2][-1] + 0.5, 20000), axes[2], np.nextafter(axes[2], np.inf),
                        np.nextafter(axes[2], -np.inf), [np.nan, -np.inf, np.inf]])
    ax_nu = axes[2].copy()
    ax_nu[1:-1] += rng.uniform(-0.2, 0.2, ax_nu.size - 2) * (ax_nu[1] - ax_nu[0])
    halos = [(axes[0][2] + 0.01, axes[1][3] + 0.2), (axes[0][0], axes[1][-1]), (axes[0][-1], axes[1][0]),
             (axes[0][-1] + 1e-3, axes[1][1]), (axes[0][3], axes[1][0] - 1e-9), (np.nan, axes[1][2]),
             (axes[0][1], axes[1][2]), (np.nextafter(axes[0][1], -np.inf), np.nextafter(axes[1][2], np.inf))]
    for ax2, search in ((axes[2], False), (axes[2], True), (ax_nu, False)):
        rgi = RGI((axes[0], axes[1], ax2), vals, bounds_error=False, fill_value=np.nan)
        for lnz, lnM in halos:
            with np.errstate(invalid='ignore'):
                want = rgi((np.full_like(x, lnz), np.full_like(x, lnM), x))
            got = host_readout((axes[0], axes[1], ax2), vals, lnz, lnM, x, force_search=search)
            keep = np.ones(x.size, dtype=bool)
            if ax2 is axes[2] and not search:
                # closed-form cell index of a uniform ln r axis: (x - r0) / step rounds either way exactly ON a node, where the
                # search (and scipy's searchsorted) is exact -- the interpolant is continuous there, so only a NaN / -inf entry
                # next to the node can tell (documented boundary tie); nodes are compared on the search path
                with np.errstate(invalid='ignore'):
                    u = (x - ax2[0]) / (ax2[1] - ax2[0])
                    keep = ~(np.abs(u - np.round(u)) < 1e-6)
            assert_close(got[keep], want[keep], f"3-D read-out at ({lnz}, {lnM}), search={search}", rtol=1e-13, atol_scale=1e-14)
    # tables with p_keys axes (4-D .. 6-D), log-valued like TabulatedProfile's (exp applied, -inf = a zero profile)
    extra = [np.linspace(2.0, 12.0, 5), np.linspace(-1.0, 1.0, 3), np.array([0.1, 0.4, 0.5, 2.0])]
    with np.errstate(divide='ignore', invalid='ignore'):
        base = np.log(synth.profile_values(axes))
    for n_extra in (1, 2, 3):
        pv = base
        for e in extra[:n_extra]:
            pv = pv[..., None] + np.log1p(0.05 * (e - e[0]))
        all_axes = (axes[0], axes[1], axes[2]) + tuple(extra[:n_extra])
        rgi = RGI(all_axes, pv, bounds_error=False)
        for ex in ([7.7, 0.3, 0.45][:n_extra], [2.0, -1.0, 0.1][:n_extra], [12.0, 1.0, 2.0][:n_extra], [12.1, 0.0, 0.3][:n_extra]):
            lnz, lnM = axes[0][3] + 0.02, axes[1][4] + 0.3
            pts = (np.full_like(x, lnz), np.full_like(x, lnM), x) + tuple(np.full_like(x, v) for v in ex)
            with np.errstate(invalid='ignore', over='ignore'):
                want = np.exp(rgi(pts))
            got = host_readout(all_axes, pv, lnz, lnM, x, extras=ex, flags=_lib.TABLE_LOG_VALUES, force_search=True)
            assert_close(got, want, f"{3 + n_extra}-D log table at extras {ex} (search)", rtol=1e-12, atol_scale=1e-14)
            got = host_readout(all_axes, pv, lnz, lnM, x, extras=ex, flags=_lib.TABLE_LOG_VALUES)
            with np.errstate(invalid='ignore'):
                u = (x - axes[2][0]) / (axes[2][1] - axes[2][0])
                keep = ~(np.abs(u - np.round(u)) < 1e-6)                          # node ties of the closed-form index, as above
            assert_close(got[keep], want[keep], f"{3 + n_extra}-D log table at extras {ex}", rtol=1e-12, atol_scale=1e-14)
    # the BASELINE table shape (10 x 10 x 500, ln r uniform): the closed-form cell index against the search, node by node
    axes = synth.table_axes()
    vals = synth.displacement_values(axes)
    xs = np.concatenate([axes[2], 0.5 * (axes[2][1:] + axes[2][:-1]), rng.uniform(axes[2][0], axes[2][-1], 50000)])
    a = host_readout(axes, vals, axes[0][4] + 0.01, axes[1][5] + 0.1, xs)
    bb = host_readout(axes, vals, axes[0][4] + 0.01, axes[1][5] + 0.1, xs, force_search=True)
    assert_close(a, bb, "uniform vs search", rtol=1e-11, atol_scale=1e-14)
    want = RGI(axes, vals, bounds_error=False)((np.full_like(xs, axes[0][4] + 0.01), np.full_like(xs, axes[1][5] + 0.1), xs))
    assert_close(bb, want, "search vs scipy", rtol=1e-13, atol_scale=1e-14)


def test_lean_readout_by_squared_radius_matches_scipy():
    """row_at_r2 -- the read-out of the default grid, particle and exact shell loops: value at a SQUARED radius, cell coordinate
    u = log2(r^2) uA + uB with the table-driven log2, v0 + t (v1 - v0) -- from the kernels' own source on the host, against scipy at
    ln r = 0.5 ln(r^2) + offset.  Away from exact nodes (deviation (vi)); r^2 = 0, inf, NaN and radii outside the table -> NaN."""
    axes = synth.table_axes()                                             # the BASELINE table: 10 x 10 x 500, ln r uniform
    vals = synth.displacement_values(axes, inject_nan=True)
    rng = np.random.default_rng(12)
    shape = (C.c_int64 * 3)(*[a.size for a in axes])
    ax = [np.ascontiguousarray(a) for a in axes]
    ptrs = (C.c_void_p * 3)(*[a.ctypes.data for a in ax])
    v = np.ascontiguousarray(vals)
    rgi = RGI(axes, vals, bounds_error=False, fill_value=np.nan)
    for lnz, lnM, offset in ((axes[0][3] + 0.02, axes[1][4] + 0.3, 0.0), (axes[0][0], axes[1][-1], 0.35),
                             (axes[0][1] + 0.01, axes[1][2] + 0.01, -1.7), (axes[0][-1] + 1e-3, axes[1][1], 0.0)):
        lnr = rng.uniform(axes[2][0] - 1.0, axes[2][-1] + 1.0, 100000) - offset
        r2 = np.concatenate([np.exp(2 * lnr), [0.0, np.inf, np.nan, -1.0, 5e-324]])
        out = np.empty_like(r2)
        ok = np.empty(r2.size, dtype=np.int32)
        _lib.check(_lib.lib().bfg_test_row_at_r2_host(3, shape, ptrs, v.ctypes.data, 0, float(lnz), float(lnM), None, float(offset),
                                                      r2.size, r2.ctypes.data, out.ctypes.data, ok.ctypes.data))
        with np.errstate(divide='ignore', invalid='ignore'):
            x = 0.5 * np.log(r2) + offset
            want = rgi((np.full_like(x, lnz), np.full_like(x, lnM), x))
            u = (x - axes[2][0]) / (axes[2][1] - axes[2][0])
            keep = ~(np.abs(u - np.round(u)) < 1e-6)                          # node ties
            edge = (np.abs(x - axes[2][0]) < 1e-9) | (np.abs(x - axes[2][-1]) < 1e-9)
        keep &= ~edge
        # a non-finite node (NaN, +-inf) makes the value non-finite on both sides -- inf + t (v1 - inf) is NaN where scipy's
        # (1 - t) inf + t v1 stays inf -- and every caller turns non-finite into "adds nothing" (HealpixRunner.py:347)
        fin = np.isfinite(want[keep])
        assert np.array_equal(np.isfinite(out[keep]), fin)
        halo_inside = bool(np.isfinite(lnz)) and axes[0][0] <= lnz <= axes[0][-1] and axes[1][0] <= lnM <= axes[1][-1]
        with np.errstate(invalid='ignore'):
            inside = halo_inside & (x[keep] >= axes[2][0]) & (x[keep] <= axes[2][-1])
        assert np.array_equal(ok[keep] == 1, inside)
        assert_close(out[keep][fin], want[keep][fin], f"row_at_r2 at ({lnz}, {lnM}), offset {offset}", rtol=1e-11, atol_scale=1e-14)
        assert np.all(np.isnan(out[-5:])) and not ok[-5:].any()
