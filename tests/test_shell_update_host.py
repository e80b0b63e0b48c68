"""
CPU: the per-pixel update of the shell kernels (shell_update<MODE, UNIFORM> in csrc/shell_kernels.cu with make_upd, table_at, the
row blend and fast_log2 -- the kernels' own source, compiled for the HOST: bfg_test_shell_update_host) against the oracle port of
the reference loop (oracle/runners_port.shell_offsets / paint_shell == HealpixRunner.py:313-355 / :449-481), halo by halo on the
halo's own query_disc pixels: displacement offsets and painted profiles to 1e-9 relative (1e-6 is the contract), zero where the
reference zeroes (beyond the model's cut, outside the table).  The halo records come from the host-compiled scalar prep
(bfg_test_shell_records_host), so the whole per-halo chain  catalogue -> record -> read-out -> update  runs from device source.
tools/sass_fingerprint.py shows no kernel changed when these functions became host-compilable.
"""
import ctypes as C
import warnings

import numpy as np

import baryonforge_b200 as b
from baryonforge_b200 import _lib, synth
from oracle import hpo
from oracle import runners_port as rp
from test_records_host import host_records


def host_update(axes, values, flags, mode, record, vec, force_search=False):
    axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
    values = np.ascontiguousarray(values, dtype=np.float64)
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    n = vec.shape[0]
    out = np.zeros((n, 3) if mode == 0 else (n,), dtype=np.float64)
    shape = (C.c_int64 * len(axes))(*[a.size for a in axes])
    ptrs = (C.c_void_p * len(axes))(*[a.ctypes.data for a in axes])
    rec = np.ascontiguousarray(record, dtype=np.float64)
    _lib.check(_lib.lib().bfg_test_shell_update_host(len(axes), shape, ptrs, values.ctypes.data, int(flags),
                                                     1 if force_search else 0, int(mode), rec.ctypes.data, None, n,
                                                     vec.ctypes.data, out.ctypes.data))
    return out


def _close(got, want, what, rtol=1e-9):
    scale = np.max(np.abs(want)) if want.size else 0.0
    err = np.abs(got - want)
    # 4e-16: the reference adds the round-off residue of normalise(vec * D) - vec (~1e-16 on the unit sphere) where the offset is
    # zero; the kernels add nothing there (deviation (iii))
    assert np.all(err <= rtol * np.abs(want) + 1e-13 * scale + 4e-16), (what, err.max(), scale)


def test_shell_update_source_on_host_matches_the_oracle_port_halo_by_halo():
    nside, n = 512, 60
    ra, dec, M, z = synth.sky_halos(n, seed=77, z=(0.05, 0.6))
    M[:4] = [10 ** 12.0, 10 ** 15.5, 10 ** 11.5, 10 ** 13]                  # table edges in M; one halo outside the table
    dec[4:6] = [89.5, -89.7]                                               # discs over the poles
    axes = synth.table_axes()
    dvals = synth.displacement_values(axes) * 30.0                         # large enough for second-order terms to matter
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=np.ones(12 * nside * nside), cosmo=synth.COSMO)
    model = b.DisplacementModel(axes, dvals, 6, synth.COSMO)               # the model cuts inside the runner's disc
    run = b.BaryonifyShell(cat, shell, 12, model, verbose=False)
    rec, _ = host_records(run, paint=False)                                # device scalar-prep source, on the host
    run.halo_records(paint=False)
    sc = run.last_scalars
    tab = rp.DisplacementTable(axes, dvals, 6)
    n_checked = n_nonzero = 0
    for j in range(n):
        one = {k: cat.cat[k][j:j + 1] for k in ('M', 'z', 'ra', 'dec')}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want_full, n_up = rp.shell_offsets(nside, one, sc["R_run"][j:j + 1], sc["D_A"][j:j + 1], sc["R_model_com"][j:j + 1], 12,
                                               tab, warn=False)
        pix = hpo.query_disc(nside, rec[j, _lib.HS_THETA], rec[j, _lib.HS_PHI], rec[j, _lib.HS_RADIUS])
        if pix.size < 4:                                                   # the < 4-pixel fallback uses the interpolation pixels
            continue
        assert pix.size == n_up
        vec = np.stack(hpo.pix2vec(nside, pix), axis=1)
        for search in (False, True):
            got = host_update(axes, dvals, 0, 0, rec[j], vec, force_search=search)
            _close(got, want_full[pix], f"halo {j} (M = {M[j]:.3g}, z = {z[j]:.3f}), search={search}")
            assert np.abs(want_full[pix][got == 0.0]).max(initial=0.0) < 4e-16     # zero where the reference zeroes (cut, outside)
        n_checked += pix.size
        n_nonzero += int(np.count_nonzero(np.any(want_full[pix] != 0, axis=1)))
    assert n_checked > 20000 and n_nonzero > 5000, (n_checked, n_nonzero)
    # PaintProfilesShell: log-valued profile table, pixel-area scale on and off
    pvals = synth.profile_values(axes)
    pm = b.ProfileModel(axes, pvals * 3.0, pvals)
    with np.errstate(divide='ignore'):
        logv = np.log(pvals)
    for pixsize in (False, True):
        prun = b.PaintProfilesShell(cat, shell, 12, pm, include_pixel_size=pixsize, verbose=False)
        prec, _ = host_records(prun, paint=True)
        prun.halo_records(paint=True)
        psc = prun.last_scalars
        ptab = rp.ProfileTable(axes, None, pvals)                              # the projected (2-D) profile, un-logged
        for j in range(0, n, 3):
            one = {k: cat.cat[k][j:j + 1] for k in ('M', 'z', 'ra', 'dec')}
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want_map, _ = rp.paint_shell(nside, one, psc["R_run"][j:j + 1], psc["D_A"][j:j + 1], 12, ptab, include_pixel_size=pixsize)
            pix = hpo.query_disc(nside, prec[j, _lib.HS_THETA], prec[j, _lib.HS_PHI], prec[j, _lib.HS_RADIUS])
            if pix.size == 0:
                continue
            vec = np.stack(hpo.pix2vec(nside, pix), axis=1)
            got = host_update(axes, logv, _lib.TABLE_LOG_VALUES, 1, prec[j], vec)
            _close(got, want_map[pix], f"paint halo {j}, pixel size {pixsize}")
