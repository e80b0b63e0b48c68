"""
-m gpu parity at the sizes SURVEY.md section 8(d) "parity at scale" names -- the largest the CPU oracle finishes in seconds:
  * PaintProfilesShell at BASELINE configs[1]'s map size (NSIDE = 1024) against the fixture the reference's own code painted;
  * BaryonifyGrid on a 256^3 grid against the oracle port (halo loop + re-binning);
  * BaryonifySnapshot on 10^7 particles against the oracle port (scipy KDTree) + the NGP deposit against np.histogramdd.
Tolerance: relative 1e-6 (fp64 path; BASELINE.json north_star), index / cell assignments bit-exact up to edge ties.
"""
import warnings

import numpy as np
import pytest

from helpers import assert_close, load

pytestmark = pytest.mark.gpu

GRID_AXES = dict(nz=6, nM=10, nr=300, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-2, r_max=2e2)


def test_paint_shell_config2_nside1024_matches_reference_fixture():
    """tests/golden/shell_paint_config2_map.npz: PaintProfilesShell.process() of the UNMODIFIED reference at NSIDE = 1024 with
    10^4 halos of configs[1]'s catalogue (oracle/make_golden.py); every 128th pixel, the map sum and the painted-pixel count."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    g = load("shell_paint_config2_map")
    nside, n, seed, stride = int(g["nside"]), int(g["n"]), int(g["seed"]), int(g["stride"])
    ra, dec, M, z = synth.sky_halos(n, seed=seed)
    axes = synth.table_axes()
    pvals = synth.profile_values(axes)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=np.zeros(12 * nside * nside), cosmo=synth.COSMO)
    run = b.PaintProfilesShell(cat, shell, g["eps_run"], b.ProfileModel(axes, pvals * 3.0, pvals), include_pixel_size=False,
                               verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = run.process()
    assert got.shape == (12 * nside * nside,)
    assert_close(got[::stride], g["out_sub"], "PaintProfilesShell NSIDE=1024 (every 128th pixel)")
    assert np.isclose(got.sum(), float(g["out_sum"]), rtol=1e-9)
    assert int(np.count_nonzero(got)) == int(g["n_painted"])           # the same pixels were painted
    assert run.last_stats["n_updates"] >= int(g["n_painted"])


def test_baryonify_grid_256_cubed_vs_oracle_port():
    """BaryonifyGrid, 256^3 cells, 3000 halos (2.8e7 halo-cell updates): offsets, update count and the re-binned map."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    N, Lbox, n, eps_run = 256, 250.0, 3000, 6
    pos, M = synth.box_halos(n, Lbox, seed=191)
    bins = (np.arange(N) + 0.5) * Lbox / N
    axes = synth.table_axes(**GRID_AXES)
    dvals = synth.displacement_values(axes) * 25.0
    gmap = np.random.default_rng(192).uniform(0, 10, (N, N, N))
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    gm = b.GriddedMap(map=gmap, redshift=0.3, bins=bins, cosmo=synth.COSMO)
    model = b.DisplacementModel(axes, dvals, 4, synth.COSMO)
    run = b.BaryonifyGrid(cat, gm, eps_run, model, verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        off, n_up = run.offsets_on_device()
        got = run.process()
    sc = run.last_scalars
    hc = {k: cat.cat[k].astype('<f4') for k in ("M", "x", "y", "z")}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        off_w, n_w = rp.grid_offsets((N, N, N), bins, hc, 1 / 1.3, sc["R_phys"], sc["R_model_com"], eps_run,
                                     rp.DisplacementTable(axes, dvals, 4), warn=False)
        want = rp.grid_regrid(gmap, off_w)
    assert int(n_up.cpu()[0]) == n_w                                   # the same (halo, cell) pairs
    assert_close(off.cpu().numpy().T, off_w, "BaryonifyGrid 256^3 offsets vs oracle port")
    assert_close(got, want, "BaryonifyGrid 256^3 map vs oracle port")
    assert np.isclose(got.sum(), gmap.sum(), rtol=1e-12)               # Map2DRunner.py:616-619


def test_baryonify_snapshot_1e7_particles_vs_oracle_port():
    """BaryonifySnapshot, 10^7 particles, 1500 halos (5.3e6 pairs): displaced positions, pair count, NGP deposit."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    from scipy.spatial import KDTree
    Lbox, n, npart, eps_run = 200.0, 1500, 10_000_000, 5
    pos, M = synth.box_halos(n, Lbox, seed=291)
    p = np.random.default_rng(292).uniform(0, Lbox, (3, npart))
    axes = synth.table_axes(**GRID_AXES)
    dvals = synth.displacement_values(axes) * 25.0
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    ps = b.ParticleSnapshot(x=p[0], y=p[1], z=p[2], M=np.ones(npart), L=Lbox, redshift=0.3, cosmo=synth.COSMO)
    model = b.DisplacementModel(axes, dvals, 4, synth.COSMO)
    run = b.BaryonifySnapshot(cat, ps, eps_run, model, verbose=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = run.process()
        grid = run.process_to_map(64)
    sc = run.last_scalars
    hc = {k: cat.cat[k].astype('<f4') for k in ("M", "x", "y", "z")}
    tree = KDTree(p.T, boxsize=Lbox, leafsize=64)                      # leafsize only changes the build time
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, n_pairs, _ = rp.baryonify_snapshot(list(p), Lbox, hc, 1 / 1.3, sc["R_phys"], sc["R_model_com"], eps_run,
                                                 rp.DisplacementTable(axes, dvals, 4), tree=tree, warn=False)
    assert run.last_stats["n_pairs"] == n_pairs                        # the same (halo, particle) pairs
    for k, name in enumerate("xyz"):
        disp_w = (want[k] - p[k] + Lbox / 2) % Lbox - Lbox / 2
        disp_g = (out[name] - p[k] + Lbox / 2) % Lbox - Lbox / 2
        assert_close(disp_g, disp_w, f"snapshot 1e7 displacement {name}", atol_scale=1e-7)
        assert_close(out[name], want[k], f"snapshot 1e7 position {name}", rtol=1e-12, atol_scale=1e-12)
    ngp = rp.make_map_ngp(want, np.ones(npart), Lbox, 64)
    assert np.abs(grid - ngp).sum() <= 4                               # a particle within 1e-12 of a cell edge may change cell
    assert grid.sum() == npart
