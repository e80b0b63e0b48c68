"""
-m gpu parity tests: the CUDA path (through the ctypes C ABI and the reference-shaped runner classes) against
  (1) the committed golden fixtures = outputs of the reference's own runner code (tests/golden, oracle/make_golden.py),
  (2) the oracle port on fresh seeded inputs,
to relative 1e-6 (fp64 path; BASELINE.json north_star), index sets bit-exact.
"""
import warnings

import numpy as np
import pytest

from helpers import assert_close, golden_names, load, port_run, product_run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("shell_bary") + golden_names("shell_paint") + golden_names("shell_anis")
                         + golden_names("grid_bary") + golden_names("grid_paint") + golden_names("grid_anis"))
def test_map_runners_match_reference_fixture(name):
    g = load(name)
    got = product_run(g)
    assert_close(got, g["out"], name)
    if "paint" in str(g["kind"]) or "anis" in str(g["kind"]):
        # painted maps span many orders of magnitude and have no cancellation: the absolute floor of the default tolerance
        # (1e-9 of the map's maximum) would hide errors in their faint pixels, so these are also held to a floor of 1e-13
        assert_close(got, g["out"], name + " (faint pixels)", atol_scale=1e-13)


@pytest.mark.parametrize("name", golden_names("shell_bary"))
def test_warp_per_halo_kernel_matches_reference_fixture(name, monkeypatch):
    """k_shell_halos_warp (one warp per small disc; the default for painting, opt-in for BaryonifyShell) against the same fixtures."""
    monkeypatch.setenv("BFG_SHELL_WARP_KERNEL", "1")
    g = load(name)
    assert_close(product_run(g), g["out"], name + " (warp-per-halo kernel)")
    monkeypatch.setenv("BFG_SHELL_WARP_KERNEL", "0")
    assert_close(product_run(g), g["out"], name + " (CTA-per-halo kernel only)")


@pytest.mark.parametrize("name", golden_names("snap"))
def test_snapshot_matches_reference_fixture(name):
    g = load(name)
    got = product_run(g)
    want = [g["out_x"], g["out_y"]] + ([g["out_z"]] if int(g["ndim"]) == 3 else [])
    for k, (a, b) in enumerate(zip(got, want)):
        # positions: relative 1e-6 on the DISPLACEMENT (out - in), i.e. absolute on the position
        src = [g["px"], g["py"], g["pz"]][k]
        L = float(g["L"])
        disp_w = (b - src + L / 2) % L - L / 2
        disp_g = (a - src + L / 2) % L - L / 2
        assert_close(disp_g, disp_w, f"{name} axis {k}", atol_scale=1e-7)
        assert_close(a, b, f"{name} pos axis {k}", rtol=1e-12, atol_scale=1e-12)


@pytest.mark.parametrize("name", golden_names("snap"))
def test_ngp_deposit_matches_histogramdd(name):
    import baryonforge_b200 as b
    g = load(name)
    ndim = int(g["ndim"])
    px = [g["out_x"], g["out_y"]] + ([g["out_z"]] if ndim == 3 else [])
    got = b.runners.deposit_ngp(px, g["pM"], float(g["L"]), 16)
    assert np.array_equal(got, g["ngp"])     # integer multiples of one particle mass: bit-exact assignment
    # edge semantics of np.histogramdd: x == L -> last bin, outside -> dropped, interior edge -> upper bin
    L = 10.0
    x = np.array([0.0, 2.5, 5.0, 10.0, 10.0000001, -1e-9, 7.5 - 1e-15])
    y = np.full_like(x, 1.0)
    want = np.histogramdd(np.vstack([x, y]).T, bins=(np.linspace(0, L, 5),) * 2, weights=np.ones_like(x))[0]
    assert np.array_equal(b.runners.deposit_ngp([x, y], np.ones_like(x), L, 4), want)


@pytest.mark.parametrize("name", golden_names("snap"))
def test_snapshot_process_to_map_matches_process_then_make_map(name):
    """BaryonifySnapshot.process_to_map (displacement + NGP deposit fused over the cell-ordered particles, no un-permute)
    == the reference's process() followed by ParticleSnapshot.make_map, for equal and for per-particle masses."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    g = load(name)
    ndim, L = int(g["ndim"]), float(g["L"])
    cosmo = synth.COSMO
    mc = dict(Omega_m=0.27 + 0.05, Omega_b=0.05, h=0.68, sigma8=0.82, n_s=0.97, w0=-1.0)
    cat = b.HaloNDCatalog(x=g["x"], y=g["y"], z=g["z"] if ndim == 3 else None, M=g["M"], redshift=g["redshift"], cosmo=cosmo)
    model = b.DisplacementModel((g["ax0"], g["ax1"], g["ax2"]), g["values"], g["eps_mod"], mc)
    ps = b.ParticleSnapshot(x=g["px"], y=g["py"], z=g["pz"] if ndim == 3 else None, M=g["pM"], L=L, redshift=g["redshift"],
                            cosmo=cosmo)
    run = b.BaryonifySnapshot(cat, ps, g["eps_run"], model, verbose=False)
    got = run.process_to_map(16)
    assert got.shape == (16,) * ndim
    # equal masses: integer multiples of one particle mass; a particle within 1e-12 of a cell edge may change cell
    assert np.abs(got - g["ngp"]).sum() <= 2 * g["pM"][0] and np.isclose(got.sum(), g["ngp"].sum(), rtol=1e-14)
    out = run.process()
    px = [out["x"], out["y"]] + ([out["z"]] if ndim == 3 else [])
    assert np.array_equal(got, b.runners.deposit_ngp(px, g["pM"], L, 16))       # same positions, same cells
    assert run.last_stats["n_pairs"] > 0
    # per-particle masses (gathered through the sort permutation)
    Mp = np.random.default_rng(5).uniform(0.5, 2.0, g["pM"].size) * 1e10
    ps2 = b.ParticleSnapshot(x=g["px"], y=g["py"], z=g["pz"] if ndim == 3 else None, M=Mp, L=L, redshift=g["redshift"],
                             cosmo=cosmo)
    run2 = b.BaryonifySnapshot(cat, ps2, g["eps_run"], model, verbose=False)
    got2 = run2.process_to_map(16)
    assert_close(got2, b.runners.deposit_ngp(px, Mp, L, 16), name + " per-particle masses", rtol=1e-12)
    ps3 = b.ParticleSnapshot(x=g["px"], y=g["py"], z=g["pz"] if ndim == 3 else None, M=None, L=L, redshift=g["redshift"],
                             cosmo=cosmo)
    with pytest.raises(AssertionError):                                       # utils/io.py:659
        b.BaryonifySnapshot(cat, ps3, g["eps_run"], model, verbose=False).process_to_map(16)


@pytest.mark.parametrize("name", golden_names("snap"))
def test_two_pass_cell_sort_gives_the_same_snapshot(name, monkeypatch):
    """The two-pass counting sort (coarse buckets, then cells; the default above 65536 cells since round 2: 26.2 vs 29.6 ms
    for 2.5e8 particles) is only a different route to the same cell list as the single-pass scatter (BFG_CELL_SORT=1)."""
    g = load(name)
    ncell = 41 if int(g["ndim"]) == 3 else 300                       # > 65536 cells, so the two-pass route is taken
    got = product_run(g, ncell=ncell)
    for mode in ("1", "3"):            # 1 = single-pass scatter, 3 = radix sort of (cell, index) pairs + gather
        monkeypatch.setenv("BFG_CELL_SORT", mode)
        want = product_run(g, ncell=ncell)
        for a, b in zip(got, want):
            assert_close(a, b, f"{name}, BFG_CELL_SORT={mode}", rtol=1e-12, atol_scale=1e-13)   # summation order of the REDs only


def test_kept_cell_list_gives_the_same_snapshots_for_successive_models():
    """keep_cells=True: the cell list survives between process() calls while `Runner.model` changes (the reference builds its
    KD-tree once in __init__ and the notebooks swap models); results equal fresh runners', and earlier results are not clobbered."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    g = load("snap_3d")
    cosmo = synth.COSMO
    mc = dict(Omega_m=0.27 + 0.05, Omega_b=0.05, h=0.68, sigma8=0.82, n_s=0.97, w0=-1.0)
    axes = (g["ax0"], g["ax1"], g["ax2"])
    cat = b.HaloNDCatalog(x=g["x"], y=g["y"], z=g["z"], M=g["M"], redshift=g["redshift"], cosmo=cosmo)
    ps = b.ParticleSnapshot(x=g["px"], y=g["py"], z=g["pz"], M=g["pM"], L=float(g["L"]), redshift=g["redshift"], cosmo=cosmo)
    models = [b.DisplacementModel(axes, g["values"] * f, g["eps_mod"], mc) for f in (1.0, 0.5, -0.7)]
    kept = b.BaryonifySnapshot(cat, ps, g["eps_run"], models[0], verbose=False, keep_cells=True)
    outs, dev_outs = [], []
    for m in models:
        kept.model = m
        dev_outs.append(kept.process_on_device())
        outs.append(kept.process())
        assert kept._cells is not None
    for m, out, d_out in zip(models, outs, dev_outs):
        want = b.BaryonifySnapshot(cat, ps, g["eps_run"], m, verbose=False).process()
        for k, name in enumerate("xyz"):
            assert_close(out[name], want[name], "kept cells " + name, rtol=1e-12, atol_scale=1e-13)
            assert_close(d_out[k].cpu().numpy(), want[name], "earlier device result " + name, rtol=1e-12, atol_scale=1e-13)
    assert np.array_equal(kept.process_to_map(16), b.BaryonifySnapshot(cat, ps, g["eps_run"], models[-1],
                                                                       verbose=False).process_to_map(16))


def test_healpix_device_geometry_matches_oracle():
    from baryonforge_b200 import healpix as dh
    from oracle import hpo
    rng = np.random.default_rng(3)
    for nside in (1, 2, 8, 64, 512):
        npix = 12 * nside * nside
        lo = 0 if nside <= 64 else npix // 2 - 5000
        hi = npix if nside <= 64 else npix // 2 + 5000
        want = np.stack(hpo.pix2vec_range(nside, lo, hi - lo))
        got = dh.pix2vec(nside, lo, hi)
        assert np.max(np.abs(got - want)) < 2e-15      # CUDA vs glibc sincos: a few ulp
        th = np.arccos(rng.uniform(-1, 1, 4000)); ph = rng.uniform(0, 2 * np.pi, 4000)
        th[:4] = [0.0, np.pi, 1e-9, np.pi - 1e-9]
        assert np.array_equal(dh.ang2pix(nside, th, ph), hpo.ang2pix(nside, th, ph))
        gp, gw = dh.interp_weights(nside, th, ph)
        wp, ww = hpo.get_interpol(nside, th, ph)
        # a direction within round-off of a pixel-centre meridian may pick the neighbouring pair with weight ~0
        same = np.all(gp == wp, axis=0)
        assert same.mean() > 0.999
        assert np.max(np.abs(gw[:, same] - ww[:, same])) < 1e-9
        assert np.allclose(gw.sum(axis=0), 1.0, atol=1e-12)


def test_ring_nest_device_matches_oracle():
    from baryonforge_b200 import healpix as dh
    from oracle import hpo
    rng = np.random.default_rng(12)
    for nside in (1, 2, 64, 4096):
        npix = 12 * nside * nside
        r = np.arange(npix) if nside <= 64 else rng.integers(0, npix, 200000)
        n = dh.reorder(nside, r, True)
        assert np.array_equal(n, hpo.ring2nest(nside, r))
        assert np.array_equal(dh.reorder(nside, n, False), r)
        th = np.arccos(rng.uniform(-1, 1, 2000)); ph = rng.uniform(0, 2 * np.pi, 2000)
        assert np.array_equal(dh.ang2pix(nside, th, ph, nest=True), hpo.ring2nest(nside, hpo.ang2pix(nside, th, ph)))


def test_query_disc_index_sets_bit_exact():
    from baryonforge_b200 import healpix as dh
    from oracle import hpo
    rng = np.random.default_rng(5)
    n_diff = 0
    n_tot = 0
    for nside in (1, 4, 32, 256, 4096):
        for k in range(60):
            theta = np.arccos(rng.uniform(-1, 1)); phi = rng.uniform(0, 2 * np.pi)
            if k % 6 == 0:
                theta = rng.choice([1e-3, np.pi - 1e-3, 0.02, np.pi - 0.02])
            rad = 10 ** rng.uniform(-3.5, 0.3) if nside < 4096 else 10 ** rng.uniform(-3.5, -1.8)
            if k == 1:
                rad = 3.2      # >= pi: the whole sphere
            want = hpo.query_disc(nside, theta, phi, rad)
            got = dh.query_disc(nside, theta, phi, rad)
            n_tot += 1
            if not np.array_equal(got, want):
                # documented boundary ties: only pixels whose centre sits within round-off of the disc edge may differ
                d = np.setxor1d(got, want)
                assert d.size <= 2, (nside, k, d.size)
                n_diff += 1
    assert n_diff <= 1, f"{n_diff} of {n_tot} discs differ"


def test_table_readout_matches_scipy():
    import baryonforge_b200 as b
    from baryonforge_b200 import healpix as dh, synth, _lib
    from scipy.interpolate import RegularGridInterpolator as RGI
    import torch
    axes = synth.table_axes(nz=7, nM=9, nr=300)
    vals = synth.displacement_values(axes, inject_nan=True)
    extra = np.linspace(2.0, 12.0, 5)
    vals4 = vals[..., None] * (1 + 0.1 * extra)[None, None, None, :]
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(axes[2][0] - 0.5, axes[2][-1] + 0.5, 3000), axes[2][[0, -1, 17]], [np.nan, -np.inf]])
    dev = torch.cuda.current_device()
    # 3-D, uniform ln r axis (closed-form cell index) and a perturbed axis (search path)
    ax_nu = axes[2].copy(); ax_nu[1:-1] += rng.uniform(-0.2, 0.2, ax_nu.size - 2) * (ax_nu[1] - ax_nu[0])
    for ax2 in (axes[2], ax_nu):
        tab = b.DeviceTable((axes[0], axes[1], ax2), vals, 0, dev)
        rgi = RGI((axes[0], axes[1], ax2), vals, bounds_error=False, fill_value=np.nan)
        for lnz, lnM in [(axes[0][2] + 0.01, axes[1][3] + 0.2), (axes[0][0], axes[1][-1]), (axes[0][-1] + 1e-3, axes[1][1])]:
            want = rgi((np.full_like(x, lnz), np.full_like(x, lnM), x))
            got = dh.table_readout(tab, lnz, lnM, x)
            assert_close(got, want, "readout3d", rtol=1e-12, atol_scale=1e-14)
    # 4-D table with one extra (p_keys) axis, log-valued
    with np.errstate(divide='ignore', invalid='ignore'):
        pv = np.log(synth.profile_values(axes)[..., None] * (1 + 0.1 * extra)[None, None, None, :])
    tab = b.DeviceTable((axes[0], axes[1], axes[2], extra), pv, _lib.TABLE_LOG_VALUES, dev)
    rgi = RGI((axes[0], axes[1], axes[2], extra), pv, bounds_error=False)
    lnz, lnM, e = axes[0][3] + 0.02, axes[1][4] + 0.3, 7.7
    with np.errstate(invalid='ignore', over='ignore'):
        want = np.exp(rgi((np.full_like(x, lnz), np.full_like(x, lnM), x, np.full_like(x, e))))
    got = dh.table_readout(tab, lnz, lnM, x, extras=[e])
    assert_close(got, want, "readout4d", rtol=1e-12, atol_scale=1e-14)


def _fresh_shell_case(nside, n, seed, eps_run, eps_mod):
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    ra, dec, M, z = synth.sky_halos(n, seed=seed)
    axes = synth.table_axes()
    vals = synth.displacement_values(axes)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=synth.shell_map(nside, seed=seed + 1), cosmo=synth.COSMO)
    model = b.DisplacementModel(axes, vals, eps_mod, synth.COSMO)
    return cat, shell, model, axes, vals


def test_baryonify_shell_config1_vs_oracle_port():
    """BASELINE.json configs[0]: NSIDE=256, 10^4 halos, table 10x10x500, epsilon_max=20 -- offsets, update count, map."""
    import baryonforge_b200 as b
    from oracle import runners_port as rp
    nside, n = 256, 10000
    cat, shell, model, axes, vals = _fresh_shell_case(nside, n, 42, 20, 20)
    run = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    rec, _ = run.halo_records(paint=False)
    sc = run.last_scalars
    tab = rp.DisplacementTable(axes, vals, 20)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        off_w, n_w = rp.shell_offsets(nside, cat.cat, sc["R_run"], sc["D_A"], sc["R_model_com"], 20, tab, warn=False)
        map_w = rp.shell_regrid(nside, shell.map, off_w)
    d_off, d_n = run.offsets_on_device()
    assert int(d_n.cpu()[0]) == n_w            # sum_j |pixind_j| : index-set sizes bit-exact
    assert_close(d_off.cpu().numpy().T, off_w, "pix_offsets")
    got = run.process()
    assert run.last_stats["n_updates"] == n_w
    assert_close(got, map_w, "new_map")
    assert np.isclose(got.sum(), shell.map.sum())


def test_shell_invariants():
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    nside = 128
    cat, shell, model, axes, vals = _fresh_shell_case(nside, 2000, 9, 20, 20)
    # zero table => identity map
    zero = b.DisplacementModel(axes, np.zeros_like(vals), 20, synth.COSMO)
    out = b.BaryonifyShell(cat, shell, 20, zero, verbose=False).process()
    assert_close(out, shell.map, "identity", rtol=1e-9, atol_scale=1e-10)   # deg<->rad round trip of the regrid
    # all-zero map is returned as the same object (HealpixRunner.py:293-294)
    zshell = b.LightconeShell(map=np.zeros(12 * nside * nside), cosmo=synth.COSMO)
    assert b.BaryonifyShell(cat, zshell, 20, model, verbose=False).process() is zshell.map
    # empty catalogue
    empty = b.HaloLightConeCatalog(ra=np.zeros(0), dec=np.zeros(0), M=np.zeros(0), z=np.zeros(0), cosmo=synth.COSMO)
    out = b.BaryonifyShell(empty, shell, 20, model, verbose=False).process()
    assert_close(out, shell.map, "empty catalogue", rtol=1e-9, atol_scale=1e-10)
    # pixel-range split: two half-sky runs add up to the full run (what ring-range sharding relies on)
    full, _ = b.BaryonifyShell(cat, shell, 20, model, verbose=False).offsets_on_device()
    npix = 12 * nside * nside
    cut = npix // 2 + 37
    lo, _ = b.BaryonifyShell(cat, shell, 20, model, verbose=False, pix_range=(0, cut)).offsets_on_device()
    hi, _ = b.BaryonifyShell(cat, shell, 20, model, verbose=False, pix_range=(cut, npix)).offsets_on_device()
    import torch
    both = torch.cat([lo, hi], dim=1)
    assert_close(both.cpu().numpy(), full.cpu().numpy(), "range split", rtol=1e-9, atol_scale=1e-12)


def test_error_behaviour_matches_reference():
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    cat, shell, model, axes, vals = _fresh_shell_case(16, 10, 1, 20, 20)
    with pytest.raises(NotImplementedError):
        b.BaryonifyShell(cat, shell, 20, model, use_ellipticity=True)

    class NoTable(object):
        def setup_interpolator(self):
            pass
    with pytest.raises(NameError):
        b.BaryonifyShell(cat, shell, 20, NoTable(), verbose=False).process()
    with pytest.raises(TypeError):
        b.PaintProfilesShell(cat, shell, 20, object(), verbose=False).process()
    with pytest.raises(ValueError):
        b.LightconeShell(map=np.zeros(12), cosmo=dict(Omega_m=0.3))


def test_fast_log2():
    """The table-driven log2 of the pixel loops: absolute error < 2e-15 over the radii range; non-normal -> NaN."""
    import torch
    from baryonforge_b200 import _lib
    rng = np.random.default_rng(0)
    x = np.concatenate([10 ** rng.uniform(-30, 30, 200000), 1 + rng.uniform(-1e-3, 1e-3, 1000), 2.0 ** np.arange(-40, 40),
                        np.nextafter(2.0 ** np.arange(-5, 5), 0), [1e-310, 0.0, np.inf, np.nan, -1.0]])
    d_x = torch.from_numpy(x).cuda()
    d_o = torch.empty_like(d_x)
    _lib.check(_lib.lib().bfg_test_fast_log2(x.size, d_x.data_ptr(), d_o.data_ptr(), _lib.current_stream()))
    got = d_o.cpu().numpy()
    ok = np.isfinite(x) & (x >= 2.3e-308)
    want = np.log2(x[ok].astype(np.longdouble)).astype(np.float64)
    err = np.abs(got[ok] - want)
    assert err.max() < 2e-15 + 2.3e-16 * np.abs(want).max(), err.max()
    assert np.all(np.isnan(got[~ok]))


def test_param_tables_p_keys_vs_oracle_port():
    """4-D tables with a per-halo extra column (`p_keys`, e.g. cdelta): BaryonCorrection.py:211-212,307-322 /
    Tabulate.py:553-590 semantics, shell baryonify + shell paint, CUDA vs the oracle port."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    nside, n = 64, 300
    ra, dec, M, z = synth.sky_halos(n, seed=77, z=(0.05, 0.5))
    rng = np.random.default_rng(78)
    cd = rng.uniform(3.0, 11.0, n)
    cd[:5] = [2.0, 12.5, 3.0, 11.0, 7.0]          # two outside the extra axis -> NaN -> zero; two on its edges
    axes = synth.table_axes(nz=8, nM=9, nr=300)
    cax = np.linspace(3.0, 11.0, 5)
    d4 = synth.displacement_values(axes)[..., None] * (0.5 + 0.1 * cax)[None, None, None, :]
    p4 = synth.profile_values(axes)[..., None] * (0.5 + 0.1 * cax)[None, None, None, :]
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO, cdelta=cd)
    shell = b.LightconeShell(map=synth.shell_map(nside, seed=79), cosmo=synth.COSMO)
    dmodel = b.DisplacementModel(axes + (cax,), d4, 20, synth.COSMO, p_keys=['cdelta'])
    run = b.BaryonifyShell(cat, shell, 20, dmodel, verbose=False)
    d_off, d_n = run.offsets_on_device()
    sc = run.last_scalars
    tab = rp.DisplacementTable(axes + (cax,), d4, 20, p_keys=['cdelta'])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        off_w, n_w = rp.shell_offsets(nside, cat.cat, sc["R_run"], sc["D_A"], sc["R_model_com"], 20, tab,
                                      extras=dict(cdelta=cd), warn=False)
    assert int(d_n.cpu()[0]) == n_w
    assert_close(d_off.cpu().numpy().T, off_w, "4-D displacement offsets")
    pmodel = b.ProfileModel(axes + (cax,), p4 * 3, p4, p_keys=['cdelta'])
    prun = b.PaintProfilesShell(cat, shell, 20, pmodel, verbose=False)
    got = prun.process()
    sc = prun.last_scalars
    ptab = rp.ProfileTable(axes + (cax,), p4 * 3, p4, p_keys=['cdelta'])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, n_w = rp.paint_shell(nside, cat.cat, sc["R_run"], sc["D_A"], 20, ptab, extras=dict(cdelta=cd))
    assert prun.last_stats["n_updates"] == n_w
    assert_close(got, want, "4-D painted map")


def test_full_size_subdomain_parity_nside4096():
    """Parity at BASELINE's full map size: a few hundred halos on the NSIDE=4096 sphere (incl. both poles), CUDA offsets
    vs the oracle port on every touched pixel; update counts bit-exact (SURVEY.md §8d 'parity at scale')."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    nside, n = 4096, 160
    ra, dec, M, z = synth.sky_halos(n, seed=4096)
    dec[:4] = [89.999, -89.9995, 89.9, -89.95]
    M[:4] = 10 ** 14.8
    axes = synth.table_axes()
    vals = synth.displacement_values(axes)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=np.broadcast_to(np.ones(1), (12 * nside * nside,)), cosmo=synth.COSMO)
    model = b.DisplacementModel(axes, vals, 20, synth.COSMO)
    run = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    d_off, d_n = run.offsets_on_device()
    sc = run.last_scalars
    tab = rp.DisplacementTable(axes, vals, 20)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        off_w, n_w = rp.shell_offsets(nside, cat.cat, sc["R_run"], sc["D_A"], sc["R_model_com"], 20, tab, warn=False)
    assert int(d_n.cpu()[0]) == n_w
    touched = np.flatnonzero(np.any(off_w != 0, axis=1))
    import torch
    got = d_off[:, torch.from_numpy(touched).to(d_off.device)].cpu().numpy().T
    assert_close(got, off_w[touched], "NSIDE=4096 offsets on touched pixels")
    # nothing outside the touched set
    assert int((d_off != 0).any(dim=0).sum().item()) <= touched.size


def test_device_scalar_prep_matches_host_formulas():
    """bfg_shell_records (device-side per-halo scalars) vs the numpy restatement of
    HealpixRunner.py:317-329 in halo_records(): D_A bit-identical (scipy PPoly order), logs bit-identical (host numpy),
    radii to 1e-12 (spline of the radius factor), angles to a few ulp (CUDA vs glibc sincos/atan2)."""
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, synth
    cat, shell, model, axes, vals = _fresh_shell_case(64, 5000, 21, 20, 7)
    model.cosmo = dict(synth.COSMO, Omega_m=0.33)            # two cosmologies -> two R200c
    cat.cat['z'][:3] = [0.0, 0.4, 0.5]
    cat.cat['dec'][:4] = [90.0, -90.0, 89.9999, 0.0]
    cat.cat['ra'][:4] = [0.0, 359.9999, 180.0, 0.0]
    pm = b.ProfileModel(axes, synth.profile_values(axes) * 3, synth.profile_values(axes))
    for paint, run in ((False, b.BaryonifyShell(cat, shell, 20, model, verbose=False)),
                       (True, b.PaintProfilesShell(cat, shell, 20, pm, include_pixel_size=True, verbose=False)),
                       (True, b.PaintProfilesShell(cat, shell, 20, pm, include_pixel_size=False, verbose=False))):
        want, _ = run.halo_records(paint=paint)
        sc_w = run.last_scalars
        got = run.device_records(paint).cpu().numpy()
        sc_g = run.last_scalars
        assert np.array_equal(got[:, _lib.HS_D], want[:, _lib.HS_D])            # D_a(z_j): same PPoly, same op order
        assert np.array_equal(got[:, _lib.HS_LNZ], want[:, _lib.HS_LNZ])
        assert np.array_equal(got[:, _lib.HS_LNM], want[:, _lib.HS_LNM])
        assert np.array_equal(got[:, _lib.HS_A], want[:, _lib.HS_A])
        for f in (_lib.HS_RADIUS, _lib.HS_RCUT, _lib.HS_SCALE, _lib.HS_LNRCOM):
            assert np.allclose(got[:, f], want[:, f], rtol=1e-12, atol=1e-13), f
        for f in (_lib.HS_VX, _lib.HS_VY, _lib.HS_VZ, _lib.HS_THETA, _lib.HS_PHI, _lib.HS_THETA_LL, _lib.HS_PHI_LL):
            d = np.abs(got[:, f] - want[:, f])
            if f == _lib.HS_PHI:      # at the poles the azimuth of the vector is round-off dominated
                d = np.minimum(d, 2 * np.pi - d)[4:]
            assert d.max() < 1e-14, (f, d.max())
        assert np.allclose(sc_g["R_run"], sc_w["R_run"], rtol=1e-12) and np.array_equal(sc_g["D_A"], sc_w["D_A"])
        if not paint:
            assert np.allclose(sc_g["R_model_com"], sc_w["R_model_com"], rtol=1e-12)


def test_parallelize_mirrors():
    """SimpleParallel / SplitJoinParallel (utils/Parallelize.py) keep their contracts on the GPU path."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    cat, shell, model, axes, vals = _fresh_shell_case(64, 800, 31, 20, 20)
    pmodel = b.ProfileModel(axes, synth.profile_values(axes) * 3, synth.profile_values(axes))
    paint = b.PaintProfilesShell(cat, shell, 20, pmodel, verbose=False)
    bary = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    outs = b.SimpleParallel([paint, bary, paint]).process()
    assert len(outs) == 3
    assert_close(outs[0], paint.process(), "SimpleParallel[0]", rtol=1e-9, atol_scale=1e-12)
    assert_close(outs[1], bary.process(), "SimpleParallel[1]", rtol=1e-9, atol_scale=1e-12)
    joined = b.SplitJoinParallel(paint, njobs=3).process()
    assert_close(joined, outs[0], "SplitJoinParallel", rtol=1e-9, atol_scale=1e-12)
    with pytest.raises(AssertionError):
        b.SplitJoinParallel(bary, njobs=2)      # Parallelize.py:206-209


def test_grid3d_tile_gather_matches_scatter_and_oracle(monkeypatch):
    """3-D grids whose geometry fits the tiling (N % 16 == 0) run the tile-centric gather kernels (grid_tile_kernels.cu):
    same offsets / painted map as the halo-centric scatter kernels (BFG_GRID_TILES=0) and as the oracle port, same update
    count, independent of the slab split; covers periodic wrap, a halo outside the table (NaN -> cleaned) and the cut."""
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    N, Lbox, n = 64, 100.0, 150
    pos, M = synth.box_halos(n, Lbox, seed=91)
    M[0] = 3e11                                      # outside the table: NaN offsets inside its cut
    pos[:, 1] = 0.01 * Lbox / N                      # cutouts that wrap around the box on every axis
    pos[:, 2] = Lbox * (1 - 1e-3)
    M[3] = 10 ** 15.4                                # a cutout of N/2 cells
    bins = (np.arange(N) + 0.5) * Lbox / N
    gaxes = synth.table_axes(nz=6, nM=10, nr=300, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-2, r_max=2e2)
    dvals = synth.displacement_values(gaxes) * 25.0
    pvals = synth.profile_values(gaxes)
    gmap = np.random.default_rng(92).uniform(0, 10, (N, N, N))
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    gm = b.GriddedMap(map=gmap, redshift=0.3, bins=bins, cosmo=synth.COSMO)
    dmodel = b.DisplacementModel(gaxes, dvals, 4, synth.COSMO)
    pmodel = b.ProfileModel(gaxes, pvals * 3.0, pvals)
    a = 1 / 1.3

    run = b.BaryonifyGrid(cat, gm, 6, dmodel, verbose=False)
    off_t, n_t = run.offsets_on_device()
    sc = run.last_scalars
    monkeypatch.setenv("BFG_GRID_TILES", "0")
    off_s, n_s = b.BaryonifyGrid(cat, gm, 6, dmodel, verbose=False).offsets_on_device()
    monkeypatch.delenv("BFG_GRID_TILES")
    assert int(n_t.cpu()[0]) == int(n_s.cpu()[0])
    assert_close(off_t.cpu().numpy(), off_s.cpu().numpy(), "tile vs scatter offsets", rtol=1e-9, atol_scale=1e-13)
    hc = dict(M=cat.cat['M'].astype('<f4'), x=cat.cat['x'].astype('<f4'), y=cat.cat['y'].astype('<f4'),
              z=cat.cat['z'].astype('<f4'))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        off_w, n_w = rp.grid_offsets((N, N, N), bins, hc, a, sc["R_phys"], sc["R_model_com"], 6,
                                     rp.DisplacementTable(gaxes, dvals, 4), warn=False)
        map_w = rp.grid_regrid(gmap, off_w)
    assert int(n_t.cpu()[0]) == n_w
    assert_close(off_t.cpu().numpy().T, off_w, "tile offsets vs oracle port")
    assert_close(run.process(), map_w, "BaryonifyGrid (tiles) vs oracle port")
    # slab split (plane ranges aligned to the tile height)
    lo, _ = b.BaryonifyGrid(cat, gm, 6, dmodel, verbose=False, plane_range=(0, 24)).offsets_on_device()
    hi, _ = b.BaryonifyGrid(cat, gm, 6, dmodel, verbose=False, plane_range=(24, N)).offsets_on_device()
    both = torch.cat([lo.reshape(3, 24, N * N), hi.reshape(3, N - 24, N * N)], dim=1).reshape(3, -1)
    assert_close(both.cpu().numpy(), off_t.cpu().numpy(), "tile slab split", rtol=1e-9, atol_scale=1e-13)
    # uneven 3-way split: rank 0's slab ends inside a tile (masked planes), the other slabs do not start on a tile boundary
    # and fall back to the scatter kernels -- the pieces still add up to the full result
    parts = []
    for plo, phi_ in ((0, 21), (21, 42), (42, N)):
        o, _ = b.BaryonifyGrid(cat, gm, 6, dmodel, verbose=False, plane_range=(plo, phi_)).offsets_on_device()
        parts.append(o.reshape(3, phi_ - plo, N * N))
    assert_close(torch.cat(parts, dim=1).reshape(3, -1).cpu().numpy(), off_t.cpu().numpy(), "uneven slab split",
                 rtol=1e-9, atol_scale=1e-13)
    # painting (model.real table)
    prun = b.PaintProfilesGrid(cat, gm, 5, pmodel, verbose=False)
    got_p = prun.process()
    monkeypatch.setenv("BFG_GRID_TILES", "0")
    got_ps = b.PaintProfilesGrid(cat, gm, 5, pmodel, verbose=False).process()
    monkeypatch.delenv("BFG_GRID_TILES")
    assert_close(got_p, got_ps, "tile vs scatter paint", rtol=1e-9, atol_scale=1e-13)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want_p, n_wp = rp.paint_grid((N, N, N), bins, hc, a, prun.last_scalars["R_phys"] / a, 5,
                                     rp.ProfileTable(gaxes, pvals * 3.0, pvals))
    assert prun.last_stats["n_updates"] == n_wp
    assert_close(got_p, want_p, "PaintProfilesGrid (tiles) vs oracle port")


def test_pipelined_process_equals_plain_process(monkeypatch):
    """The latitude-chunked end-to-end path of BaryonifyShell.process (halo loop, re-binning and download of finished rings
    overlapped) returns the same map as the plain path and as the oracle port -- including its fallback when the re-binning
    moves mass further than the assumed margin."""
    import baryonforge_b200 as b
    from oracle import runners_port as rp
    nside, n = 128, 3000
    cat, shell, model, axes, vals = _fresh_shell_case(nside, n, 55, 20, 20)
    monkeypatch.setenv("BFG_PIPELINE", "0")
    plain = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    want = plain.process()
    assert not plain.last_stats.get("pipelined", False)
    monkeypatch.setenv("BFG_PIPELINE", "1")
    run = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    run.PIPELINE_MIN_HALOS = 0
    got = run.process()
    assert run.last_stats["pipelined"] and run.last_stats["n_updates"] == plain.last_stats["n_updates"]
    assert 0 < run.last_stats["max_offset"] < run.PIPELINE_MARGIN_RAD
    assert_close(got, want, "pipelined vs plain", rtol=1e-9, atol_scale=1e-12)
    sc = run.last_scalars
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        map_w = rp.baryonify_shell(nside, shell.map, cat.cat, sc["R_run"], sc["D_A"], sc["R_model_com"], 20,
                                   rp.DisplacementTable(axes, vals, 20), warn=False)
    assert_close(got, map_w, "pipelined vs oracle port")
    # margin violated on purpose -> the whole map is downloaded again at the end
    fb = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    fb.PIPELINE_MIN_HALOS = 0
    fb.PIPELINE_MARGIN_RAD = 1e-12
    assert_close(fb.process(), want, "pipelined fallback", rtol=1e-9, atol_scale=1e-12)
    # few chunks / many chunks
    for K in (2, 19):
        r2 = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
        r2.PIPELINE_MIN_HALOS = 0
        r2.PIPELINE_CHUNKS = K
        assert_close(r2.process(), want, f"pipelined K={K}", rtol=1e-9, atol_scale=1e-12)


def test_full_size_properties_nside4096(monkeypatch):
    """BASELINE's map size (NSIDE = 4096, 2.0e8 pixels) with 3e5 halos -- large enough that process() takes the pipelined
    path on its own.  Size-independent properties (SURVEY.md §8d 'parity at scale'): mass conservation (the reference's own
    assert), pipelined == plain path, update count == sum of device disc sizes, linearity in the input map, and a zero
    table leaves the map unchanged."""
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import healpix as dh, synth
    nside, n = 4096, 300000
    ra, dec, M, z = synth.sky_halos(n, seed=77)
    axes = synth.table_axes()
    vals = synth.displacement_values(axes)
    npix = 12 * nside * nside
    hmap = synth.shell_map(nside, seed=78)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=hmap, cosmo=synth.COSMO)
    model = b.DisplacementModel(axes, vals, 20, synth.COSMO)
    run = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    got = run.process()
    assert run.last_stats["pipelined"]
    assert np.isclose(run.last_stats["new_sum"], run.last_stats["old_sum"], rtol=1e-12)
    n_up = run.last_stats["n_updates"]
    # update count == sum_j |query_disc_j| (discs this large never take the < 4-pixel fallback)
    rec = run.device_records(False)
    counts = torch.empty(n, dtype=torch.int64, device=rec.device)
    from baryonforge_b200 import _lib
    _lib.check(_lib.lib().bfg_healpix_disc_counts(nside, n, rec.data_ptr(), counts.data_ptr(), _lib.current_stream()))
    assert int(counts.sum().item()) == n_up and int(counts.min().item()) >= 4
    monkeypatch.setenv("BFG_PIPELINE", "0")
    plain_run = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    plain = plain_run.process()
    monkeypatch.delenv("BFG_PIPELINE")
    assert plain_run.last_stats["n_updates"] == n_up
    assert_close(got, plain, "pipelined vs plain at NSIDE=4096", rtol=1e-9, atol_scale=1e-12)
    del plain
    # linearity in the map: the offsets do not depend on it
    shell3 = b.LightconeShell(map=3.0 * hmap, cosmo=synth.COSMO)
    got3 = b.BaryonifyShell(cat, shell3, 20, model, verbose=False).process()
    assert_close(got3, 3.0 * got, "linearity", rtol=1e-9, atol_scale=1e-12)
    del got3
    zero = b.DisplacementModel(axes, np.zeros_like(vals), 20, synth.COSMO)
    ident = b.BaryonifyShell(cat, shell, 20, zero, verbose=False).process()
    # not bit-exact in the reference either: the re-binning goes radians -> degrees -> radians (HealpixRunner.py:358,361),
    # which at 0.86-arcmin pixels leaves ~1e-9 of a pixel's mass on its neighbours
    assert_close(ident, hmap, "zero table = identity", rtol=1e-6, atol_scale=1e-8)


def test_grid_param_tables_p_keys_vs_oracle_port():
    """4-D tables with a per-halo extra column on grids: 3-D (tile-centric gather) and 2-D (scatter kernel) BaryonifyGrid and
    3-D PaintProfilesGrid against the oracle port (BaryonCorrection.py:211-212,307-322 / Tabulate.py:553-590)."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    rng = np.random.default_rng(5)
    gaxes = synth.table_axes(nz=6, nM=9, nr=200, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-2, r_max=2e2)
    cax = np.linspace(3.0, 11.0, 5)
    d4 = synth.displacement_values(gaxes)[..., None] * 25.0 * (0.5 + 0.1 * cax)[None, None, None, :]
    p4 = synth.profile_values(gaxes)[..., None] * (0.5 + 0.1 * cax)[None, None, None, :]
    a = 1 / 1.3
    for ndim, N, Lbox, n in ((3, 32, 60.0, 40), (2, 64, 120.0, 60)):
        pos, M = synth.box_halos(n, Lbox, seed=60 + ndim, ndim=ndim)
        cd = rng.uniform(3.0, 11.0, n)
        cd[:3] = [2.0, 11.0, 3.0]                     # one outside the extra axis (NaN -> cleaned), two on its edges
        bins = (np.arange(N) + 0.5) * Lbox / N
        gmap = rng.uniform(0, 10, (N,) * ndim)
        cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2] if ndim == 3 else None, M=M, redshift=0.3, cosmo=synth.COSMO,
                              cdelta=cd)
        gm = b.GriddedMap(map=gmap, redshift=0.3, bins=bins, cosmo=synth.COSMO)
        hc = {k: cat.cat[k].astype('<f4') for k in ('M', 'x', 'y', 'z')}
        ex = dict(cdelta=cat.cat['cdelta'].astype('<f4'))
        run = b.BaryonifyGrid(cat, gm, 5, b.DisplacementModel(gaxes + (cax,), d4, 4, synth.COSMO, p_keys=['cdelta']),
                              verbose=False)
        got = run.process()
        sc = run.last_scalars
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = rp.baryonify_grid(gmap, bins, hc, a, sc["R_phys"], sc["R_model_com"], 5,
                                     rp.DisplacementTable(gaxes + (cax,), d4, 4, p_keys=['cdelta']), extras=ex, warn=False)
        assert_close(got, want, f"{ndim}-D BaryonifyGrid with p_keys")
        if ndim == 3:
            prun = b.PaintProfilesGrid(cat, gm, 4, b.ProfileModel(gaxes + (cax,), p4 * 3, p4, p_keys=['cdelta']), verbose=False)
            gotp = prun.process()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                wantp, _ = rp.paint_grid((N,) * 3, bins, hc, a, prun.last_scalars["R_phys"] / a, 4,
                                         rp.ProfileTable(gaxes + (cax,), p4 * 3, p4, p_keys=['cdelta']), extras=ex)
            assert_close(gotp, wantp, "3-D PaintProfilesGrid with p_keys")


def test_snapshot_particle_on_halo_centre_becomes_nan():
    """a particle sitting exactly on a (float32-rounded) halo centre has x_hat = 0/0; the reference
    zeroes the non-finite displacement but not the direction, so the particle comes back with NaN coordinates
    (SnapshotRunner.py:252-260).  The CUDA path reproduces that, and everything else matches the oracle port."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    Lbox, n, n_part = 60.0, 30, 4000
    pos, M = synth.box_halos(n, Lbox, seed=81)
    pos32 = pos.astype('f4').astype('f8')
    rng = np.random.default_rng(82)
    p = rng.uniform(0, Lbox, (3, n_part))
    p[:, 0] = pos32[:, 5]                          # exactly on halo 5
    p[:, 1] = pos32[:, 9]
    gaxes = synth.table_axes(nz=6, nM=10, nr=300, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-2, r_max=2e2)
    dvals = synth.displacement_values(gaxes) * 25.0
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    ps = b.ParticleSnapshot(x=p[0], y=p[1], z=p[2], M=np.ones(n_part), L=Lbox, redshift=0.3, cosmo=synth.COSMO)
    run = b.BaryonifySnapshot(cat, ps, 4, b.DisplacementModel(gaxes, dvals, 5, synth.COSMO), verbose=False)
    out = run.process()
    sc = run.last_scalars
    hc = {k: cat.cat[k].astype('<f4') for k in ('M', 'x', 'y', 'z')}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want, n_pairs, _ = rp.baryonify_snapshot([p[0], p[1], p[2]], Lbox, hc, 1 / 1.3, sc["R_phys"], sc["R_model_com"], 4,
                                                 rp.DisplacementTable(gaxes, dvals, 5), warn=False)
    assert run.last_stats["n_pairs"] == n_pairs
    for k, name in enumerate(("x", "y", "z")):
        assert np.isnan(want[k][0]) and np.isnan(want[k][1])            # the reference's own behaviour
        assert np.array_equal(np.isnan(out[name]), np.isnan(want[k]))
        ok = ~np.isnan(want[k])
        assert np.max(np.abs(out[name][ok] - want[k][ok])) < 1e-9


def test_box_device_records_match_host_formulas():
    """bfg_box_records (device-side per-halo scalars of the grid / snapshot runners) vs the numpy restatement of
    Map2DRunner.py:484-520 / SnapshotRunner.py:219-228 in halo_records(): centre cells, cutout sizes, float32 ln M and
    positions bit-identical, radii to 1e-13 (cbrt vs pow)."""
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, synth
    from baryonforge_b200.runners import _box_device_records
    dev = torch.device('cuda', torch.cuda.current_device())
    gaxes = synth.table_axes(nz=6, nM=10, nr=300, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-2, r_max=2e2)
    dm = b.DisplacementModel(gaxes, synth.displacement_values(gaxes), 4, dict(synth.COSMO, Omega_m=0.33))
    pm = b.ProfileModel(gaxes, synth.profile_values(gaxes) * 3, synth.profile_values(gaxes))
    exact = (_lib.HB_X, _lib.HB_Y, _lib.HB_Z, _lib.HB_NSIZE, _lib.HB_CX, _lib.HB_CY, _lib.HB_CZ, _lib.HB_LNZ, _lib.HB_LNM,
             _lib.HB_DX, _lib.HB_DY, _lib.HB_DZ)
    close = (_lib.HB_RQ, _lib.HB_RCUT, _lib.HB_LNRCOM, _lib.HB_PAINTCUT)
    for ndim, N, Lbox in ((3, 64, 200.0), (2, 128, 200.0)):
        n = 20000
        pos, M = synth.box_halos(n, Lbox, seed=4, ndim=ndim)
        pos[:, 0] = 0.0; pos[:, 1] = Lbox * (1 - 1e-7)         # box edges: centre cells 0 and N - 1
        M[2], M[3] = 1e8, 1e17                                 # cutout sizes clipped to 2 and to N / 2
        bins = (np.arange(N) + 0.5) * Lbox / N
        pos[0, 4] = bins[10] + 0.5 * (bins[1] - bins[0])       # (as close as float32 allows to) a tie between two cells
        cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2] if ndim == 3 else None, M=M, redshift=0.3, cosmo=synth.COSMO)
        gm = b.GriddedMap(map=np.zeros((N,) * ndim), redshift=0.3, bins=bins, cosmo=synth.COSMO)
        for paint, model, cls in ((False, dm, b.BaryonifyGrid), (True, pm, b.PaintProfilesGrid)):
            run = cls(cat, gm, 6, model, verbose=False)
            want, _ = run.halo_records(paint=paint)
            sc_w = run.last_scalars
            d_rec, d_aux = _box_device_records(run, cat.cat, 0.3, ndim, paint, True, np.max(bins) / 2, dev, bins=bins,
                                               res=gm.res)
            got, aux = d_rec.cpu().numpy(), d_aux.cpu().numpy()
            for f in exact:
                assert np.array_equal(got[:, f], want[:, f]), (ndim, paint, f)
            for f in close:
                assert np.allclose(got[:, f], want[:, f], rtol=1e-13, atol=1e-14, equal_nan=True), (ndim, paint, f)   # ln R ~ 0
            assert np.allclose(aux[0], sc_w["R_phys"], rtol=1e-13)
            if not paint:
                assert np.allclose(aux[1], sc_w["R_model_com"], rtol=1e-13)
        ps = b.ParticleSnapshot(x=np.zeros(1), y=np.zeros(1), z=np.zeros(1) if ndim == 3 else None, M=1.0, L=Lbox,
                                redshift=0.3, cosmo=synth.COSMO)
        srun = b.BaryonifySnapshot(cat, ps, 5, dm, verbose=False)
        want, _ = srun.halo_records()
        d_rec, _ = _box_device_records(srun, cat.cat, 0.3, ndim, False, False, Lbox / 2, dev)
        got = d_rec.cpu().numpy()
        for f in (_lib.HB_X, _lib.HB_Y, _lib.HB_Z, _lib.HB_LNZ, _lib.HB_LNM):
            assert np.array_equal(got[:, f], want[:, f]), ("snap", ndim, f)
        for f in (_lib.HB_RQ, _lib.HB_RCUT, _lib.HB_LNRCOM):
            assert np.allclose(got[:, f], want[:, f], rtol=1e-13, atol=1e-14), ("snap", ndim, f)


def test_snapshot_raw_record_path_matches_the_per_field_path(monkeypatch):
    """BaryonifySnapshot.process() moves the reference's particle container (one structured array of 32-byte M, x, y, z
    records, utils/io.py:588) over the host link as raw bytes and replaces x, y(, z) inside the records on the device
    (bfg_snap_build_cells_strided + bfg_snap_apply_records).  Same answer as the per-field staging (BFG_SNAP_RAW=0), with the
    other fields passed through untouched; several staging chunks; pinned and pageable results; other record layouts."""
    import baryonforge_b200 as b
    from baryonforge_b200 import runners, synth
    n, Lbox = 300000, 60.0
    rng = np.random.default_rng(77)
    p = rng.uniform(0, Lbox, (3, n))
    Mp = rng.uniform(0.5, 2.0, n)
    pos, M = synth.box_halos(40, Lbox, seed=78)
    gaxes = synth.table_axes(nz=10, nM=10, nr=500, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-3, r_max=3e2)
    model = b.DisplacementModel(gaxes, synth.displacement_values(gaxes), 5.0, synth.COSMO)
    monkeypatch.setattr(runners, "_RAW_CHUNK", 100003)                   # 12 ragged chunks through the 3-buffer ring
    for ndim in (3, 2):
        cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2] if ndim == 3 else None, M=M, redshift=0.3, cosmo=synth.COSMO)
        ps = b.ParticleSnapshot(x=p[0], y=p[1], z=p[2] if ndim == 3 else None, M=Mp, L=Lbox, redshift=0.3, cosmo=synth.COSMO)
        assert runners._record_layout(ps.cat, ['x', 'y']) == dict(M=0, x=1, y=2, z=3)
        run = b.BaryonifySnapshot(cat, ps, 5.0, model, verbose=False)
        monkeypatch.setenv("BFG_SNAP_RAW", "0")
        want = run.process()
        want_map = run.process_to_map(32)
        monkeypatch.setenv("BFG_SNAP_RAW", "1")
        got = run.process()
        assert got.dtype == ps.cat.dtype and got.shape == ps.cat.shape
        assert np.abs(want["x"] - ps.cat["x"]).max() > 1e-6              # something moved
        for nm in ("x", "y", "z")[:ndim]:
            assert_close(got[nm], want[nm], f"raw records, {ndim}-D, {nm}", rtol=1e-12, atol_scale=1e-13)
        assert np.array_equal(got["M"], Mp) and (ndim == 3 or np.array_equal(got["z"], ps.cat["z"]))
        assert_close(run.process_to_map(32), want_map, f"raw records, {ndim}-D, process_to_map", rtol=1e-12)
        # a second result does not clobber the first (each is its own recycled page-locked buffer) ...
        keep = got.copy()
        run.model = b.DisplacementModel(gaxes, synth.displacement_values(gaxes) * -0.5, 5.0, synth.COSMO)
        other = run.process()
        assert np.array_equal(got, keep) and not np.array_equal(other["x"], got["x"])
        # ... and results above BFG_PINNED_RESULT_MAX_GB come back through the staging ring into ordinary memory
        monkeypatch.setattr(runners, "_RAW_RESULT_MAX_BYTES", 0)
        again = run.process()
        for nm in ps.cat.dtype.names:
            assert_close(again[nm], other[nm], f"pageable result, {nm}", rtol=1e-12, atol_scale=1e-13)
        monkeypatch.setattr(runners, "_RAW_RESULT_MAX_BYTES", 1 << 40)
        # the same particles with the fields in another order take the raw path too; float32 masses do not (per-field path)
        for dt in ([('x', 'f8'), ('z', 'f8'), ('M', 'f8'), ('y', 'f8')], [('M', 'f4'), ('x', 'f8'), ('y', 'f8'), ('z', 'f8')]):
            alt = np.zeros(n, dt)
            for nm in alt.dtype.names:
                alt[nm] = ps.cat[nm]
            ps_alt = b.ParticleSnapshot(x=p[0], y=p[1], z=p[2] if ndim == 3 else None, M=Mp, L=Lbox, redshift=0.3,
                                        cosmo=synth.COSMO)
            ps_alt.cat = alt
            assert (runners._record_layout(alt, ['x', 'y']) is not None) == (dt[0][0] == 'x')
            out = b.BaryonifySnapshot(cat, ps_alt, 5.0, run.model, verbose=False).process()
            assert out.dtype == alt.dtype
            for nm in ("x", "y", "z")[:ndim]:
                assert_close(out[nm], other[nm], f"layout {dt[0][0]}..., {nm}", rtol=1e-12, atol_scale=1e-13)
            assert np.array_equal(out["M"], alt["M"])


@pytest.mark.parametrize("nside", [16, 64, 256])
def test_shell_regrid_small_angle_path_matches_literal_chain_and_oracle(nside, monkeypatch):
    """k_shell_regrid's default path (regrid_target_fast: small-angle azimuth / colatitude, ring table) against the literal
    acos / atan2 / degrees chain on the device (BFG_REGRID_LITERAL=1) and against the oracle port's shell_regrid: pixel-sized,
    sub-pixel and zero offsets, a few huge ones and the polar pixels (which fall back to the literal chain), sparse map.
    Tolerance 1e-7 relative: next to the poles the LITERAL chain's arccos(z / |v|) carries ~1e-9 of round-off in a weight."""
    import torch
    from baryonforge_b200 import _lib
    from oracle import runners_port as rp
    npix = 12 * nside * nside
    rng = np.random.default_rng(400 + nside)
    pixsize = np.sqrt(4 * np.pi / npix)
    off = rng.normal(0, 1.0, (3, npix)) * pixsize * rng.choice([0.0, 0.01, 0.3, 2.5], npix)[None, :]
    big = rng.choice(npix, 50, replace=False)
    off[:, big] = rng.normal(0, 0.3, (3, 50))                                 # far beyond the small-angle guards
    m = rng.uniform(0, 10, npix) * (rng.random(npix) < 0.9)
    d_map, d_off = torch.from_numpy(m).cuda(), torch.from_numpy(np.ascontiguousarray(off)).cuda()
    outs = {}
    for literal in ("1", "0"):
        monkeypatch.setenv("BFG_REGRID_LITERAL", literal)
        d_new = torch.zeros(npix, dtype=torch.float64, device="cuda")
        _lib.check(_lib.lib().bfg_shell_regrid(nside, d_map.data_ptr(), d_off.data_ptr(), d_new.data_ptr(), 0, npix,
                                               _lib.current_stream()))
        outs[literal] = d_new.cpu().numpy()
        # the range form used by the pipelined end-to-end path: two source ranges into the same output
        d_new2 = torch.zeros(npix, dtype=torch.float64, device="cuda")
        cut = (npix // 3) & ~3
        for a, b in ((0, cut), (cut, npix)):
            _lib.check(_lib.lib().bfg_shell_regrid_range(nside, d_map.data_ptr(), d_off.data_ptr(), npix, d_new2.data_ptr(), a, b,
                                                         _lib.current_stream()))
        assert_close(d_new2.cpu().numpy(), outs[literal], f"regrid_range, literal={literal}", rtol=1e-12, atol_scale=1e-14)
    want = rp.shell_regrid(nside, m, np.ascontiguousarray(off.T))
    assert_close(outs["1"], want, f"literal chain vs oracle, NSIDE={nside}", rtol=1e-7, atol_scale=1e-10)
    assert_close(outs["0"], want, f"small-angle path vs oracle, NSIDE={nside}", rtol=1e-7, atol_scale=1e-10)
    assert_close(outs["0"], outs["1"], f"small-angle path vs literal chain, NSIDE={nside}", rtol=1e-7, atol_scale=1e-10)
    assert np.isclose(outs["0"].sum(), m.sum(), rtol=1e-12)                   # mass conservation (HealpixRunner.py:368-370)
