"""CPU: host-side mirror of the reference interface -- containers, per-halo scalar prep, sharding helpers."""
import os
import sys

import numpy as np
import pytest

import baryonforge_b200 as b
from baryonforge_b200 import _lib, cosmology, parallel, synth
from baryonforge_b200.runners import _nearest_bin
from oracle import hpo


def test_containers_mirror_reference_layouts():
    ra, dec, M, z = synth.sky_halos(50)
    dec[0] = 90.0
    with pytest.warns(UserWarning):
        cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO, cdelta=np.full(50, 7.0))
    assert cat.cat.dtype.names == ('M', 'z', 'ra', 'dec', 'cdelta') and cat.cat['M'].dtype == np.float64
    assert cat.cat['dec'][0] == 90 - 1e-8                      # io.py:65-68
    assert len(cat[:10].cat) == 10 and cat[:10].cat['cdelta'][3] == 7.0
    pos, Mb = synth.box_halos(20, 100.0)
    nd = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=Mb, redshift=0.2, cosmo=synth.COSMO)
    assert nd.cat.dtype['M'].str == '>f4'                      # io.py:204-205 big-endian float32
    with pytest.raises(ValueError):
        b.HaloNDCatalog(x=pos[0], y=pos[1], M=Mb, redshift=0.2, cosmo=dict(h=0.7))
    sh = b.LightconeShell(map=np.zeros(12 * 8 * 8), cosmo=synth.COSMO)
    assert sh.NSIDE == 8
    with pytest.raises(ValueError):
        b.LightconeShell(cosmo=synth.COSMO)
    N = 16
    bins = (np.arange(N) + 0.5) * 100 / N
    gm = b.GriddedMap(map=np.zeros((N, N, N)), redshift=0, bins=bins, cosmo=synth.COSMO)
    assert (gm.Npix, gm.is2D) == (N, False) and np.isclose(gm.L, 100.0) and np.isclose(gm.res, 100 / N)
    assert gm.inds.shape == (N, N, N) and len(gm.grid) == 3    # built lazily, same content as io.py:463-470
    ps = b.ParticleSnapshot(x=pos[0], y=pos[1], M=1.0, L=100.0, redshift=0, cosmo=synth.COSMO)
    assert ps.is2D and ps.cat.dtype['x'] == np.float64


def test_nearest_bin_equals_argmin():
    rng = np.random.default_rng(0)
    N = 64
    bins = (np.arange(N) + 0.5) * 200.0 / N
    x = np.concatenate([rng.uniform(-5, 205, 5000), bins, bins + 200.0 / N / 2, [0.0, 200.0]]).astype('f4').astype('f8')
    want = np.array([np.argmin(np.abs(bins - v)) for v in x])
    assert np.array_equal(_nearest_bin(bins, x), want)


def test_background_fallback_matches_pyccl_shim():
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims")
    sys.path.insert(0, shim)
    try:
        import pyccl as ccl
    finally:
        sys.path.remove(shim)
    c = ccl.Cosmology(Omega_c=0.26, Omega_b=0.04, h=0.7, sigma8=0.8, n_s=0.96, w0=-1.0)
    bg = cosmology.Background(0.30, 0.04, 0.7, -1.0)
    a = 1 / (1 + np.linspace(0.0, 3.0, 40))
    assert np.allclose(bg.angular_diameter_distance(a), ccl.angular_diameter_distance(c, a), rtol=1e-9)
    M = np.geomspace(1e11, 1e16, 9)
    want = np.array([ccl.halos.massdef.MassDef(200, 'critical').get_radius(c, m, 0.7) for m in M])
    assert np.allclose(cosmology.radius_of_mass(bg, M, 0.7), want, rtol=1e-12)


def test_shell_records_follow_the_reference_formulas():
    ra, dec, M, z = synth.sky_halos(200)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=np.ones(12 * 16 * 16), cosmo=synth.COSMO)
    axes = synth.table_axes()
    model = b.DisplacementModel(axes, synth.displacement_values(axes), 7, dict(synth.COSMO, Omega_m=0.33))
    run = b.BaryonifyShell(cat, shell, 20, model, verbose=False)
    rec, ex = run.halo_records(paint=False)
    assert ex is None and rec.shape == (200, _lib.HALO_STRIDE)
    a = 1 / (1 + z)
    assert np.array_equal(rec[:, _lib.HS_LNZ], np.log(1 / a)) and np.array_equal(rec[:, _lib.HS_LNM], np.log(M))
    theta, phi = np.pi / 2 - np.radians(dec), np.radians(ra)
    v = np.array([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)]).T
    assert np.array_equal(rec[:, :3], v)
    for j in (0, 17, 199):      # pointing(vec) as the healpy wrapper builds it
        t, p = hpo.vec2pointing(v[j])
        assert rec[j, _lib.HS_THETA] == t and rec[j, _lib.HS_PHI] == p
    sc = run.last_scalars
    assert np.array_equal(rec[:, _lib.HS_RADIUS], sc["R_run"] * 20 / sc["D_A"])
    assert np.array_equal(rec[:, _lib.HS_RCUT], 7 * sc["R_model_com"])
    assert not np.allclose(sc["R_model_com"] * a, sc["R_run"])     # two cosmologies -> two R200c (§10 #6)


def test_pixel_ranges_and_halo_assignment_cover_everything():
    nside = 32
    npix = 12 * nside * nside
    for world in (1, 2, 3, 8):
        rng_ = parallel.pixel_ranges(nside, world)
        assert rng_[0][0] == 0 and rng_[-1][1] == npix and all(a[1] == b_[0] for a, b_ in zip(rng_, rng_[1:]))
    pix = np.arange(npix)
    t, _ = hpo.pix2ang(nside, pix)
    ring = parallel.ring_of_pixel(nside, pix)
    z_ring = np.array([hpo.lib().hpo_ring2z(nside, int(r)) for r in range(1, 4 * nside)])
    assert np.allclose(np.cos(t), z_ring[ring - 1], atol=1e-12)
    # every (halo, pixel) pair of the full run is owned by a rank that was given the halo
    rng = np.random.default_rng(2)
    theta = np.arccos(rng.uniform(-1, 1, 300)); phi = rng.uniform(0, 2 * np.pi, 300)
    rad = 10 ** rng.uniform(-2.5, -0.3, 300)
    world = 4
    masks = [parallel.halos_touching_pixel_range(nside, theta, rad, lo, hi) for lo, hi in parallel.pixel_ranges(nside, world)]
    for j in range(300):
        p = hpo.query_disc(nside, theta[j], phi[j], rad[j])
        if p.size < 4:
            p = np.union1d(p, hpo.get_interpol(nside, [theta[j]], [phi[j]])[0][:, 0])
        for r, (lo, hi) in enumerate(parallel.pixel_ranges(nside, world)):
            if np.any((p >= lo) & (p < hi)):
                assert masks[r][j], (j, r)


def test_plane_assignment():
    N = 64
    rng = np.random.default_rng(3)
    cen = rng.integers(0, N, 500); ns = 2 * rng.integers(1, N // 4 + 1, 500)
    for lo, hi in parallel.plane_ranges(N, 4):
        m = parallel.halos_touching_planes(N, cen, ns, lo, hi)
        for j in range(500):
            planes = (np.arange(cen[j] - ns[j] // 2, cen[j] + ns[j] // 2)) % N
            assert m[j] == bool(np.any((planes >= lo) & (planes < hi)))


def test_radius_factor_spline_reproduces_get_radius():
    """The spline the device record kernel evaluates: R_delta(M, a) = cbrt(M) * g(ln(1+z)) to < 1e-12 (z_max = 30)."""
    bg = cosmology.runner_cosmology(synth.COSMO, with_w0=True)
    rng = np.random.default_rng(4)
    for z_max in (0.5, 3.0, 30.0):
        g = cosmology.radius_factor_spline(bg, None, z_max)
        z = np.concatenate([rng.uniform(0, z_max, 20000), [0.0, z_max]])
        M = 10 ** rng.uniform(11, 16, z.size)
        a = 1 / (1 + z)
        want = cosmology.radius_of_mass(bg, M, a)
        got = np.cbrt(M) * g(np.log(1 / a))
        assert np.max(np.abs(got / want - 1)) < 2e-12
    # matter density helper of the anisotropic painters (HealpixRunner.py:580 / Map2DRunner.py:886)
    assert np.isclose(cosmology.rho_matter(bg, 0.5, is_comoving=True), cosmology.rho_matter(bg, 1.0))
    assert np.isclose(cosmology.rho_matter(bg, 0.5), 8 * cosmology.rho_matter(bg, 1.0))


def test_get_parameter_and_runner_pickling():
    """utils/Tabulate.py:66-96 `_get_parameter` on the stand-in models; runners stay picklable (Parallelize.py:47-49)."""
    import pickle
    from baryonforge_b200.tables import get_parameter
    axes = synth.table_axes()
    inner = b.ProfileModel(axes, None, synth.profile_values(axes), proj_cutoff=37.5)
    assert get_parameter(inner, 'proj_cutoff') == 37.5 and get_parameter(inner, 'no_such_key') is None

    class Wrapper(object):            # a profile wrapping another profile, like TabulatedProfile(model=...)
        def __init__(self, m):
            self.model = m

        def real(self):
            pass

        def projected(self):
            pass
    assert get_parameter(Wrapper(inner), 'proj_cutoff') == 37.5
    ra, dec, M, z = synth.sky_halos(20)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=np.ones(12 * 4 * 4), cosmo=synth.COSMO, redshift=0.3)
    run = b.PaintProfilesAnisShell(cat, shell, 5, inner, inner, inner, 1.5, 0.2, verbose=False)
    back = pickle.loads(pickle.dumps(run))
    assert back.background_val == 1.5 and back.global_tracer_fraction == 0.2 and back.epsilon_max == 5
    assert type(back).__mro__[1].__name__ == 'DefaultRunner'
    with pytest.raises(NotImplementedError):
        b.PaintProfilesAnisShell(cat, shell, 5, inner, inner, inner, 1.5, 0.2, use_ellipticity=True)
    pos, Mb = synth.box_halos(5, 50.0)
    nd = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=Mb, redshift=0.2, cosmo=synth.COSMO)
    gm3 = b.GriddedMap(map=np.ones((8, 8, 8)), redshift=0.2, bins=(np.arange(8) + 0.5) * 50 / 8, cosmo=synth.COSMO)
    grid_run = pickle.loads(pickle.dumps(b.PaintProfilesAnisGrid(nd, gm3, 5, inner, inner, inner, 1.0, 0.1, verbose=False)))
    with pytest.raises(AssertionError, match="2D maps"):      # Map2DRunner.py:847, raised before any GPU work
        grid_run.process()


def test_first_pixel_at_colatitude_is_a_ring_start():
    """Host helper of the pipelined BaryonifyShell.process: first pixel of the first ring at or south of a colatitude."""
    rng = np.random.default_rng(6)
    for nside in (1, 2, 16, 128):
        npix = 12 * nside * nside
        theta, _ = hpo.pix2ang(nside, np.arange(npix))
        ring = parallel.ring_of_pixel(nside, np.arange(npix))
        starts = np.flatnonzero(np.diff(ring, prepend=0) > 0)          # first pixel of every ring
        assert parallel.first_pixel_at_colatitude(nside, 0.0) == 0
        assert parallel.first_pixel_at_colatitude(nside, np.pi) == npix
        for t in rng.uniform(1e-3, np.pi - 1e-3, 300):
            p = parallel.first_pixel_at_colatitude(nside, t)
            assert p == npix or p in starts
            # every pixel before p lies strictly north of t, the ring starting at p does not
            if p > 0:
                assert theta[p - 1] < t + 1e-12
            if p < npix:
                assert theta[p] >= t - 1e-12


def test_public_helper_methods_of_the_runner_base_classes():
    """build_Rmat / coord_array / pick_indices / enforce_periodicity / compute_distance keep the reference's behaviour
    (HealpixRunner.py:179-233, Map2DRunner.py:281-375,400-429, SnapshotRunner.py:103-158), checked on known answers."""
    import baryonforge_b200 as b
    sh = object.__new__(b.runners.DefaultRunner)
    A, ref = np.array([0.0, 2.0]), np.array([3.0, 0.0])
    R = sh.build_Rmat(A, ref)
    assert np.allclose(A, [0, 1]) and np.allclose(ref, [1, 0])                 # normalised in place
    assert np.allclose(R, [[0, -1], [1, 0]])
    xy = sh.coord_array(np.arange(6).reshape(2, 3), 10 * np.arange(6).reshape(2, 3))
    assert xy.shape == (6, 2) and np.array_equal(xy[4], [4, 40])
    gr = object.__new__(b.runners.DefaultRunnerGrid)
    assert np.allclose(gr.build_Rmat(np.array([1.0, 0.0]), 1.0), np.eye(2))    # round halo: identity
    S = gr.build_Rmat(np.array([2.0, 0.0]), 0.5)                               # galsim Shear(q = 0.5, beta = 0)
    g = np.tanh(0.5 * np.log(2.0))
    assert np.allclose(S, np.array([[1 + g, 0], [0, 1 - g]]) / np.sqrt(1 - g * g)) and np.isclose(np.linalg.det(S), 1.0)
    with pytest.raises(NotImplementedError):
        gr.build_Rmat(np.array([1.0, 0.0, 0.0]), 0.5)
    with pytest.raises(ValueError):
        gr.build_Rmat(np.array([1.0]), 0.5)
    assert np.array_equal(gr.pick_indices(1, 3, 10), [8, 9, 0, 1, 2, 3])
    assert np.array_equal(gr.pick_indices(9, 2, 10), [7, 8, 9, 0])
    assert np.array_equal(gr.coord_array(np.ones((2, 2)), np.zeros((2, 2))), np.array([[1, 0]] * 4))
    sn = object.__new__(b.runners.DefaultRunnerSnapshot)

    class _PS(object):
        L = 10.0
    sn.ParticleSnapshot = _PS()
    assert np.array_equal(sn.enforce_periodicity(np.array([6.0, -6.0, 5.0, -5.0, 1.0])), [-4.0, 4.0, 5.0, -5.0, 1.0])
    assert np.allclose(sn.compute_distance(np.array([9.0]), np.array([-8.0]), np.array([0.5])), np.sqrt(1 + 4 + 0.25))


def test_table_cache_is_keyed_by_identity_not_by_a_recyclable_id():
    """`Runner.model = NewModel` in a loop (examples/10_...ipynb cell 15): a new model must never be served the previous
    model's device table, even when CPython gives it the address of the collected one."""
    from baryonforge_b200.runners import _Ident, _TableCache

    class FakeTable(object):
        def __init__(self, tag):
            self.tag, self.closed = tag, False

        def close(self):
            self.closed = True

    class Model(object):
        def __init__(self, tag):
            self.interp_d = [tag]

    from baryonforge_b200 import runners
    del runners._SHARED_TABLES[:]
    cache = _TableCache()
    made = []

    def table_for(m):
        def make():
            made.append(FakeTable(m.interp_d[0]))
            return made[-1]
        return cache.get((_Ident(m), _Ident(m.interp_d)), make)

    m = Model(0)
    t0 = table_for(m)
    assert table_for(m) is t0 and len(made) == 1                      # same model, same table object: re-used
    for k in range(1, 50):                                            # fresh models, the old ones become garbage
        m = Model(k)
        t = table_for(m)
        assert t.tag == k, "stale table served for a new model"
        assert not t.closed and not made[-2].closed                   # the last 8 tables stay resident (shared by all runners) ...
        assert len(made) < 9 or made[-9].closed                       # ... older ones are released
    assert _TableCache().get((_Ident(m), _Ident(m.interp_d)), lambda: None) is t   # another runner's cache object: same table
    m.interp_d = [99]                                                 # setup_interpolator() re-run on the same model
    assert table_for(m).tag == 99
    assert _Ident(m) == _Ident(m) and _Ident(m) != _Ident(Model(1)) and (_Ident(m), '2D') != (_Ident(m), '3D')
    assert _Ident(None) == _Ident(None) and _Ident(m) != 0


def test_parallel_driver_mirrors_keep_the_reference_contract_without_a_gpu():
    """SimpleParallel / SplitJoinParallel (utils/Parallelize.py:8-113, 116-320) with a duck-typed painting runner: output
    order, njobs semantics, the seeded reshuffle + ceil(N / njobs) chunks, include_pixel_size not forwarded,
    Baryonify* runners refused (:206-209)."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth

    class CountingPainter(object):
        """process() = histogram of halo masses over 8 'pixels' -- additive over halos like a painting runner."""
        def __init__(self, HaloLightConeCatalog, LightconeShell, epsilon_max, model, use_ellipticity=False, mass_def=None,
                     include_pixel_size=False, verbose=True):
            self.HaloLightConeCatalog, self.LightconeShell = HaloLightConeCatalog, LightconeShell
            self.cosmo = HaloLightConeCatalog.cosmology
            self.epsilon_max, self.model, self.use_ellipticity, self.mass_def = epsilon_max, model, use_ellipticity, mass_def
            self.include_pixel_size, self.verbose = include_pixel_size, verbose

        def process(self):
            cat = self.HaloLightConeCatalog.cat
            return np.bincount((cat['ra'] // 45).astype(int), weights=cat['M'], minlength=8) + self.LightconeShell.map[:8]

    ra, dec, M, z = synth.sky_halos(103, seed=9)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=np.zeros(12 * 4 * 4), cosmo=synth.COSMO)
    whole = CountingPainter(cat, shell, 5, None, include_pixel_size=True)
    sj = b.SplitJoinParallel(whole, njobs=4)
    assert sj.njobs == 4 and len(sj.Runner_list) == 4
    sizes = [len(r.HaloLightConeCatalog.cat) for r in sj.Runner_list]
    assert sizes == [26, 26, 26, 25]                                             # ceil(103 / 4) per split
    perm = np.random.default_rng(42).choice(103, size=103, replace=False)         # Parallelize.py:255
    assert np.array_equal(sj.Runner_list[0].HaloLightConeCatalog.cat['M'], M[perm[:26]])
    assert all(r.include_pixel_size is False and r.verbose is False for r in sj.Runner_list)   # :271 drops it
    assert all(np.all(r.LightconeShell.map == 0) for r in sj.Runner_list)
    np.testing.assert_allclose(sj.process(), whole.process(), rtol=1e-12)
    # njobs = -1: the reference forks cpu_count() workers (Parallelize.py:210); on the GPU one pass does the same sum
    assert b.SplitJoinParallel(whole).njobs == 1
    a, c = CountingPainter(cat[:10], shell, 5, None), CountingPainter(cat[10:30], shell, 5, None)
    outs = b.SimpleParallel([a, c, a]).process()
    assert len(outs) == 3 and np.array_equal(outs[0], a.process()) and np.array_equal(outs[1], c.process())
    assert b.SimpleParallel([a, c, a], njobs=2).njobs == 2 and b.SimpleParallel([a, c]).njobs == 2
    axes = synth.table_axes()
    model = b.DisplacementModel(axes, synth.displacement_values(axes), 20, synth.COSMO)
    with pytest.raises(AssertionError):
        b.SplitJoinParallel(b.BaryonifyShell(cat, shell, 20, model, verbose=False), njobs=2)


def test_out_of_table_warnings_once_per_catalogue():
    """BaryonCorrection.py:378-389 words the out-of-range warnings per halo; the GPU runners emit them once per process()
    from the catalogue's extremes -- same wording, and a broken model never breaks process()."""
    import warnings
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from baryonforge_b200.runners import _table_range_messages, _warn_table_range
    axes = synth.table_axes(z_min=0.01, z_max=1.0, M_min=1e12, M_max=10 ** 15.5)
    model = b.DisplacementModel(axes, synth.displacement_values(axes), 20, synth.COSMO)
    inside = _table_range_messages(model, np.array([0.1, 0.5]), np.array([1e13, 1e15]))
    assert inside == []
    msgs = _table_range_messages(model, np.array([0.1, 1.2]), np.array([3e11, 1e15]))
    assert len(msgs) == 2
    assert msgs[0].startswith("Requested redshift range [0.1, 1.2] outside table's range [")
    assert msgs[1].startswith("Requested log_Mass range [") and "outside table's range [" in msgs[1]
    assert _table_range_messages(model, 0.3, np.array([1e13], dtype='>f4')) == []          # box catalogues: scalar z, f32 M
    assert _table_range_messages(model, np.zeros(0), np.zeros(0)) == []
    assert _table_range_messages(object(), np.array([5.0]), np.array([1.0])) == []         # no table: nothing to say
    assert _table_range_messages(model, np.array(['x']), np.array([1.0])) == []            # garbage in: swallowed
    with pytest.warns(UserWarning, match="Requested redshift range"):
        _warn_table_range(model, np.array([2.0]), np.array([1e13]))
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        _warn_table_range(model, np.array([0.2]), np.array([1e13]))                        # in range: silent


def test_record_layout_and_chunk_fractions(monkeypatch):
    """Host logic of the round-2 end-to-end paths: which particle catalogues cross the link as raw 32-byte records, and where the
    latitude chunks of the pipelined shell paths start."""
    from baryonforge_b200 import runners
    ref = np.zeros(10, [('M', 'f8'), ('x', 'f8'), ('y', 'f8'), ('z', 'f8')])          # utils/io.py:588
    assert runners._record_layout(ref, ['x', 'y', 'z']) == dict(M=0, x=1, y=2, z=3)
    assert runners._record_layout(np.zeros(4, [('z', 'f8'), ('y', 'f8'), ('x', 'f8'), ('M', 'f8')]), ['x', 'y']) == dict(z=0, y=1, x=2, M=3)
    assert ref.view(np.float64).shape == (40,)                                       # the raw view the staging copies
    for bad in (np.zeros(10, [('M', 'f4'), ('x', 'f8'), ('y', 'f8'), ('z', 'f8')]),   # float32 masses
                np.zeros(10, [('M', '>f8'), ('x', 'f8'), ('y', 'f8'), ('z', 'f8')]),  # big-endian field
                np.zeros(10, [('M', 'f8'), ('x', 'f8'), ('y', 'f8'), ('z', 'f8'), ('id', 'i8')]),
                np.zeros(10, [('M', 'f8'), ('x', 'f8'), ('y', 'f8')]),
                ref[::2], np.zeros((2, 5), ref.dtype), np.zeros(10)):
        assert runners._record_layout(bad, ['x', 'y']) is None
    assert runners._record_layout(np.zeros(3, [('M', 'f8'), ('x', 'f8'), ('y', 'f8'), ('w', 'f8')]), ['x', 'y', 'z']) is None
    for K in (1, 4, 6, 8, 12, 24):
        fr = runners._chunk_fractions(K)
        assert len(fr) == K and fr[0] == 0.0 and all(b > a for a, b in zip(fr[:-1], fr[1:])) and fr[-1] < 1.0
        if K >= 6:
            assert fr[-3:] == [0.90, 0.96, 0.99]                                      # the exposed tail is a 1 % piece
    monkeypatch.setenv("BFG_PIPELINE_TAPER", "0")
    assert runners._chunk_fractions(8) == [k / 8 for k in range(8)]


def test_host_thread_pool_and_buffer_release_work_without_a_gpu(monkeypatch):
    """The persistent host thread pool runs every chunk exactly once (ragged tail included, results in place), follows
    BFG_HOST_THREADS, and release_host_buffers() is safe to call in a process that never touched a GPU."""
    from baryonforge_b200 import runners
    out = np.zeros(200003)
    runners._parallel_chunks(lambda sl: out.__setitem__(sl, out[sl] + np.arange(sl.start, sl.stop)), out.size, chunk=4096)
    assert np.array_equal(out, np.arange(out.size, dtype=float))
    monkeypatch.setenv("BFG_HOST_THREADS", "3")
    assert runners._host_threads() == 3 and runners._pool() is runners._pool()
    monkeypatch.setenv("BFG_HOST_THREADS", "1")
    seen = []
    runners._parallel_chunks(lambda sl: seen.append((sl.start, sl.stop)), 10, chunk=4)
    assert seen == [(0, 4), (4, 8), (8, 10)]                                  # one thread: in order, on the caller's thread
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    monkeypatch.delenv("BFG_HOST_THREADS")
    assert 1 <= runners._host_threads() <= 16
    runners.release_host_buffers()
    assert runners._PINNED_FREE == {} and runners._PINNED_SCRATCH == {}


def test_both_bench_arms_name_the_same_config():
    """bench.shell_config builds `config` for the GPU arm and for `--impl reference`: equal dicts by construction, and equal to what
    the recorded GPU lines of this round carry (profiles/r2_bench_n1_final.json, N = 8 line)."""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    argv, sys.argv = sys.argv, ['bench.py']
    try:
        import bench
        args = bench.parse()
    finally:
        sys.argv = argv
    n1 = json.load(open(os.path.join(root, "profiles", "r2_bench_n1_final.json")))["config"]
    n8 = json.load(open(os.path.join(root, "profiles", "r2_bench_n8_final_grid_leg_failed.json")))["config"]
    assert bench.shell_config(args, 1) == n1 == bench.shell_config(args, 1, float(n1["n_updates_per_step"]), False, False)
    assert bench.shell_config(args, 8) == n8 == bench.shell_config(args, 8, float(n8["n_updates_per_step"]), True, False)
    assert bench.sharding_text(4, p2p=False) == "RING pixel ranges x4, overlap halos replicated, NCCL all-reduce of partial maps"
    args.mass_function = True
    assert bench.shell_config(args, 1)["n_updates_per_step"] is None            # only counted workloads are named
