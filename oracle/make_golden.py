"""
oracle/make_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/*.npz.

Runs the REFERENCE'S OWN runner code (imported unmodified from /root/reference, with the `healpy` / `pyccl`
stand-ins of oracle/shims on sys.path because those third-party packages are not installed here) on small seeded
synthetic inputs and stores inputs + outputs.  The reference tree does not exist on the GPU box, so these fixtures
are what pins oracle/runners_port.py and the CUDA path there.

    python oracle/make_golden.py            # rewrites every fixture (needs /root/reference)
    python oracle/make_golden.py anis       # only the fixtures whose name contains 'anis'

Inputs follow SURVEY.md §8(d) (distributions of the reference's tests/test_healpix.py:29-55, tests/defaults.py:5).
The reference objects that need real pyccl to be *built* (Baryonification2D/3D, TabulatedProfile) are created
without running their constructors and given the attributes setup_interpolator() would leave behind
(Profiles/BaryonCorrection.py:307-323, utils/Tabulate.py:261-271) from analytic synthetic tables.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("BFG_REFERENCE", "/root/reference")
for p in (ROOT, os.path.join(HERE, "shims"), REF):
    if p not in sys.path:
        sys.path.insert(0, p)

from baryonforge_b200 import synth  # noqa: E402  (input generators only)


def import_reference():
    warnings.filterwarnings("ignore")
    import BaryonForge as bfg  # noqa: F401
    import pyccl as ccl
    from scipy.interpolate import RegularGridInterpolator as RGI
    from BaryonForge.Profiles.BaryonCorrection import Baryonification2D, Baryonification3D
    from BaryonForge.utils.Tabulate import TabulatedProfile
    return bfg, ccl, RGI, Baryonification2D, Baryonification3D, TabulatedProfile


MODEL_COSMO = dict(Omega_c=0.27, Omega_b=0.05, h=0.68, sigma8=0.82, n_s=0.97)   # != runner cosmology (§10 #6)


def ref_displacement_model(axes, values, epsilon_max, cls_name='2D', Rdelta_sampling=False):
    bfg, ccl, RGI, B2, B3, TP = import_reference()
    m = object.__new__(B2 if cls_name == '2D' else B3)
    m.cosmo = ccl.Cosmology(matter_power_spectrum='linear', **MODEL_COSMO)
    m.epsilon_max = epsilon_max
    m.mass_def = ccl.halos.massdef.MassDef(200, 'critical')
    m.p_keys = []
    m.raw_input_d = values
    m.raw_input_z_range, m.raw_input_M_range, m.raw_input_r_range = axes
    m.interp_d = RGI(tuple(axes), values, bounds_error=False, fill_value=np.nan)
    m.Rdelta_sampling = Rdelta_sampling
    return m


def ref_profile_model(axes, raw3D, raw2D):
    bfg, ccl, RGI, B2, B3, TP = import_reference()
    m = object.__new__(TP)
    ccl.halos.profiles.HaloProfile.__init__(m, mass_def=ccl.halos.massdef.MassDef(200, 'critical'))
    with np.errstate(divide='ignore', invalid='ignore'):
        m.interp3D = RGI(tuple(axes), np.log(raw3D), bounds_error=False)
        m.interp2D = RGI(tuple(axes), np.log(raw2D), bounds_error=False)
    m.raw_input_3D, m.raw_input_2D = raw3D, raw2D
    m.raw_input_z_range, m.raw_input_M_range, m.raw_input_r_range = axes
    return m


# ---------------------------------------------------------------------------------------------------------------
# case definitions (shared with tests/ through the stored inputs)
# ---------------------------------------------------------------------------------------------------------------
def shell_catalog(n, seed, z=(0.03, 0.5), edge_cases=True):
    ra, dec, M, zz = synth.sky_halos(n, seed=seed, z=z)
    if edge_cases and n >= 40:
        dec[0], dec[1] = 89.93, -89.95          # discs that swallow a pole
        ra[2], ra[3] = 0.01, 359.99             # discs straddling phi = 0
        M[4], M[5] = 3e11, 8e15                 # outside the table's mass range -> NaN -> zero
        zz[6] = 1.2                             # outside the table's redshift range
        dec[7] = 90.0                           # exactly on the pole: clipped by the catalogue (io.py:65-68)
        M[8:12] = 10 ** 12.05                   # tiny discs -> the <4-pixel fallback
        zz[8:12] = 0.49
        M[12], zz[12] = 10 ** 15.5, 0.031       # a huge disc
    return ra, dec, M, zz


def per_halo_scalars_shell(ccl, cosmo_dict, model, M, z):
    """What the reference computes per halo through pyccl (inputs of the port)."""
    from scipy import interpolate
    cosmo = ccl.Cosmology(Omega_c=cosmo_dict['Omega_m'] - cosmo_dict['Omega_b'], Omega_b=cosmo_dict['Omega_b'],
                          h=cosmo_dict['h'], sigma8=cosmo_dict['sigma8'], n_s=cosmo_dict['n_s'], w0=cosmo_dict['w0'],
                          matter_power_spectrum='linear')
    md = ccl.halos.massdef.MassDef(200, 'critical')
    a = 1 / (1 + z)
    z_t = np.linspace(0, np.max(z) + 0.1, 1000)
    D_a = interpolate.CubicSpline(z_t, ccl.angular_diameter_distance(cosmo, 1 / (1 + z_t)))
    R_run = np.array([md.get_radius(cosmo, M[j], a[j]) for j in range(M.size)])
    D_A = np.array([D_a(z[j]) for j in range(M.size)])
    R_mod = None
    if model is not None:
        R_mod = np.array([model.mass_def.get_radius(model.cosmo, M[j], a[j]) / a[j] for j in range(M.size)])
    return R_run, D_A, R_mod


def per_halo_scalars_box(ccl, cosmo_dict, model, M32, redshift):
    cosmo = ccl.Cosmology(Omega_c=cosmo_dict['Omega_m'] - cosmo_dict['Omega_b'], Omega_b=cosmo_dict['Omega_b'],
                          h=cosmo_dict['h'], sigma8=cosmo_dict['sigma8'], n_s=cosmo_dict['n_s'],
                          matter_power_spectrum='linear')
    md = ccl.halos.massdef.MassDef(200, 'critical')
    a = 1 / (1 + redshift)
    R_phys = np.array([md.get_radius(cosmo, M32[j], a) for j in range(M32.size)])
    R_mod = None
    if model is not None:
        R_mod = np.array([model.mass_def.get_radius(model.cosmo, M32[j], a) / a for j in range(M32.size)])
    return R_phys, R_mod


def save(name, **arrays):
    only = sys.argv[1:]                      # optional name filters: python oracle/make_golden.py anis
    if only and not any(f in name for f in only):
        return
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(out, **arrays)
    print("wrote", out, "%.1f KB" % (os.path.getsize(out) / 1024))


def main():
    bfg, ccl, RGI, B2, B3, TP = import_reference()
    from BaryonForge.Runners import (BaryonifyShell, PaintProfilesShell, BaryonifyGrid, PaintProfilesGrid,
                                     BaryonifySnapshot)
    from BaryonForge.utils.io import (HaloLightConeCatalog, HaloNDCatalog, LightconeShell, GriddedMap,
                                      ParticleSnapshot)
    cosmo = synth.COSMO

    # ---------------- shells
    axes = synth.table_axes(nz=10, nM=10, nr=500)
    dvals = synth.displacement_values(axes, inject_nan=True)
    pvals = synth.profile_values(axes)

    def shell_bary(name, nside, n, seed, eps_run, eps_mod, rdelta=False, map_lo=0.0):
        ra, dec, M, z = shell_catalog(n, seed)
        hmap = synth.shell_map(nside, seed=seed + 1, lo=map_lo, hi=10.0)
        hmap[::7] = 0.0                                    # zero pixels are skipped by the regrid (§10 #2)
        ax = axes
        vals = dvals
        if rdelta:
            ax = (axes[0], axes[1], np.log(np.geomspace(1e-3, 10, 500)))
            vals = synth.displacement_values((axes[0], axes[1], ax[2] + 0.0))
        model = ref_displacement_model(ax, vals, eps_mod, '2D', rdelta)
        cat = HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=cosmo)
        shell = LightconeShell(map=hmap, cosmo=cosmo)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            new_map = BaryonifyShell(cat, shell, eps_run, model, verbose=False).process()
        R_run, D_A, R_mod = per_halo_scalars_shell(ccl, cosmo, model, cat.cat['M'], cat.cat['z'])
        save(name, kind="shell_bary", nside=nside, ra=cat.cat['ra'], dec=cat.cat['dec'], M=cat.cat['M'], z=cat.cat['z'],
             map=hmap, ax0=ax[0], ax1=ax[1], ax2=ax[2], values=vals, eps_run=eps_run, eps_mod=eps_mod, rdelta=rdelta,
             R_run=R_run, D_A=D_A, R_mod=R_mod, out=new_map, model_cosmo=np.array(list(MODEL_COSMO.values())))

    shell_bary("shell_bary_n64", 64, 400, 11, 20, 6)
    shell_bary("shell_bary_n32_signed", 32, 150, 12, 10, 20, map_lo=-10.0)
    shell_bary("shell_bary_n32_rdelta", 32, 150, 13, 20, 8, rdelta=True)

    def shell_bary_config1(name, nside=256, n=10000, seed=42, eps_run=20, eps_mod=20, stride=8):
        """BASELINE.json configs[0] at full size -- BaryonifyShell NSIDE=256, 10^4 halos of the reference's own test
        distribution (tests/test_healpix.py:29-32), table 10x10x500, epsilon_max=20 -- run by the reference's code.  Inputs
        are regenerated from the seeds; the fixture keeps the per-halo pyccl scalars and every `stride`-th output pixel."""
        ra, dec, M, z = synth.sky_halos(n, seed=seed)
        hmap = synth.shell_map(nside, seed=seed + 1)
        vals = synth.displacement_values(axes)
        model = ref_displacement_model(axes, vals, eps_mod, '2D', False)
        cat = HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=cosmo)
        shell = LightconeShell(map=hmap, cosmo=cosmo)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            new_map = BaryonifyShell(cat, shell, eps_run, model, verbose=False).process()
        R_run, D_A, R_mod = per_halo_scalars_shell(ccl, cosmo, model, cat.cat['M'], cat.cat['z'])
        save(name, kind="shell_bary_config1", nside=nside, n=n, seed=seed, eps_run=eps_run, eps_mod=eps_mod, stride=stride,
             R_run=R_run, D_A=D_A, R_mod=R_mod, out_sub=new_map[::stride], out_sum=new_map.sum(), map_sum=hmap.sum(),
             n_changed=np.int64(np.count_nonzero(new_map != hmap)))

    shell_bary_config1("shell_bary_config1")

    def shell_paint(name, nside, n, seed, eps_run, pixsize):
        ra, dec, M, z = shell_catalog(n, seed)
        model = ref_profile_model(axes, pvals * 3.0, pvals)
        cat = HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=cosmo)
        shell = LightconeShell(map=np.zeros(12 * nside * nside), cosmo=cosmo)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            new_map = PaintProfilesShell(cat, shell, eps_run, model, include_pixel_size=pixsize, verbose=False).process()
        R_run, D_A, _ = per_halo_scalars_shell(ccl, cosmo, None, cat.cat['M'], cat.cat['z'])
        save(name, kind="shell_paint", nside=nside, ra=cat.cat['ra'], dec=cat.cat['dec'], M=cat.cat['M'], z=cat.cat['z'],
             ax0=axes[0], ax1=axes[1], ax2=axes[2], raw2D=pvals, raw3D=pvals * 3.0, eps_run=eps_run, pixsize=pixsize,
             R_run=R_run, D_A=D_A, out=new_map)

    def shell_paint_config2_map(name, nside=1024, n=10000, seed=42, eps_run=20, stride=128):
        """BASELINE.json configs[1] (PaintProfilesShell, NSIDE=1024) at the full map size with 10^4 of its halos -- what the
        reference's code can paint in seconds.  Inputs regenerate from the seeds; every `stride`-th output pixel is kept."""
        ra, dec, M, z = synth.sky_halos(n, seed=seed)
        model = ref_profile_model(axes, pvals * 3.0, pvals)
        cat = HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=cosmo)
        shell = LightconeShell(map=np.zeros(12 * nside * nside), cosmo=cosmo)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            new_map = PaintProfilesShell(cat, shell, eps_run, model, include_pixel_size=False, verbose=False).process()
        R_run, D_A, _ = per_halo_scalars_shell(ccl, cosmo, None, cat.cat['M'], cat.cat['z'])
        save(name, kind="shell_paint_config2", nside=nside, n=n, seed=seed, eps_run=eps_run, stride=stride, R_run=R_run, D_A=D_A,
             out_sub=new_map[::stride], out_sum=new_map.sum(), n_painted=np.int64(np.count_nonzero(new_map)))

    shell_paint_config2_map("shell_paint_config2_map")

    shell_paint("shell_paint_n64", 64, 400, 21, 20, False)
    shell_paint("shell_paint_n32_pixsize", 32, 150, 22, 10, True)

    # ---------------- anisotropic shell painter (HealpixRunner.py:484-640)
    def shell_anis(name, nside, n, seed, eps_run, pixsize, halo_fraction, z_shell, proj_cutoff=50.0):
        """halo_fraction: target rho_halos / rho_m -- < 1 leaves a uniform background, > 1 leaves none (and pixels
        no halo reaches keep Mtot = 0, the `where = Mtot > 0` branch)."""
        from BaryonForge.Runners import PaintProfilesAnisShell
        from scipy import interpolate
        ra, dec, M, z = shell_catalog(n, seed)
        cat = HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=cosmo)
        hmap = synth.shell_map(nside, seed=seed + 1, lo=0.0, hi=10.0)
        shell = LightconeShell(map=hmap, cosmo=cosmo, redshift=z_shell)
        ccosmo = ccl.Cosmology(Omega_c=cosmo['Omega_m'] - cosmo['Omega_b'], Omega_b=cosmo['Omega_b'], h=cosmo['h'],
                               sigma8=cosmo['sigma8'], n_s=cosmo['n_s'], w0=cosmo['w0'], matter_power_spectrum='linear')
        z_t = np.linspace(0, np.max(cat.cat['z']) + 0.1, 1000)
        dD = float(interpolate.CubicSpline(z_t, ccl.angular_diameter_distance(ccosmo, 1 / (1 + z_t)))(z_shell))
        rho_m = float(ccosmo.rho_x(1 / (z_shell + 1), species='matter', is_comoving=False))
        # scale the total-mass table so that the halos carry `halo_fraction` of the mean matter density
        mt0 = ref_profile_model(axes, pvals * 3.0, pvals)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m0 = PaintProfilesShell(cat, shell, eps_run, mt0, include_pixel_size=True, verbose=False).process()
        pixarea = 4 * np.pi / m0.size
        dV = pixarea * ((dD + 2 * proj_cutoff) ** 3 - dD ** 3)
        amp = halo_fraction * rho_m * dV * m0.size / np.sum(m0)
        mtot2D = pvals * amp
        tracer2D = pvals ** 0.8 * 40.0                      # a different radial shape than the painted profile
        mtot = ref_profile_model(axes, mtot2D * 3.0, mtot2D)
        mtot.proj_cutoff = proj_cutoff
        tracer = ref_profile_model(axes, tracer2D * 3.0, tracer2D)
        paint2D = pvals * 1e24                              # halo term of the same order as the background term
        model = ref_profile_model(axes, paint2D * 3.0, paint2D)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = PaintProfilesAnisShell(cat, shell, eps_run, model, tracer, mtot, 2.5, 0.3,
                                         include_pixel_size=pixsize, verbose=False).process()
        R_run, D_A, _ = per_halo_scalars_shell(ccl, cosmo, None, cat.cat['M'], cat.cat['z'])
        save(name, kind="shell_anis", nside=nside, ra=cat.cat['ra'], dec=cat.cat['dec'], M=cat.cat['M'], z=cat.cat['z'],
             map=hmap, ax0=axes[0], ax1=axes[1], ax2=axes[2], raw2D=paint2D, tracer2D=tracer2D, mtot2D=mtot2D,
             eps_run=eps_run, pixsize=pixsize, z_shell=z_shell, proj_cutoff=proj_cutoff, background_val=2.5,
             global_tracer_fraction=0.3, dD=dD, rho_m=rho_m, R_run=R_run, D_A=D_A, out=out)

    shell_anis("shell_anis_n32_background", 32, 150, 51, 10, False, 0.35, 0.3)
    shell_anis("shell_anis_n32_overfull", 32, 150, 52, 6, True, 2.0, 0.25)

    # ---------------- grids
    gaxes = synth.table_axes(nz=6, nM=10, nr=300, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-2, r_max=2e2)
    gd = synth.displacement_values(gaxes, inject_nan=False)
    gd[:, :, :] *= 25.0                                    # cell-scale displacements at res ~ 1.5 Mpc
    gp = synth.profile_values(gaxes)

    def grid_case(name, ndim, N, Lbox, n, seed, eps_run, eps_mod, redshift, paint, ell=False):
        pos, M = synth.box_halos(n, Lbox, seed=seed, ndim=ndim)
        M[0] = 3e11                                        # outside the table: NaN-poisons its cells (§10 #5)
        pos[:, 1] = 0.01 * Lbox / N                        # hugging the box corner: periodic wrap of the cutout
        pos[:, 2] = Lbox * (1 - 1e-3)
        bins = (np.arange(N) + 0.5) * Lbox / N
        gmap = np.random.default_rng(seed + 1).uniform(0, 10, (N,) * ndim)
        ekw = {}
        if ell:
            erng = np.random.default_rng(seed + 7)
            ekw = dict(q_ell=erng.uniform(0.4, 1.0, n), A_ell=erng.normal(size=(n, 2)))
            ekw['q_ell'][1] = 1.0 - 1e-6                       # the small-eta series branch of build_Rmat
        cat = HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2] if ndim == 3 else None, M=M, redshift=redshift, cosmo=cosmo, **ekw)
        gm = GriddedMap(map=gmap, redshift=redshift, bins=bins, cosmo=cosmo)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if paint:
                model = ref_profile_model(gaxes, gp * 3.0, gp)
                out = PaintProfilesGrid(cat, gm, eps_run, model, use_ellipticity=ell, verbose=False).process()
                R_phys, R_mod = per_halo_scalars_box(ccl, cosmo, None, cat.cat['M'], redshift)
            else:
                model = ref_displacement_model(gaxes, gd, eps_mod, '2D' if ndim == 2 else '3D')
                out = BaryonifyGrid(cat, gm, eps_run, model, use_ellipticity=ell, verbose=False).process()
                R_phys, R_mod = per_halo_scalars_box(ccl, cosmo, model, cat.cat['M'], redshift)
        extra = dict(raw2D=gp, raw3D=gp * 3.0) if paint else dict(values=gd, R_mod=R_mod, eps_mod=eps_mod)
        if ell:
            extra.update(q_ell=cat.cat['q_ell'].astype('<f4'), A_ell=cat.cat['A_ell'].astype('<f4'))
        save(name, kind="grid_paint" if paint else "grid_bary", ndim=ndim, N=N, L=Lbox, redshift=redshift,
             M=cat.cat['M'].astype('<f4'), x=cat.cat['x'].astype('<f4'), y=cat.cat['y'].astype('<f4'),
             z=cat.cat['z'].astype('<f4'), map=gmap, ax0=gaxes[0], ax1=gaxes[1], ax2=gaxes[2], eps_run=eps_run,
             R_phys=R_phys, out=out, **extra)

    grid_case("grid_bary_2d", 2, 128, 200.0, 150, 31, 10, 5, 0.3, False)
    grid_case("grid_bary_3d", 3, 40, 80.0, 50, 32, 6, 4, 0.3, False)
    grid_case("grid_paint_2d", 2, 128, 200.0, 150, 33, 6, None, 0.3, True)
    grid_case("grid_paint_3d", 3, 40, 80.0, 50, 34, 4, None, 0.0, True)
    grid_case("grid_bary_2d_ell", 2, 96, 150.0, 100, 35, 10, 5, 0.3, False, ell=True)
    grid_case("grid_paint_2d_ell", 2, 96, 150.0, 100, 36, 6, None, 0.3, True, ell=True)

    # ---------------- anisotropic grid painter (Map2DRunner.py:833-1015), 2-D maps only
    def grid_anis(name, N, Lbox, n, seed, eps_run, redshift, pixsize, halo_fraction, ell=False, proj_cutoff=40.0):
        from BaryonForge.Runners import PaintProfilesAnisGrid
        pos, M = synth.box_halos(n, Lbox, seed=seed, ndim=2)
        M[0] = 3e11
        pos[:, 1] = 0.01 * Lbox / N
        pos[:, 2] = Lbox * (1 - 1e-3)
        bins = (np.arange(N) + 0.5) * Lbox / N
        gmap = np.random.default_rng(seed + 1).uniform(0, 10, (N, N))
        ekw = {}
        if ell:
            erng = np.random.default_rng(seed + 7)
            ekw = dict(q_ell=erng.uniform(0.4, 1.0, n), A_ell=erng.normal(size=(n, 2)))
        cat = HaloNDCatalog(x=pos[0], y=pos[1], z=None, M=M, redshift=redshift, cosmo=cosmo, **ekw)
        gm = GriddedMap(map=gmap, redshift=redshift, bins=bins, cosmo=cosmo)
        ccosmo = ccl.Cosmology(Omega_c=cosmo['Omega_m'] - cosmo['Omega_b'], Omega_b=cosmo['Omega_b'], h=cosmo['h'],
                               sigma8=cosmo['sigma8'], n_s=cosmo['n_s'], matter_power_spectrum='linear')
        rho_m = float(ccosmo.rho_x(1 / (redshift + 1), species='matter', is_comoving=True))
        mt0 = ref_profile_model(gaxes, gp * 3.0, gp)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            m0 = PaintProfilesGrid(cat, gm, eps_run, mt0, use_ellipticity=ell, include_pixel_size=False,
                                   verbose=False).process()
        amp = halo_fraction * rho_m * (2 * proj_cutoff) / np.average(m0)
        mtot2D = gp * amp
        tracer2D = gp ** 0.8 * 40.0
        mtot = ref_profile_model(gaxes, mtot2D * 3.0, mtot2D)
        mtot.proj_cutoff = proj_cutoff
        tracer = ref_profile_model(gaxes, tracer2D * 3.0, tracer2D)
        paint2D = gp * 1e18                                 # halo term of the same order as the background term
        model = ref_profile_model(gaxes, paint2D * 3.0, paint2D)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = PaintProfilesAnisGrid(cat, gm, eps_run, model, tracer, mtot, 2.5, 0.3, include_pixel_size=pixsize,
                                        use_ellipticity=ell, verbose=False).process()
        R_phys, _ = per_halo_scalars_box(ccl, cosmo, None, cat.cat['M'], redshift)
        extra = {}
        if ell:
            extra.update(q_ell=cat.cat['q_ell'].astype('<f4'), A_ell=cat.cat['A_ell'].astype('<f4'))
        save(name, kind="grid_anis", ndim=2, N=N, L=Lbox, redshift=redshift, M=cat.cat['M'].astype('<f4'),
             x=cat.cat['x'].astype('<f4'), y=cat.cat['y'].astype('<f4'), z=cat.cat['z'].astype('<f4'), map=gmap,
             ax0=gaxes[0], ax1=gaxes[1], ax2=gaxes[2], raw2D=paint2D, tracer2D=tracer2D, mtot2D=mtot2D, eps_run=eps_run,
             pixsize=pixsize, proj_cutoff=proj_cutoff, background_val=2.5, global_tracer_fraction=0.3, rho_m=rho_m,
             R_phys=R_phys, out=out, **extra)

    grid_anis("grid_anis_2d_background", 96, 150.0, 100, 61, 6, 0.3, True, 0.35)
    grid_anis("grid_anis_2d_overfull_ell", 96, 150.0, 100, 62, 4, 0.3, False, 2.0, ell=True)

    # ---------------- snapshots
    def snap_case(name, ndim, n_part, Lbox, n, seed, eps_run, eps_mod, redshift):
        rng = np.random.default_rng(seed)
        pos, M = synth.box_halos(n, Lbox, seed=seed + 1, ndim=ndim)
        M[0] = 3e11
        pos[:, 1] = 0.02
        p = rng.uniform(0, Lbox, (ndim, n_part))
        # a clustered component: blobs around the first halos
        k = n_part // 3
        own = rng.integers(0, n, k)
        p[:, :k] = (pos[:, own].astype('f4').astype('f8') + rng.normal(0, 1.5, (ndim, k))) % Lbox
        Mp = np.full(n_part, 1e10)
        cat = HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2] if ndim == 3 else None, M=M, redshift=redshift, cosmo=cosmo)
        ps = ParticleSnapshot(x=p[0], y=p[1], z=p[2] if ndim == 3 else None, M=Mp, L=Lbox, redshift=redshift, cosmo=cosmo)
        model = ref_displacement_model(gaxes, gd, eps_mod, '3D')
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = BaryonifySnapshot(cat, ps, eps_run, model, verbose=False).process()
        R_phys, R_mod = per_halo_scalars_box(ccl, cosmo, model, cat.cat['M'], redshift)
        ps2 = ParticleSnapshot(x=out['x'], y=out['y'], z=out['z'] if ndim == 3 else None, M=Mp, L=Lbox,
                               redshift=redshift, cosmo=cosmo)
        ngp = ps2.make_map(16)
        save(name, kind="snap", ndim=ndim, L=Lbox, redshift=redshift, M=cat.cat['M'].astype('<f4'),
             x=cat.cat['x'].astype('<f4'), y=cat.cat['y'].astype('<f4'), z=cat.cat['z'].astype('<f4'),
             px=p[0], py=p[1], pz=p[2] if ndim == 3 else np.zeros(0), pM=Mp, ax0=gaxes[0], ax1=gaxes[1], ax2=gaxes[2],
             values=gd, eps_run=eps_run, eps_mod=eps_mod, R_phys=R_phys, R_mod=R_mod, out_x=out['x'], out_y=out['y'],
             out_z=out['z'], ngp=ngp)

    snap_case("snap_3d", 3, 30000, 100.0, 60, 41, 4, 5, 0.3)
    snap_case("snap_2d", 2, 20000, 100.0, 60, 42, 4, 5, 0.3)

    # ---------------- P(k) of a particle set: the notebook cells that follow BaryonifySnapshot.process()
    pk_case("pk_nb10_n48", 48, 30, 60000, 205.0 / 0.6711, 7)
    pk_case("pk_nb10_n64", 64, 45, 150000, 100.0, 8)


def notebook_cells(name):
    import json
    nb = json.load(open(os.path.join(REF, "examples", name)))
    return ["".join(c["source"]) for c in nb["cells"]]


def pk_case(name, Ngrd, Nk, n_part, Lbox, seed):
    """
    Executes the reference's own P(k) code -- cells 1, 12 and the `for factor in [1, 8]` body of cell 15 of
    examples/10_Reproduce_Schneider_deltaPk.ipynb, read from the notebook file -- on a seeded particle set.  Only the two
    size constants of cell 12 (Ngrd = 256, Nk = 180) are replaced.
    """
    cells = notebook_cells("10_Reproduce_Schneider_deltaPk.ipynb")
    c1, c12, c15 = cells[1], cells[12], cells[15]
    assert "def numba_histogram3d" in c1 and "kinds  = np.floor" in c12 and "for factor in [1, 8]:" in c15
    c12 = c12.replace("Ngrd   = 256", "Ngrd   = %d" % Ngrd).replace("Nk     = 180", "Nk     = %d" % Nk)
    assert "Ngrd   = %d" % Ngrd in c12 and "Nk     = %d" % Nk in c12
    lines = c15.split("\n")
    i0 = next(i for i, l in enumerate(lines) if l.strip() == "for factor in [1, 8]:")
    i1 = next(i for i in range(i0, len(lines)) if lines[i].strip().startswith("PkB  = np.bincount"))
    body = [l[4:] for l in lines[i0:i1 + 1]] + ["    RES[factor] = PkB"]
    body = [l.replace("; del FFTB; gc.collect()", "").replace("; del MapB; gc.collect()", "") for l in body]

    class _Snap(object):
        pass
    p = synth.pk_particles(n_part, Lbox, seed)
    Snap = _Snap()
    Snap.L = Lbox
    Snap.cat = dict(x=p[:, 0], y=p[:, 1], z=p[:, 2])
    from numba import njit
    ns = dict(np=np, njit=njit, Snap=Snap, RES={})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(c1, ns)
        exec(c12, ns)
        ns["Part_B"] = ns["Part_D"]
        exec("\n".join(body), ns)
    save(name, kind="pk", Ngrd=Ngrd, Nk=Nk, n_part=n_part, L=Lbox, seed=seed, kbins=ns["kbins"], klin=ns["klin"],
         k_c=ns["k_c"], k_cen=ns["k_cen"], pk_f1=ns["RES"][1], pk_f8=ns["RES"][8],
         kinds_sum=np.int64(ns["kinds"][ns["kmsk"]].sum()))


if __name__ == "__main__":
    main()
