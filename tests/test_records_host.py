"""
CPU: the per-halo scalar prep of the shell runners (shell_record_one in csrc/records_kernels.cu -- the source k_shell_records runs:
the reference's CubicSpline of D_A in scipy's PPoly operation order, R200c from the splined radius factor, hp.ang2vec, the pointing
healpy's query_disc wrapper rebuilds; HealpixRunner.py:317-330, BaryonCorrection.py:371,398-399,410) compiled for the HOST
(bfg_test_shell_records_host), against the numpy restatement in DefaultRunner.halo_records() and against scipy / the cosmology
module directly.  tools/sass_fingerprint.py shows that k_shell_records' SASS is unchanged by the host-compilable form.
"""
import numpy as np

import baryonforge_b200 as b
from baryonforge_b200 import _lib, cosmology, synth


def host_records(run, paint):
    """What DefaultRunner.device_records() hands to bfg_shell_records, built with numpy, through the host entry."""
    cat = run.HaloLightConeCatalog.cat
    n = cat.size
    cols = np.empty((6, n))
    M, z = cat['M'], cat['z']
    cols[0], cols[1], cols[2], cols[3] = M, z, cat['ra'], cat['dec']
    np.log(1 / (1 / (1 + z)), out=cols[4])                                       # BaryonCorrection.py:371
    np.log(M, out=cols[5])                                                       # :398
    pack, n_DA, n_g = run._spline_pack(paint, float(np.max(z)))
    base = pack.ctypes.data
    o_DAc = 8 * n_DA
    o_gx = o_DAc + 8 * 4 * (n_DA - 1)
    o_grun = o_gx + 8 * n_g
    o_gmod = o_grun + 8 * 4 * (n_g - 1)
    pixarea = 4 * np.pi / run.LightconeShell.map.size
    rec = np.zeros((n, _lib.HALO_STRIDE))
    aux = np.zeros((3, n))
    _lib.check(_lib.lib().bfg_test_shell_records_host(
        n, cols.ctypes.data, 1 if paint else 0, float(run.epsilon_max), 0.0 if paint else float(run.model.epsilon_max),
        pixarea if (paint and run.include_pixel_size) else 0.0, n_DA, base, base + o_DAc, n_g, base + o_gx, base + o_grun,
        None if paint else base + o_gmod, rec.ctypes.data, aux.ctypes.data))
    return rec, aux


def test_shell_record_source_on_host_matches_the_numpy_restatement_and_scipy():
    n = 20000
    ra, dec, M, z = synth.sky_halos(n, seed=21)
    z[:3] = [0.0, 0.4, 0.5]
    dec[:4] = [90.0, -90.0, 89.9999, 0.0]
    ra[:4] = [0.0, 359.9999, 180.0, 0.0]
    axes = synth.table_axes()
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=np.ones(12 * 64 * 64), cosmo=synth.COSMO)
    model = b.DisplacementModel(axes, synth.displacement_values(axes), 7, dict(synth.COSMO, Omega_m=0.33))   # two cosmologies
    pm = b.ProfileModel(axes, synth.profile_values(axes) * 3, synth.profile_values(axes))
    for paint, run in ((False, b.BaryonifyShell(cat, shell, 20, model, verbose=False)),
                       (True, b.PaintProfilesShell(cat, shell, 20, pm, include_pixel_size=True, verbose=False)),
                       (True, b.PaintProfilesShell(cat, shell, 20, pm, include_pixel_size=False, verbose=False))):
        want, _ = run.halo_records(paint=paint)
        sc = run.last_scalars
        got, aux = host_records(run, paint)
        for f in (_lib.HS_D, _lib.HS_LNZ, _lib.HS_LNM, _lib.HS_A):               # D_a(z_j): same PPoly, same operation order
            assert np.array_equal(got[:, f], want[:, f]), f
        for f in (_lib.HS_RADIUS, _lib.HS_RCUT, _lib.HS_SCALE, _lib.HS_LNRCOM):  # radius factor: 2048-node spline of g(ln(1+z))
            assert np.allclose(got[:, f], want[:, f], rtol=1e-12, atol=1e-13), f
        for f in (_lib.HS_VX, _lib.HS_VY, _lib.HS_VZ, _lib.HS_THETA, _lib.HS_PHI, _lib.HS_THETA_LL, _lib.HS_PHI_LL):
            d = np.abs(got[:, f] - want[:, f])
            if f == _lib.HS_PHI:                                                 # at the poles the azimuth is round-off dominated
                d = np.minimum(d, 2 * np.pi - d)[4:]
            assert d.max() < 1e-14, (f, d.max())
        assert np.all(got[:, _lib.HS_SKIP] == 0.0)
        assert np.allclose(aux[0], sc["R_run"], rtol=1e-12) and np.array_equal(aux[1], sc["D_A"])
        if not paint:
            assert np.allclose(aux[2], sc["R_model_com"], rtol=1e-12)
        # and against the sources themselves: the reference's CubicSpline object and the radius of the mass definition
        cosmo = cosmology.runner_cosmology(run.cosmo, with_w0=True)
        DA = cosmology.D_A_spline_to(cosmo, float(np.max(z)))
        assert np.array_equal(got[:, _lib.HS_D], DA(z))                          # scipy evaluates its own PPoly: bit-identical
        R = cosmology.radius_of_mass(cosmo, M, 1 / (1 + z), run.mass_def)
        assert np.allclose(got[:, _lib.HS_RADIUS], R * run.epsilon_max / got[:, _lib.HS_D], rtol=1e-12)
