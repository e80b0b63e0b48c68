"""Shared test plumbing: golden-case loading, oracle-port drivers, product-object builders, tolerances."""
import os
import warnings

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance for the fp64 path: relative 1e-6 on displacements / map values.  The absolute floor covers
# entries whose true value is a cancellation residue (e.g. nw_vec - vec ~ 1e-16) and scales with the data.
RTOL = 1e-6


def assert_close(got, want, what="", rtol=RTOL, atol_scale=1e-9):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs ({nan_g.sum()} vs {nan_w.sum()})"
    ok = ~nan_w
    scale = np.max(np.abs(want[ok])) if ok.any() else 0.0
    err = np.abs(got[ok] - want[ok])
    tol = rtol * np.abs(want[ok]) + atol_scale * scale
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()} / {bad.size} entries off; worst abs err {err.max():.3e} "
                           f"(scale {scale:.3e}), worst rel {np.max(err / np.maximum(np.abs(want[ok]), 1e-300)):.3e}")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: (z[k].item() if z[k].ndim == 0 else z[k]) for k in z.files}


def golden_names(kind):
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".npz"):
            z = np.load(os.path.join(GOLDEN, f))
            if str(z["kind"]) == kind:
                out.append(f[:-4])
    return out


# ---- oracle port drivers ----------------------------------------------------------------------------------
def port_run(g):
    """Run oracle/runners_port.py on a golden case's inputs; returns the same output(s) the reference produced."""
    from oracle import runners_port as rp
    kind = g["kind"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if kind == "shell_bary":
            tab = rp.DisplacementTable((g["ax0"], g["ax1"], g["ax2"]), g["values"], g["eps_mod"], bool(g["rdelta"]))
            cat = dict(M=g["M"], z=g["z"], ra=g["ra"], dec=g["dec"])
            return rp.baryonify_shell(int(g["nside"]), g["map"], cat, g["R_run"], g["D_A"], g["R_mod"], g["eps_run"], tab)
        if kind == "shell_paint":
            tab = rp.ProfileTable((g["ax0"], g["ax1"], g["ax2"]), g["raw3D"], g["raw2D"])
            cat = dict(M=g["M"], z=g["z"], ra=g["ra"], dec=g["dec"])
            return rp.paint_shell(int(g["nside"]), cat, g["R_run"], g["D_A"], g["eps_run"], tab, bool(g["pixsize"]))[0]
        if kind in ("grid_bary", "grid_paint"):
            ndim, N, L = int(g["ndim"]), int(g["N"]), float(g["L"])
            bins = (np.arange(N) + 0.5) * L / N
            a = 1 / (1 + g["redshift"])
            cat = dict(M=g["M"], x=g["x"], y=g["y"], z=g["z"])
            ell = (g["q_ell"], g["A_ell"]) if "q_ell" in g else None
            if kind == "grid_bary":
                tab = rp.DisplacementTable((g["ax0"], g["ax1"], g["ax2"]), g["values"], g["eps_mod"])
                return rp.baryonify_grid(g["map"], bins, cat, a, g["R_phys"], g["R_mod"], g["eps_run"], tab, ell=ell)
            tab = rp.ProfileTable((g["ax0"], g["ax1"], g["ax2"]), g["raw3D"], g["raw2D"])
            return rp.paint_grid((N,) * ndim, bins, cat, a, g["R_phys"] / a, g["eps_run"], tab, ell=ell)[0]
        if kind == "shell_anis":
            axes = (g["ax0"], g["ax1"], g["ax2"])
            cat = dict(M=g["M"], z=g["z"], ra=g["ra"], dec=g["dec"])
            return rp.paint_anis_shell(int(g["nside"]), g["map"], cat, g["R_run"], g["D_A"], g["eps_run"],
                                       rp.ProfileTable(axes, None, g["raw2D"]), rp.ProfileTable(axes, None, g["tracer2D"]),
                                       rp.ProfileTable(axes, None, g["mtot2D"]), float(g["dD"]), float(g["rho_m"]),
                                       float(g["proj_cutoff"]), float(g["background_val"]),
                                       float(g["global_tracer_fraction"]), bool(g["pixsize"]))[0]
        if kind == "grid_anis":
            axes = (g["ax0"], g["ax1"], g["ax2"])
            N, L = int(g["N"]), float(g["L"])
            bins = (np.arange(N) + 0.5) * L / N
            a = 1 / (1 + g["redshift"])
            cat = dict(M=g["M"], x=g["x"], y=g["y"], z=g["z"])
            ell = (g["q_ell"], g["A_ell"]) if "q_ell" in g else None
            return rp.paint_anis_grid(g["map"], bins, cat, a, g["R_phys"] / a, g["eps_run"],
                                      rp.ProfileTable(axes, None, g["raw2D"]), rp.ProfileTable(axes, None, g["tracer2D"]),
                                      rp.ProfileTable(axes, None, g["mtot2D"]), float(g["rho_m"]), float(g["proj_cutoff"]),
                                      float(g["background_val"]), float(g["global_tracer_fraction"]), bool(g["pixsize"]),
                                      ell=ell)[0]
        if kind == "snap":
            ndim = int(g["ndim"])
            a = 1 / (1 + g["redshift"])
            tab = rp.DisplacementTable((g["ax0"], g["ax1"], g["ax2"]), g["values"], g["eps_mod"])
            px = [g["px"], g["py"]] + ([g["pz"]] if ndim == 3 else [])
            cat = dict(M=g["M"], x=g["x"], y=g["y"], z=g["z"])
            out, _, _ = rp.baryonify_snapshot(px, float(g["L"]), cat, a, g["R_phys"], g["R_mod"], g["eps_run"], tab)
            return out
    raise ValueError(kind)


# ---- product objects from a golden case --------------------------------------------------------------------
MODEL_COSMO_KEYS = ("Omega_c", "Omega_b", "h", "sigma8", "n_s")


def product_run(g, **gpu_kwargs):
    """Run the CUDA path (baryonforge_b200 runners) on a golden case's inputs through the reference-shaped API."""
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    cosmo = synth.COSMO
    kind = g["kind"]
    axes = (g["ax0"], g["ax1"], g["ax2"])
    mc = dict(Omega_m=0.27 + 0.05, Omega_b=0.05, h=0.68, sigma8=0.82, n_s=0.97, w0=-1.0)   # oracle/make_golden.py MODEL_COSMO
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if kind == "shell_bary":
            model = b.DisplacementModel(axes, g["values"], g["eps_mod"], mc, Rdelta_sampling=bool(g["rdelta"]))
            cat = b.HaloLightConeCatalog(ra=g["ra"], dec=g["dec"], M=g["M"], z=g["z"], cosmo=cosmo)
            shell = b.LightconeShell(map=g["map"], cosmo=cosmo)
            return b.BaryonifyShell(cat, shell, g["eps_run"], model, verbose=False, **gpu_kwargs).process()
        if kind == "shell_paint":
            model = b.ProfileModel(axes, g["raw3D"], g["raw2D"])
            cat = b.HaloLightConeCatalog(ra=g["ra"], dec=g["dec"], M=g["M"], z=g["z"], cosmo=cosmo)
            shell = b.LightconeShell(map=np.zeros(12 * int(g["nside"]) ** 2), cosmo=cosmo)
            return b.PaintProfilesShell(cat, shell, g["eps_run"], model, include_pixel_size=bool(g["pixsize"]),
                                        verbose=False, **gpu_kwargs).process()
        if kind in ("grid_bary", "grid_paint"):
            ndim, N, L = int(g["ndim"]), int(g["N"]), float(g["L"])
            bins = (np.arange(N) + 0.5) * L / N
            ell = "q_ell" in g
            ekw = dict(q_ell=g["q_ell"], A_ell=g["A_ell"]) if ell else {}
            cat = b.HaloNDCatalog(x=g["x"], y=g["y"], z=g["z"] if ndim == 3 else None, M=g["M"],
                                  redshift=g["redshift"], cosmo=cosmo, **ekw)
            gm = b.GriddedMap(map=g["map"], redshift=g["redshift"], bins=bins, cosmo=cosmo)
            if kind == "grid_bary":
                model = b.DisplacementModel(axes, g["values"], g["eps_mod"], mc)
                return b.BaryonifyGrid(cat, gm, g["eps_run"], model, use_ellipticity=ell, verbose=False,
                                       **gpu_kwargs).process()
            model = b.ProfileModel(axes, g["raw3D"], g["raw2D"])
            return b.PaintProfilesGrid(cat, gm, g["eps_run"], model, use_ellipticity=ell, verbose=False,
                                       **gpu_kwargs).process()
        if kind == "shell_anis":
            model = b.ProfileModel(axes, None, g["raw2D"])
            tracer = b.ProfileModel(axes, None, g["tracer2D"])
            mtot = b.ProfileModel(axes, None, g["mtot2D"], proj_cutoff=float(g["proj_cutoff"]))
            cat = b.HaloLightConeCatalog(ra=g["ra"], dec=g["dec"], M=g["M"], z=g["z"], cosmo=cosmo)
            shell = b.LightconeShell(map=g["map"], cosmo=cosmo, redshift=float(g["z_shell"]))
            return b.PaintProfilesAnisShell(cat, shell, g["eps_run"], model, tracer, mtot, float(g["background_val"]),
                                            float(g["global_tracer_fraction"]), include_pixel_size=bool(g["pixsize"]),
                                            verbose=False, **gpu_kwargs).process()
        if kind == "grid_anis":
            N, L = int(g["N"]), float(g["L"])
            bins = (np.arange(N) + 0.5) * L / N
            ell = "q_ell" in g
            ekw = dict(q_ell=g["q_ell"], A_ell=g["A_ell"]) if ell else {}
            model = b.ProfileModel(axes, None, g["raw2D"])
            tracer = b.ProfileModel(axes, None, g["tracer2D"])
            mtot = b.ProfileModel(axes, None, g["mtot2D"], proj_cutoff=float(g["proj_cutoff"]))
            cat = b.HaloNDCatalog(x=g["x"], y=g["y"], z=None, M=g["M"], redshift=g["redshift"], cosmo=cosmo, **ekw)
            gm = b.GriddedMap(map=g["map"], redshift=g["redshift"], bins=bins, cosmo=cosmo)
            return b.PaintProfilesAnisGrid(cat, gm, g["eps_run"], model, tracer, mtot, float(g["background_val"]),
                                           float(g["global_tracer_fraction"]), include_pixel_size=bool(g["pixsize"]),
                                           use_ellipticity=ell, verbose=False, **gpu_kwargs).process()
        if kind == "snap":
            ndim = int(g["ndim"])
            cat = b.HaloNDCatalog(x=g["x"], y=g["y"], z=g["z"] if ndim == 3 else None, M=g["M"],
                                  redshift=g["redshift"], cosmo=cosmo)
            ps = b.ParticleSnapshot(x=g["px"], y=g["py"], z=g["pz"] if ndim == 3 else None, M=g["pM"], L=float(g["L"]),
                                    redshift=g["redshift"], cosmo=cosmo)
            model = b.DisplacementModel(axes, g["values"], g["eps_mod"], mc)
            out = b.BaryonifySnapshot(cat, ps, g["eps_run"], model, verbose=False, **gpu_kwargs).process()
            return [out["x"], out["y"]] + ([out["z"]] if ndim == 3 else [])
    raise ValueError(kind)
