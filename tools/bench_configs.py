#!/usr/bin/env python
"""
tools/bench_configs.py -- device-resident timings of the other BASELINE.json configs (1 GPU), for profiles/.

  C2  PaintProfilesShell  NSIDE=1024, 10^5 halos                         (16 B / update)
  C3  BaryonifyGrid       N^3 cells (default 1024), 10^6 halos, eps=20    (48 B / update; regrid 160 B / cell)
  C4  BaryonifySnapshot   n_part particles (default 2.5e8 = one GPU's share of 2e9), halos at the same number density
                          as 3e6 in (1000 Mpc)^3, displacement + NGP deposit   (72 B / pair)
Inputs are generated on the device (torch.rand = plumbing); every timed call goes through the C ABI.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def events(fn, warm=1, reps=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def snapshot_pipeline(n_part, snap_eps=5.0, device=0, reps=2, peak=6650.0, slab=None):
    """
    C4 (one GPU's share): BaryonifySnapshot on n_part uniform particles + halos at the number density of 3e6 per
    (1000 Mpc)^3 (M = 10^U(12,15.5)), then the NGP deposit -- cell list, halo loop, apply + un-permute, deposit, all through
    the C ABI with device-resident inputs.  Returns per-phase CUDA-event times and particles/s.

    slab = None: the share is its own periodic cube of side (n_part / density)^(1/3) (what one GPU of the box holds, as an
    independent problem).  slab = (rank, world): ONE periodic box of side (world n_part / density)^(1/3) (1000 Mpc for 8 x
    2.5e8), this rank owning the particles with x in [rank, rank + 1) L / world; the 3e6 halos of the whole box are drawn once
    (same seed on every rank) and the rank keeps those whose search sphere reaches its slab (replicated overlap halos); the
    NGP grid is full-size on every rank and the caller sums the partial grids (NCCL all-reduce).
    """
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, synth
    from baryonforge_b200.runners import _upload_records, _sort_records
    from baryonforge_b200.tables import displacement_table_of
    L = _lib.lib()
    dev = torch.device("cuda", device)
    st = torch.cuda.current_stream().cuda_stream
    n_part = int(n_part)
    dens = 2e9 / 1000.0 ** 3
    if slab is None:
        Lbox = (n_part / dens) ** (1 / 3.)
        x_lo, x_hi = 0.0, Lbox
    else:
        Lbox = (slab[1] * n_part / dens) ** (1 / 3.)
        x_lo, x_hi = Lbox * slab[0] / slab[1], Lbox * (slab[0] + 1) / slab[1]
    n_halo = int(round(3e6 * (Lbox / 1000.0) ** 3))
    pos, M = synth.box_halos(n_halo, Lbox, seed=42)
    gaxes = synth.table_axes(nz=10, nM=10, nr=500, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-3, r_max=3e2)
    model = b.DisplacementModel(gaxes, synth.displacement_values(gaxes), snap_eps, synth.COSMO)
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    ps = b.ParticleSnapshot(x=np.zeros(1), y=np.zeros(1), z=np.zeros(1), M=1.0, L=Lbox, redshift=0.3, cosmo=synth.COSMO)
    run = b.BaryonifySnapshot(cat, ps, snap_eps, model, verbose=False)
    rec, _ = run.halo_records()
    n_halo_box = n_halo
    if slab is not None:        # overlap halos: everything whose search sphere reaches the slab (periodic in x)
        xc, rq = rec[:, _lib.HB_X], rec[:, _lib.HB_RQ]
        mid, half = 0.5 * (x_lo + x_hi), 0.5 * (x_hi - x_lo)
        dxm = np.abs((xc - mid + Lbox / 2) % Lbox - Lbox / 2)
        rec = np.ascontiguousarray(rec[dxm <= half + rq])
        n_halo = rec.shape[0]
    ncell = run._pick_ncell(rec[:, _lib.HB_RQ], n_part * (1 if slab is None else slab[1]), 3, Lbox)
    tab = displacement_table_of(model, device)
    d_rec = _upload_records(rec, dev)
    d_rec, _ = _sort_records(d_rec, None, 1, Lbox, 16, 3)
    g = torch.Generator(device=dev); g.manual_seed(1)
    d_p = [torch.rand(n_part, dtype=torch.float64, device=dev, generator=g) * Lbox for _ in range(3)]
    if slab is not None:
        g.manual_seed(1 + slab[0])
        d_p[0] = x_lo + torch.rand(n_part, dtype=torch.float64, device=dev, generator=g) * (x_hi - x_lo)
    d_s = [torch.empty(n_part, dtype=torch.float64, device=dev) for _ in range(3)]
    d_start = torch.empty(ncell ** 3 + 1, dtype=torch.int64, device=dev)
    d_order = torch.empty(n_part, dtype=torch.int64, device=dev)
    d_tot = torch.zeros((3, n_part), dtype=torch.float64, device=dev)
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    d_m = torch.ones(n_part, dtype=torch.float64, device=dev)
    Ng = 512
    d_grid = torch.zeros(Ng ** 3, dtype=torch.float64, device=dev)

    def build():
        _lib.check(L.bfg_snap_build_cells(3, n_part, d_p[0].data_ptr(), d_p[1].data_ptr(), d_p[2].data_ptr(), Lbox, ncell,
                                          d_start.data_ptr(), d_order.data_ptr(), d_s[0].data_ptr(), d_s[1].data_ptr(),
                                          d_s[2].data_ptr(), st))

    def halos():
        d_tot.zero_()
        _lib.check(L.bfg_snap_offsets(tab.handle, 3, n_part, d_s[0].data_ptr(), d_s[1].data_ptr(), d_s[2].data_ptr(), Lbox,
                                      ncell, d_start.data_ptr(), n_halo, d_rec.data_ptr(), None, 0, d_tot.data_ptr(),
                                      d_n.data_ptr(), st))
    d_o = [torch.empty(n_part, dtype=torch.float64, device=dev) for _ in range(3)]

    def apply_dep():
        _lib.check(L.bfg_snap_apply(3, n_part, d_s[0].data_ptr(), d_s[1].data_ptr(), d_s[2].data_ptr(), d_tot.data_ptr(),
                                    d_order.data_ptr(), Lbox, d_o[0].data_ptr(), d_o[1].data_ptr(), d_o[2].data_ptr(), st))
        d_grid.zero_()
        _lib.check(L.bfg_snap_deposit_ngp(3, n_part, d_o[0].data_ptr(), d_o[1].data_ptr(), d_o[2].data_ptr(), d_m.data_ptr(),
                                          Lbox, Ng, d_grid.data_ptr(), st))

    def apply_dep_fused():      # process_to_map: deposit straight from the cell-ordered particles, no un-permute
        d_grid.zero_()
        _lib.check(L.bfg_snap_apply_deposit(3, n_part, d_s[0].data_ptr(), d_s[1].data_ptr(), d_s[2].data_ptr(),
                                            d_tot.data_ptr(), d_order.data_ptr(), None, 1.0, Lbox, Ng, d_grid.data_ptr(), st))
    ms_b = events(build, warm=1, reps=reps)
    ms_h = events(halos, warm=1, reps=reps)
    ms_a = events(apply_dep, warm=1, reps=reps)
    mass_a = float(d_grid.sum().item())
    npairs = int(d_n.cpu()[0])
    tot_ms = ms_b + ms_h + ms_a
    out = dict(n_part=n_part, L=Lbox, halos=n_halo, halos_in_box=n_halo_box, ncell=ncell, eps=snap_eps, pairs=npairs,
               slab=None if slab is None else [x_lo, x_hi], grid_handle=d_grid,
               host_inputs=dict(halo_x=rec[:, _lib.HB_X].copy(), halo_y=rec[:, _lib.HB_Y].copy(), halo_z=rec[:, _lib.HB_Z].copy(),
                                halo_M=np.exp(rec[:, _lib.HB_LNM]), x_lo=x_lo, x_hi=x_hi, gaxes=gaxes),
               build_cells_ms=ms_b, halo_loop_ms=ms_h, apply_deposit_ms=ms_a,
               particles_per_s=n_part / tot_ms * 1e3, pairs_per_s=npairs / ms_h * 1e3,
               halo_loop_alg_GBs=72 * npairs / ms_h / 1e6, halo_loop_frac=72 * npairs / ms_h / 1e6 / peak,
               deposited_mass=mass_a)
    try:
        ms_f = events(apply_dep_fused, warm=1, reps=reps)
        out.update(apply_deposit_fused_ms=ms_f, particles_per_s_to_map=n_part / (ms_b + ms_h + ms_f) * 1e3,
                   deposited_mass_fused=float(d_grid.sum().item()))
    except Exception as e:      # the catalogue-returning pipeline above stays the reported number
        out["apply_deposit_fused_error"] = str(e)[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c2,c3,c4")
    ap.add_argument("--grid-n", type=int, default=1024)
    ap.add_argument("--grid-halos", type=int, default=1000000)
    ap.add_argument("--npart", type=int, default=250000000)
    ap.add_argument("--snap-eps", type=float, default=5.0)
    args = ap.parse_args()
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, synth
    from baryonforge_b200.runners import _upload_records, _sort_records
    from baryonforge_b200.tables import displacement_table_of, profile_table_of
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
        if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
    out = {}

    if "c2" in args.which:
        nside, n = 1024, 100000
        ra, dec, M, z = synth.sky_halos(n, seed=42)
        axes = synth.table_axes()
        model = b.ProfileModel(axes, synth.profile_values(axes) * 3, synth.profile_values(axes))
        cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
        shell = b.LightconeShell(map=np.zeros(12 * nside * nside), cosmo=synth.COSMO)
        run = b.PaintProfilesShell(cat, shell, 20, model, verbose=False)
        rec, _ = run.halo_records(paint=True)
        tab = profile_table_of(model, '2D', 0)
        d_rec = _upload_records(rec, dev)
        d_rec, _ = _sort_records(d_rec, None, 0, b.runners.SKY_BAND_RAD)
        d_map = torch.zeros(12 * nside * nside, dtype=torch.float64, device=dev)
        d_n = torch.zeros(1, dtype=torch.int64, device=dev)

        def f():
            d_map.zero_()
            _lib.check(L.bfg_shell_paint(tab.handle, nside, n, d_rec.data_ptr(), None, 0, d_map.data_ptr(), 0, d_map.numel(),
                                         d_n.data_ptr(), st))
        ms = events(f)
        nu = int(d_n.cpu()[0])
        out["c2_paint_shell"] = dict(nside=nside, halos=n, updates=nu, ms=ms, updates_per_s=nu / ms * 1e3,
                                     alg_GBs=16 * nu / ms / 1e6, frac=16 * nu / ms / 1e6 / peak)
        import time
        t0 = time.perf_counter(); run.process(); torch.cuda.synchronize(); t1 = time.perf_counter()
        t0 = time.perf_counter(); run.process(); torch.cuda.synchronize(); t1 = time.perf_counter()
        out["c2_paint_shell"]["e2e_ms"] = 1e3 * (t1 - t0)

    if "c3" in args.which:
        N, n, Lbox = args.grid_n, args.grid_halos, 1000.0 * args.grid_n / 1024
        pos, M = synth.box_halos(n, Lbox, seed=42)
        bins = (np.arange(N) + 0.5) * Lbox / N
        gaxes = synth.table_axes(nz=10, nM=10, nr=500, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-3, r_max=3e2)
        model = b.DisplacementModel(gaxes, synth.displacement_values(gaxes) * 10, 20, synth.COSMO)
        cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
        gm = b.GriddedMap(map=np.broadcast_to(np.zeros(1), (N, N, N)), redshift=0.3, bins=bins, cosmo=synth.COSMO)
        run = b.BaryonifyGrid(cat, gm, 20, model, verbose=False)
        rec, _ = run.halo_records(paint=False)
        tab = displacement_table_of(model, 0)
        d_rec = _upload_records(rec, dev)
        d_rec, _ = _sort_records(d_rec, None, 1, Lbox, 16, 3)
        ncell = N ** 3
        d_off = torch.zeros((3, ncell), dtype=torch.float64, device=dev)
        d_map = torch.rand(ncell, dtype=torch.float64, device=dev) * 10
        d_new = torch.zeros(ncell, dtype=torch.float64, device=dev)
        d_n = torch.zeros(1, dtype=torch.int64, device=dev)
        d_s = torch.zeros(2, dtype=torch.float64, device=dev)

        def f1():
            d_off.zero_()
            _lib.check(L.bfg_grid_offsets(tab.handle, 3, N, float(gm.res), n, d_rec.data_ptr(), None, 0, 0, d_off.data_ptr(), 0, N,
                                          d_n.data_ptr(), st))

        def f2():
            d_new.zero_()
            _lib.check(L.bfg_grid_regrid(3, N, d_map.data_ptr(), d_off.data_ptr(), d_new.data_ptr(), 0, N, st))
        ms1 = events(f1)
        ms2 = events(f2)
        nu = int(d_n.cpu()[0])
        _lib.check(L.bfg_sum_f64(d_new.data_ptr(), ncell, d_s.data_ptr(), st))
        _lib.check(L.bfg_sum_f64(d_map.data_ptr(), ncell, d_s.data_ptr() + 8, st))
        sums = d_s.cpu().numpy()
        out["c3_grid_bary"] = dict(N=N, halos=n, updates=nu, offsets_ms=ms1, regrid_ms=ms2,
                                   updates_per_s=nu / ms1 * 1e3, offsets_alg_GBs=48 * nu / ms1 / 1e6,
                                   offsets_frac=48 * nu / ms1 / 1e6 / peak, regrid_alg_GBs=160 * ncell / ms2 / 1e6,
                                   regrid_frac=160 * ncell / ms2 / 1e6 / peak,
                                   mass_conserved=bool(np.isclose(sums[0], sums[1])), nan_cells=int(torch.isnan(d_off).sum().item()))
        del d_off, d_map, d_new

    if "c4" in args.which:
        out["c4_snapshot"] = snapshot_pipeline(args.npart, args.snap_eps, 0, 2, peak)
        out["c4_snapshot"].pop("grid_handle", None)
        out["c4_snapshot"].pop("host_inputs", None)

    print(json.dumps(out))


if __name__ == "__main__":
    main()
