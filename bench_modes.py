"""
bench_modes.py -- the other BASELINE.json configs as `bench.py --config ...` lines (same JSON contract as the headline line).

  lightcone  configs[4]: 20 BaryonifyShell shells, NSIDE=4096, 10^6 halos per shell, on N GPUs.  Two ways to use the GPUs are
             timed through the runner API with host buffers and the faster one is the line's `e2e`:
               "shell-per-gpu"  every rank runs whole shells (the reference's model: utils/Parallelize.py:92-113 gives each
                                shell to one joblib worker) -- no communication at all;
               "ring-sharded"   all ranks cooperate on every shell (RING pixel ranges, fused re-binning + NVLink exchange).
  paint      configs[1]: PaintProfilesShell, NSIDE=1024, 10^5 halos, one GPU.

The CPU legs use the oracle port (oracle/runners_port.py), as the headline line does.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _sync_all(torch, dist, world):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(torch, dist, world, dev, x):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def _sum_over_ranks(torch, dist, world, dev, x):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


# ---------------------------------------------------------------------------------------------------------------------
def run_lightcone(args, bench):
    """bench = the bench.py module (peaks, ClockSampler, emit, claim_stdout, cpu legs)."""
    bench.claim_stdout()
    import torch
    import torch.distributed as dist
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, parallel, synth
    from baryonforge_b200.tables import displacement_table_of
    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    nside, n_shells, n_halo, eps = args.nside, args.shells, args.halos, args.eps
    npix = 12 * nside * nside
    axes = synth.table_axes()
    vals = synth.displacement_values(axes)
    model = b.DisplacementModel(axes, vals, eps, synth.COSMO)
    # one catalogue per shell (seed 42 + i: same distributions as the headline line), one input map shared by the shells
    cats = []
    for i in range(n_shells):
        ra, dec, M, z = synth.sky_halos(n_halo, seed=42 + i)
        cats.append(b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO))
    pinned_map = torch.empty(npix, dtype=torch.float64, pin_memory=True)
    pinned_map.numpy()[:] = synth.shell_map(nside, seed=7)
    shell = b.LightconeShell(map=pinned_map.numpy(), cosmo=synth.COSMO)
    mine = [i for i in range(n_shells) if i % world == rank]
    sampler = bench.ClockSampler(local)

    def per_gpu_pass():
        """shell-per-gpu: this rank's shells, one process() each (the single-GPU pipelined path)."""
        n_up = 0
        for i in mine:
            run = b.BaryonifyShell(cats[i], shell, eps, model, verbose=False, device=local)
            out = run.process()
            n_up += run.last_stats["n_updates"]
            del out
        return n_up

    lo, hi = parallel.pixel_ranges(nside, world)[rank]

    def sharded_pass():
        """ring-sharded: every shell on all ranks."""
        n_up = 0
        for i in range(n_shells):
            run = b.BaryonifyShell(cats[i], shell, eps, model, verbose=False, device=local, pix_range=(lo, hi))
            out = run.process()
            n_up += run.last_stats["n_updates"]          # already summed over the ranks
            del out
        return n_up

    results = {}
    # -- strategy A: shell per GPU ------------------------------------------------------------------------------------
    b.BaryonifyShell(cats[mine[0] if mine else 0], shell, eps, model, verbose=False, device=local).process()   # warm-up
    _sync_all(torch, dist, world)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    n_up_a = per_gpu_pass()
    torch.cuda.synchronize()
    dt_a = _max_over_ranks(torch, dist, world, dev, time.perf_counter() - t0)
    n_up_a = _sum_over_ranks(torch, dist, world, dev, n_up_a)
    results["shell-per-gpu"] = dict(seconds=dt_a, updates=n_up_a)
    # -- strategy B: every shell ring-sharded over all ranks ----------------------------------------------------------
    if world > 1:
        _sync_all(torch, dist, world)
        b.BaryonifyShell(cats[0], shell, eps, model, verbose=False, device=local, pix_range=(lo, hi)).process()   # warm-up
        _sync_all(torch, dist, world)
        t0 = time.perf_counter()
        n_up_b = sharded_pass()
        torch.cuda.synchronize()
        dt_b = _max_over_ranks(torch, dist, world, dev, time.perf_counter() - t0)
        results["ring-sharded"] = dict(seconds=dt_b, updates=float(n_up_b))
    clocks = sampler.stop() if rank == 0 else None

    # -- device-resident: this rank's shells back to back, records + map already in HBM (CUDA events) -------------------
    table = displacement_table_of(model, local)
    st = torch.cuda.current_stream().cuda_stream
    d_map = pinned_map.to(dev)
    d_off = torch.empty((3, npix), dtype=torch.float64, device=dev)
    d_new = torch.empty(npix, dtype=torch.float64, device=dev)
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    recs = []
    for i in mine:
        run = b.BaryonifyShell(cats[i], shell, eps, model, verbose=False, device=local)
        recs.append(run.device_records(paint=False, dev=dev))
    d_sorted = torch.empty_like(recs[0]) if recs else None
    launches = 0

    def device_pass():
        nonlocal launches
        tot = 0
        for d_rec in recs:
            d_off.zero_(); d_new.zero_()
            _lib.check(L.bfg_halo_sort(0, d_rec.shape[0], d_rec.data_ptr(), d_sorted.data_ptr(), None, None, 0,
                                       b.runners.SKY_BAND_RAD, 0.0, 3, st))
            _lib.check(L.bfg_shell_offsets(table.handle, nside, d_rec.shape[0], d_sorted.data_ptr(), None, 0, d_off.data_ptr(),
                                           0, npix, d_n.data_ptr(), st))
            _lib.check(L.bfg_shell_regrid(nside, d_map.data_ptr(), d_off.data_ptr(), d_new.data_ptr(), 0, npix, st))
            launches += 4
        return tot
    device_pass()
    _sync_all(torch, dist, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record(); device_pass(); e1.record()
    torch.cuda.synchronize()
    ms_dev = _max_over_ranks(torch, dist, world, dev, e0.elapsed_time(e1))
    n_launch = launches

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    best = min(results, key=lambda k: results[k]["seconds"])
    n_up = results["shell-per-gpu"]["updates"]
    peak, peak_src = bench.peaks()
    line = {"metric": "halo-pixel updates/s (BaryonifyShell lightcone, 20 shells)", "value": n_up / (ms_dev * 1e-3),
            "unit": "halo-pixel updates/s", "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"lightcone of {n_shells} BaryonifyShell shells, NSIDE={nside} npix={npix}, halos/shell={n_halo} "
                                   f"(seeds 42..{41 + n_shells}), table=10x10x500 epsilon_max={eps:g}, one shared input map U(0,10)",
                       "n_updates_per_step": int(n_up), "step": "the whole lightcone (all shells once)",
                       "device_resident_strategy": "shell-per-gpu (rank r runs shells r, r + N, ...)",
                       "l2_policy": "working set per shell (4.8 GB offsets + 3.2 GB maps) >> 126 MB L2; no flush needed"},
            "clocks": clocks, "gpu_launches": n_launch,
            "e2e": {"value": n_up / results[best]["seconds"], "unit": "halo-pixel updates/s", "strategy": best,
                    "shells_per_s": n_shells / results[best]["seconds"], "seconds_per_lightcone": results[best]["seconds"],
                    "h2d_bytes_per_step": int(n_shells * (npix * 8 + 6 * n_halo * 8)), "d2h_bytes_per_step": int(n_shells * npix * 8),
                    "strategies": {k: {"seconds_per_lightcone": v["seconds"], "shells_per_s": n_shells / v["seconds"],
                                       "updates_per_s": v["updates"] / v["seconds"]} for k, v in results.items()},
                    "includes": "per shell: host staging of the catalogue columns, H2D (pinned map + columns), device scalar prep, "
                                "sort, halo loop, re-binning (+ exchange when ring-sharded), D2H of the new map into host memory; "
                                "results stay with the rank that computed them (shell-per-gpu) or land in one shared host map "
                                "(ring-sharded)"},
            "roofline": None, "note": "roofline and cpu_baseline of the per-shell kernel: the default bench line (same shell workload)"}
    bench.emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def _paint_cpu_worker(job):
    nside, eps, sl, cat, R_run, D_A, axes, pvals = job
    import warnings
    from oracle import runners_port as rp
    tab = rp.ProfileTable(axes, pvals * 3.0, pvals)
    sub = {k: cat[k][sl] for k in ("M", "z", "ra", "dec")}
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, n_up = rp.paint_shell(nside, sub, R_run[sl], D_A[sl], eps, tab, False)
    return n_up, time.perf_counter() - t0


def run_paint(args, bench):
    """configs[1]: PaintProfilesShell NSIDE=1024, 10^5 halos, one B200."""
    bench.claim_stdout()
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, synth
    from baryonforge_b200.tables import profile_table_of
    assert int(os.environ.get("WORLD_SIZE", "1")) == 1, "--config paint is a single-GPU configuration (BASELINE configs[1])"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    nside, n, eps = 1024, 100000, 20.0
    npix = 12 * nside * nside
    ra, dec, M, z = synth.sky_halos(n, seed=42)
    axes = synth.table_axes()
    pvals = synth.profile_values(axes)
    model = b.ProfileModel(axes, pvals * 3.0, pvals)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    pinned = torch.zeros(npix, dtype=torch.float64, pin_memory=True)
    shell = b.LightconeShell(map=pinned.numpy(), cosmo=synth.COSMO)
    run = b.PaintProfilesShell(cat, shell, eps, model, include_pixel_size=False, verbose=False, device=0)
    # device-resident
    d_rec = run.device_records(paint=True, dev=dev)
    d_sorted = torch.empty_like(d_rec)
    tab = profile_table_of(model, '2D', 0)
    d_map = torch.zeros(npix, dtype=torch.float64, device=dev)
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    d_flush = torch.empty(64 * 1024 * 1024, dtype=torch.float64, device=dev)      # 512 MB > 126 MB L2

    def step(kev=None):
        d_flush.zero_()                                                           # the 100 MB map fits the L2: flush between steps
        d_map.zero_()
        _lib.check(L.bfg_halo_sort(0, n, d_rec.data_ptr(), d_sorted.data_ptr(), None, None, 0, b.runners.SKY_BAND_RAD, 0.0, 3, st))
        if kev:
            kev[0].record()
        _lib.check(L.bfg_shell_paint(tab.handle, nside, n, d_sorted.data_ptr(), None, 0, d_map.data_ptr(), 0, npix,
                                     d_n.data_ptr(), st))
        if kev:
            kev[1].record()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(0)
    sampler.start()
    kevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the flush is not part of the step: time the steps individually and sum
    ms_steps = []
    for k in range(args.steps):
        d_flush.zero_(); d_map.zero_()
        e0.record()
        _lib.check(L.bfg_halo_sort(0, n, d_rec.data_ptr(), d_sorted.data_ptr(), None, None, 0, b.runners.SKY_BAND_RAD, 0.0, 3, st))
        kevs[k][0].record()
        _lib.check(L.bfg_shell_paint(tab.handle, nside, n, d_sorted.data_ptr(), None, 0, d_map.data_ptr(), 0, npix,
                                     d_n.data_ptr(), st))
        kevs[k][1].record()
        e1.record()
        torch.cuda.synchronize()
        ms_steps.append(e0.elapsed_time(e1))
    clocks = sampler.stop()
    n_up = int(d_n.cpu()[0])
    ms_step = float(np.mean(ms_steps))
    ms_kernel = float(np.mean([a.elapsed_time(bb) for a, bb in kevs]))
    # e2e through the runner API
    for _ in range(2):
        out = run.process()
    ref = d_map.cpu().numpy()
    e2e_err = float(np.max(np.abs(out - ref)) / np.max(np.abs(ref)))
    del out
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        out = run.process()
        ts.append(time.perf_counter() - t0)
        del out
    dt = float(np.mean(ts))
    peak, peak_src = bench.peaks()
    achieved = 16.0 * n_up / (ms_kernel * 1e-3) / 1e9
    line = {"metric": "halo-pixel updates/s (PaintProfilesShell)", "value": n_up / (ms_step * 1e-3), "unit": "halo-pixel updates/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PaintProfilesShell NSIDE={nside} npix={npix} halos={n} table=10x10x500 epsilon_max={eps:g} "
                                   "catalogue=10^U(12,15.5) empty map", "n_updates_per_step": n_up,
                       "l2_policy": "512 MB written between timed steps (the 100 MB map would otherwise stay in the 126 MB L2)"},
            "clocks": clocks, "gpu_launches": 3 * args.steps,
            "roofline": {"bound": "hbm", "kernel": "k_shell_halos<paint>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "alg_bytes_per_update": 16.0,
                         "alg_bytes_per_launch": 16.0 * n_up, "kernel_ms": ms_kernel,
                         "note": "a 1.1e8-update launch lasts ~1.5 ms: per-halo set-up and the tail, not bandwidth, bound it"},
            "e2e": {"value": n_up / dt, "unit": "halo-pixel updates/s", "ms_per_step": 1e3 * dt,
                    "h2d_bytes_per_step": int(6 * n * 8), "d2h_bytes_per_step": int(npix * 8), "parity_vs_device_step": e2e_err}}
    if not args.no_cpu_baseline:
        sc = bench.cpu_scalars(cat, b.DisplacementModel(axes, synth.displacement_values(axes), eps, synth.COSMO), eps)
        sl = slice(0, min(20000, n))
        n_cpu, dt_cpu = _paint_cpu_worker((nside, eps, sl, cat.cat, sc["R_run"], sc["D_A"], axes, pvals))
        line["cpu_baseline"] = {"value": n_cpu / dt_cpu, "unit": "halo-pixel updates/s", "cores": 1, "kind": "port",
                                "sample": f"first {sl.stop} halos on the full NSIDE={nside} map, oracle/runners_port.paint_shell, "
                                          f"{dt_cpu:.1f} s, {n_cpu} updates"}
    bench.emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# Legs that the default `bench.py` line carries next to the headline (and that `--config grid` prints as a line of its own):
# BASELINE.json configs[2] (BaryonifyGrid 1024^3, slab-sharded) and configs[4] (the 20-shell lightcone), so that the driver's
# own N = 1, 2, 4, 8 runs measure them at the stated scale.
# ---------------------------------------------------------------------------------------------------------------------
def _host_map_with_pinned_slab(N, lo, hi, d_slab):
    """A full (N, N, N) float64 host map of which only this rank's planes [lo, hi) are ever touched: they are filled from
    `d_slab` and page-locked (bfg_host_register), so the runner's upload of orig_map[lo:hi] is a DMA from pinned memory while
    the other ranks' planes stay untouched zero pages (8 ranks x 8.6 GB of real host memory otherwise)."""
    from baryonforge_b200 import _lib
    full = np.zeros((N, N, N))
    full[lo:hi] = d_slab.reshape(hi - lo, N, N).cpu().numpy()
    slab = full[lo:hi]
    _lib.check(_lib.lib().bfg_host_register(slab.ctypes.data, slab.nbytes))
    return full, slab.ctypes.data


def _grid_cpu_sample(N, Lbox, n_halo_total, eps, gaxes, vals, n_sample=120):
    """Oracle port of the BaryonifyGrid halo loop (oracle/runners_port.grid_offsets, Map2DRunner.py:474-586) on the first
    n_sample halos of the same catalogue on the full N^3 grid (the offsets array is lazily zeroed memory)."""
    import warnings
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    pos, M = synth.box_halos(n_halo_total, Lbox, seed=42)
    sl = slice(0, n_sample)
    bins = (np.arange(N) + 0.5) * Lbox / N
    cat = b.HaloNDCatalog(x=pos[0][sl], y=pos[1][sl], z=pos[2][sl], M=M[sl], redshift=0.3, cosmo=synth.COSMO)
    model = b.DisplacementModel(gaxes, vals, eps, synth.COSMO)
    gm = b.GriddedMap(map=np.broadcast_to(np.zeros(1), (N, N, N)), redshift=0.3, bins=bins, cosmo=synth.COSMO)
    run = b.BaryonifyGrid(cat, gm, eps, model, verbose=False)
    run.halo_records(paint=False)
    sc = run.last_scalars
    hc = {k: cat.cat[k].astype('<f4') for k in ("M", "x", "y", "z")}
    tab = rp.DisplacementTable(gaxes, vals, eps)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, n_up = rp.grid_offsets((N, N, N), bins, hc, 1 / 1.3, sc["R_phys"], sc["R_model_com"], eps, tab, warn=False)
    return n_up, time.perf_counter() - t0


def grid_leg(args, bench, rank, world, local, cpu_baseline=True, reps=2):
    """configs[2]: BaryonifyGrid, 3-D periodic box of N^3 = 1024^3 cells, 10^6 halos, axis-0 slabs over the ranks.
    Device-resident pass (CUDA events, max over ranks) + end to end through BaryonifyGrid.process() with a host map."""
    import torch
    import torch.distributed as dist
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, parallel, synth
    from baryonforge_b200.runners import _upload_records, _sort_records
    from baryonforge_b200.tables import displacement_table_of
    dev = torch.device("cuda", local)
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    N, n, eps = int(args.grid_n), int(args.grid_halos), 20.0
    Lbox = 1000.0 * N / 1024
    lo, hi = parallel.plane_ranges(N, world)[rank]
    pos, M = synth.box_halos(n, Lbox, seed=42)
    bins = (np.arange(N) + 0.5) * Lbox / N
    gaxes = synth.table_axes(nz=10, nM=10, nr=500, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-3, r_max=3e2)
    vals = synth.displacement_values(gaxes) * 10
    model = b.DisplacementModel(gaxes, vals, eps, synth.COSMO)
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    nloc = (hi - lo) * N * N
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)
    d_map = torch.rand(nloc, dtype=torch.float64, device=dev, generator=g) * 10
    host_map, reg_addr = _host_map_with_pinned_slab(N, lo, hi, d_map)
    try:
        gm = b.GriddedMap(map=host_map, redshift=0.3, bins=bins, cosmo=synth.COSMO)
        run = b.BaryonifyGrid(cat, gm, eps, model, verbose=False, device=local, plane_range=None if world == 1 else (lo, hi))
        # ---- device-resident pass ----
        rec, _ = run.halo_records(paint=False)
        if world > 1:
            rec = np.ascontiguousarray(rec[parallel.halos_touching_planes(N, rec[:, _lib.HB_CX], rec[:, _lib.HB_NSIZE], lo, hi)])
        tab = displacement_table_of(model, local)
        d_rec, _ = _sort_records(_upload_records(rec, dev), None, 1, Lbox, 16, 3)
        d_off = torch.zeros((3, nloc), dtype=torch.float64, device=dev)
        d_new = torch.zeros(N ** 3, dtype=torch.float64, device=dev)
        d_own = torch.empty(nloc, dtype=torch.float64, device=dev) if world > 1 else None
        d_n = torch.zeros(1, dtype=torch.int64, device=dev)
        d_s = torch.zeros(2, dtype=torch.float64, device=dev)

        def step(ev=None):
            d_off.zero_(); d_new.zero_()
            if ev:
                ev[0].record()
            _lib.check(L.bfg_grid_offsets(tab.handle, 3, N, float(gm.res), rec.shape[0], d_rec.data_ptr(), None, 0, 0,
                                          d_off.data_ptr(), lo, hi, d_n.data_ptr(), st))
            if ev:
                ev[1].record()
            _lib.check(L.bfg_grid_regrid(3, N, d_map.data_ptr(), d_off.data_ptr(), d_new.data_ptr(), lo, hi, st))
            if ev:
                ev[2].record()
            if world > 1:      # the CIC deposit of a slab reaches into its neighbours: sum the partial maps, keep the own slab
                dist.reduce_scatter_tensor(d_own, d_new, op=dist.ReduceOp.SUM)
            res = d_own if world > 1 else d_new
            _lib.check(L.bfg_sum_f64(res.data_ptr(), res.numel(), d_s.data_ptr(), st))
            _lib.check(L.bfg_sum_f64(d_map.data_ptr(), nloc, d_s.data_ptr() + 8, st))
            if ev:
                ev[3].record()

        step()
        _sync_all(torch, dist, world)
        times = []
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            _sync_all(torch, dist, world)
            step(ev)
            _sync_all(torch, dist, world)
            times.append([ev[0].elapsed_time(ev[3]), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])])
        t = torch.tensor(np.mean(np.array(times), axis=0), dtype=torch.float64, device=dev)
        cnt = torch.tensor([float(d_n.cpu()[0]), float(rec.shape[0])], dtype=torch.float64, device=dev)
        sums = d_s.clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt)
            dist.all_reduce(sums)
        n_up = float(cnt[0])
        mass_ok = bool(np.isclose(float(sums[0]), float(sums[1]), rtol=1e-9))
        stride = max(1, nloc // 100000)
        ref_s = (d_own if world > 1 else d_new)[::stride].cpu().numpy()
        del d_off, d_new, d_own, d_rec
        torch.cuda.empty_cache()
        # ---- end to end: BaryonifyGrid.process() on the host map (this rank's planes page-locked) ----
        out = run.process()                               # warm-up: result buffers / shared host map
        got_s = np.asarray(out)[lo:hi].reshape(-1)[::stride].copy()    # a view would keep the result's buffer alive
        e2e_err = float(np.max(np.abs(got_s - ref_s) / (np.abs(ref_s) + 1e-3 * np.max(np.abs(ref_s)))))
        del out
        ts = []
        for _ in range(reps):
            _sync_all(torch, dist, world)
            t0 = time.perf_counter()
            out = run.process()
            del out
            ts.append(time.perf_counter() - t0)
        dt = _max_over_ranks(torch, dist, world, dev, float(np.mean(ts)))
    finally:
        _lib.lib().bfg_host_unregister(reg_addr)
    peak, peak_src = bench.peaks()
    ms_pass, ms_loop = float(t[0]), float(t[1])
    achieved = 48.0 * (n_up / world) / (ms_loop * 1e-3) / 1e9
    leg = {"metric": "halo-cell updates/s (BaryonifyGrid)", "value": n_up / (ms_pass * 1e-3), "unit": "halo-cell updates/s",
           "n_gpus": world, "ms_per_step": ms_pass, "scaling": "strong", "dtype": "f64",
           "config": {"workload": f"BaryonifyGrid {N}^3 cells, L = {Lbox:.0f} Mpc, {n} halos (M = 10^U(12,15.5)), table=10x10x500, "
                                  f"epsilon_max={eps:g}, map U(0,10)", "n_updates_per_step": int(n_up),
                      "sharding": "none" if world == 1 else f"axis-0 slabs x{world}, halos replicated by cutout overlap "
                                  f"({int(cnt[1])} records over all ranks), partial maps summed by NCCL reduce-scatter (every rank keeps its slab)"},
           "phases_ms": {"halo_loop (binning + k_tile_gather)": ms_loop, "regrid (k_grid_regrid)": float(t[2]),
                         "exchange_and_sums": float(t[3])},
           "mass_conserved": mass_ok,
           "roofline": {"bound": "hbm", "kernel": "k_tile_gather (+ tile binning, row blend)", "achieved": achieved, "peak": peak,
                        "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                        "alg_bytes_per_update": 48.0, "kernel_ms": ms_loop,
                        "note": "48 B/update is the reference dataflow (3 f64 read-modify-writes per halo-cell update, SURVEY 8d). The "
                                "tile-centric gather keeps a cell's 3 accumulators in registers and writes each cell once, so DRAM "
                                "traffic is ~58 GB per pass against 17.8 TB algorithmic (profiles/r1_tile_gather_ncu_summary.txt) and "
                                "`frac` > 1 says only that; the kernel is arithmetic-bound (issue 68 %, FP64 pipe 38 %)"},
           "e2e": {"value": n_up / dt, "unit": "halo-cell updates/s", "ms_per_step": 1e3 * dt,
                   "h2d_bytes_per_step": int(N ** 3 * 8 + world * 4 * n * 4), "d2h_bytes_per_step": int(N ** 3 * 8),
                   "parity_vs_device_step": e2e_err,
                   "api": "BaryonifyGrid.process(): host map in (each rank's planes page-locked), new map out in host memory"
                          + ("" if world == 1 else " (one page-locked map shared by the ranks of the box, each rank writes its slab)")}}
    if cpu_baseline and rank == 0 and world == 1:
        try:
            n_cpu, dt_cpu = _grid_cpu_sample(N, Lbox, n, eps, gaxes, vals)
            leg["cpu_baseline"] = {"value": n_cpu / dt_cpu, "unit": "halo-cell updates/s", "cores": 1, "kind": "port",
                                   "sample": f"first 120 halos of the same catalogue on the full {N}^3 grid, halo loop only "
                                             f"(oracle/runners_port.grid_offsets), {dt_cpu:.1f} s, {n_cpu} updates"}
        except Exception as exc:
            leg["cpu_baseline"] = {"error": str(exc)[:300]}
    return leg


def lightcone_leg(args, bench, rank, world, local, n_shells=20):
    """configs[4]: a lightcone of 20 BaryonifyShell shells (NSIDE and halos per shell of the headline), whole shells per GPU
    (the reference's model, utils/Parallelize.py:92-113; measured faster than ring-sharding every shell: profiles/r2_lightcone_n8.json),
    end to end through process() with host buffers.  Returns seconds per lightcone, max over ranks."""
    import torch
    import torch.distributed as dist
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    dev = torch.device("cuda", local)
    nside, n_halo, eps = args.nside, args.halos, args.eps
    npix = 12 * nside * nside
    axes = synth.table_axes()
    model = b.DisplacementModel(axes, synth.displacement_values(axes), eps, synth.COSMO)
    mine = [i for i in range(n_shells) if i % world == rank]
    cats = {}
    for i in mine:
        ra, dec, M, z = synth.sky_halos(n_halo, seed=42 + i)
        cats[i] = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    pinned_map = torch.empty(npix, dtype=torch.float64, pin_memory=True)
    pinned_map.numpy()[:] = synth.shell_map(nside, seed=7)
    shell = b.LightconeShell(map=pinned_map.numpy(), cosmo=synth.COSMO)
    b.BaryonifyShell(cats[mine[0]], shell, eps, model, verbose=False, device=local).process()   # warm-up
    _sync_all(torch, dist, world)
    t0 = time.perf_counter()
    n_up = 0
    for i in mine:
        run = b.BaryonifyShell(cats[i], shell, eps, model, verbose=False, device=local)
        out = run.process()
        n_up += run.last_stats["n_updates"]
        del out
    torch.cuda.synchronize()
    dt = _max_over_ranks(torch, dist, world, dev, time.perf_counter() - t0)
    n_up = _sum_over_ranks(torch, dist, world, dev, n_up)
    return {"metric": "halo-pixel updates/s (BaryonifyShell lightcone)", "value": n_up / dt, "unit": "halo-pixel updates/s",
            "n_gpus": world, "scaling": "strong", "dtype": "f64", "seconds_per_lightcone": dt, "shells_per_s": n_shells / dt,
            "config": {"workload": f"lightcone of {n_shells} BaryonifyShell shells, NSIDE={nside}, halos/shell={n_halo} (seeds 42.."
                                   f"{41 + n_shells}), table=10x10x500 epsilon_max={eps:g}, one shared input map U(0,10)",
                       "n_updates_per_step": int(n_up), "strategy": "shell-per-gpu (rank r runs shells r, r + N, ...), no collective"},
            "e2e": {"value": n_up / dt, "unit": "halo-pixel updates/s", "h2d_bytes_per_step": int(n_shells * (npix * 8 + 6 * n_halo * 8)),
                    "d2h_bytes_per_step": int(n_shells * npix * 8),
                    "api": "one BaryonifyShell.process() per shell: host catalogue + pinned host map in, new map out in host memory"}}


def run_grid(args, bench):
    """`bench.py --config grid`: the grid leg as a JSON line of its own (same contract as the headline line)."""
    bench.claim_stdout()
    import torch
    import torch.distributed as dist
    from baryonforge_b200 import parallel
    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    leg = grid_leg(args, bench, rank, world, local, cpu_baseline=not args.no_cpu_baseline, reps=max(1, args.steps))
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.barrier()
    if rank == 0:
        line = {"metric": leg["metric"], "value": leg["value"], "unit": leg["unit"], "n_gpus": world, "steps": max(1, args.steps),
                "warmup": 1, "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "clocks": clocks,
                "gpu_launches": 8 * max(1, args.steps)}
        line.update({k: v for k, v in leg.items() if k not in line})
        line["config"]["l2_policy"] = "working set (26 GB offsets + 17 GB maps) >> 126 MB L2; no flush needed"
        bench.emit(line)
    if world > 1:
        dist.destroy_process_group()
