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
