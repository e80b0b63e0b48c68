#!/usr/bin/env python
"""
bench.py -- headline benchmark of the runner hot path (driver contract in the task statement).

Workload (BASELINE.json `metric` / configs[4] per shell): ONE BaryonifyShell shell at NSIDE=4096
(201 326 592 pixels) with 10^6 synthetic halos (M = 10^U(12,15.5), z = U(0.4,0.5), uniform sky, seed 42), table
10x10x500, epsilon_max = 20, map U(0,10).  A "step" = one full pass of the hot path over the catalogue: halo loop
(fused disc / separation / table / accumulate kernel) + re-binning + mass-conservation sums.

  metric `value`  : halo-pixel updates / s, inputs already resident in HBM (CUDA events, max over ranks)
  `e2e`           : the same through BaryonifyShell(...).process() with HOST (pinned) buffers: host scalar prep,
                    H2D of halo records + map, kernels, D2H of the new map -- all inside the timed region
  `roofline`      : fused halo-loop kernel, algorithmic 48 B per update / its CUDA-event time, against
                    MEASURED_PEAKS.json hbm_gbs
  `cpu_baseline`  : oracle/runners_port.py (the reference's per-halo Python loop, restated) on a halo subsample
  --impl reference: that same CPU path on every host core (one process per core, disjoint halo subsamples -- the
                    reference's own parallel model for Baryonify runners is one process per map, Parallelize.py:206-209)

N > 1 (torchrun): the shell is sharded by RING pixel range, overlap halos replicated, the re-binning deposits straight into
the owning rank's slice over NVLink peer memory ("scaling": "strong": the total work is fixed).  Rank 0 then recomputes the shell
un-sharded and the line carries `parity_vs_n1` (max relative difference over a strided pixel sample + the sums).
--config lightcone | paint: the other BASELINE configs (bench_modes.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_UPDATE = 48.0        # 3 x f64 read-modify-write of pix_offsets
# Per-round ncu evidence of the dominant kernel (written by tools/ncu_summary.py --json from the round's `ncu --set full`
# capture of THIS bench command): DRAM bytes per launch, FP64-pipe instructions per update, pipe utilisations.  bench.py
# only reports what that file holds for the workload it is running -- nothing here is a literal.
NCU_FACTS = os.path.join(ROOT, "profiles", "shell_halos_ncu_facts.json")
B200_SMS = 148
FP64_LANES_PER_SM_CLK = 64         # 4 sub-partitions x 16 lanes: one FP64 warp instruction per 2 cycles per scheduler


def ncu_facts(workload):
    """The committed ncu facts for `workload` (exact string match on config.workload), else {}."""
    try:
        facts = json.load(open(NCU_FACTS))
    except Exception:
        return {}
    for f in facts.get("captures", []):
        if f.get("workload") == workload:
            return f
    return {}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nside", type=int, default=4096)
    ap.add_argument("--halos", type=int, default=1000000)
    ap.add_argument("--eps", type=float, default=20.0)
    ap.add_argument("--cpu-sample", type=int, default=3000, help="halos in the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sort", action="store_true", help="skip the device-side sky ordering of halos (L2 locality)")
    ap.add_argument("--mass-function", action="store_true", help="steeper dn/dlogM ~ M^-0.9 catalogue variant")
    ap.add_argument("--gather-result", action="store_true",
                    help="N > 1: all-gather the new map onto every rank inside the device-resident step (round 1's step).  "
                         "Default: the map stays distributed over the ranks that own its slices, as in the end-to-end path, "
                         "which downloads each slice over its own PCIe link and never gathers on the device")
    ap.add_argument("--no-particles", action="store_true",
                    help="skip the secondary metric (BaryonifySnapshot particles displaced/s, weak scaling)")
    ap.add_argument("--particles-per-gpu", type=int, default=250000000)
    ap.add_argument("--grid-n", type=int, default=1024, help="grid leg / --config grid: cells per axis")
    ap.add_argument("--grid-halos", type=int, default=1000000, help="grid leg / --config grid: halos in the box")
    ap.add_argument("--leg-timeout", type=float, default=240.0, help="seconds after which a hanging extra leg is abandoned")
    ap.add_argument("--no-extra-configs", action="store_true",
                    help="skip the BASELINE configs[2] (BaryonifyGrid 1024^3) and configs[4] (20-shell lightcone) legs of the default line")
    ap.add_argument("--config", default="shell", choices=["shell", "lightcone", "paint", "grid"],
                    help="shell = the headline line (one BaryonifyShell shell, BASELINE metric); lightcone = configs[4] (20 shells "
                         "on N GPUs); paint = configs[1] (PaintProfilesShell NSIDE=1024, 10^5 halos) -- bench_modes.py")
    ap.add_argument("--shells", type=int, default=20, help="--config lightcone: shells in the lightcone")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(args):
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    ra, dec, M, z = synth.sky_halos(args.halos, seed=42, mass_function=args.mass_function)
    axes = synth.table_axes()
    vals = synth.displacement_values(axes)
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    model = b.DisplacementModel(axes, vals, args.eps, synth.COSMO)
    return cat, model, axes, vals


def workload_name(args):
    return (f"BaryonifyShell NSIDE={args.nside} npix={12 * args.nside ** 2} halos={args.halos} table=10x10x500 "
            f"epsilon_max={args.eps:g} catalogue={'dn/dlogM~M^-0.9' if args.mass_function else '10^U(12,15.5)'} map=U(0,10)")


L2_POLICY = "working set (4.8 GB offsets + 3.2 GB maps per step) >> 126 MB L2; no flush needed"
# sum_j |query_disc_j| of the seeded workloads, as counted by the device AND by the oracle (tests/test_gpu_parity.py::
# test_full_size_properties_nside4096): lets the reference arm -- which only runs a sample -- name the same `config`
N_UPDATES_KNOWN = {
    "BaryonifyShell NSIDE=4096 npix=201326592 halos=1000000 table=10x10x500 epsilon_max=20 catalogue=10^U(12,15.5) map=U(0,10)":
        17746618419,
}


def sharding_text(world, p2p=True, gather=False):
    if world == 1:
        return "none"
    return f"RING pixel ranges x{world}, overlap halos replicated, " + (
        "re-binning fused with the exchange (fp64 REDs into the owner's slice over NVLink peer memory) + " + (
            "NCCL all-gather of slices" if gather
            else "new map left distributed over the ranks that own its slices (as in the end-to-end path)")
        if p2p else "NCCL all-reduce of partial maps")


def shell_config(args, world, n_up=None, p2p=True, gather=False):
    """`config` of the headline line; both arms build it here, so the reference arm runs on the GPU arm's `config` by construction."""
    w = workload_name(args)
    return {"workload": w, "n_updates_per_step": int(n_up) if n_up is not None else N_UPDATES_KNOWN.get(w),
            "l2_policy": L2_POLICY, "sharding": sharding_text(world, p2p, gather)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arms (oracle port)
# ---------------------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    nside, eps, sl, cat, R_run, D_A, R_mod, axes, vals = job
    import warnings
    from oracle import runners_port as rp
    tab = rp.DisplacementTable(axes, vals, eps)
    sub = {k: cat[k][sl] for k in ("M", "z", "ra", "dec")}
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, n_up = rp.shell_offsets(nside, sub, R_run[sl], D_A[sl], R_mod[sl], eps, tab, warn=True)
    return n_up, time.perf_counter() - t0


def cpu_scalars(cat, model, eps):
    """Per-halo scalars (inputs shared by both arms), from the product's host prep -- no GPU involved."""
    import baryonforge_b200 as b
    shell = b.LightconeShell(map=np.zeros(12), cosmo=cat.cosmo)
    run = b.BaryonifyShell(cat, shell, eps, model, verbose=False)
    run.halo_records(paint=False)
    return run.last_scalars


def cpu_regrid_sample(nside, n_pix=4000000):
    """The reference's re-binning step (HealpixRunner.py:357-365, restated in oracle/runners_port.shell_regrid) on the first
    n_pix pixels of the map with zero offsets: pix2vec, vec2ang, get_interp_weights, scatter.  Returns (pixels, seconds)."""
    from oracle import hpo
    from oracle import runners_port as rp
    npix = 12 * nside * nside
    n_pix = min(n_pix, npix)
    m = np.random.default_rng(7).uniform(0, 10, n_pix)
    t0 = time.perf_counter()
    vec = np.stack(hpo.pix2vec_range(nside, 0, n_pix), axis=1)
    dnorm = np.sqrt(np.sum(np.square(vec), axis=1))
    theta = np.arccos(vec[:, 2] / dnorm)
    phi = np.arctan2(vec[:, 1], vec[:, 0])
    phi[phi < 0] += 2 * np.pi
    lon, lat = np.degrees(phi), 90.0 - np.degrees(theta)
    c_pix, c_w = rp._interp_weights_lonlat(nside, lon, lat)
    new_map = np.zeros(npix)
    hpo.regrid_scatter(new_map, m, np.ascontiguousarray(c_pix.T), np.ascontiguousarray(c_w.T))
    return n_pix, time.perf_counter() - t0


def cpu_baseline(args, cat, model, axes, vals, n_sample):
    sc = cpu_scalars(cat, model, args.eps)
    sl = slice(0, min(n_sample, len(cat)))
    n_up, dt = _cpu_worker((args.nside, args.eps, sl, cat.cat, sc["R_run"], sc["D_A"], sc["R_model_com"], axes, vals))
    out = {"value": n_up / dt, "unit": "halo-pixel updates/s", "cores": 1, "kind": "port",
           "sample": f"first {sl.stop} halos of the same catalogue on the full NSIDE={args.nside} map, HALO LOOP ONLY "
                     f"(oracle/runners_port.shell_offsets; the re-binning is timed separately below), {dt:.1f} s, "
                     f"{n_up} updates, {sl.stop / dt:.0f} halos/s"}
    try:    # the other half of process(): the re-binning, on a pixel sample, so that a whole-shell CPU time can be stated
        n_pix, dt_r = cpu_regrid_sample(args.nside)
        npix = 12 * args.nside ** 2
        loop_s = len(cat) / (sl.stop / dt)
        out["regrid"] = {"pixels_per_s": n_pix / dt_r, "sample": f"first {n_pix} pixels, zero offsets, {dt_r:.1f} s, 1 core"}
        out["whole_shell_estimate_s"] = {"halo_loop": loop_s, "regrid": npix / (n_pix / dt_r),
                                         "note": "1 core, scaled linearly from the two samples; the GPU e2e line covers both"}
    except Exception as e:
        out["regrid"] = {"error": str(e)[:200]}
    return out


def run_reference(args):
    """--impl reference: the CPU path on all host cores, one process per core, disjoint halo subsamples per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    claim_stdout()
    import multiprocessing as mp
    cat, model, axes, vals = make_inputs(args)
    sc = cpu_scalars(cat, model, args.eps)
    cores = os.cpu_count() or 1
    # ~4.8 GB of lazily-touched pix_offsets per worker at NSIDE=4096: cap workers by RAM
    try:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2 ** 30
    except Exception:
        mem_gb = 64
    per_worker_gb = 12 * args.nside ** 2 * 24 / 2 ** 30 + 1.0
    workers = int(max(1, min(cores, mem_gb * 0.6 // per_worker_gb)))
    per = max(50, min(1200, len(cat) // (workers * (args.steps + args.warmup) + 1)))
    ctx = mp.get_context("fork")
    times, ups = [], []
    with ctx.Pool(workers) as pool:
        k = 0
        for step in range(args.warmup + args.steps):
            jobs = []
            for w in range(workers):
                sl = slice(k, k + per); k += per
                jobs.append((args.nside, args.eps, sl, cat.cat, sc["R_run"], sc["D_A"], sc["R_model_com"], axes, vals))
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt); ups.append(sum(r[0] for r in res))
    value = sum(ups) / sum(times)
    sample = (f"{per} halos per worker x {workers} workers per step on the full NSIDE={args.nside} map, halo loop only, "
              f"oracle port of the reference's Python loop")
    line = {"impl": "reference", "metric": "halo-pixel updates/s (BaryonifyShell)", "value": value,
            "unit": "halo-pixel updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shell_config(args, max(1, int(args.gpus))), "timing": "wall clock around a multiprocessing map",
            "cpu_baseline": {"value": value, "unit": "halo-pixel updates/s", "cores": workers, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "halo-pixel updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: anything libraries write to fd 1 (NCCL prints its "NCCL version ..." banner
    there) is sent to stderr, and emit() writes the line to the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def particles_e2e(args, world, rank, local, dev, c4, host_in, barrier):
    """particles displaced/s through BaryonifySnapshot.process() / process_to_map() with HOST inputs and outputs: this rank's
    slab of the box as a ParticleSnapshot (structured array), the overlap halos as a HaloNDCatalog."""
    import torch
    import torch.distributed as dist
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    n = int(args.particles_per_gpu)
    try:
        avail = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0] / 2 ** 20
    except Exception:
        avail = 64.0
    need = world * n * 8 * 12 / 2 ** 30          # per rank: x, y, z (24 B) + input catalogue (32 B) + output (32 B) + slack
    if need > 0.6 * avail:
        return {"skipped": f"host RAM: {need:.0f} GB needed for {world} x {n} particles, {avail:.0f} GB available"}
    Lbox = float(c4["L"])
    rng = np.random.default_rng(1000 + rank)
    x = host_in["x_lo"] + rng.random(n) * (host_in["x_hi"] - host_in["x_lo"])
    y, z = rng.random(n) * Lbox, rng.random(n) * Lbox
    ps = b.ParticleSnapshot(x=x, y=y, z=z, M=1.0, L=Lbox, redshift=0.3, cosmo=synth.COSMO)
    del x, y, z
    gaxes = host_in["gaxes"]
    model = b.DisplacementModel(gaxes, synth.displacement_values(gaxes), 5.0, synth.COSMO)
    cat = b.HaloNDCatalog(x=host_in["halo_x"], y=host_in["halo_y"], z=host_in["halo_z"], M=host_in["halo_M"], redshift=0.3,
                          cosmo=synth.COSMO)
    run = b.BaryonifySnapshot(cat, ps, 5.0, model, verbose=False, device=local)
    out = run.process()                           # warm-up (pinned staging buffers, table)
    del out
    barrier()
    t0 = time.perf_counter()
    out = run.process()
    dt_cat = time.perf_counter() - t0
    moved = float(np.abs(out["x"][:1000000] - ps.cat["x"][:1000000]).max())
    del out
    run.process_to_map(512)                       # warm-up (the grid's page-locked result buffer), like process() above
    barrier()
    t0 = time.perf_counter()
    grid = run.process_to_map(512)
    dt_map = time.perf_counter() - t0
    ok = bool(abs(float(grid.sum()) - n) < 0.5)
    del grid
    t = torch.tensor([dt_cat, dt_map], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"value": world * n / float(t[0]), "unit": "particles/s", "seconds_per_pass": float(t[0]),
            "h2d_bytes_per_step": int(world * n * 32), "d2h_bytes_per_step": int(world * n * 32),
            "api": "BaryonifySnapshot.process(): host structured array in (pageable, the reference's 32-byte M, x, y, z records, moved "
                   "as raw bytes), structured array of displaced particles out",
            "to_map": {"value": world * n / float(t[1]), "seconds_per_pass": float(t[1]),
                       "api": "BaryonifySnapshot.process_to_map(512): host particles in, host NGP grid out (per rank; N > 1: the "
                              "partial grids still have to be summed, parallel.deposit_ngp_all)",
                       "deposited_mass_equals_particle_count": ok},
            "pairs": int(run.last_stats.get("n_pairs", 0)), "max_displacement_checked": moved}


def particles_cpu_baseline(gaxes, n_cpu=2000000):
    """The reference's BaryonifySnapshot on the host (oracle/runners_port.baryonify_snapshot: scipy KDTree + the Python halo
    loop) on a sub-box of the same particle and halo density: 2e6 particles in (100 Mpc)^3 with 3000 halos."""
    import warnings
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    from oracle import runners_port as rp
    from scipy.spatial import KDTree
    dens = 2e9 / 1000.0 ** 3
    Lc = (n_cpu / dens) ** (1 / 3.)
    n_halo = int(round(3e6 * (Lc / 1000.0) ** 3))
    pos, M = synth.box_halos(n_halo, Lc, seed=42)
    p = np.random.default_rng(5).uniform(0, Lc, (3, n_cpu))
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    ps = b.ParticleSnapshot(x=p[0], y=p[1], z=p[2], M=1.0, L=Lc, redshift=0.3, cosmo=synth.COSMO)
    model = b.DisplacementModel(gaxes, synth.displacement_values(gaxes), 5.0, synth.COSMO)
    run = b.BaryonifySnapshot(cat, ps, 5.0, model, verbose=False)
    run.halo_records()
    sc = run.last_scalars
    hc = {k: cat.cat[k].astype('<f4') for k in ("M", "x", "y", "z")}
    t0 = time.perf_counter()
    tree = KDTree(p.T, boxsize=Lc, leafsize=1000)                 # examples/10: KDTree_kwargs = {'leafsize': 1e3}
    t_tree = time.perf_counter() - t0
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, n_pairs, _ = rp.baryonify_snapshot(list(p), Lc, hc, 1 / 1.3, sc["R_phys"], sc["R_model_com"], 5.0,
                                              rp.DisplacementTable(gaxes, synth.displacement_values(gaxes), 5.0), tree=tree,
                                              warn=False)
    t_loop = time.perf_counter() - t0
    return {"value": n_cpu / t_loop, "unit": "particles/s", "cores": 1, "kind": "port",
            "sample": f"{n_cpu} particles in a ({Lc:.0f} Mpc)^3 periodic box (the workload's particle and halo density), {n_halo} halos, "
                      f"halo loop {t_loop:.1f} s ({n_pairs} pairs); KD-tree build {t_tree:.1f} s stated separately "
                      "(the reference builds it once per snapshot, SnapshotRunner.py:95-100)",
            "kdtree_build_s": t_tree}


def run_b200(args):
    claim_stdout()
    import torch
    import torch.distributed as dist
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, parallel, synth
    from baryonforge_b200.tables import displacement_table_of

    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L = _lib.lib()

    cat, model, axes, vals = make_inputs(args)
    nside = args.nside
    npix = 12 * nside * nside
    lo, hi = parallel.pixel_ranges(nside, world)[rank]
    # host inputs live in pinned memory (contract: H2D from pinned host memory inside the e2e region)
    pinned_map = torch.empty(npix, dtype=torch.float64, pin_memory=True)
    pinned_map.numpy()[:] = synth.shell_map(nside, seed=7)
    shell = b.LightconeShell(map=pinned_map.numpy(), cosmo=synth.COSMO)
    runner = b.BaryonifyShell(cat, shell, args.eps, model, verbose=False, device=local,
                              pix_range=None if world == 1 else (lo, hi), sort_halos=not args.no_sort)

    # ---- device-resident step ------------------------------------------------------------------------------
    # halo records: per-halo scalars computed on the device (bfg_shell_records); every rank holds the whole catalogue
    # (128 B per halo) and the halo-loop kernel skips the halos whose disc cannot touch its pixel range
    d_rec = runner.device_records(paint=False, dev=dev)
    n_rec = d_rec.shape[0]
    table = displacement_table_of(model, local)
    d_rec_sorted = torch.empty_like(d_rec)
    d_map = pinned_map[lo:hi].to(dev)
    d_off = torch.empty((3, hi - lo), dtype=torch.float64, device=dev)
    d_new = torch.empty(npix, dtype=torch.float64, device=dev)
    peers = runner._peer_slices(npix, dev) if world > 1 else None
    own = peers.own_tensor() if peers is not None else None
    token = torch.zeros(1, device=dev)
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    d_sums = torch.zeros(2, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    launches = [0]

    def step(timed_kernel=None):
        d_off.zero_()
        if peers is None:
            d_new.zero_()
        else:
            own.zero_()
        if args.no_sort:
            d_use = d_rec
        else:   # locality ordering of the halo records is part of the step
            if world == 1:
                _lib.check(L.bfg_halo_sort(0, n_rec, d_rec.data_ptr(), d_rec_sorted.data_ptr(), None, None, 0,
                                           b.runners.SKY_BAND_RAD, 0.0, 3, st))
            else:   # + per-rank compaction: the halos of other ranks go last and end the halo loop
                _lib.check(L.bfg_halo_sort_owned(nside, lo, hi, n_rec, d_rec.data_ptr(), d_rec_sorted.data_ptr(), None,
                                                 None, 0, b.runners.SKY_BAND_RAD, st))
            d_use = d_rec_sorted
            launches[0] += 2 if world == 1 else 3   # own kernels only: keys, gather (+ mark); CUB's radix passes not counted
        if timed_kernel is not None:
            timed_kernel[0].record()
        _lib.check(L.bfg_shell_offsets(table.handle, nside, n_rec, d_use.data_ptr(), None, 0, d_off.data_ptr(),
                                       lo, hi, d_n.data_ptr(), st))
        if timed_kernel is not None:
            timed_kernel[1].record()
        if peers is not None:
            # fused re-binning + exchange over NVLink peer memory, then the owned slices are gathered into the full map
            dist.all_reduce(token)
            _lib.check(L.bfg_shell_regrid_p2p(nside, d_map.data_ptr(), d_off.data_ptr(), lo, hi, world, rank, peers.h_bounds,
                                              peers.h_slices, None, st))
            dist.all_reduce(token)
            if (not args.gather_result):
                _lib.check(L.bfg_sum_f64(own.data_ptr(), hi - lo, d_sums.data_ptr(), st))
            else:
                full = parallel.gather_owned_ranges(own, npix)
                _lib.check(L.bfg_sum_f64(full.data_ptr(), npix, d_sums.data_ptr(), st))
        else:
            _lib.check(L.bfg_shell_regrid(nside, d_map.data_ptr(), d_off.data_ptr(), d_new.data_ptr(), lo, hi, st))
            if world > 1:
                dist.all_reduce(d_new, op=dist.ReduceOp.SUM)
            _lib.check(L.bfg_sum_f64(d_new.data_ptr(), npix, d_sums.data_ptr(), st))
        _lib.check(L.bfg_sum_f64(d_map.data_ptr(), hi - lo, d_sums.data_ptr() + 8, st))
        launches[0] += 4

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    n_up_local = int(d_n.cpu()[0])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches[0] = 0
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step(kev[k])
    ev[1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev[0].elapsed_time(ev[1])
    ms_kernel = float(np.mean([a.elapsed_time(bb) for a, bb in kev]))
    t = torch.tensor([ms_total, ms_kernel, float(n_up_local)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, ms_kernel, n_up = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        n_up = float(n_up_local)
    ms_step = ms_total / args.steps
    value = n_up / (ms_step * 1e-3)
    # mass conservation of the timed step (HealpixRunner.py:368-370): sum(new map) == sum(old map) over ALL ranks
    d_chk = d_sums.clone()
    if world > 1:
        if peers is not None and (not args.gather_result):
            dist.all_reduce(d_chk)                   # both entries are per-slice sums
        else:
            dist.all_reduce(d_chk[1:])               # the new map was gathered / all-reduced: only the old map's sum is partial
    sums = d_chk.cpu().numpy()
    assert np.isclose(sums[0], sums[1], rtol=1e-9), f"mass not conserved: {sums[0]!r} != {sums[1]!r}"
    n_launch = launches[0]

    # ---- N > 1: is the sharded result the single-GPU result?  (driver-visible multi-GPU parity) ------------------------
    # Rank 0 recomputes the shell un-sharded on its own GPU (same seeds, same kernels, pix range = whole map) outside the
    # timed region and compares it with the map the sharded step produced: every pixel of a strided sample + the sums.
    parity_vs_n1 = None
    sample_stride = max(1, npix // 200000)
    step_sample = None if world > 1 else d_new[::sample_stride].clone()
    if world > 1:
        if peers is not None:
            sharded = parallel.gather_owned_ranges(own, npix)
        else:
            sharded = d_new
        step_sample = sharded[::sample_stride].clone()
        if rank == 0:
            d_map_full = pinned_map.to(dev)
            d_off_full = torch.zeros((3, npix), dtype=torch.float64, device=dev)
            d_ref = torch.zeros(npix, dtype=torch.float64, device=dev)
            d_n1 = torch.zeros(1, dtype=torch.int64, device=dev)
            _lib.check(L.bfg_halo_sort(0, n_rec, d_rec.data_ptr(), d_rec_sorted.data_ptr(), None, None, 0,
                                       b.runners.SKY_BAND_RAD, 0.0, 3, st))
            _lib.check(L.bfg_shell_offsets(table.handle, nside, n_rec, d_rec_sorted.data_ptr(), None, 0, d_off_full.data_ptr(),
                                           0, npix, d_n1.data_ptr(), st))
            _lib.check(L.bfg_shell_regrid(nside, d_map_full.data_ptr(), d_off_full.data_ptr(), d_ref.data_ptr(), 0, npix, st))
            stride = max(1, npix // 200000)
            a_s, r_s = sharded[::stride], d_ref[::stride]
            scale = float(r_s.abs().max())
            err = float(((a_s - r_s).abs() / (r_s.abs() + 1e-3 * scale)).max())
            parity_vs_n1 = {"max_rel_err": err, "n_sample": int(a_s.numel()), "stride": int(stride),
                            "sum_rel_err": float(abs(float(sharded.sum()) - float(d_ref.sum())) / abs(float(d_ref.sum()))),
                            "n_updates_equal": bool(int(d_n1.cpu()[0]) == int(n_up)),
                            "what": "new map of the sharded step vs the same shell computed un-sharded on rank 0's GPU"}
            assert err < 1e-9 and parity_vs_n1["n_updates_equal"], f"sharded result differs from N=1: {parity_vs_n1}"
            del d_map_full, d_off_full, d_ref
        del sharded
        torch.cuda.empty_cache()
        dist.barrier()

    # ---- end-to-end through the reference-shaped API -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            out = runner.process()
        # the API result must be the device-resident step's result (same map on every rank at N > 1)
        ref_s = step_sample.cpu().numpy()
        got_s = np.asarray(out)[::sample_stride].copy()
        e2e_err = float(np.max(np.abs(got_s - ref_s) / (np.abs(ref_s) + 1e-3 * np.max(np.abs(ref_s)))))
        assert e2e_err < 1e-9, f"process() result differs from the device-resident step: {e2e_err}"
        del out
        barrier()
        t0 = time.perf_counter()
        n_e2e = 3
        iter_ms = []
        for _ in range(n_e2e):
            ti = time.perf_counter()
            out = runner.process()
            del out   # a held result keeps its pinned buffer; the next call would then pin a fresh 1.6 GB (~0.8 s)
            iter_ms.append(round(1e3 * (time.perf_counter() - ti), 1))
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": n_up / float(tt[0]), "unit": "halo-pixel updates/s",
               "h2d_bytes_per_step": int(npix * 8 + world * 6 * n_rec * 8), "d2h_bytes_per_step": int(npix * 8),
               "bytes_note": "whole job: every rank uploads its map slice + the 6 catalogue columns and downloads its slice of the new map",
               "ms_per_step": 1e3 * float(tt[0]), "host_prep_ms": 1e3 * runner.last_timing.get("host_prep_s", 0.0),
               "host_threads": b.runners._host_threads(), "parity_vs_device_step": e2e_err,
               "iter_ms": iter_ms, "phases_ms": {k: round(1e3 * v, 2) for k, v in runner.last_timing.items()},
               "includes": "host staging of raw catalogue columns + numpy ln(1+z), ln M; H2D (pinned map + 6 columns); device scalar prep, sort, halo loop, re-binning, exchange (N>1); D2H of the new map"}

    # ---- secondary metric of BASELINE.json: particles displaced/s (BaryonifySnapshot + NGP deposit) ---------------
    # weak scaling: every rank owns one x-slab share of the 2e9-particle box (configs[3]: 2.5e8 particles per GPU at N = 8)
    particles = None
    if not args.no_particles:
        d_off.resize_(0); d_new.resize_(0); d_map.resize_(0); d_rec_sorted.resize_(0)     # give the HBM back (names stay bound)
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        pk, _ = peaks()
        barrier()
        # N > 1: ONE periodic box (1000 Mpc for 8 x 2.5e8 particles), rank r owns the x-slab [r, r + 1) L / N, overlap halos
        # replicated, every rank deposits into a full-size NGP grid and the partial grids are summed over NVLink
        c4 = bench_configs.snapshot_pipeline(args.particles_per_gpu, 5.0, local, 2, pk,
                                             slab=None if world == 1 else (rank, world))
        reduce_ms = 0.0
        d_grid = c4.pop("grid_handle")
        host_in = c4.pop("host_inputs")
        if world > 1:
            barrier()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record(); dist.all_reduce(d_grid); eb.record()
            torch.cuda.synchronize()
            reduce_ms = ea.elapsed_time(eb)
            launches_nccl = 1
        grid_mass = float(d_grid.sum().item())
        del d_grid
        tot_ms = c4["build_cells_ms"] + c4["halo_loop_ms"] + c4["apply_deposit_ms"] + reduce_ms
        fused_ms = c4["build_cells_ms"] + c4["halo_loop_ms"] + c4["apply_deposit_fused_ms"] + reduce_ms \
            if "apply_deposit_fused_ms" in c4 else float("inf")
        tp = torch.tensor([tot_ms, fused_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        particles = {"metric": "particles displaced/s (BaryonifySnapshot + NGP deposit)",
                     "value": world * args.particles_per_gpu / (float(tp[0]) * 1e-3), "unit": "particles/s",
                     "scaling": "weak", "ms_per_pass": float(tp[0]),
                     "config": {"workload": (
                         f"BaryonifySnapshot, ONE periodic box L = {c4['L']:.0f} Mpc with {world * args.particles_per_gpu} uniform "
                         f"particles and {c4['halos_in_box']} halos (M = 10^U(12,15.5)); " + (
                             "one GPU" if world == 1 else
                             f"x-slabs over {world} GPUs ({args.particles_per_gpu} particles each, rank 0 keeps {c4['halos']} halos "
                             "incl. the replicated overlap halos), partial NGP grids summed by NCCL all-reduce") +
                         f"; epsilon_max=5, cell list {c4['ncell']}^3, NGP deposit 512^3; device-resident (particles generated on "
                         "the device), CUDA events, max over ranks"),
                         "deposited_mass_equals_particle_count": bool(abs(grid_mass - world * args.particles_per_gpu) < 0.5)},
                     "phases_ms_rank0": dict({k: c4[k] for k in ("build_cells_ms", "halo_loop_ms", "apply_deposit_ms")},
                                             ngp_allreduce_ms=reduce_ms),
                     "pairs_per_s_rank0": c4["pairs_per_s"], "halo_loop_alg_frac_rank0": c4["halo_loop_frac"]}
        # ---- end to end through the reference-shaped API: host ParticleSnapshot (one structured array, utils/io.py:588) ->
        # BaryonifySnapshot.process() -> structured array of displaced particles on the host; and process_to_map -> host grid
        try:
            particles["e2e"] = particles_e2e(args, world, rank, local, dev, c4, host_in, barrier)
        except Exception as exc:
            particles["e2e"] = {"error": str(exc)[:300]}
        if rank == 0 and not args.no_cpu_baseline:
            try:
                particles["cpu_baseline"] = particles_cpu_baseline(host_in["gaxes"])
            except Exception as exc:
                particles["cpu_baseline"] = {"error": str(exc)[:300]}
        if np.isfinite(float(tp[1])):
            # BaryonifySnapshot.process_to_map: the same cell list and halo loop, then the NGP deposit straight from the
            # cell-ordered particles (no scatter back to the caller's order) -- for callers that only need the grid
            particles["to_map"] = {"value": world * args.particles_per_gpu / (float(tp[1]) * 1e-3), "unit": "particles/s",
                                   "ms_per_pass": float(tp[1]), "apply_deposit_fused_ms_rank0": c4["apply_deposit_fused_ms"],
                                   "deposited_mass_matches": bool(c4["deposited_mass_fused"] == c4["deposited_mass"])}

    if world > 1:
        dist.barrier()

    def make_line():
        peak, peak_src = peaks()
        upd_per_launch = n_up_local if world == 1 else n_up / world
        achieved = ALG_BYTES_PER_UPDATE * upd_per_launch / (ms_kernel * 1e-3) / 1e9
        facts = ncu_facts(workload_name(args)) if world == 1 else {}
        # The resources that bind this kernel (profiles/README.md, round 2) are not HBM: (1) the L2's fp64 atomic unit -- ncu
        # lts__d_atomic_input_cycles_active of the committed capture -- and (2) the arithmetic itself: with the REDs compiled out the
        # kernel is only ~10 % faster.  For (2) the ceiling is what the FP64 pipe could do if it issued nothing but this loop's FP64
        # instructions, at the SM clock measured DURING the timed region: SMs x 64 FP64 lanes per clock x clock / (FP64-pipe
        # instructions per update, counted in the committed SASS of the pixel loop).
        fp64_per_upd = facts.get("fp64_inst_per_update")
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz")
        binding = {"resource": "L2 fp64 atomic unit, co-limited by the arithmetic (FP64 pipe + instruction issue)",
                   "l2_atomic_unit_busy_frac": (facts.get("l2_atomic_input_active_pct") or 0) / 100.0 or None,
                   "l2_red_sectors_per_launch": facts.get("l2_red_sectors"),
                   "kernel_ms_without_reds": facts.get("kernel_ms_without_reds"),
                   "fp64_inst_per_update": fp64_per_upd, "fp64_inst_source": facts.get("fp64_inst_source"),
                   "sm_clock_mhz": sm_mhz, "ncu_source": facts.get("source")}
        if world > 1:
            binding = {"note": "per-kernel ncu facts are captured at N = 1 (profiles/shell_halos_ncu_facts.json); see that line"}
        if world == 1 and fp64_per_upd and sm_mhz:
            ceil_ups = B200_SMS * FP64_LANES_PER_SM_CLK * sm_mhz * 1e6 / fp64_per_upd
            binding.update({"fp64_ceiling_updates_s": ceil_ups,
                            "frac_fp64": upd_per_launch / (ms_kernel * 1e-3) / ceil_ups,
                            "ncu_fp64_pipe_active_pct": facts.get("fp64_pipe_active_pct"),
                            "ncu_issue_active_pct": facts.get("issue_active_pct")})
        line = {"metric": "halo-pixel updates/s (BaryonifyShell)", "value": value, "unit": "halo-pixel updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": shell_config(args, world, n_up, peers is not None, bool(args.gather_result)),
                "clocks": clocks, "gpu_launches": n_launch,
                "roofline": {"bound": "hbm", "kernel": "k_shell_halos<baryonify> (fused disc/separation/table/accumulate)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": facts.get("dram_bytes_per_launch"), "traffic_source": facts.get("source"),
                             "peak_source": peak_src, "alg_bytes_per_update": ALG_BYTES_PER_UPDATE,
                             "alg_bytes_per_launch": ALG_BYTES_PER_UPDATE * upd_per_launch,
                             "kernel_ms": ms_kernel, "binding": binding,
                             "note": "algorithmic bytes = the reference dataflow's 3 f64 read-modify-writes per update (SURVEY 8d). "
                                     "With sky-ordered halos those REDs are absorbed by the 126 MB L2 (`traffic` = measured DRAM "
                                     "bytes of one launch, ~1/25 of algorithmic), so `frac` can exceed 1 and says nothing about "
                                     "efficiency; read `binding`: the L2 atomic unit is `l2_atomic_unit_busy_frac` busy, and the "
                                     "arithmetic alone (`kernel_ms_without_reds`) is almost as slow as the whole kernel"},
                "e2e": e2e, "particles": particles}
        if parity_vs_n1 is not None:
            line["parity_vs_n1"] = parity_vs_n1
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, cat, model, axes, vals, args.cpu_sample)
        return line

    line = make_line() if rank == 0 else None

    # ---- BASELINE configs[2] and configs[4] at their stated sizes, as legs of this line (so that the driver's N = 1..8 runs
    # measure them).  The headline line is complete at this point; a watchdog prints it and ends the process if a leg hangs
    # (a collective some rank never reaches), and an exception inside a leg only marks that leg.
    if not args.no_extra_configs:
        import gc
        import threading
        import bench_modes
        del d_off, d_new, d_map, d_rec_sorted
        state = {"leg": None, "deadline": None}

        def watchdog():
            while state["deadline"] is not None:
                time.sleep(1.0)
                dl = state["deadline"]
                if dl is not None and time.monotonic() > dl:
                    if line is not None:
                        line[state["leg"]] = {"error": f"leg exceeded {args.leg_timeout} s and was abandoned"}
                        emit(line)
                    os._exit(0)
        for name, fn in (("grid", bench_modes.grid_leg), ("lightcone", bench_modes.lightcone_leg)):
            gc.collect()
            torch.cuda.empty_cache()
            b.runners.release_host_buffers()       # the previous leg's idle page-locked buffers (8 GB particle / grid results)
            barrier()
            state["leg"], state["deadline"] = name, time.monotonic() + args.leg_timeout
            if name == "grid":
                th = threading.Thread(target=watchdog, daemon=True)
                th.start()
            try:
                kw = {"cpu_baseline": not args.no_cpu_baseline} if name == "grid" else {}
                res = fn(args, sys.modules[__name__], rank, world, local, **kw)
            except Exception as exc:       # a leg must never take the headline line down with it
                res = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
            if line is not None:
                line[name] = res
        state["leg"], state["deadline"] = "extra_legs_exit_barrier", time.monotonic() + 120.0
        if world > 1:
            dist.barrier()
        state["deadline"] = None
    elif world > 1:
        dist.barrier()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "lightcone":
        import bench_modes
        bench_modes.run_lightcone(args, sys.modules[__name__])
    elif args.config == "paint":
        import bench_modes
        bench_modes.run_paint(args, sys.modules[__name__])
    elif args.config == "grid":
        import bench_modes
        bench_modes.run_grid(args, sys.modules[__name__])
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
