#!/usr/bin/env python
"""
tools/bench_grid_sharded.py -- BASELINE.json configs[2]: BaryonifyGrid on a 3-D periodic box of N^3 cells (default 1024)
with 10^6 halos, slab-sharded over the GPUs of one box (one process per GPU, torchrun).

Per rank: axis-0 slab [plane_lo, plane_hi) of the map and of the offsets; halos whose cutout touches the slab
(parallel.halos_touching_planes); halo loop (tile-centric gather) -> re-binning into a full-size partial map -> NCCL
all-reduce (the CIC deposit of a slab reaches into the neighbouring slabs).  Device-resident inputs, CUDA events, max over
ranks; rank 0 prints one JSON line.  N = 1 runs the same code without the all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_grid_sharded.py
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid-n", type=int, default=1024)
    ap.add_argument("--halos", type=int, default=1000000)
    ap.add_argument("--eps", type=float, default=20.0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import baryonforge_b200 as b
    from baryonforge_b200 import _lib, parallel, synth
    from baryonforge_b200.runners import _upload_records, _sort_records
    from baryonforge_b200.tables import displacement_table_of

    rank, world, local = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    N, n = args.grid_n, args.halos
    Lbox = 1000.0 * N / 1024
    lo, hi = parallel.plane_ranges(N, world)[rank]
    pos, M = synth.box_halos(n, Lbox, seed=42)
    bins = (np.arange(N) + 0.5) * Lbox / N
    gaxes = synth.table_axes(nz=10, nM=10, nr=500, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-3, r_max=3e2)
    model = b.DisplacementModel(gaxes, synth.displacement_values(gaxes) * 10, args.eps, synth.COSMO)
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    gm = b.GriddedMap(map=np.broadcast_to(np.zeros(1), (N, N, N)), redshift=0.3, bins=bins, cosmo=synth.COSMO)
    run = b.BaryonifyGrid(cat, gm, args.eps, model, verbose=False, device=local, plane_range=(lo, hi))
    rec, _ = run.halo_records(paint=False)
    keep = parallel.halos_touching_planes(N, rec[:, _lib.HB_CX], rec[:, _lib.HB_NSIZE], lo, hi)
    rec = np.ascontiguousarray(rec[keep])
    tab = displacement_table_of(model, local)
    d_rec = _upload_records(rec, dev)
    d_rec, _ = _sort_records(d_rec, None, 1, Lbox, 16, 3)
    nloc = (hi - lo) * N * N
    d_off = torch.zeros((3, nloc), dtype=torch.float64, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(100 + rank)
    d_map = torch.rand(nloc, dtype=torch.float64, device=dev, generator=g) * 10
    d_new = torch.zeros(N ** 3, dtype=torch.float64, device=dev)
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    d_s = torch.zeros(2, dtype=torch.float64, device=dev)

    def step(ev=None):
        d_off.zero_()
        d_new.zero_()
        if ev:
            ev[0].record()
        _lib.check(L.bfg_grid_offsets(tab.handle, 3, N, float(gm.res), rec.shape[0], d_rec.data_ptr(), None, 0, 0,
                                      d_off.data_ptr(), lo, hi, d_n.data_ptr(), st))
        if ev:
            ev[1].record()
        _lib.check(L.bfg_grid_regrid(3, N, d_map.data_ptr(), d_off.data_ptr(), d_new.data_ptr(), lo, hi, st))
        if ev:
            ev[2].record()
        if world > 1:
            dist.all_reduce(d_new, op=dist.ReduceOp.SUM)
        _lib.check(L.bfg_sum_f64(d_new.data_ptr(), N ** 3, d_s.data_ptr(), st))
        _lib.check(L.bfg_sum_f64(d_map.data_ptr(), nloc, d_s.data_ptr() + 8, st))
        if ev:
            ev[3].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    barrier()
    times = []
    for _ in range(args.reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(); step(ev); z.record()
        barrier()
        times.append([a.elapsed_time(z), ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])])
    t = torch.tensor(np.median(np.array(times), axis=0), dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(d_n.cpu()[0]), float(rec.shape[0])], dtype=torch.float64, device=dev)
    src_sum = d_s[1:2].clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        dist.all_reduce(src_sum, op=dist.ReduceOp.SUM)
    ok = bool(np.isclose(float(d_s[0]), float(src_sum[0])))
    if rank == 0:
        line = {"workload": f"BaryonifyGrid {N}^3 cells, {n} halos, epsilon_max={args.eps:g}, axis-0 slabs x{world}, "
                            "tile-centric gather + CIC re-binning + NCCL all-reduce of partial maps; device-resident",
                "n_gpus": world, "updates": cnt[0].item(), "halo_records_summed_over_ranks": cnt[1].item(),
                "ms_per_pass": t[0].item(), "halo_loop_ms": t[1].item(), "regrid_ms": t[2].item(),
                "reduce_and_sums_ms": t[3].item(), "updates_per_s": cnt[0].item() / (t[0].item() * 1e-3),
                "mass_conserved": ok}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
