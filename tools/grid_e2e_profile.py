#!/usr/bin/env python
"""tools/grid_e2e_profile.py -- where the host time of BaryonifyGrid.process() goes at 1024^3 (cProfile, one GPU)."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    import bench_modes
    N, n, eps = int(os.environ.get("GRID_N", "1024")), 1000000, 20.0
    Lbox = 1000.0 * N / 1024
    dev = torch.device("cuda", 0)
    pos, M = synth.box_halos(n, Lbox, seed=42)
    bins = (np.arange(N) + 0.5) * Lbox / N
    gaxes = synth.table_axes(nz=10, nM=10, nr=500, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-3, r_max=3e2)
    model = b.DisplacementModel(gaxes, synth.displacement_values(gaxes) * 10, eps, synth.COSMO)
    cat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=M, redshift=0.3, cosmo=synth.COSMO)
    d_map = torch.rand(N ** 3, dtype=torch.float64, device=dev) * 10
    host_map, addr = bench_modes._host_map_with_pinned_slab(N, 0, N, d_map)
    print("torch sees the registered map as pinned:", torch.from_numpy(host_map).is_pinned(), flush=True)
    gm = b.GriddedMap(map=host_map, redshift=0.3, bins=bins, cosmo=synth.COSMO)
    run = b.BaryonifyGrid(cat, gm, eps, model, verbose=False, device=0)
    run.process()
    t0 = time.perf_counter()
    run.process()
    print("process():", time.perf_counter() - t0, "s", flush=True)
    pr = cProfile.Profile()
    pr.enable()
    run.process()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)


if __name__ == "__main__":
    main()
