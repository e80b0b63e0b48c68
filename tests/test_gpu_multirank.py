"""-m gpu, needs >= 2 GPUs: ring-range / slab sharding over NCCL gives the single-GPU answer (to fp64 summation order)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    import baryonforge_b200 as b
    from baryonforge_b200 import synth
    nside = 256
    ra, dec, M, z = synth.sky_halos(6000, seed=8, z=(0.1, 0.5))
    axes = synth.table_axes()
    cat = b.HaloLightConeCatalog(ra=ra, dec=dec, M=M, z=z, cosmo=synth.COSMO)
    shell = b.LightconeShell(map=synth.shell_map(nside, seed=9), cosmo=synth.COSMO)
    dmodel = b.DisplacementModel(axes, synth.displacement_values(axes), 20, synth.COSMO)
    pmodel = b.ProfileModel(axes, synth.profile_values(axes) * 3, synth.profile_values(axes))
    N, Lbox = 96, 150.0
    pos, Mb = synth.box_halos(400, Lbox, seed=10)
    bins = (np.arange(N) + 0.5) * Lbox / N
    gaxes = synth.table_axes(nz=6, nM=10, nr=300, z_min=0.0, z_max=1.0, z_linear=True, r_min=1e-2, r_max=2e2)
    gmodel = b.DisplacementModel(gaxes, synth.displacement_values(gaxes) * 20, 5, synth.COSMO)
    gcat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=Mb, redshift=0.3, cosmo=synth.COSMO)
    gm = b.GriddedMap(map=np.random.default_rng(11).uniform(0, 10, (N, N, N)), redshift=0.3, bins=bins, cosmo=synth.COSMO)
    return nside, cat, shell, dmodel, pmodel, N, gcat, gm, gmodel


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import baryonforge_b200 as b
    from baryonforge_b200 import parallel
    parallel.init_from_env("nccl")
    nside, cat, shell, dmodel, pmodel, N, gcat, gm, gmodel = _inputs()
    pr = parallel.pixel_ranges(nside, world)[rank]
    out1 = b.BaryonifyShell(cat, shell, 20, dmodel, verbose=False, device=rank, pix_range=pr).process()
    out2 = b.PaintProfilesShell(cat, shell, 20, pmodel, verbose=False, device=rank, pix_range=pr).process()
    out3 = b.BaryonifyGrid(gcat, gm, 6, gmodel, verbose=False, device=rank,
                           plane_range=parallel.plane_ranges(N, world)[rank]).process()
    if rank == 0:
        q.put((out1, out2, out3))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_sharded_runs_match_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import baryonforge_b200 as b
    from helpers import assert_close
    nside, cat, shell, dmodel, pmodel, N, gcat, gm, gmodel = _inputs()
    want1 = b.BaryonifyShell(cat, shell, 20, dmodel, verbose=False).process()
    want2 = b.PaintProfilesShell(cat, shell, 20, pmodel, verbose=False).process()
    want3 = b.BaryonifyGrid(gcat, gm, 6, gmodel, verbose=False).process()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got1, got2, got3 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert_close(got1, want1, "sharded BaryonifyShell", rtol=1e-9, atol_scale=1e-12)
    assert_close(got2, want2, "sharded PaintProfilesShell", rtol=1e-9, atol_scale=1e-12)
    assert_close(got3, want3, "sharded BaryonifyGrid", rtol=1e-9, atol_scale=1e-12)
