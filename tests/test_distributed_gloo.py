"""CPU, world_size 2 over gloo: the N>1 host logic (range ownership, partial-map reduce, range gather)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import warnings
    from baryonforge_b200 import parallel, synth
    from oracle import runners_port as rp
    nside = 32
    npix = 12 * nside * nside
    lo, hi = parallel.pixel_ranges(nside, world)[rank]
    # the oracle port plays the part of the per-rank kernel: each rank handles only the halos assigned to it
    ra, dec, M, z = synth.sky_halos(120, seed=4, z=(0.05, 0.5))
    axes = synth.table_axes()
    tab = rp.DisplacementTable(axes, synth.displacement_values(axes), 20)
    cat = dict(M=M, z=z, ra=ra, dec=dec)
    R = 1.2 * (M / 1e14) ** (1 / 3.); D = 1200.0 * z / 0.45; Rc = R * (1 + z)
    theta = np.pi / 2 - np.radians(dec)
    keep = parallel.halos_touching_pixel_range(nside, theta, R * 20 / D, lo, hi)
    sub = {k: v[keep] for k, v in cat.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        off, _ = rp.shell_offsets(nside, sub, R[keep], D[keep], Rc[keep], 20, tab, warn=False)
        full, _ = rp.shell_offsets(nside, cat, R, D, Rc, 20, tab, warn=False)
    owned = torch.from_numpy(off[lo:hi, 0].copy())
    gathered = parallel.gather_owned_ranges(owned, npix)
    ok_gather = np.allclose(gathered.numpy(), full[:, 0], rtol=1e-12, atol=1e-18)
    # partial full-size maps: all-reduce == sum
    hmap = synth.shell_map(nside, seed=3)
    part = np.zeros(npix); part[lo:hi] = hmap[lo:hi]
    red, src_sum = parallel.reduce_partial_map(torch.from_numpy(part), torch.from_numpy(hmap[lo:hi].copy()))
    ok_reduce = np.allclose(red.numpy(), hmap) and np.isclose(float(src_sum), hmap.sum())
    # SimpleParallel under torch.distributed: the runner list is split round-robin over the ranks, every rank gets all outputs
    class _Job(object):
        def __init__(self, i):
            self.i, self.ran_on = i, None

        def process(self):
            self.ran_on = dist.get_rank()
            return np.full(7, float(self.i)) + np.arange(7)
    jobs = [_Job(i) for i in range(5)]
    outs = parallel.SimpleParallel(jobs).process()
    ok_simple = (len(outs) == 5 and all(np.array_equal(o, np.full(7, float(i)) + np.arange(7)) for i, o in enumerate(outs))
                 and [j.ran_on for j in jobs] == [rank if i % world == rank else None for i in range(5)])
    # "do all ranks share one machine?" -- the gate of the CUDA-IPC / shared-host-map path: yes here; no as soon as one rank
    # reports another node (BFG_FAKE_NODE stands in for a second host); no when torchrun says the node holds fewer ranks
    ok_node = parallel.single_node_group() is True
    parallel._SINGLE_NODE.clear()
    os.environ["BFG_FAKE_NODE"] = "node%d" % rank
    ok_node = ok_node and parallel.single_node_group() is False
    parallel._SINGLE_NODE.clear()
    os.environ["BFG_FAKE_NODE"] = ""
    os.environ["LOCAL_WORLD_SIZE"] = "1"
    ok_node = ok_node and parallel.single_node_group() is False
    del os.environ["LOCAL_WORLD_SIZE"]
    parallel._SINGLE_NODE.clear()
    # the slab-sharded grid path only takes the reduce-scatter + shared-host-map route inside an NCCL group on one machine:
    # under gloo the gate says no before anything touches a device, and process() falls back to the all-reduce
    import baryonforge_b200 as b
    N = 16
    pos, Mh = synth.box_halos(5, 100.0, seed=1)
    gcat = b.HaloNDCatalog(x=pos[0], y=pos[1], z=pos[2], M=Mh, redshift=0.3, cosmo=synth.COSMO)
    gm = b.GriddedMap(map=np.ones((N, N, N)), redshift=0.3, bins=(np.arange(N) + 0.5) * 100.0 / N, cosmo=synth.COSMO)
    glo, ghi = parallel.plane_ranges(N, world)[rank]
    grun = b.BaryonifyGrid(gcat, gm, 5, b.DisplacementModel(axes, synth.displacement_values(axes), 5, synth.COSMO),
                           verbose=False, plane_range=(glo, ghi))
    ok_node = ok_node and grun._slab_exchange(N, 3, glo, ghi, None) is None
    q.put((rank, bool(ok_gather), bool(ok_reduce and ok_simple and ok_node), int(keep.sum())))
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_g, ok_r, nkeep in res:
        assert ok_g and ok_r, (rank, ok_g, ok_r)
        assert 0 < nkeep < 120      # each rank got a strict subset of the halos
