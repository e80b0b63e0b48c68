"""
Multi-GPU sharding of one map across the GPUs of a box -- the replacement for BaryonForge/utils/Parallelize.py
(joblib/loky: SimpleParallel :8-113, SplitJoinParallel :116-320), SURVEY.md §8(e).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch):
  * shells: the RING pixel index space is cut into `world` contiguous ranges of equal pixel count; rank r owns
    pix_offsets[lo_r:hi_r].  Every halo whose disc touches a rank's rings is given to that rank (halo records are
    128 B, so overlap halos are simply replicated) -- no communication in the halo loop.  The re-binning is fused
    with its exchange step: each rank owns a slice of the new map that its peers have mapped through CUDA IPC
    (PeerSlices), and bfg_shell_regrid_p2p deposits every displaced pixel straight into the owner's slice (fp64 REDs
    over NVLink for the few that cross a range border); an all_gather then assembles the full map.  Fallback
    (BFG_EXCHANGE=allreduce, gloo, >8 ranks): full-size partial maps + ONE all-reduce(sum, fp64).
  * grids: the same with slabs of axis-0 planes.
  * painting: each rank paints its own range; ranges are concatenated with all_gather.
The reference can only split painting runs across halos and sums whole maps in the parent (Parallelize.py:318);
results here are independent of the number of ranks up to fp64 summation order (§10 #14).
"""
import numpy as np

__all__ = ['single_node_group', 'SegmentsExhausted', 'pixel_ranges', 'plane_ranges', 'ring_of_pixel', 'halos_touching_pixel_range', 'halos_touching_planes',
           'reduce_partial_map', 'gather_owned_ranges', 'init_from_env', 'PeerSlices', 'SharedHostMaps', 'SimpleParallel',
           'SplitJoinParallel', 'snapshot_slab', 'deposit_ngp_all']


def pixel_ranges(nside, world):
    """`world` contiguous RING ranges [lo, hi) of (nearly) equal pixel count covering the whole map."""
    npix = 12 * nside * nside
    edges = [(npix * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def plane_ranges(N, world):
    edges = [(N * r) // world for r in range(world + 1)]
    return [(edges[r], edges[r + 1]) for r in range(world)]


def ring_of_pixel(nside, pix):
    """Ring number (1 .. 4 nside - 1) of RING-ordered pixels (host copy of the device pix2ring)."""
    pix = np.asarray(pix, dtype=np.int64)
    npix, ncap = 12 * nside * nside, 2 * nside * (nside - 1)
    ring = np.empty(pix.shape, dtype=np.int64)
    north = pix < ncap
    south = pix >= npix - ncap
    eq = ~(north | south)
    ring[north] = (1 + np.floor(np.sqrt(1 + 2 * pix[north] + 0.5)).astype(np.int64)) >> 1
    ring[eq] = (pix[eq] - ncap) // (4 * nside) + nside
    q = npix - pix[south]
    ring[south] = 4 * nside - ((1 + np.floor(np.sqrt(2 * q - 1 + 0.5)).astype(np.int64)) >> 1)
    return ring


def _ring_z(nside, ring):
    ring = np.asarray(ring, dtype=np.float64)
    npix = 12.0 * nside * nside
    f2 = 4.0 / npix
    f1 = 2 * nside * f2
    z = np.where(ring < nside, 1 - ring * ring * f2,
                 np.where(ring <= 3 * nside, (2 * nside - ring) * f1, (4 * nside - ring) ** 2 * f2 - 1))
    return z


def first_pixel_at_colatitude(nside, theta):
    """First RING pixel of the first ring whose colatitude is >= theta (npix when theta lies south of the last ring)."""
    npix = 12 * nside * nside
    if theta <= 0:
        return 0
    if theta >= np.pi:
        return npix
    z = np.cos(theta)
    lo, hi = 1, 4 * nside          # smallest ring r in [1, 4 nside - 1] with z_ring(r) <= z  (z decreases with r)
    while lo < hi:
        mid = (lo + hi) // 2
        if mid <= 4 * nside - 1 and float(_ring_z(nside, mid)) <= z:
            hi = mid
        else:
            lo = mid + 1
    r = lo
    if r > 4 * nside - 1:
        return npix
    ncap = 2 * nside * (nside - 1)
    if r < nside:
        return 2 * r * (r - 1)
    if r < 3 * nside:
        return ncap + (r - nside) * 4 * nside
    q = 4 * nside - r
    return npix - 2 * q * (q + 1)


def halos_touching_pixel_range(nside, theta, radius, lo, hi, margin_rings=2):
    """
    Boolean mask of the halos whose disc (colatitude theta, angular radius) can contain a pixel of [lo, hi).
    Conservative (a margin of rings each side): a halo given to a rank that owns none of its pixels costs only time.
    """
    if hi <= lo:
        return np.zeros(np.shape(theta), dtype=bool)
    r_top, r_bot = ring_of_pixel(nside, np.array([lo, hi - 1]))
    r_top = max(1, int(r_top) - margin_rings)
    r_bot = min(4 * nside - 1, int(r_bot) + margin_rings)
    th_top = 0.0 if r_top == 1 else np.arccos(np.clip(_ring_z(nside, r_top), -1, 1))
    th_bot = np.pi if r_bot == 4 * nside - 1 else np.arccos(np.clip(_ring_z(nside, r_bot), -1, 1))
    theta, radius = np.asarray(theta), np.asarray(radius)
    return (theta + radius >= th_top) & (theta - radius <= th_bot)


def halos_touching_planes(N, centre, nsize, lo, hi):
    """Mask of halos whose cutout [centre - nsize/2, centre + nsize/2) (periodic, axis 0) meets planes [lo, hi)."""
    centre = np.asarray(centre, dtype=np.int64)
    half = np.asarray(nsize, dtype=np.int64) // 2
    start = centre - half
    length = 2 * half
    # distance from `start` forward (periodic) to the first owned plane
    d = (lo - start) % N
    return (d < length) | (((start - lo) % N) < (hi - lo))


def bind_to_gpu_numa_node(device_index):
    """
    Restrict this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node and that node's
    cpulist), so that page-locked buffers it allocates or first touches -- staging buffers, result maps, its slice of the shared
    host map -- are local to the GPU's PCIe root.  One process per GPU without this lets Linux place eight ranks' buffers on
    one socket: measured 11 GB/s per rank for the result download at N = 8 against 55 GB/s at N = 1.
    BFG_NUMA_BIND=0 disables it.  Returns the node number, or None when nothing was changed.
    """
    import os
    if os.environ.get("BFG_NUMA_BIND", "1") != "1" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b_ = part.partition("-")
            cpus.update(range(int(a), int(b_ or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def init_from_env(backend=None):
    """torch.distributed bootstrap from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun)."""
    import os
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and torch.cuda.is_available():
        bind_to_gpu_numa_node(local)
    return rank, world, local


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def reduce_partial_map(partial_full_map, owned_source_map):
    """all-reduce(sum) of the per-rank partial full-size maps; also returns sum(original map) over all ranks."""
    dist = _dist()
    src_sum = owned_source_map.sum().reshape(1)
    if dist is not None:
        dist.all_reduce(partial_full_map, op=dist.ReduceOp.SUM)
        dist.all_reduce(src_sum, op=dist.ReduceOp.SUM)
    return partial_full_map, src_sum[0]


def all_reduce_sum(t):
    """In-place sum of a small device tensor over the ranks (no-op without an initialised process group)."""
    dist = _dist()
    if dist is not None and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_owned_ranges(owned, total):
    """Concatenate every rank's owned (contiguous, rank-ordered) slice into the full array on every rank."""
    import torch
    dist = _dist()
    if dist is None:
        assert owned.numel() == total
        return owned
    world = dist.get_world_size()
    if total % world == 0 and owned.numel() * world == total:
        # pixel_ranges / plane_ranges give equal slices whenever `world` divides the map: one collective, no staging
        full = torch.empty(total, dtype=owned.dtype, device=owned.device)
        dist.all_gather_into_tensor(full, owned.contiguous())
        return full
    sizes = [torch.zeros(1, dtype=torch.int64, device=owned.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([owned.numel()], dtype=torch.int64, device=owned.device))
    sizes = [int(s.item()) for s in sizes]
    assert sum(sizes) == total
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=owned.dtype, device=owned.device)
    buf[:owned.numel()] = owned
    parts = [torch.empty(pad, dtype=owned.dtype, device=owned.device) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)])


_SINGLE_NODE = {}


def single_node_group():
    """True when every rank of the default process group runs on THIS machine (same hostname and boot id), i.e. CUDA IPC
    handles and /proc/<pid>/fd paths are meaningful between the ranks.  Collective; cached per process group size."""
    import os
    import socket
    dist = _dist()
    if dist is None:
        return True
    key = (dist.get_world_size(), dist.get_rank())
    if key not in _SINGLE_NODE:
        lws = os.environ.get("LOCAL_WORLD_SIZE")
        try:
            boot = open("/proc/sys/kernel/random/boot_id").read().strip()
        except OSError:
            boot = ""
        mine = (socket.gethostname(), boot, os.environ.get("BFG_FAKE_NODE", ""))
        got = [None] * dist.get_world_size()
        dist.all_gather_object(got, mine)
        same = all(g == got[0] for g in got)
        if lws is not None and int(lws) != dist.get_world_size():
            same = False
        _SINGLE_NODE[key] = same
    return _SINGLE_NODE[key]


class SegmentsExhausted(RuntimeError):
    """Every shared host segment is still referenced by a result map of an earlier process() call (raised on ALL ranks)."""


class PeerSlices(object):
    """
    Each rank's slice of a sharded output map, allocated IPC-exportable and mapped into every other process of the box, so
    a kernel can deposit straight into the owner's HBM over NVLink (bfg_shell_regrid_p2p).  One process per GPU.
    """

    def __init__(self, bounds, rank, world, device):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        self.bounds, self.rank, self.world, self.device = list(bounds), rank, world, device
        self.n_own = bounds[rank + 1] - bounds[rank]
        own = C.c_void_p()
        _lib.check(L.bfg_shared_alloc(C.byref(own), 8 * max(self.n_own, 1), device))
        self._own = own
        handle = (C.c_ubyte * 64)()
        _lib.check(L.bfg_ipc_export(own, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=torch.device('cuda', device))
        allh = [torch.empty(64, dtype=torch.uint8, device=mine.device) for _ in range(world)]
        dist.all_gather(allh, mine)
        self._peers = []
        ptrs = []
        ok = 1
        for r in range(world):
            if r == rank:
                ptrs.append(own.value)
                continue
            raw = bytes(allh[r].cpu().tolist())
            buf = (C.c_ubyte * 64).from_buffer_copy(raw)
            p = C.c_void_p()
            if L.bfg_ipc_import(buf, C.byref(p)) != 0:   # e.g. the peer lives on another node / no P2P between the GPUs
                ok = 0
                break
            self._peers.append(p)
            ptrs.append(p.value)
        # every rank must have mapped every peer before anybody launches a kernel that writes peer memory
        vote = torch.tensor([ok], dtype=torch.int32, device=mine.device)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN)
        if int(vote.cpu()[0]) == 0:
            self.close()
            raise OSError("peer slices could not be mapped by every rank (CUDA IPC); use the all-reduce exchange")
        self.h_bounds = (C.c_int64 * (world + 1))(*self.bounds)
        self.h_slices = (C.c_void_p * world)(*ptrs)

    def own_tensor(self):
        """The owned slice as a torch tensor (no copy) -- for zeroing, summing, gathering."""
        import torch

        class _Arr(object):
            pass
        a = _Arr()
        a.__cuda_array_interface__ = dict(shape=(self.n_own,), typestr='<f8', data=(self._own.value, False), version=3,
                                          strides=None)
        return torch.as_tensor(a, device=torch.device('cuda', self.device))

    def close(self):
        from . import _lib
        L = _lib.lib()
        for p in self._peers:
            L.bfg_ipc_close(p)
        self._peers = []
        if self._own is not None:
            L.bfg_shared_free(self._own)
            self._own = None


class SharedHostMaps(object):
    """
    Result maps in host memory that EVERY process of the box has mapped (memfd, page-locked through bfg_host_register):
    each rank copies only its owned slice device -> host, in parallel over its own PCIe link, and after a barrier all
    ranks hold the complete map without a device all-gather and without N full-size D2H copies.  Replaces the parent
    process summing whole maps returned by pickling (utils/Parallelize.py:315-318).

    A segment is handed out again once the numpy array given to the caller (and every view of it) has been garbage-
    collected on ALL ranks; acquire() is collective.
    """

    MAX_SEGMENTS = 6

    def __init__(self, numel, rank, world, device, own_range=None):
        self.numel, self.rank, self.world, self.device = int(numel), rank, world, device
        self.own_range = own_range   # (lo, hi) elements this rank will write: first-touched here, so they are NUMA-local
        self.segs = []        # dicts: mm (mmap), addr, free (bool, local view)

    def _new_segment(self):
        import ctypes as C
        import mmap
        import os
        import torch.distributed as dist
        from . import _lib
        nbytes = 8 * self.numel
        info = [None, None]
        fd = -1
        if self.rank == 0:
            fd = os.memfd_create("bfg_b200_map")
            os.ftruncate(fd, nbytes)
            info = [os.getpid(), fd]
        dist.broadcast_object_list(info, src=0)
        import torch
        mm, addr, ok = None, 0, 1
        try:
            if self.rank != 0:
                fd = os.open(f"/proc/{info[0]}/fd/{info[1]}", os.O_RDWR)   # same PID namespace (one box, torchrun)
            mm = mmap.mmap(fd, nbytes)      # MAP_SHARED; mmap keeps its own duplicate of the descriptor
            addr = C.addressof(C.c_char.from_buffer(mm))
            if self.own_range is not None:
                # first touch: the pages of the slice this rank's GPU will write are allocated on the node this process
                # runs on (bind_to_gpu_numa_node); page-locking below would otherwise fault them in wherever the first
                # registering rank happens to run
                lo_b, hi_b = 8 * int(self.own_range[0]), 8 * int(self.own_range[1])
                np.frombuffer(mm, dtype=np.uint8)[lo_b:hi_b] = 0
        except Exception:
            ok = 0
        t_sync = torch.zeros(1, device=torch.device('cuda', self.device))
        dist.all_reduce(t_sync)             # every rank has touched its slice ...
        t_sync.cpu()
        try:
            if ok:
                _lib.check(_lib.lib().bfg_host_register(addr, nbytes))   # ... before anybody pins the whole segment
        except Exception:
            ok = 0
        # every rank must agree before anybody relies on the segment (this also fences rank 0's descriptor)
        flag = torch.tensor([ok], dtype=torch.int32, device=torch.device('cuda', self.device))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if fd >= 0:
            os.close(fd)
        if int(flag.cpu()[0]) == 0:
            if ok:
                _lib.lib().bfg_host_unregister(addr)
            if mm is not None:
                try:
                    mm.close()
                except Exception:
                    pass
            raise OSError("shared host map could not be mapped by every rank")
        self.segs.append(dict(mm=mm, addr=addr, free=True))
        return len(self.segs) - 1

    def acquire(self):
        """Collective: pick a segment no rank still exposes to its caller; returns (index, host address)."""
        import torch
        import torch.distributed as dist
        flags = torch.zeros(self.MAX_SEGMENTS, dtype=torch.int32, device=torch.device('cuda', self.device))
        mine = [1 if (i < len(self.segs) and self.segs[i]['free']) else 0 for i in range(self.MAX_SEGMENTS)]
        flags.copy_(torch.tensor(mine, dtype=torch.int32))
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        common = flags.cpu().tolist()
        idx = next((i for i, f in enumerate(common) if f), None)
        if idx is None:
            # len(self.segs) is the same on every rank (segments are created collectively), so all ranks raise together
            if len(self.segs) >= self.MAX_SEGMENTS:
                raise SegmentsExhausted(f"{self.MAX_SEGMENTS} result maps of earlier process() calls are still referenced")
            idx = self._new_segment()
        self.segs[idx]['free'] = False
        return idx, self.segs[idx]['addr']

    def export(self, idx, shape=None):
        """The whole segment as a numpy array; the segment returns to the pool when the array and its views are gone."""
        import ctypes as C
        import weakref
        seg = self.segs[idx]
        owner = (C.c_double * self.numel).from_buffer(seg['mm'])
        weakref.finalize(owner, seg.__setitem__, 'free', True)
        arr = np.frombuffer(owner, dtype=np.float64)
        return arr if shape is None else arr.reshape(shape)

    def close(self):
        from . import _lib
        for seg in self.segs:
            try:
                _lib.lib().bfg_host_unregister(seg['addr'])
                seg['mm'].close()
            except Exception:
                pass            # a caller still holds the map: the mapping stays valid until it is dropped
        self.segs = []


# =====================================================================================================================
# Drop-in mirrors of BaryonForge/utils/Parallelize.py, so user scripts that wrap runners in them keep running
# =====================================================================================================================
class SimpleParallel(object):
    """
    BaryonForge/utils/Parallelize.py:8-113: run a list of independent runners, return their outputs in input order.
    The reference forks one joblib/loky process per runner; here the runners execute one after another on the GPU(s)
    of this process (round-robin over `devices` if given) -- each `process()` is already parallel inside.
    Under torch.distributed (one process per GPU) the list is split across ranks and outputs are exchanged, so every
    rank returns the full ordered list.
    """

    def __init__(self, Runner_list, njobs=-1, devices=None):
        self.Runner_list = Runner_list
        self.njobs = len(Runner_list) if njobs == -1 else min(njobs, len(Runner_list))
        self.devices = devices

    def single_run(self, i, Runner):
        return i, Runner.process()

    def process(self):
        dist = _dist()
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        mine = {}
        for i, Runner in enumerate(self.Runner_list):
            if i % world != rank:
                continue
            if self.devices:
                Runner.device = self.devices[i % len(self.devices)]
            mine[i] = Runner.process()
        if dist is None:
            return [mine[i] for i in range(len(self.Runner_list))]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        merged = {}
        for g in gathered:
            merged.update(g)
        return [merged[i] for i in range(len(self.Runner_list))]


class SplitJoinParallel(object):
    """
    BaryonForge/utils/Parallelize.py:116-320: split a PAINTING runner's halo catalogue into `njobs` chunks, paint each on
    an empty shell and sum the maps.  Kept for API compatibility: on the GPU the split buys nothing (the halo loop is
    already data-parallel), so the chunks run back to back and are summed exactly as the reference sums them --
    including its quirks: halos are reshuffled with default_rng(seed), `include_pixel_size` is not forwarded (:271), and
    Baryonify* runners are refused (:206-209).
    """

    def __init__(self, Runner, njobs=-1, seed=42):
        from .runners import BaryonifyShell, BaryonifyGrid, BaryonifySnapshot
        text = f"Runner of type {type(Runner)} is not supported for SplitJoinParallel."
        assert not isinstance(Runner, (BaryonifyGrid, BaryonifyShell, BaryonifySnapshot)), text
        self.Runner = Runner
        self.seed = seed
        self.njobs = 1 if njobs == -1 else njobs
        self.Runner_list = self.split_run(self.Runner)

    def split_run(self, Runner):
        HaloCat, Shell = Runner.HaloLightConeCatalog, Runner.LightconeShell
        Ntotal = len(HaloCat.cat)
        Npersplit = int(np.ceil(Ntotal / self.njobs))
        HaloCat = HaloCat[np.random.default_rng(self.seed).choice(Ntotal, size=Ntotal, replace=False)]
        empty_shell = type(Shell)(map=np.zeros_like(Shell.map), cosmo=Runner.cosmo)
        out = []
        for i in range(self.njobs):
            sub = HaloCat[i * Npersplit:(i + 1) * Npersplit]
            out.append(type(Runner)(sub, empty_shell, Runner.epsilon_max, Runner.model, Runner.use_ellipticity,
                                    Runner.mass_def, verbose=False))
        return out

    def single_run(self, Runner):
        return Runner.process()

    def process(self):
        outputs = [np.array(r.process()) for r in self.Runner_list]
        return np.sum(outputs, axis=0)


# =====================================================================================================================
# particle snapshots across ranks (BASELINE config 4): slabs in x by particle position
# =====================================================================================================================
def snapshot_slab(ps, rank, world):
    """
    The particles of `ps` (a ParticleSnapshot) whose x lies in this rank's slab [rank, rank+1) * L / world, as a new
    ParticleSnapshot of the SAME periodic box, plus their indices in the original arrays.  Every halo within reach of a
    slab is needed by that rank; BaryonifySnapshot simply gets the whole (small) halo catalogue -- halos far from the slab
    find empty cells.  Particles are displaced by at most the table's range, so no particle exchange happens before the
    displacement is applied; the NGP deposit of the displaced particles is summed over ranks (deposit_ngp_all).
    """
    from .io import ParticleSnapshot
    L = ps.L
    x = ps.cat['x']
    lo, hi = L * rank / world, L * (rank + 1) / world
    sel = np.flatnonzero((x >= lo) & (x < hi) if rank < world - 1 else (x >= lo))
    sub = ParticleSnapshot(x=ps.cat['x'][sel], y=ps.cat['y'][sel], z=None if ps.is2D else ps.cat['z'][sel],
                           M=ps.cat['M'][sel], L=L, redshift=ps.redshift, cosmo=ps.cosmo)
    return sub, sel


def deposit_ngp_all(coords, mass, L, N_grid, device=None):
    """NGP deposit of this rank's particles followed by an all-reduce(sum) of the partial grids over the ranks."""
    import torch
    from .runners import deposit_ngp
    grid = deposit_ngp(coords, mass, L, N_grid, device=device)
    dist = _dist()
    if dist is None:
        return grid
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device)) \
        if dist.get_backend() == "nccl" else torch.device('cpu')
    t = torch.from_numpy(np.ascontiguousarray(grid)).to(dev)
    dist.all_reduce(t)
    return t.cpu().numpy()
