"""
Runners: the drop-in mirror of BaryonForge/Runners (same class names, constructor signatures, attributes and
process() return values), executing the per-halo loop and the re-binning on a B200 through libbfg_b200.so.

    BaryonifyShell / PaintProfilesShell   <- BaryonForge/Runners/HealpixRunner.py:160-177, :252-373, :390-483
    BaryonifyGrid / PaintProfilesGrid     <- BaryonForge/Runners/Map2DRunner.py:255-278, :431-621, :676-829
    BaryonifySnapshot                     <- BaryonForge/Runners/SnapshotRunner.py:83-100, :176-274

Host work per process(): the per-halo scalars the reference computes at the top of every loop iteration, done once
with numpy (cosmology.py), packed into 128-byte halo records.  Everything per (halo, pixel|cell|particle) happens in
CUDA.  PyTorch is used only for device memory, streams and (parallel.py) torch.distributed.
There is no CPU fallback: without a GPU / the shared library, process() raises.
"""
import os
import time
import warnings
import weakref
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib, cosmology
from .tables import displacement_table_of, get_parameter, profile_table_of

__all__ = ['DefaultRunner', 'BaryonifyShell', 'PaintProfilesShell', 'PaintProfilesAnisShell', 'DefaultRunnerGrid',
           'BaryonifyGrid', 'PaintProfilesGrid', 'PaintProfilesAnisGrid', 'DefaultRunnerSnapshot', 'BaryonifySnapshot',
           'deposit_ngp']


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _lib.BFGError("baryonforge_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


def _to_device(arr, device, dtype=None):
    torch = _torch()
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=dtype))
    return t.to(device, non_blocking=True)


def _upload_records(rec, dev):
    """Halo records to the device as [n][16].  `halo_records` builds them field-major (each field one contiguous numpy
    vector); that buffer is copied as is and transposed on the device, which is much cheaper than 16 strided host writes."""
    torch = _torch()
    n = rec.shape[0]
    if n > 0 and rec.T.flags.c_contiguous and not rec.flags.c_contiguous:
        d_t = torch.from_numpy(rec.T).to(dev, non_blocking=True)   # asynchronous when the storage is pinned
        d_rec = torch.empty((n, _lib.HALO_STRIDE), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().bfg_transpose_offsets(_lib.ptr(d_t), _lib.ptr(d_rec), n, _lib.HALO_STRIDE,
                                                    _lib.current_stream()))
        return d_rec
    return _to_device(rec, dev)


def _sort_records(d_rec, d_ext, mode, p0, p1=0.0, ndim=3, owned=None):
    """Locality ordering on the device (bfg_halo_sort); returns the re-ordered (records, extras).
    owned = (nside, pix_lo, pix_hi): sky ordering that also moves the halos of other ranks behind the owned ones and
    marks them (bfg_halo_sort_owned), so the halo loop of a sharded run stops there."""
    torch = _torch()
    n = d_rec.shape[0]
    if n < 2 and owned is None:
        return d_rec, d_ext
    if n == 0:
        return d_rec, d_ext
    out = torch.empty_like(d_rec)
    out_e = None if d_ext is None else torch.empty_like(d_ext)
    if owned is not None:
        _lib.check(_lib.lib().bfg_halo_sort_owned(int(owned[0]), int(owned[1]), int(owned[2]), n, _lib.ptr(d_rec),
                                                  _lib.ptr(out), _lib.ptr(d_ext), _lib.ptr(out_e),
                                                  0 if d_ext is None else d_ext.shape[1], float(p0),
                                                  _lib.current_stream()))
        return out, out_e
    _lib.check(_lib.lib().bfg_halo_sort(mode, n, _lib.ptr(d_rec), _lib.ptr(out), _lib.ptr(d_ext), _lib.ptr(out_e),
                                        0 if d_ext is None else d_ext.shape[1], float(p0), float(p1), ndim,
                                        _lib.current_stream()))
    return out, out_e


_SIDE_STREAMS = {}


def _side_stream(dev, which=0):
    """Copy streams per device: 0 = uploads (H2D of the map underneath the halo loop), 1 = downloads of finished map parts."""
    torch = _torch()
    key = (dev.index, which)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


_PINNED_FREE = {}      # numel -> list of idle pinned float64 tensors
_PEER_SLICES = {}      # (npix, world, rank, device) -> parallel.PeerSlices or None (per process, shared by all runners)
_HOST_MAPS = {}        # (npix, world, rank, device) -> parallel.SharedHostMaps or None
_SPLINE_PACKS = []     # last few (key, spline pack) of BaryonifyShell._spline_pack, shared by all runner objects


def _pinned_result(numel):
    """
    A float64 pinned host buffer for a D2H result, as (tensor, numpy view).  Buffers are recycled as soon as the numpy
    array handed to the caller (and every view of it) is garbage-collected, so steady-state calls never pay the ~0.5 s/GB
    page-locking of a fresh allocation; a caller that keeps N results alive simply owns N buffers.
    """
    torch = _torch()
    free = _PINNED_FREE.setdefault(numel, [])
    t = free.pop() if free else torch.empty(numel, dtype=torch.float64, pin_memory=True)
    arr = t.numpy()

    def _recycle(tensor=t, bucket=free):
        if len(bucket) < 4:
            bucket.append(tensor)
    weakref.finalize(arr, _recycle)
    return t, arr


_PINNED_SCRATCH = {}   # numel -> list of idle pinned float64 staging tensors (halo-record batches)


def release_host_buffers():
    """
    Give the IDLE page-locked buffers of this process back to the system: recycled result buffers (_pinned_result), staging
    rings (_take_scratch) and what torch's pinned-memory allocator caches behind them.  Buffers a caller still holds (results of
    earlier process() calls) and the shared host maps of sharded runs are not touched.  For long-lived processes that move on to
    a workload of a different size (the pools are keyed by size: a 2.5e8-particle run leaves 8 GB result buffers behind).
    """
    _PINNED_FREE.clear()
    _PINNED_SCRATCH.clear()
    import gc
    gc.collect()
    try:
        import torch
        torch._C._host_emptyCache()
    except Exception:      # older torch: the cached pinned blocks stay with its allocator
        pass


def _take_scratch(numel):
    torch = _torch()
    free = _PINNED_SCRATCH.setdefault(numel, [])
    return free.pop() if free else torch.empty(numel, dtype=torch.float64, pin_memory=True)


def _give_scratch(tensors):
    for t in tensors:
        bucket = _PINNED_SCRATCH.setdefault(t.numel(), [])
        if len(bucket) < 8:
            bucket.append(t)


def _host_threads():
    """Host threads of this process: BFG_HOST_THREADS, else min(16, cores / processes of this box (torchrun))."""
    if "BFG_HOST_THREADS" in os.environ:
        return int(os.environ["BFG_HOST_THREADS"])
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    return max(1, min(16, (os.cpu_count() or 1) // local_world))


_POOL = {}


def _pool():
    """The process-wide host thread pool (BFG_HOST_THREADS, default min(16, cores / local ranks)).  It is created once: starting
    16 threads for every process() call cost ~1 ms of the 4 ms host staging of a 10^6-halo catalogue."""
    nth = max(1, _host_threads())
    ex = _POOL.get(nth)
    if ex is None:
        ex = _POOL[nth] = ThreadPoolExecutor(max_workers=nth, thread_name_prefix="bfg-host")
    return ex


def _parallel_chunks(fn, n, chunk=65536):
    """Run fn(slice) over [0, n) in chunks on the host thread pool."""
    nthreads = _host_threads()
    slices = [slice(i, min(i + chunk, n)) for i in range(0, n, chunk)]
    if nthreads <= 1 or len(slices) <= 1:
        for sl in slices:
            fn(sl)
        return
    list(_pool().map(fn, slices))


_FIELD_CHUNK = 1 << 22      # elements per staging chunk (32 MB of float64)


def _fields_to_device(cat, names, dev):
    """
    Columns of a structured catalogue (strided 8-byte fields) -> contiguous float64 device tensors.  The reference keeps particles
    as ONE structured array (utils/io.py:588), so every column is a strided view: np.ascontiguousarray + a pageable copy moves
    2 GB per column at ~1 GB/s.  Here host threads gather chunks of the field into a ring of page-locked buffers while the
    previous chunks are in flight to the device (asynchronous copies on the current stream).
    """
    torch = _torch()
    n = cat.shape[0]
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in names]
    if n == 0:
        return outs
    nbuf = 4
    bufs = [_take_scratch(_FIELD_CHUNK) for _ in range(nbuf)]
    evs = [None] * nbuf
    stream = torch.cuda.current_stream()
    nth = max(1, _host_threads())
    ex = _pool()
    k = 0
    for col, name in enumerate(names):
        src = cat[name]
        for a in range(0, n, _FIELD_CHUNK):
            b = min(n, a + _FIELD_CHUNK)
            slot = k % nbuf
            if evs[slot] is not None:
                evs[slot].synchronize()                  # the copy that last used this buffer has left the host
            dst = bufs[slot].numpy()[:b - a]
            step = -(-(b - a) // nth)
            list(ex.map(lambda i: np.copyto(dst[i:i + step], src[a + i:min(a + i + step, b)], casting='unsafe'),
                        range(0, b - a, step)))
            outs[col][a:b].copy_(bufs[slot][:b - a], non_blocking=True)
            evs[slot] = stream.record_event()
            k += 1
    for e in evs:
        if e is not None:
            e.synchronize()
    _give_scratch(bufs)
    return outs


_RAW_CHUNK = 1 << 24        # doubles per staging chunk of a raw record copy (128 MB)
_RAW_RESULT_MAX_BYTES = int(float(os.environ.get("BFG_PINNED_RESULT_MAX_GB", "24")) * 2 ** 30)


def _record_layout(cat, names):
    """
    Slot (in doubles) of every field when `cat` is what the reference's ParticleSnapshot holds (utils/io.py:588): ONE
    C-contiguous structured array of 32-byte records made of four native float64 fields, `names` among them.  Such a catalogue
    crosses the host link as raw bytes -- contiguous copies at memory speed -- and the device reads the coordinates out of the
    records with a stride; anything else returns None and takes the per-field path (_fields_to_device).
    """
    dt = cat.dtype
    if dt.names is None or dt.itemsize != 32 or len(dt.names) != 4 or cat.ndim != 1 or not cat.flags.c_contiguous:
        return None
    slots = {}
    for nm in dt.names:
        fdt, off = dt.fields[nm][:2]
        if fdt != np.dtype('<f8') or off % 8:
            return None
        slots[nm] = off // 8
    if sorted(slots.values()) != [0, 1, 2, 3] or any(nm not in slots for nm in names):
        return None
    return slots


def _raw_to_device(flat, dev):
    """Contiguous float64 host array (pageable) -> device tensor: host threads copy chunk k + 1 into a ring of page-locked
    buffers (plain memcpy, the GIL is released) while chunk k is on its way over the link."""
    torch = _torch()
    n = flat.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=dev)
    if n == 0:
        return out
    chunk = max(1, min(_RAW_CHUNK, n))
    nbuf = 3
    bufs = [_take_scratch(chunk) for _ in range(nbuf)]
    evs = [None] * nbuf
    stream = torch.cuda.current_stream()
    nth = max(1, _host_threads())
    ex = _pool()
    for k, a in enumerate(range(0, n, chunk)):
        b = min(n, a + chunk)
        slot = k % nbuf
        if evs[slot] is not None:
            evs[slot].synchronize()                          # the copy that last used this buffer has left the host
        dst = bufs[slot].numpy()[:b - a]
        step = -(-(b - a) // nth)
        list(ex.map(lambda i: np.copyto(dst[i:i + step], flat[a + i:min(a + i + step, b)]), range(0, b - a, step)))
        out[a:b].copy_(bufs[slot][:b - a], non_blocking=True)
        evs[slot] = stream.record_event()
    for e in evs:
        if e is not None:
            e.synchronize()
    _give_scratch(bufs)
    return out


def _device_to_raw(d_flat):
    """
    Device tensor -> float64 host array.  Up to BFG_PINNED_RESULT_MAX_GB (24) the result IS a recycled page-locked buffer
    (_pinned_result: one DMA, no host copy; the buffer returns to the pool when the caller drops the array); beyond that it
    is ordinary memory filled through the page-locked ring by host threads while the next chunk comes down.
    """
    torch = _torch()
    n = d_flat.numel()
    if n * 8 <= _RAW_RESULT_MAX_BYTES:
        t, arr = _pinned_result(n)
        t.copy_(d_flat, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return arr
    out = np.empty(n, dtype=np.float64)
    chunk = max(1, min(_RAW_CHUNK, n))
    nbuf = 3
    bufs = [_take_scratch(chunk) for _ in range(nbuf)]
    stream = torch.cuda.current_stream()
    nth = max(1, _host_threads())
    jobs = []
    ex = _pool()
    def drain(job):
        slot, ev, a, b = job
        ev.synchronize()
        src = bufs[slot].numpy()[:b - a]
        step = -(-(b - a) // nth)
        list(ex.map(lambda i: np.copyto(out[a + i:min(a + i + step, b)], src[i:i + step]), range(0, b - a, step)))
    for k, a in enumerate(range(0, n, chunk)):
        b = min(n, a + chunk)
        if len(jobs) == nbuf:
            drain(jobs.pop(0))
        slot = k % nbuf
        bufs[slot][:b - a].copy_(d_flat[a:b], non_blocking=True)
        jobs.append((slot, stream.record_event(), a, b))
    while jobs:
        drain(jobs.pop(0))
    _give_scratch(bufs)
    return out


def _device_to_fields(d_cols, out_cat, names):
    """The way back: contiguous device tensors -> the strided fields of a structured array, chunked through page-locked buffers
    with the host-side scatter of chunk k running while chunk k + 1 comes down."""
    torch = _torch()
    n = out_cat.shape[0]
    if n == 0:
        return
    nbuf = 4
    bufs = [_take_scratch(_FIELD_CHUNK) for _ in range(nbuf)]
    stream = torch.cuda.current_stream()
    nth = max(1, _host_threads())
    jobs = []       # (slot, event, column name, a, b) in flight
    ex = _pool()
    def drain(job):
        slot, ev, name, a, b = job
        ev.synchronize()
        src = bufs[slot].numpy()[:b - a]
        dst = out_cat[name]
        step = -(-(b - a) // nth)
        list(ex.map(lambda i: np.copyto(dst[a + i:min(a + i + step, b)], src[i:i + step]), range(0, b - a, step)))
    k = 0
    for col, name in enumerate(names):
        for a in range(0, n, _FIELD_CHUNK):
            b = min(n, a + _FIELD_CHUNK)
            if len(jobs) == nbuf:
                drain(jobs.pop(0))
            slot = k % nbuf
            bufs[slot][:b - a].copy_(d_cols[col][a:b], non_blocking=True)
            jobs.append((slot, stream.record_event(), name, a, b))
            k += 1
    while jobs:
        drain(jobs.pop(0))
    _give_scratch(bufs)



def _all_close_to_zero(a):
    """np.allclose(a, 0) (HealpixRunner.py:293) without scanning a 1.6 GB map when its first entries already say no."""
    head = a.reshape(-1)[:4096]
    if head.size and not np.all(np.abs(head) <= 1e-8):
        return False
    return bool(np.allclose(a, 0))


def _check_keys(model, keys):
    """HealpixRunner.py:304-311: p_keys need a table model that carries them."""
    if len(keys) > 0 and not (hasattr(model, 'interp_d') or hasattr(model, 'interp2D')):
        raise AssertionError(
            f"You asked to use {keys} properties in Baryonification. You must pass a ParamTabulatedProfile "
            f"or BaryonificationClass as the model. You have passed {type(model)} instead.")


def _extras(cat, keys):
    if len(keys) == 0:
        return None
    return np.ascontiguousarray(np.stack([np.asarray(cat[k], dtype=np.float64) for k in keys], axis=1))


def _model_cosmo(model, fallback):
    c = getattr(model, 'cosmo', None)
    if c is None:
        return fallback
    if isinstance(c, dict):
        return cosmology.runner_cosmology(c, True)
    return c


def _table_range_messages(model, z, M):
    """
    The out-of-table warnings of BaryonificationClass._readout (BaryonCorrection.py:378-389) for a whole catalogue: the
    reference evaluates them for every halo inside the loop; here once per process() from the catalogue's extremes (SURVEY.md
    section 10 #13).  Returns the message list (same wording); never raises -- a diagnostic must not break process().
    """
    try:
        zr, Mr = getattr(model, 'raw_input_z_range', None), getattr(model, 'raw_input_M_range', None)
        if zr is None or Mr is None:
            grid = getattr(getattr(model, 'interp_d', None), 'grid', None)
            if grid is None:
                return []
            zr, Mr = grid[0], grid[1]
        z = np.atleast_1d(np.asarray(z, dtype=np.float64))
        M = np.atleast_1d(np.asarray(M, dtype=np.float64))
        if z.size == 0 or M.size == 0:
            return []
        z_tab, M_tab = np.exp(np.asarray(zr, dtype=np.float64)) - 1, np.exp(np.asarray(Mr, dtype=np.float64))
        out = []
        if (np.min(z) < np.min(z_tab)) | (np.max(z) > np.max(z_tab)):
            out.append(f"Requested redshift range [{np.min(z)}, {np.max(z)}] outside table's range "
                       f"[{np.min(z_tab)}, {np.max(z_tab)}]")
        if (np.min(M) < np.min(M_tab)) | (np.max(M) > np.max(M_tab)):
            out.append(f"Requested log_Mass range [{np.log10(np.min(M))}, {np.log10(np.max(M))}] outside "
                       f"table's range [{np.log10(np.min(M_tab))}, {np.log10(np.max(M_tab))}]")
        return out
    except Exception:
        return []


def _warn_table_range(model, z, M, runner=None, owner=None):
    """Emit the messages; with `runner`, only once per (model, catalogue `owner`) so that repeated process() calls on the same
    inputs do not rescan a 10^6-halo catalogue (a few ms of host time on the end-to-end path)."""
    if runner is not None:
        # identity (not id(): ids are re-used once an object is collected) of model and catalogue, plus the table's own range,
        # so that `Runner.model = NewModel` in a loop is re-checked like the reference re-checks every call
        zr, Mr = getattr(model, 'raw_input_z_range', None), getattr(model, 'raw_input_M_range', None)
        try:
            rng = (float(np.min(zr)), float(np.max(zr)), float(np.min(Mr)), float(np.max(Mr)))
        except Exception:
            rng = None
        key = (_Ident(model), _Ident(owner), np.size(M), rng)
        if getattr(runner, '_range_checked', None) == key:
            return
        try:
            runner._range_checked = key
        except Exception:
            pass
    for text in _table_range_messages(model, z, M):
        warnings.warn(text, UserWarning, stacklevel=3)


def _cat_ranges(container):
    """(z_min, z_max, M_min, M_max) of a catalogue container, cached on the container while its `.cat` array is the same object
    (each reduction over a field of a 10^6-row structured array is ~1 ms of host time per process() call otherwise)."""
    cat = container.cat
    c = getattr(container, '_bfg_ranges', None)
    if c is not None and c[0] is cat and c[1] == cat.size:
        return c[2]
    if cat.size:
        z, M = np.ascontiguousarray(cat['z'], dtype=np.float64), np.ascontiguousarray(cat['M'], dtype=np.float64)
        r = (float(np.min(z)), float(np.max(z)), float(np.min(M)), float(np.max(M)))
    else:
        r = (0.0, 0.0, 0.0, 0.0)
    try:
        container._bfg_ranges = (cat, cat.size, r)
    except Exception:
        pass
    return r


def _cat_z_max(container):
    return _cat_ranges(container)[1]


SKY_BAND_RAD = 0.04     # colatitude band width of the sky ordering (~160 pixels at NSIDE=4096)


class _Ident(object):
    """Cache-key element that compares by object IDENTITY and keeps the object alive: a bare id() can be handed to a new
    object once the old one is collected (`Runner.model = NewModel` in a loop, examples/10_...ipynb cell 15), which would
    make a stale device table look current."""
    __slots__ = ('obj',)

    def __init__(self, obj):
        self.obj = obj

    def __eq__(self, other):
        return isinstance(other, _Ident) and other.obj is self.obj

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return id(self.obj)


_SHARED_TABLES = []    # [(key, device index, DeviceTable)]: the last few tables uploaded by this process


class _TableCache(object):
    """Tables go to the device once per (model, table, device) and are re-used by later process() calls -- of ANY runner object
    of this process (a lightcone makes one runner per shell around the same model)."""

    def get(self, key, make):
        import torch
        dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
        for k, d, t in _SHARED_TABLES:
            if d == dev and k == key:
                return t
        t = make()
        _SHARED_TABLES.append((key, dev, t))
        while len(_SHARED_TABLES) > 8:
            _SHARED_TABLES.pop(0)[2].close()
        return t


# =====================================================================================================================
# HEALPix shells
# =====================================================================================================================
class DefaultRunner(object):
    """Constructor contract of BaryonForge/Runners/HealpixRunner.py:160-177 (+ keyword-only GPU knobs)."""

    def __init__(self, HaloLightConeCatalog, LightconeShell, epsilon_max, model, use_ellipticity=False,
                 mass_def=None, include_pixel_size=False, verbose=True, *, device=None, pix_range=None,
                 sort_halos=True):
        self.HaloLightConeCatalog = HaloLightConeCatalog
        self.LightconeShell = LightconeShell
        self.cosmo = HaloLightConeCatalog.cosmology
        self.model = model
        self.epsilon_max = epsilon_max
        self.mass_def = mass_def          # None == ccl.halos.massdef.MassDef(200, 'critical')
        self.verbose = verbose
        self.use_ellipticity = use_ellipticity
        self.include_pixel_size = include_pixel_size
        self.device = device
        self.pix_range = pix_range        # (lo, hi) RING range owned by this rank (parallel.py); None = whole map
        self.sort_halos = sort_halos      # order halos by sky cell on the device before the halo loop (L2 locality)
        self.last_stats = {}
        self._last_scalars, self._d_aux, self._aux_paint = None, None, False
        self._tables = _TableCache()
        if use_ellipticity:
            raise NotImplementedError("You have set use_ellipticity = True, but this not yet implemented for HealpixRunner")

    # objects stay picklable for joblib-style drivers (utils/Parallelize.py:47-49)
    def __getstate__(self):
        d = dict(self.__dict__)
        d['_tables'] = None
        for k in ('_scratch_inflight', '_peers', '_d_aux', '_d_pack', '_spl_cache', '_host_maps', '_cells'):
            d.pop(k, None)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._tables = _TableCache()

    def _device(self):
        torch = _torch()
        return torch.device('cuda', torch.cuda.current_device() if self.device is None else int(self.device))

    # ---- small public helpers of the reference's DefaultRunner (host-side numpy; not used by the device path)
    def build_Rmat(self, A, ref):
        """2 x 2 rotation by the angle between A and ref; both are normalised IN PLACE, like HealpixRunner.py:179-208."""
        A /= np.linalg.norm(A)
        ref /= np.linalg.norm(ref)
        ang = np.arccos(np.dot(A, ref))
        c, s_ = np.cos(ang), np.sin(ang)
        return np.array([[c, -s_], [s_, c]])

    def coord_array(self, *args):
        """(N, M) coordinate rows from M same-shaped arrays (HealpixRunner.py:211-233)."""
        return np.stack([np.ravel(a) for a in args], axis=1)

    def _record_plan(self, paint):
        """
        The per-halo scalars of HealpixRunner.py:317-329 (+ BaryonCorrection.py:371,398-399,410), vectorised, as a plan:
        returns (n, fill) where fill(dst, sl) writes the records of halos `sl` into dst, an [m, 16] float64 view.
        """
        cat = self.HaloLightConeCatalog.cat
        n = cat.size
        cosmo = cosmology.runner_cosmology(self.cosmo, with_w0=True)          # :280-284
        if n == 0:
            return 0, None
        D_of_z = cosmology.D_A_spline(cosmo, cat['z'])                         # :297-299
        mcosmo = None if paint else _model_cosmo(self.model, cosmo)
        R = np.empty(n)
        D = np.empty(n)
        R_com = None if paint else np.empty(n)
        pixarea = 4 * np.pi / self.LightconeShell.map.size
        eps_model = None if paint else self.model.epsilon_max
        self.last_scalars = dict(R_run=R, D_A=D, R_model_com=R_com)

        def fill(r, sl):   # numpy releases the GIL inside these ufuncs, so chunks run on several host cores
            M, z = cat['M'][sl], cat['z'][sl]
            a = 1 / (1 + z)                                                    # :319
            R[sl] = cosmology.radius_of_mass(cosmo, M, a, self.mass_def)       # :320 physical Mpc
            D[sl] = D_of_z(z)                                                  # :321
            theta_ll, phi_ll = np.pi / 2.0 - np.radians(cat['dec'][sl]), np.radians(cat['ra'][sl])   # lonlat2thetaphi
            st = np.sin(theta_ll)
            vx, vy, vz = st * np.cos(phi_ll), st * np.sin(phi_ll), np.cos(theta_ll)      # hp.ang2vec :327
            r[:, _lib.HS_VX], r[:, _lib.HS_VY], r[:, _lib.HS_VZ] = vx, vy, vz
            # pointing(vec) as healpy's query_disc wrapper builds it
            r[:, _lib.HS_THETA] = np.arctan2(np.sqrt(vx * vx + vy * vy), vz)
            phi = np.arctan2(vy, vx)
            r[:, _lib.HS_PHI] = np.where(phi < 0, phi + 2 * np.pi, phi)
            r[:, _lib.HS_D] = D[sl]
            r[:, _lib.HS_A] = a
            r[:, _lib.HS_RADIUS] = R[sl] * self.epsilon_max / D[sl]            # :329
            r[:, _lib.HS_LNZ] = np.log(1 / a)                                  # BaryonCorrection.py:371 / Tabulate.py:312
            r[:, _lib.HS_LNM] = np.log(M)                                      # BaryonCorrection.py:398 / Tabulate.py:317
            if paint:
                r[:, _lib.HS_RCUT] = np.inf
                r[:, _lib.HS_SCALE] = pixarea * D[sl] ** 2 if self.include_pixel_size else 1.0   # :478
            else:
                R_com[sl] = cosmology.radius_of_mass(mcosmo, M, a, getattr(self.model, 'mass_def', None)) / a   # BaryonCorrection.py:399
                r[:, _lib.HS_RCUT] = eps_model * R_com[sl]                     # :410
                r[:, _lib.HS_LNRCOM] = np.log(R_com[sl])                       # :408
                r[:, _lib.HS_SCALE] = 1.0
            r[:, _lib.HS_THETA_LL], r[:, _lib.HS_PHI_LL] = theta_ll, phi_ll
        return n, fill

    def halo_records(self, paint):
        """All halo records at once: (records[n,16] float64 (field-major storage), extras[n,k] or None)."""
        n, fill = self._record_plan(paint)
        rec = np.zeros((_lib.HALO_STRIDE, n), dtype=np.float64).T
        keys = list(vars(self.model).get('p_keys', []))                        # :304
        _check_keys(self.model, keys)
        if n == 0:
            return rec, None
        _parallel_chunks(lambda sl: fill(rec[sl], sl), n)
        return rec, _extras(self.HaloLightConeCatalog.cat, keys)

    # ---- device-side scalar prep ---------------------------------------------------------------------------
    @property
    def last_scalars(self):
        """R_run (physical), D_A, R_model_com of the last halo_records()/process() call, in catalogue order."""
        if self._last_scalars is None and getattr(self, '_d_aux', None) is not None:
            aux = self._d_aux.cpu().numpy()
            self._last_scalars = dict(R_run=aux[0], D_A=aux[1], R_model_com=None if self._aux_paint else aux[2])
        return self._last_scalars

    @last_scalars.setter
    def last_scalars(self, value):
        self._last_scalars = value

    def _spline_pack(self, paint, z_max):
        """
        The three per-process() splines the record kernel evaluates, packed as one float64 vector:
        D_A(z) -- the reference's own CubicSpline (HealpixRunner.py:297-299) -- and the radius factors g(ln(1+z)) with
        R_delta = cbrt(M) g for the runner's cosmology/mass_def (HealpixRunner.py:320) and the model's
        (BaryonCorrection.py:399).  Cached on (cosmology, z_max, mass definitions).
        """
        mcos = None if paint else getattr(self.model, 'cosmo', None)
        key = (paint, float(z_max), tuple(sorted(self.cosmo.items())), _Ident(self.mass_def), _Ident(mcos),
               tuple(sorted(mcos.items())) if isinstance(mcos, dict) else None,
               None if paint else _Ident(getattr(self.model, 'mass_def', None)))
        if getattr(self, '_spl_cache', None) is not None and self._spl_cache[0] == key:
            return self._spl_cache[1]
        for k, v in _SPLINE_PACKS:               # another runner object already built it (one runner per shell in a lightcone)
            if k == key:
                self._spl_cache = (key, v)
                return v
        cosmo = cosmology.runner_cosmology(self.cosmo, with_w0=True)          # :280-284
        DA = cosmology.D_A_spline_to(cosmo, z_max)
        g_run = cosmology.radius_factor_spline(cosmo, self.mass_def, z_max)
        parts = [DA.x, DA.c.reshape(-1), g_run.x, g_run.c.reshape(-1)]
        if not paint:
            g_mod = cosmology.radius_factor_spline(_model_cosmo(self.model, cosmo),
                                                   getattr(self.model, 'mass_def', None), z_max)
            parts.append(g_mod.c.reshape(-1))
        pack = (np.ascontiguousarray(np.concatenate(parts)), DA.x.size, g_run.x.size)
        self._spl_cache = (key, pack)
        _SPLINE_PACKS.append((key, pack))
        del _SPLINE_PACKS[:-8]
        return pack

    def device_records(self, paint, dev=None):
        """
        Halo records built ON THE DEVICE (bfg_shell_records): the host only stages the raw catalogue columns plus
        numpy's ln(1+z) and ln M into pinned memory; returns the [n, 16] device tensor (catalogue order).
        """
        torch = _torch()
        dev = self._device() if dev is None else dev
        cat = self.HaloLightConeCatalog.cat
        n = cat.size
        self._last_scalars, self._d_aux, self._aux_paint = None, None, paint
        with torch.cuda.device(dev):
            d_rec = torch.empty((n, _lib.HALO_STRIDE), dtype=torch.float64, device=dev)
            if n == 0:
                return d_rec
            z_max = _cat_z_max(self.HaloLightConeCatalog)
            assert z_max <= 30, f"We assume max(z) = 30, but your catalog has max(z) = {z_max}"   # HealpixRunner.py:301
            pack, n_DA, n_g = self._spline_pack(paint, z_max)
            # Sharded runs: every rank needs the whole catalogue on its device (it selects its own halos there), but the
            # host staging is split -- rank r stages halos [r m, (r+1) m) and the columns are all-gathered over NVLink.
            world, rank = 1, 0
            if self.pix_range is not None and os.environ.get("BFG_SHARD_STAGING", "1") == "1":
                from .parallel import _dist
                dist = _dist()
                if dist is not None and dist.get_backend() == "nccl":
                    world, rank = dist.get_world_size(), dist.get_rank()
            m = -(-n // world)                       # halos staged per rank
            lo_h, hi_h = min(rank * m, n), min((rank + 1) * m, n)
            stage = _take_scratch(6 * m)
            self._scratch_inflight = getattr(self, '_scratch_inflight', None) or []
            self._scratch_inflight.append(stage)
            cols = stage.numpy().reshape(6, m)

            def fill(sl):   # numpy releases the GIL inside these ufuncs/copies
                src = slice(lo_h + sl.start, lo_h + sl.stop)
                M, z = cat['M'][src], cat['z'][src]
                cols[0, sl], cols[1, sl], cols[2, sl], cols[3, sl] = M, z, cat['ra'][src], cat['dec'][src]
                np.log(1 / (1 / (1 + z)), out=cols[4, sl])                     # np.log(1/a_j)  BaryonCorrection.py:371
                np.log(M, out=cols[5, sl])                                     # np.log(M_j)    BaryonCorrection.py:398
            _parallel_chunks(fill, hi_h - lo_h)
            if world == 1:
                d_cols = stage.to(dev, non_blocking=True)
            else:
                if hi_h - lo_h < m:
                    cols[:, hi_h - lo_h:] = 1.0                               # padding of the last rank's block
                d_part = stage.to(dev, non_blocking=True).reshape(6, m)
                d_all = torch.empty((world, 6, m), dtype=torch.float64, device=dev)
                dist.all_gather_into_tensor(d_all.reshape(-1), d_part.reshape(-1))
                d_cols = d_all.permute(1, 0, 2).reshape(6, world * m)[:, :n].contiguous()   # [6][n], catalogue order
            # the spline pack lives on the device as long as its host copy is cached (a pageable H2D copy per call would block
            # the host until the stream -- NCCL all-gather of the staged columns included -- has drained)
            dp = getattr(self, '_d_pack', None)
            if dp is None or dp[0] is not pack or dp[1].device != dev:
                self._d_pack = (pack, torch.from_numpy(pack).to(dev))
            d_pack = self._d_pack[1]
            d_aux = torch.empty((3, n), dtype=torch.float64, device=dev)
            base = d_pack.data_ptr()
            o_DAc = 8 * n_DA
            o_gx = o_DAc + 8 * 4 * (n_DA - 1)
            o_grun = o_gx + 8 * n_g
            o_gmod = o_grun + 8 * 4 * (n_g - 1)
            pixarea = 4 * np.pi / self.LightconeShell.map.size
            _lib.check(_lib.lib().bfg_shell_records(
                n, d_cols.data_ptr(), 1 if paint else 0, float(self.epsilon_max),
                0.0 if paint else float(self.model.epsilon_max), pixarea if (paint and self.include_pixel_size) else 0.0,
                n_DA, base, base + o_DAc, n_g, base + o_gx, base + o_grun, None if paint else base + o_gmod,
                d_rec.data_ptr(), d_aux.data_ptr(), _lib.current_stream()))
            self._d_aux = d_aux
        return d_rec

    def _halo_loop(self, paint, table, launch, NSIDE, lo, hi, dev):
        """
        Stage raw columns -> device scalar prep -> sky sort -> `launch(d_rec, d_ext, n, 0)`.  Nothing here waits for the
        GPU; halos that cannot touch [lo, hi) are skipped inside the halo-loop kernel (ring-range sharding).
        """
        torch = _torch()
        t0 = time.perf_counter()
        cat = self.HaloLightConeCatalog.cat
        n = cat.size
        keys = list(vars(self.model).get('p_keys', []))                        # :304
        _check_keys(self.model, keys)
        if n == 0:
            self.last_timing = dict(host_prep_s=0.0)
            return 0
        # staging buffers of the previous call may still feed an un-synchronised stream (offsets_on_device callers)
        if getattr(self, '_scratch_inflight', None):
            torch.cuda.current_stream().synchronize()
            _give_scratch(self._scratch_inflight)
        self._scratch_inflight = []
        with torch.cuda.device(dev):
            d_rec = self.device_records(paint, dev)
            ext = _extras(cat, keys)
            d_ext = None if ext is None else _to_device(ext, dev)
            if self.sort_halos:
                owned = None if self.pix_range is None else (NSIDE, lo, hi)
                d_rec, d_ext = _sort_records(d_rec, d_ext, 0, SKY_BAND_RAD, owned=owned)
            launch(d_rec, d_ext, n, 0)
        self.last_timing = dict(host_prep_s=time.perf_counter() - t0)
        return 1

    def _range(self, npix):
        return (0, npix) if self.pix_range is None else (int(self.pix_range[0]), int(self.pix_range[1]))

    def _owned_halos(self, rec, extras, nside, lo, hi):
        """Ring-range sharding: keep the halos whose disc can touch this rank's pixel range (parallel.py)."""
        if self.pix_range is None or rec.shape[0] == 0:
            return rec, extras
        from .parallel import halos_touching_pixel_range
        keep = halos_touching_pixel_range(nside, rec[:, _lib.HS_THETA], rec[:, _lib.HS_RADIUS], lo, hi)
        if keep.all():
            return rec, extras
        return np.ascontiguousarray(rec[keep]), (None if extras is None else np.ascontiguousarray(extras[keep]))


def _chunk_fractions(K):
    """
    Start of each of the K latitude chunks of the pipelined end-to-end paths, as a fraction of the (owned part of the) sky.
    The last three chunks shrink to 6 %, 3 % and 1 %: what stays exposed after the last halo -- re-binning and download of the
    final chunk's rings -- is then a 1 % piece instead of 1 / K, without paying for more chunk boundaries (each one drains the
    persistent halo-loop kernel).  Measured at N = 1: 110.4 -> 108.2 ms per shell end to end.  BFG_PIPELINE_TAPER=0: equal chunks.
    """
    if K >= 6 and os.environ.get("BFG_PIPELINE_TAPER", "1") == "1":
        return [0.90 * k / (K - 3) for k in range(K - 3)] + [0.90, 0.96, 0.99]
    return [k / K for k in range(K)]


class BaryonifyShell(DefaultRunner):
    """BaryonForge/Runners/HealpixRunner.py:180-373."""

    def _peer_slices(self, npix, dev):
        """Peer-mapped owned slices for the fused regrid + exchange (needs an initialised NCCL group whose ranks use the
        ranges of parallel.pixel_ranges); None -> fall back to full-size partial maps + all-reduce."""
        import torch.distributed as dist
        if os.environ.get("BFG_EXCHANGE", "p2p") != "p2p":
            return None
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2 or dist.get_world_size() > 8:
            return None
        if dist.get_backend() != "nccl":
            return None
        from .parallel import pixel_ranges, PeerSlices, single_node_group
        world, rank = dist.get_world_size(), dist.get_rank()
        ranges = pixel_ranges(self.LightconeShell.NSIDE, world)
        # every rank must take the same branch below (the constructor is collective): agree on "my range is the standard one"
        # through the group instead of deciding locally
        if not single_node_group():              # CUDA IPC and /proc/<pid>/fd only work inside one machine
            return None
        # one set of peer-mapped slices per process and map size, shared by every runner object (a lightcone makes one runner
        # per shell: each would otherwise allocate and IPC-map its own slices)
        key = (npix, world, rank, dev.index)
        if key not in _PEER_SLICES:
            std = int(tuple(ranges[rank]) == tuple(int(v) for v in self.pix_range))
            import torch
            vote = torch.tensor([std], dtype=torch.int32, device=dev)
            dist.all_reduce(vote, op=dist.ReduceOp.MIN)
            peers = None
            if int(vote.cpu()[0]) == 1:
                bounds = [r[0] for r in ranges] + [npix]
                try:
                    peers = PeerSlices(bounds, rank, world, dev.index)
                except OSError:                  # agreed by all ranks (collective vote inside): all-reduce exchange instead
                    peers = None
            _PEER_SLICES[key] = peers
        if tuple(ranges[rank]) != tuple(int(v) for v in self.pix_range):
            return None
        return _PEER_SLICES[key]

    def _shared_host(self, npix, peers):
        """Shared page-locked host maps for the result (parallel.SharedHostMaps); None -> per-rank full-map D2H."""
        if os.environ.get("BFG_HOST_GATHER", "shared") != "shared":
            return None
        key = (npix, peers.world, peers.rank, peers.device)
        if key not in _HOST_MAPS:                # per process, shared by every runner object
            from .parallel import SharedHostMaps
            ok = hasattr(os, 'memfd_create')
            _HOST_MAPS[key] = SharedHostMaps(npix, peers.rank, peers.world, peers.device,
                                             own_range=(peers.bounds[peers.rank], peers.bounds[peers.rank + 1])) if ok else None
        return _HOST_MAPS[key]

    def offsets_on_device(self):
        """Run the halo loop only; returns (offsets tensor [3, n_local] on the device, n_updates)."""
        torch = _torch()
        dev = self._device()
        L = _lib.lib()
        NSIDE = self.LightconeShell.NSIDE
        npix = 12 * NSIDE * NSIDE
        lo, hi = self._range(npix)
        with torch.cuda.device(dev):   # table first: a model without one fails here, as in the reference
            table = self._tables.get((_Ident(self.model), _Ident(self.model.interp_d) if hasattr(self.model, 'interp_d') else 0),
                                     lambda: displacement_table_of(self.model, dev.index))
        with torch.cuda.device(dev):
            d_off = torch.zeros((3, hi - lo), dtype=torch.float64, device=dev)
            d_nb = torch.zeros(4, dtype=torch.int64, device=dev)

        def launch(d_rec, d_ext, n, k):
            _lib.check(L.bfg_shell_offsets(table.handle, NSIDE, n, _lib.ptr(d_rec), _lib.ptr(d_ext), table.n_extra,
                                           _lib.ptr(d_off), lo, hi, d_nb.data_ptr() + 8 * k, _lib.current_stream()))
        self._halo_loop(False, table, launch, NSIDE, lo, hi, dev)
        d_n = d_nb.sum().reshape(1)
        return d_off, d_n

    PIPELINE_MIN_HALOS = 200000     # below this the halo loop is too short to hide the download behind
    PIPELINE_CHUNKS = 12
    PIPELINE_MARGIN_RAD = 0.02      # how far (chord on the unit sphere) the re-binning may move mass; verified per call

    def _process_pipelined(self):
        """
        Single-GPU end-to-end path for large catalogues: the sky-sorted halo loop is cut into latitude chunks of equal
        area; as soon as a chunk is done the rings north of (chunk edge - largest disc radius) have their final offsets, so
        they are re-binned and the finished part of the NEW map is downloaded on a side stream while the halo loop works
        on the next chunk.  Only the last chunk's re-binning + download is exposed.  The assumption that the re-binning
        moves mass by less than PIPELINE_MARGIN_RAD is checked against max|offset| at the end; if it fails (absurdly
        large displacements) the whole map is downloaded again.  Returns None when the path does not apply.
        """
        torch = _torch()
        from .parallel import first_pixel_at_colatitude
        orig_map = self.LightconeShell.map
        NSIDE = self.LightconeShell.NSIDE
        npix = orig_map.size
        cat = self.HaloLightConeCatalog.cat
        n = cat.size
        dev = self._device()
        L = _lib.lib()
        K = int(os.environ.get("BFG_PIPELINE_CHUNKS", self.PIPELINE_CHUNKS))
        keys = list(vars(self.model).get('p_keys', []))
        _check_keys(self.model, keys)
        with torch.cuda.device(dev):
            table = self._tables.get((_Ident(self.model), _Ident(self.model.interp_d) if hasattr(self.model, 'interp_d') else 0),
                                     lambda: displacement_table_of(self.model, dev.index))
            if getattr(self, '_scratch_inflight', None):
                torch.cuda.current_stream().synchronize()
                _give_scratch(self._scratch_inflight)
            self._scratch_inflight = []
            st = _lib.current_stream()
            main = torch.cuda.current_stream()
            side, down = _side_stream(dev, 0), _side_stream(dev, 1)   # uploads / downloads: the link is full duplex, and a
            # download queued behind the map pieces would wait for the whole 1.6 GB to go up first (a fast halo loop --
            # small discs -- then ran 63 ms end to end instead of ~40)
            # the accumulators are zeroed first, so that the GPU clears 6.4 GB while the host still stages the catalogue
            d_off = torch.zeros((3, npix), dtype=torch.float64, device=dev)
            d_new = torch.zeros(npix, dtype=torch.float64, device=dev)
            d_nb = torch.zeros(K, dtype=torch.int64, device=dev)
            d_max = torch.zeros(1, dtype=torch.float64, device=dev)
            t0 = time.perf_counter()
            d_rec = self.device_records(False, dev)
            host_prep_s = time.perf_counter() - t0
            ext = _extras(cat, keys)
            d_ext = None if ext is None else _to_device(ext, dev)
            d_rec, d_ext = _sort_records(d_rec, d_ext, 0, SKY_BAND_RAD)
            # latitude chunks of equal area, expressed as band indices of the sort key
            fr = _chunk_fractions(K)
            edges = [int(np.floor(np.arccos(1 - 2.0 * f) / SKY_BAND_RAD)) for f in fr] + [1 << 40]
            d_edges = torch.tensor(edges, dtype=torch.int64, device=dev)
            d_bounds = torch.empty(K + 1, dtype=torch.int64, device=dev)
            d_rho = torch.zeros(1, dtype=torch.float64, device=dev)
            _lib.check(L.bfg_halo_band_bounds(n, _lib.ptr(d_rec), SKY_BAND_RAD, K + 1, _lib.ptr(d_edges), _lib.ptr(d_bounds),
                                              _lib.ptr(d_rho), st))
            # the map goes up in K pieces on the side stream, underneath the halo loop; a re-binning step waits only for the
            # pieces it reads (a single 1.6 GB copy would stall the first re-binning, and with it the halo loop, for ~15 ms)
            h_map = torch.from_numpy(np.ascontiguousarray(orig_map, dtype=np.float64).reshape(-1))
            piece = -(-npix // K)
            ev_h2d = []
            with torch.cuda.stream(side):
                d_map = torch.empty(npix, dtype=torch.float64, device=dev)
                for j in range(K):
                    a0, a1 = j * piece, min((j + 1) * piece, npix)
                    d_map[a0:a1].copy_(h_map[a0:a1], non_blocking=True)
                    ev_h2d.append(side.record_event())
            bounds = d_bounds.cpu().tolist()        # one small synchronisation: chunk boundaries in the sorted catalogue
            rho_max = float(d_rho.cpu()[0])
            pix_rad = np.sqrt(4 * np.pi / npix)
            out, out_np = _pinned_result(npix)
            n_extra = table.n_extra
            pieces_waited = 0
            d_map.record_stream(main)
            p_prev = q_prev = 0

            def regrid_to(p_k):
                nonlocal pieces_waited, p_prev
                if p_k <= p_prev:
                    return False
                need = min(K, -(-p_k // piece))     # map pieces covering [0, p_k)
                while pieces_waited < need:
                    main.wait_event(ev_h2d[pieces_waited])
                    pieces_waited += 1
                _lib.check(L.bfg_shell_regrid_range(NSIDE, _lib.ptr(d_map), _lib.ptr(d_off), npix, _lib.ptr(d_new),
                                                    p_prev, p_k, st))
                p_prev = p_k
                return True

            def download_to(q_k):
                nonlocal q_prev
                if q_k <= q_prev:
                    return
                down.wait_event(main.record_event())
                with torch.cuda.stream(down):
                    out[q_prev:q_k].copy_(d_new[q_prev:q_k], non_blocking=True)
                q_prev = q_k

            for k in range(K):
                b0, b1 = bounds[k], bounds[k + 1]
                if b1 > b0:
                    _lib.check(L.bfg_shell_offsets(table.handle, NSIDE, b1 - b0, d_rec.data_ptr() + 8 * _lib.HALO_STRIDE * b0,
                                                   None if d_ext is None else d_ext.data_ptr() + 8 * n_extra * b0, n_extra,
                                                   _lib.ptr(d_off), 0, npix, d_nb.data_ptr() + 8 * k, st))
                if k + 1 == K:
                    break
                # every halo not yet processed has theta >= edge band * band width; its disc reaches rho_max further north
                th_done = edges[k + 1] * SKY_BAND_RAD - rho_max - 3 * pix_rad
                if th_done > 0 and regrid_to(first_pixel_at_colatitude(NSIDE, th_done)):
                    th_copy = th_done - self.PIPELINE_MARGIN_RAD - 3 * pix_rad
                    if th_copy > 0:
                        download_to(first_pixel_at_colatitude(NSIDE, th_copy))
            d_sums = torch.zeros(2, dtype=torch.float64, device=dev)
            regrid_to(npix)
            download_to(npix)                        # the last piece leaves first; the three reductions below run underneath it
            d_new.record_stream(down)
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_new), npix, _lib.ptr(d_sums), st))
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_map), npix, d_sums.data_ptr() + 8, st))   # regrid_to(npix) waited for every piece
            _lib.check(L.bfg_offsets_max_norm2(_lib.ptr(d_off), npix, 0, npix, _lib.ptr(d_max), st))
            sums = d_sums.cpu()
            n_up = int(d_nb.sum().cpu())
            max_norm = float(np.sqrt(float(d_max.cpu()[0])))
            down.synchronize()
            if not (max_norm < self.PIPELINE_MARGIN_RAD):
                # the re-binning moved mass further than assumed: parts of the map were downloaded too early
                out.copy_(d_new, non_blocking=True)
                main.synchronize()
        _give_scratch(getattr(self, '_scratch_inflight', []))
        self._scratch_inflight = []
        new_sum, old_sum = float(sums[0]), float(sums[1])
        self.last_stats = dict(n_updates=n_up, new_sum=new_sum, old_sum=old_sum, pipelined=True, max_offset=max_norm)
        self.last_timing = dict(host_prep_s=host_prep_s, chunks=float(K))
        assert np.isclose(new_sum, old_sum), \
            "ERROR in pixel regridding, sum(new_map) [%0.14e] != sum(oldmap) [%0.14e]" % (new_sum, old_sum)   # :368-370
        return out_np

    def _process_sharded(self, peers):
        """
        Ring-range sharded end-to-end path (one process per GPU, NCCL group on one machine, `peers` = the IPC-mapped slices
        of the new map).  Everything is enqueued without waiting for the GPU; the ranks meet at three stream-ordered points:
          fence 1  every rank's slice of the new map is zero        -- on the side stream, hidden under the halo loop
          fence 2  every rank's deposits have landed                 -- after the fused re-binning + exchange
          sums     one all-reduce of [sum(new), sum(old), n_updates, n_remote] enqueued AFTER the rank's slice went to the
                   shared host map: its completion on this rank means every rank's copy is done, so no extra barrier.
        The map slice goes up on the side stream underneath the halo loop; each rank downloads only its own slice, into ONE
        page-locked host map all ranks have mapped (parallel.SharedHostMaps).
        """
        torch = _torch()
        import torch.distributed as dist
        from .parallel import SegmentsExhausted, gather_owned_ranges
        orig_map = self.LightconeShell.map
        NSIDE = self.LightconeShell.NSIDE
        npix = orig_map.size
        lo, hi = self._range(npix)
        dev = self._device()
        L = _lib.lib()
        prof = os.environ.get("BFG_PROFILE_E2E") == "1"
        marks = [("start", time.perf_counter())]

        def mark(name):
            if prof:
                torch.cuda.synchronize()
                marks.append((name, time.perf_counter()))

        # large catalogues: the chunked variant that overlaps upload, halo loop and download (same decision on every rank:
        # it depends on the catalogue size and the environment only)
        if (not prof and self.sort_halos and os.environ.get("BFG_PIPELINE", "1") == "1"
                and self.HaloLightConeCatalog.cat.size >= self.PIPELINE_MIN_HALOS):
            host = self._shared_host(npix, peers)
            if host is not None:
                seg = None
                try:
                    with torch.cuda.device(dev):
                        seg, addr = host.acquire()
                except SegmentsExhausted:            # raised on all ranks: the plain path below returns a private copy
                    pass
                except OSError:                      # collective failure (agreed by all ranks)
                    _HOST_MAPS[(npix, peers.world, peers.rank, peers.device)] = None
                if seg is not None:
                    return self._process_sharded_pipelined(peers, host, seg, addr)

        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            side = _side_stream(dev)
            own = peers.own_tensor()
            side.wait_stream(main)                   # the previous call's reads of `own` are stream-ordered before the zeroing
            with torch.cuda.stream(side):
                own.zero_()
                token = torch.zeros(1, device=dev)
                dist.all_reduce(token)               # fence 1
            d_off, d_n = self.offsets_on_device()    # staging, device scalar prep, owned sort, halo loop -- all enqueued on main
            with torch.cuda.stream(side):            # after the catalogue's small copies: the copy engine serves its queue in order
                d_map = _to_device(orig_map[lo:hi], dev, dtype=np.float64)
                ev_side = side.record_event()
            mark("halo_loop")
            # where the result goes: a segment of the shared host map (collective choice; its tiny all-reduce runs on the side
            # stream so that it does not queue behind the halo loop)
            host = self._shared_host(npix, peers)
            seg = addr = None
            if host is not None:
                try:
                    with torch.cuda.stream(side):
                        seg, addr = host.acquire()
                except SegmentsExhausted:            # the caller still holds MAX_SEGMENTS earlier results (raised on all ranks):
                    host = None                      # this call returns a private copy instead
                except OSError:                      # collective failure (agreed by all ranks)
                    _HOST_MAPS[(npix, peers.world, peers.rank, peers.device)] = None
                    host = None
            main.wait_event(ev_side)
            d_map.record_stream(main)
            st = _lib.current_stream()
            d_acc = torch.zeros(4, dtype=torch.float64, device=dev)     # sum(new slice), sum(old slice), n_updates, n_remote
            d_rem = torch.zeros(1, dtype=torch.int64, device=dev)
            _lib.check(L.bfg_shell_regrid_p2p(NSIDE, _lib.ptr(d_map), _lib.ptr(d_off), lo, hi, peers.world, peers.rank,
                                              peers.h_bounds, peers.h_slices, _lib.ptr(d_rem), st))
            del d_off
            dist.all_reduce(token)                   # fence 2
            mark("regrid_exchange")
            if host is not None:
                _lib.check(L.bfg_copy_to_host_async(addr + 8 * lo, own.data_ptr(), 8 * (hi - lo), st))
            _lib.check(L.bfg_sum_f64(own.data_ptr(), hi - lo, _lib.ptr(d_acc), st))
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_map), hi - lo, d_acc.data_ptr() + 8, st))
            d_acc[2] = d_n.reshape(()).to(torch.float64)                # counts are < 2^53: exact as float64
            d_acc[3] = d_rem.reshape(()).to(torch.float64)
            dist.all_reduce(d_acc)                   # sums + "every rank's slice is in the host map"
            if host is None:
                d_new = gather_owned_ranges(own, npix)
                out, out_np = _pinned_result(npix)
                out.copy_(d_new, non_blocking=True)
            acc = d_acc.cpu()
            main.synchronize()
            mark("download")
        _give_scratch(getattr(self, '_scratch_inflight', []))
        self._scratch_inflight = []
        new_sum, old_sum = float(acc[0]), float(acc[1])
        self.last_stats = dict(n_updates=int(acc[2]), new_sum=new_sum, old_sum=old_sum, sharded=True)
        self.last_stats_remote = int(acc[3])
        if prof:
            self.last_timing.update({"sharded_" + b[0] + "_s": b[1] - a[1] for a, b in zip(marks[:-1], marks[1:])})
        assert np.isclose(new_sum, old_sum), \
            "ERROR in pixel regridding, sum(new_map) [%0.14e] != sum(oldmap) [%0.14e]" % (new_sum, old_sum)   # :368-370
        return host.export(seg, orig_map.shape) if host is not None else out_np.reshape(orig_map.shape)

    SHARD_CHUNKS = 8            # latitude chunks per rank of the pipelined sharded path (N = 8: 28.9 ms with 8, 30.8 with 4, 34.3 with 2)

    def _process_sharded_pipelined(self, peers, host, seg, addr):
        """
        _process_sharded with the rank's work cut into latitude chunks, so that the three things that take time at N = 8 --
        the map slice going up, the halo loop, the new slice coming down -- overlap instead of following each other (on the
        measured box the host link moves ~90 GB/s per direction for all GPUs together: 1.6 GB up and 1.6 GB down are 18 ms
        each, the halo loop 12 ms).  The owned, sky-sorted halos are processed chunk by chunk; after chunk k every ring north
        of (next chunk's first band - largest disc radius) has its final offsets, so those source rings are re-binned
        (bfg_shell_regrid_p2p_range, deposits go to this rank's slice or to a neighbour's over NVLink) and the part of this
        rank's slice that can no longer change -- a margin further north, and not the strip next to the northern neighbour,
        which keeps receiving that neighbour's deposits until fence 2 -- is copied to the shared host map on a second copy
        stream while the next chunk computes.  The margin assumption (no pixel is moved by PIPELINE_MARGIN_RAD or more) is
        verified from max|offset| over ALL ranks; if it fails every rank downloads its whole slice again.
        """
        torch = _torch()
        import torch.distributed as dist
        from .parallel import first_pixel_at_colatitude, ring_of_pixel, _ring_z
        orig_map = self.LightconeShell.map
        NSIDE = self.LightconeShell.NSIDE
        npix = orig_map.size
        lo, hi = self._range(npix)
        nloc = hi - lo
        dev = self._device()
        L = _lib.lib()
        cat = self.HaloLightConeCatalog.cat
        n = cat.size
        K = max(1, int(os.environ.get("BFG_SHARD_CHUNKS", self.SHARD_CHUNKS)))
        keys = list(vars(self.model).get('p_keys', []))
        _check_keys(self.model, keys)
        pix_rad = np.sqrt(4 * np.pi / npix)

        def colat_of_pixel(p):
            if p <= 0:
                return 0.0
            if p >= npix:
                return np.pi
            return float(np.arccos(np.clip(_ring_z(NSIDE, ring_of_pixel(NSIDE, np.array([p]))[0]), -1, 1)))

        with torch.cuda.device(dev):
            table = self._tables.get((_Ident(self.model), _Ident(self.model.interp_d) if hasattr(self.model, 'interp_d') else 0),
                                     lambda: displacement_table_of(self.model, dev.index))
            if getattr(self, '_scratch_inflight', None):
                torch.cuda.current_stream().synchronize()
                _give_scratch(self._scratch_inflight)
            self._scratch_inflight = []
            main = torch.cuda.current_stream()
            up, down = _side_stream(dev, 0), _side_stream(dev, 1)
            st = _lib.current_stream()
            own = peers.own_tensor()
            up.wait_stream(main)                     # the previous call's reads of `own` precede the zeroing
            h_map = torch.from_numpy(np.ascontiguousarray(orig_map, dtype=np.float64).reshape(-1))
            piece = -(-nloc // K)
            ev_h2d = []
            with torch.cuda.stream(up):
                own.zero_()
                token = torch.zeros(1, device=dev)
                dist.all_reduce(token)               # fence 1: every rank's slice is zero before any deposit
            d_off = torch.zeros((3, nloc), dtype=torch.float64, device=dev)    # cleared while the host stages the catalogue
            t0 = time.perf_counter()
            d_rec = self.device_records(False, dev)  # the (small) catalogue copies go first: the copy engine serves its queue in
            ext = _extras(cat, keys)                 # order, and behind the map upload they would hold the halo loop back
            d_ext = None if ext is None else _to_device(ext, dev)
            d_rec, d_ext = _sort_records(d_rec, d_ext, 0, SKY_BAND_RAD, owned=(NSIDE, lo, hi))
            with torch.cuda.stream(up):
                d_map = torch.empty(nloc, dtype=torch.float64, device=dev)
                for j in range(K):
                    a0, a1 = j * piece, min((j + 1) * piece, nloc)
                    if a1 > a0:
                        d_map[a0:a1].copy_(h_map[lo + a0:lo + a1], non_blocking=True)
                    ev_h2d.append(up.record_event())
            # chunk k = owned halos whose colatitude band lies in [edges[k], edges[k+1]); the last edge (2^20) = "all owned"
            th_cut = [colat_of_pixel(lo + int(nloc * f)) for f in _chunk_fractions(K)[1:]]
            edges = [0] + [int(np.floor(t / SKY_BAND_RAD)) for t in th_cut] + [1 << 20]
            d_edges = torch.tensor(edges, dtype=torch.int64, device=dev)
            d_bounds = torch.empty(K + 1, dtype=torch.int64, device=dev)
            d_rho = torch.zeros(1, dtype=torch.float64, device=dev)
            _lib.check(L.bfg_halo_band_bounds(n, _lib.ptr(d_rec), SKY_BAND_RAD, K + 1, _lib.ptr(d_edges), _lib.ptr(d_bounds),
                                              _lib.ptr(d_rho), st))
            d_nb = torch.zeros(K, dtype=torch.int64, device=dev)
            d_rem = torch.zeros(1, dtype=torch.int64, device=dev)
            d_max = torch.zeros(1, dtype=torch.float64, device=dev)
            d_acc = torch.zeros(5, dtype=torch.float64, device=dev)     # sum(new), sum(old), n_updates, n_remote, margin violated
            bounds = d_bounds.cpu().tolist()         # the one early synchronisation: chunk boundaries in the sorted catalogue
            rho_max = float(d_rho.cpu()[0])
            host_prep_s = time.perf_counter() - t0
            bounds[0] = 0
            n_extra = table.n_extra
            d_map.record_stream(main)
            # the strip next to the northern neighbour keeps receiving its deposits until fence 2
            q_top = lo if peers.rank == 0 else min(hi, max(lo, first_pixel_at_colatitude(
                NSIDE, colat_of_pixel(lo) + self.PIPELINE_MARGIN_RAD + 3 * pix_rad)))
            p_prev, q_prev, pieces_waited = lo, q_top, 0

            def regrid_to(p_k):
                nonlocal p_prev, pieces_waited
                if p_k <= p_prev:
                    return False
                need = min(K, -(-(p_k - lo) // piece))     # map pieces covering [lo, p_k); the first one also orders fence 1
                while pieces_waited < need:
                    main.wait_event(ev_h2d[pieces_waited])
                    pieces_waited += 1
                _lib.check(L.bfg_shell_regrid_p2p_range(NSIDE, _lib.ptr(d_map), _lib.ptr(d_off), lo, hi, p_prev, p_k,
                                                        peers.world, peers.rank, peers.h_bounds, peers.h_slices,
                                                        _lib.ptr(d_rem), st))
                p_prev = p_k
                return True

            def download(a, b, stream):
                if b > a:
                    _lib.check(L.bfg_copy_to_host_async(addr + 8 * a, own.data_ptr() + 8 * (a - lo), 8 * (b - a),
                                                        stream.cuda_stream))

            for k in range(K):
                b0, b1 = bounds[k], bounds[k + 1]
                if b1 > b0:
                    _lib.check(L.bfg_shell_offsets(table.handle, NSIDE, b1 - b0, d_rec.data_ptr() + 8 * _lib.HALO_STRIDE * b0,
                                                   None if d_ext is None else d_ext.data_ptr() + 8 * n_extra * b0, n_extra,
                                                   _lib.ptr(d_off), lo, hi, d_nb.data_ptr() + 8 * k, st))
                if k + 1 == K:
                    break
                # every owned halo not yet processed has theta >= edge band * band width; its disc reaches rho_max further north
                th_done = edges[k + 1] * SKY_BAND_RAD - rho_max - 3 * pix_rad
                p_k = min(hi, max(lo, first_pixel_at_colatitude(NSIDE, th_done))) if th_done > 0 else lo
                if regrid_to(p_k):
                    th_copy = th_done - self.PIPELINE_MARGIN_RAD - 3 * pix_rad
                    q_k = min(hi, max(lo, first_pixel_at_colatitude(NSIDE, th_copy))) if th_copy > 0 else lo
                    if q_k > q_prev:
                        down.wait_event(main.record_event())
                        download(q_prev, q_k, down)
                        q_prev = q_k
            regrid_to(hi)
            if pieces_waited == 0:                   # nothing re-binned yet (empty range): still order fence 1 before fence 2
                main.wait_event(ev_h2d[0])
            dist.all_reduce(token)                   # fence 2: every rank's deposits have landed
            down.wait_event(main.record_event())
            download(lo, q_top, down)                # the northern strip and everything not yet copied
            download(max(q_prev, q_top), hi, down)
            _lib.check(L.bfg_sum_f64(own.data_ptr(), nloc, _lib.ptr(d_acc), st))
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_map), nloc, d_acc.data_ptr() + 8, st))
            _lib.check(L.bfg_offsets_max_norm2(_lib.ptr(d_off), nloc, 0, nloc, _lib.ptr(d_max), st))
            d_acc[2] = d_nb.sum().to(torch.float64)                     # counts are < 2^53: exact as float64
            d_acc[3] = d_rem.reshape(()).to(torch.float64)
            d_acc[4] = (~(d_max.reshape(()) < self.PIPELINE_MARGIN_RAD ** 2)).to(torch.float64)
            main.wait_stream(down)
            dist.all_reduce(d_acc)                   # sums + "every rank's slice is in the host map" + "any margin violated?"
            acc = d_acc.cpu()
            if float(acc[4]) > 0:
                # somewhere the re-binning moved mass further than assumed: parts were copied too early -- on every rank again
                download(lo, hi, main)
                dist.all_reduce(token)
                token.cpu()
            main.synchronize()
        _give_scratch(getattr(self, '_scratch_inflight', []))
        self._scratch_inflight = []
        new_sum, old_sum = float(acc[0]), float(acc[1])
        self.last_stats = dict(n_updates=int(acc[2]), new_sum=new_sum, old_sum=old_sum, sharded=True, pipelined=True,
                               margin_violated=bool(float(acc[4]) > 0))
        self.last_stats_remote = int(acc[3])
        self.last_timing = dict(host_prep_s=host_prep_s, chunks=float(K))
        assert np.isclose(new_sum, old_sum), \
            "ERROR in pixel regridding, sum(new_map) [%0.14e] != sum(oldmap) [%0.14e]" % (new_sum, old_sum)   # :368-370
        return host.export(seg, orig_map.shape)

    def process(self):
        torch = _torch()
        orig_map = self.LightconeShell.map
        NSIDE = self.LightconeShell.NSIDE
        if _all_close_to_zero(orig_map):             # :293-294 returns the input object
            return orig_map
        zlo, zhi, Mlo, Mhi = _cat_ranges(self.HaloLightConeCatalog)      # the messages only need the catalogue's extremes
        _warn_table_range(self.model, np.array([zlo, zhi]), np.array([Mlo, Mhi]), self, self.HaloLightConeCatalog.cat)
        if (self.pix_range is None and self.sort_halos and os.environ.get("BFG_PIPELINE", "1") == "1"
                and os.environ.get("BFG_PROFILE_E2E") != "1"
                and self.HaloLightConeCatalog.cat.size >= self.PIPELINE_MIN_HALOS):
            return self._process_pipelined()
        dev = self._device()
        L = _lib.lib()
        npix = orig_map.size
        lo, hi = self._range(npix)
        if self.pix_range is not None:
            peers = self._peer_slices(npix, dev)
            if peers is not None:
                return self._process_sharded(peers)
        prof = os.environ.get("BFG_PROFILE_E2E") == "1"
        t_start = time.perf_counter()
        with torch.cuda.device(dev):
            # the halo loop is enqueued first; the map is only needed by the re-binning, so its H2D copy runs on a side
            # stream underneath the halo loop (a pageable source blocks the host, not the GPU)
            d_off, d_n = self.offsets_on_device()
            if prof:
                torch.cuda.synchronize(); t_loop = time.perf_counter()
            side = _side_stream(dev)
            with torch.cuda.stream(side):
                d_map = _to_device(orig_map[lo:hi], dev, dtype=np.float64)
            torch.cuda.current_stream().wait_stream(side)
            d_map.record_stream(torch.cuda.current_stream())
            if prof:
                torch.cuda.synchronize(); t_h2d = time.perf_counter()
            st = _lib.current_stream()
            d_map_sum = None
            # (the NVLink peer-memory exchange is _process_sharded; here: single GPU, or full-size partial maps + all-reduce)
            d_new = torch.zeros(npix, dtype=torch.float64, device=dev)
            _lib.check(L.bfg_shell_regrid(NSIDE, _lib.ptr(d_map), _lib.ptr(d_off), _lib.ptr(d_new), lo, hi, st))
            del d_off
            if self.pix_range is not None:
                from .parallel import reduce_partial_map
                d_new, d_map_sum = reduce_partial_map(d_new, d_map)
            d_sums = torch.zeros(2, dtype=torch.float64, device=dev)
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_new), npix, _lib.ptr(d_sums), st))
            if d_map_sum is None:
                _lib.check(L.bfg_sum_f64(_lib.ptr(d_map), hi - lo, d_sums.data_ptr() + 8, st))
            else:
                d_sums[1] = d_map_sum
            if prof:
                torch.cuda.synchronize(); t_regrid = time.perf_counter()
            out, out_np = _pinned_result(npix)
            if prof:
                t_alloc = time.perf_counter()
            out.copy_(d_new, non_blocking=True)
            sums = d_sums.cpu()
            n_up = int(d_n.cpu()[0])
            torch.cuda.current_stream().synchronize()
        _give_scratch(getattr(self, '_scratch_inflight', []))
        self._scratch_inflight = []
        new_sum, old_sum = float(sums[0]), float(sums[1])
        self.last_stats = dict(n_updates=n_up, new_sum=new_sum, old_sum=old_sum)
        if prof:
            t_end = time.perf_counter()
            self.last_timing.update(halo_loop_total_s=t_loop - t_start, h2d_map_s=t_h2d - t_loop,
                                    regrid_reduce_s=t_regrid - t_h2d, pinned_alloc_s=t_alloc - t_regrid,
                                    d2h_s=t_end - t_alloc, total_s=t_end - t_start)
        assert np.isclose(new_sum, old_sum), \
            "ERROR in pixel regridding, sum(new_map) [%0.14e] != sum(oldmap) [%0.14e]" % (new_sum, old_sum)   # :368-370
        return out_np


class PaintProfilesShell(DefaultRunner):
    """BaryonForge/Runners/HealpixRunner.py:376-483."""

    def paint_on_device(self):
        """The halo loop only: returns (painted map of the owned range on the device, update-count tensor)."""
        torch = _torch()
        assert self.model is not None, "You must provide a model"
        dev = self._device()
        L = _lib.lib()
        NSIDE = self.LightconeShell.NSIDE
        npix = self.LightconeShell.map.size
        lo, hi = self._range(npix)
        with torch.cuda.device(dev):
            table = self._tables.get((_Ident(self.model), _Ident(getattr(self.model, 'interp2D', None))),
                                     lambda: profile_table_of(self.model, '2D', dev.index))
        with torch.cuda.device(dev):
            d_new = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
            d_nb = torch.zeros(4, dtype=torch.int64, device=dev)

        def launch(d_rec, d_ext, n, k):
            _lib.check(L.bfg_shell_paint(table.handle, NSIDE, n, _lib.ptr(d_rec), _lib.ptr(d_ext), table.n_extra,
                                         _lib.ptr(d_new), lo, hi, d_nb.data_ptr() + 8 * k, _lib.current_stream()))
        self._halo_loop(True, table, launch, NSIDE, lo, hi, dev)
        with torch.cuda.device(dev):
            return d_new, d_nb.sum().reshape(1)

    def process(self):
        torch = _torch()
        assert self.model is not None, "You must provide a model"
        dev = self._device()
        L = _lib.lib()
        NSIDE = self.LightconeShell.NSIDE
        npix = self.LightconeShell.map.size
        lo, hi = self._range(npix)
        with torch.cuda.device(dev):
            table = self._tables.get((_Ident(self.model), _Ident(getattr(self.model, 'interp2D', None))),
                                     lambda: profile_table_of(self.model, '2D', dev.index))
        with torch.cuda.device(dev):
            d_new = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
            d_nb = torch.zeros(4, dtype=torch.int64, device=dev)

        def launch(d_rec, d_ext, n, k):
            _lib.check(L.bfg_shell_paint(table.handle, NSIDE, n, _lib.ptr(d_rec), _lib.ptr(d_ext), table.n_extra,
                                         _lib.ptr(d_new), lo, hi, d_nb.data_ptr() + 8 * k, _lib.current_stream()))
        self._halo_loop(True, table, launch, NSIDE, lo, hi, dev)
        with torch.cuda.device(dev):
            d_n = d_nb.sum().reshape(1)
            if self.pix_range is not None:
                from .parallel import gather_owned_ranges
                d_new = gather_owned_ranges(d_new, npix)
            out, out_np = _pinned_result(npix)
            out.copy_(d_new, non_blocking=True)
            n_up = int(d_n.cpu()[0])
            torch.cuda.current_stream().synchronize()
        _give_scratch(getattr(self, '_scratch_inflight', []))
        self._scratch_inflight = []
        self.last_stats = dict(n_updates=n_up)
        return out_np


class PaintProfilesAnisShell(DefaultRunner):
    """
    BaryonForge/Runners/HealpixRunner.py:484-640: paints `model` weighted by the share of each pixel's total mass that
    the halo's `Tracer_model` profile accounts for, times the input map; mass not in halos is a uniform background.
    Three tables are needed on the device: model.interp2D, Tracer_model.interp2D and Mtot_model.interp2D.
    """

    def __init__(self, HaloLightConeCatalog, LightConeShell, epsilon_max, model, Tracer_model, Mtot_model,
                 background_val, global_tracer_fraction, mass_def=None, include_pixel_size=False, use_ellipticity=False,
                 verbose=True, *, device=None, pix_range=None, sort_halos=True):
        self.Tracer_model = Tracer_model
        self.Mtot_model = Mtot_model
        self.background_val = background_val
        self.global_tracer_fraction = global_tracer_fraction
        super().__init__(HaloLightConeCatalog, LightConeShell, epsilon_max, model, use_ellipticity, mass_def,
                         include_pixel_size, verbose, device=device, pix_range=pix_range, sort_halos=sort_halos)
        self._tables2 = _TableCache()

    def __setstate__(self, d):
        super().__setstate__(d)
        self._tables2 = _TableCache()

    def __getstate__(self):
        d = super().__getstate__()
        d['_tables2'] = None
        return d

    def process(self):
        torch = _torch()
        dev = self._device()
        L = _lib.lib()
        cosmo = cosmology.runner_cosmology(self.cosmo, with_w0=True)               # :535-540
        orig_map = self.LightconeShell.map
        NSIDE = self.LightconeShell.NSIDE
        npix = orig_map.size
        pixarea = 4 * np.pi / npix                                                 # :545
        lo, hi = self._range(npix)
        cat = self.HaloLightConeCatalog.cat
        keys = list(vars(self.model).get('p_keys', []))                            # :552
        _check_keys(self.model, keys)
        # total-mass map of the halos (:565-570): a PaintProfilesShell pass of Mtot_model, kept on the device
        mrun = PaintProfilesShell(self.HaloLightConeCatalog, self.LightconeShell, self.epsilon_max, self.Mtot_model,
                                  self.use_ellipticity, self.mass_def, True, self.verbose, device=self.device,
                                  pix_range=self.pix_range, sort_halos=self.sort_halos)
        d_mtot, _ = mrun.paint_on_device()
        dL = 2 * get_parameter(self.Mtot_model, 'proj_cutoff')                     # :573
        z_m = float(np.max(cat['z'])) if cat.size else 0.0
        dD = float(cosmology.D_A_spline_to(cosmo, z_m)(self.LightconeShell.redshift))   # :546-549, :574
        dV = pixarea * ((dD + dL) ** 3 - dD ** 3)                                  # :575
        with torch.cuda.device(dev):
            d_s = torch.zeros(1, dtype=torch.float64, device=dev)
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_mtot), hi - lo, _lib.ptr(d_s), _lib.current_stream()))
            if self.pix_range is not None:
                from .parallel import all_reduce_sum
                all_reduce_sum(d_s)
            mtot_sum = float(d_s.cpu()[0])
        rho_halos = mtot_sum / (dV * npix)                                         # :576
        rho_m = float(cosmology.rho_matter(cosmo, 1 / (self.LightconeShell.redshift + 1), is_comoving=False))   # :580
        drho_m = float(np.clip(rho_m - rho_halos, 0, None))                        # :581
        mtot_add = dV * drho_m                                                     # :582
        if self.verbose:
            print(f"Inputted halos contribute {100*(rho_halos/rho_m):0.2f}% of the total matter density.")
            print(f"Remaining density is assigned to a uniform background.")
        if rho_halos > rho_m:
            warnings.warn("Inputted halos contribute more mass than is available for this mean matter density."
                          "Your Mtot_model profiles are either too extended or you are using the wrong cosmology.")
        with torch.cuda.device(dev):
            t_paint = self._tables.get((_Ident(self.model), _Ident(getattr(self.model, 'interp2D', None))),
                                       lambda: profile_table_of(self.model, '2D', dev.index))
            t_tracer = self._tables2.get((_Ident(self.Tracer_model), _Ident(getattr(self.Tracer_model, 'interp2D', None))),
                                         lambda: profile_table_of(self.Tracer_model, '2D', dev.index))
            d_orig = _to_device(orig_map[lo:hi], dev, dtype=np.float64)
            d_new = torch.zeros(hi - lo, dtype=torch.float64, device=dev)
            d_nb = torch.zeros(4, dtype=torch.int64, device=dev)

        def launch(d_rec, d_ext, n, k):
            _lib.check(L.bfg_shell_paint_anis(t_paint.handle, t_tracer.handle, NSIDE, n, _lib.ptr(d_rec), _lib.ptr(d_ext),
                                              t_paint.n_extra, _lib.ptr(d_mtot), float(mtot_add), _lib.ptr(d_orig),
                                              _lib.ptr(d_new), lo, hi, d_nb.data_ptr() + 8 * k, _lib.current_stream()))
        self._halo_loop(True, t_paint, launch, NSIDE, lo, hi, dev)
        with torch.cuda.device(dev):
            # uniform-background term (:633-636)
            _lib.check(L.bfg_anis_background(hi - lo, _lib.ptr(d_mtot), float(mtot_add), _lib.ptr(d_orig),
                                             float(self.background_val * self.global_tracer_fraction), 1.0,
                                             _lib.ptr(d_new), _lib.current_stream()))
            d_n = d_nb.sum().reshape(1)
            if self.pix_range is not None:
                from .parallel import gather_owned_ranges
                d_new = gather_owned_ranges(d_new, npix)
            out, out_np = _pinned_result(npix)
            out.copy_(d_new, non_blocking=True)
            n_up = int(d_n.cpu()[0])
            torch.cuda.current_stream().synchronize()
        _give_scratch(getattr(self, '_scratch_inflight', []))
        self._scratch_inflight = []
        self.last_stats = dict(n_updates=n_up, rho_halos=rho_halos, rho_m=rho_m, dV=dV, dD=dD)
        return out_np.reshape(orig_map.shape)


# =====================================================================================================================
# periodic grids
# =====================================================================================================================
def _nearest_bin(bins, x):
    """np.argmin(np.abs(bins - x_j)) for every halo (Map2DRunner.py:512-513), without the n_halo x N matrix."""
    N = bins.size
    res = bins[1] - bins[0]
    c = np.clip(np.floor((x - bins[0]) / res + 0.5), 0, N - 1).astype(np.int64)
    best = np.clip(c - 1, 0, N - 1)
    dbest = np.abs(bins[best] - x)
    for o in (0, 1):                      # ascending index order + strict '<' == argmin's first-minimum rule
        cand = np.clip(c + o, 0, N - 1)
        d = np.abs(bins[cand] - x)
        take = d < dbest
        best = np.where(take, cand, best)
        dbest = np.where(take, d, dbest)
    return best


def _box_device_records(runner, cat, redshift, ndim, paint, grid, rq_clip, dev, bins=None, res=1.0):
    """
    Box (grid / snapshot) halo records built on the device (bfg_box_records): the host stages the float32-rounded
    M, x, y, z and numpy's float32 ln M; returns (records [n, 16] on the device, aux [2, n]: R_phys, R_model_com).
    """
    torch = _torch()
    n = cat.size
    d_rec = torch.empty((n, _lib.HALO_STRIDE), dtype=torch.float64, device=dev)
    d_aux = torch.empty((2, n), dtype=torch.float64, device=dev)
    if n == 0:
        return d_rec, d_aux
    cosmo = cosmology.runner_cosmology(runner.cosmo, with_w0=False)               # Map2DRunner.py:462-465 (no w0)
    a = 1 / (1 + redshift)
    one = np.ones(1)
    g_run = float(cosmology.radius_of_mass(cosmo, one, a, runner.mass_def)[0])
    g_mod = 0.0
    if not paint:
        g_mod = float(cosmology.radius_of_mass(_model_cosmo(runner.model, cosmo), one, a,
                                               getattr(runner.model, 'mass_def', None))[0])
    stage = _take_scratch(5 * n)
    cols = stage.numpy().reshape(5, n)
    names = ['x', 'y', 'z'][:ndim]

    def fill(sl):
        M32 = cat['M'][sl].astype('<f4')                                          # io.py:204-205
        cols[0, sl] = M32
        cols[4, sl] = np.log(M32)                                                 # float32 log, SURVEY §10 #8
        for k, name in enumerate(names):
            cols[1 + k, sl] = cat[name][sl].astype('<f4')
        for k in range(ndim, 3):
            cols[1 + k, sl] = 0.0
    _parallel_chunks(fill, n)
    d_cols = stage.to(dev, non_blocking=True)
    d_bins = None if bins is None else _to_device(bins, dev, dtype=np.float64)
    _lib.check(_lib.lib().bfg_box_records(
        n, d_cols.data_ptr(), ndim, 1 if grid else 0, 1 if paint else 0, float(a), float(np.log(1 / a)), g_run, g_mod,
        float(runner.epsilon_max), 0.0 if paint else float(runner.model.epsilon_max), float(res), float(rq_clip),
        0 if bins is None else int(bins.size), _lib.ptr(d_bins), d_rec.data_ptr(), d_aux.data_ptr(), _lib.current_stream()))
    torch.cuda.current_stream().synchronize()     # the pinned staging buffer goes straight back to the pool
    _give_scratch([stage])
    return d_rec, d_aux


class DefaultRunnerGrid(object):
    """Constructor contract of BaryonForge/Runners/Map2DRunner.py:255-278 (+ keyword-only GPU knobs)."""

    def __init__(self, HaloNDCatalog, GriddedMap, epsilon_max, model, use_ellipticity=False,
                 mass_def=None, include_pixel_size=True, verbose=True, *, device=None, plane_range=None):
        self.HaloNDCatalog = HaloNDCatalog
        self.GriddedMap = GriddedMap
        self.cosmo = HaloNDCatalog.cosmology
        self.model = model
        self.epsilon_max = epsilon_max
        self.mass_def = mass_def
        self.verbose = verbose
        self.use_ellipticity = use_ellipticity
        self.include_pixel_size = include_pixel_size
        self.device = device
        self.plane_range = plane_range    # (lo, hi) axis-0 planes owned by this rank; None = whole grid
        self.last_stats = {}
        self._tables = _TableCache()
        if use_ellipticity:               # Map2DRunner.py:272-278
            names = HaloNDCatalog.cat.dtype.names
            assert 'q_ell' in names, "The 'q_ell' column is missing, but you set use_ellipticity = True"
            if not GriddedMap.is2D:
                assert 'c_ell' in names, "The 'c_ell' column is missing, but you set use_ellipticity = True"
            assert 'A_ell' in names, "The 'A_ell' column is missing, but you set use_ellipticity = True"

    __getstate__ = DefaultRunner.__getstate__
    __setstate__ = DefaultRunner.__setstate__
    _device = DefaultRunner._device

    def _planes(self, N):
        return (0, N) if self.plane_range is None else (int(self.plane_range[0]), int(self.plane_range[1]))

    def _owned_halos(self, rec, extras, N, lo, hi):
        """Slab sharding: keep the halos whose cutout touches this rank's axis-0 planes (parallel.py)."""
        if self.plane_range is None or rec.shape[0] == 0:
            return rec, extras
        from .parallel import halos_touching_planes
        keep = halos_touching_planes(N, rec[:, _lib.HB_CX], rec[:, _lib.HB_NSIZE], lo, hi)
        if keep.all():
            return rec, extras
        return np.ascontiguousarray(rec[keep]), (None if extras is None else np.ascontiguousarray(extras[keep]))

    def _records_on_device(self, paint, dev):
        """
        (records, extras, n) on the device, box-cell ordered, built ON THE DEVICE (bfg_box_records).  Slab-sharded runs hand
        every rank the whole catalogue as well: the tile binning / the scatter kernels only take the part of a cutout that
        lies in the owned planes, so a halo that does not reach the slab costs one record read (the host-side filter of
        BFG_DEVICE_RECORDS=0 took ~0.15 s per call for 10^6 halos).
        """
        gm = self.GriddedMap
        ndim, N = (2 if gm.is2D else 3), gm.Npix
        lo, hi = self._planes(N)
        if os.environ.get("BFG_DEVICE_RECORDS", "1") != "1":
            rec, extras = self.halo_records(paint)
            rec, extras = self._owned_halos(rec, extras, N, lo, hi)
            d_rec = _upload_records(rec, dev)
        else:
            if self.use_ellipticity and not gm.is2D:
                if paint:
                    raise ValueError("use_ellipticity is not implemented for 3D maps")               # Map2DRunner.py:801
                raise NotImplementedError("Currently not able to ellipticities with 3D maps.")       # Map2DRunner.py:571
            cat = self.HaloNDCatalog.cat
            bins = np.asarray(gm.bins, dtype=np.float64)
            d_rec, d_aux = _box_device_records(self, cat, self.HaloNDCatalog.redshift, ndim, paint, True,
                                               np.max(bins) / 2, dev, bins=bins, res=gm.res)
            aux = d_aux.cpu().numpy()
            self.last_scalars = dict(R_phys=aux[0], R_model_com=None if paint else aux[1])
            if ndim == 2 and cat.size:
                dxy = d_rec[:, _lib.HB_DX:_lib.HB_DX + 2]
                assert bool((dxy <= gm.res).all().item()), "Halo offsets are larger than res"        # :522
            keys = list(vars(self.model).get('p_keys', []))
            _check_keys(self.model, keys)
            extras = _extras(cat, keys)
            if self.use_ellipticity:
                Rmat = self.shear_matrices()
                extras = Rmat if extras is None else np.ascontiguousarray(np.hstack([extras, Rmat]))
        d_ext = None if extras is None else _to_device(extras, dev)
        n = d_rec.shape[0]
        d_rec, d_ext = _sort_records(d_rec, d_ext, 1, float(gm.L), 16, ndim)
        return d_rec, d_ext, n

    def halo_records(self, paint):
        """Per-halo scalars of Map2DRunner.py:484-520 / :727-760 (+ BaryonCorrection.py:371,398-399,410), vectorised."""
        if self.use_ellipticity and not self.GriddedMap.is2D:
            if paint:
                raise ValueError("use_ellipticity is not implemented for 3D maps")               # Map2DRunner.py:801
            raise NotImplementedError("Currently not able to ellipticities with 3D maps.")       # Map2DRunner.py:571
        cat = self.HaloNDCatalog.cat
        n = cat.size
        bins = np.asarray(self.GriddedMap.bins, dtype=np.float64)
        res = self.GriddedMap.res
        ndim = 2 if self.GriddedMap.is2D else 3
        cosmo = cosmology.runner_cosmology(self.cosmo, with_w0=False)             # :462-465 (no w0)
        a = 1 / (1 + self.HaloNDCatalog.redshift)                                 # :490
        rec = np.zeros((_lib.HALO_STRIDE, n), dtype=np.float64).T
        if n == 0:
            return rec, None
        R_phys_all = np.empty(n)
        R_mod_all = None if paint else np.empty(n)
        mcosmo = None if paint else _model_cosmo(self.model, cosmo)
        bmax = np.max(bins)
        eps_model = None if paint else self.model.epsilon_max
        mdef_model = None if paint else getattr(self.model, 'mass_def', None)
        names = ['x', 'y', 'z'][:ndim]

        def fill(sl):   # chunks run on the host thread pool (numpy releases the GIL in these kernels)
            r = rec[sl]
            M32 = cat['M'][sl].astype('<f4')                                      # io.py:204-205
            M = M32.astype(np.float64)
            R_phys = cosmology.radius_of_mass(cosmo, M, a, self.mass_def)         # :491
            R_phys_all[sl] = R_phys
            if paint:
                R_com = R_phys / a                                                # :734
                Nf = 2 * self.epsilon_max * R_com / res                           # :740
                r[:, _lib.HB_PAINTCUT] = R_com * self.epsilon_max                 # :815
                r[:, _lib.HB_RQ] = R_com * self.epsilon_max
                r[:, _lib.HB_RCUT] = np.inf
            else:
                R_q = np.clip(self.epsilon_max * R_phys / a, 0, bmax / 2)         # :492-493
                Nf = 2 * R_q / res                                                # :500
                r[:, _lib.HB_RQ] = R_q
                R_mod = cosmology.radius_of_mass(mcosmo, M, a, mdef_model) / a    # BaryonCorrection.py:399
                R_mod_all[sl] = R_mod
                r[:, _lib.HB_RCUT] = eps_model * R_mod
                r[:, _lib.HB_LNRCOM] = np.log(R_mod)
            Nsize = ((Nf // 2).astype(np.int64)) * 2                              # :501
            Nsize = np.clip(Nsize, 2, bins.size // 2)                             # :503
            r[:, _lib.HB_NSIZE] = Nsize
            r[:, _lib.HB_LNZ] = np.log(1 / a)
            r[:, _lib.HB_LNM] = np.log(M32).astype(np.float64)                    # float32 log, §10 #8
            for k, name in enumerate(names):
                x = cat[name][sl].astype('<f4').astype(np.float64)
                cen = _nearest_bin(bins, x)
                r[:, _lib.HB_X + k] = x
                r[:, _lib.HB_CX + k] = cen
                r[:, _lib.HB_DX + k] = bins[cen] - x                              # :519-520
        _parallel_chunks(fill, n, chunk=32768)
        self.last_scalars = dict(R_phys=R_phys_all, R_model_com=R_mod_all)
        if ndim == 2:
            dx, dy = rec[:, _lib.HB_DX], rec[:, _lib.HB_DY]
            assert np.all((dx <= res) & (dy <= res)), "Halo offsets are larger than res"   # :522
        keys = list(vars(self.model).get('p_keys', []))
        _check_keys(self.model, keys)
        extras = _extras(cat, keys)
        if self.use_ellipticity:
            Rmat = self.shear_matrices()
            extras = Rmat if extras is None else np.ascontiguousarray(np.hstack([extras, Rmat]))
        return rec, extras

    # ---- small public helpers of the reference's DefaultRunnerGrid (host-side numpy; the device path uses
    #      shear_matrices() and the cutout index math of csrc/grid_kernels.cu instead)
    def build_Rmat(self, A, q):
        """Shear matrix of one halo from its orientation A (normalised IN PLACE) and axis ratio q (Map2DRunner.py:281-350)."""
        A /= np.linalg.norm(A)
        if len(A) == 1:
            raise ValueError("Can't rotate a 1-dimensional vector")
        if len(A) == 3:
            raise NotImplementedError("This method has not yet been verified. Use 2D ellipticity method instead")
        beta = np.arccos(np.dot(A, np.array([1., 0.])))
        eta = -np.log(q)
        if eta > 1e-4:
            eta2g = np.tanh(0.5 * eta) / eta
        else:
            etasq = eta * eta
            eta2g = 0.5 + etasq * ((-1 / 24) + etasq * (1 / 240))
        g = eta2g * eta * np.exp(2j * beta)
        return np.array([[1 + g.real, g.imag], [g.imag, 1 - g.real]]) / np.sqrt(1 - np.abs(g) ** 2)

    def coord_array(self, *args):
        """(N, M) coordinate rows from M same-shaped arrays (Map2DRunner.py:353-375)."""
        return np.stack([np.ravel(a) for a in args], axis=1)

    def pick_indices(self, center, width, Npix):
        """Periodic cutout indices center - width .. center + width - 1 (Map2DRunner.py:400-429; one wrap, like the reference)."""
        inds = np.arange(center - width, center + width)
        inds = np.where(inds < 0, inds + Npix, inds)
        return np.where(inds >= Npix, inds - Npix, inds)

    def shear_matrices(self):
        """
        build_Rmat(A_ell, q_ell) of every halo (Map2DRunner.py:281-350,495-498,533), row-major [n, 4], with the reference's
        float32 arithmetic (HaloNDCatalog columns are float32): A normalised twice in float32, beta = arccos(A_x),
        eta = -log(q) and eta2g in float32, then g = eta2g*eta*exp(2i beta) in complex128.
        """
        cat = self.HaloNDCatalog.cat
        q = cat['q_ell'].astype('<f4')
        assert np.all(q > 0), "The axis ratio of a halo is not positive"                       # Map2DRunner.py:532
        A = cat['A_ell'].astype('<f4')
        A = A / np.sqrt(np.sum(A ** 2, axis=1))[:, None]                                       # :497
        A = A / np.sqrt(np.sum(A * A, axis=1))[:, None]                                        # A /= np.linalg.norm(A) :310
        beta = np.arccos(A[:, 0].astype(np.float64))                                           # arccos(dot(A, [1, 0]))
        eta = -np.log(q)
        with np.errstate(divide='ignore', invalid='ignore'):
            etasq = eta * eta
            eta2g = np.where(eta > 1e-4, np.tanh(np.float32(0.5) * eta) / eta,
                             np.float32(0.5) + etasq * (np.float32(-1 / 24) + etasq * np.float32(1 / 240)))
        g = (eta2g * eta).astype(np.float64) * np.exp(2j * beta)
        g1, g2 = g.real, g.imag
        det = np.sqrt(1 - np.abs(g) ** 2)
        return np.ascontiguousarray(np.stack([(1 + g1) / det, g2 / det, g2 / det, (1 - g1) / det], axis=1))


class BaryonifyGrid(DefaultRunnerGrid):
    """BaryonForge/Runners/Map2DRunner.py:353-621."""

    def offsets_on_device(self):
        torch = _torch()
        dev = self._device()
        L = _lib.lib()
        gm = self.GriddedMap
        ndim, N = (2 if gm.is2D else 3), gm.Npix
        lo, hi = self._planes(N)
        with torch.cuda.device(dev):
            table = self._tables.get((_Ident(self.model), _Ident(getattr(self.model, 'interp_d', None))),
                                     lambda: displacement_table_of(self.model, dev.index))
        with torch.cuda.device(dev):
            d_rec, d_ext, n_rec = self._records_on_device(False, dev)
            nloc = (hi - lo) * N ** (ndim - 1)
            d_off = torch.zeros((ndim, nloc), dtype=torch.float64, device=dev)
            d_n = torch.zeros(1, dtype=torch.int64, device=dev)
            n_cols = 0 if d_ext is None else d_ext.shape[1]
            _lib.check(L.bfg_grid_offsets(table.handle, ndim, N, float(gm.res), n_rec, _lib.ptr(d_rec),
                                          _lib.ptr(d_ext), n_cols, 1 if self.use_ellipticity else 0, _lib.ptr(d_off), lo, hi,
                                          _lib.ptr(d_n), _lib.current_stream()))
        return d_off, d_n

    def process(self):
        torch = _torch()
        gm = self.GriddedMap
        orig_map = gm.map
        ndim, N = (2 if gm.is2D else 3), gm.Npix
        lo, hi = self._planes(N)
        dev = self._device()
        L = _lib.lib()
        _warn_table_range(self.model, self.HaloNDCatalog.redshift, self.HaloNDCatalog.cat['M'], self, self.HaloNDCatalog.cat)
        with torch.cuda.device(dev):
            # halo loop first (asynchronous); the map is only needed by the re-binning, so its upload runs on a side stream
            # underneath the halo loop (a pageable source blocks the host, not the GPU)
            d_off, d_n = self.offsets_on_device()
            side = _side_stream(dev)
            with torch.cuda.stream(side):
                d_map = _to_device(orig_map[lo:hi], dev, dtype=np.float64)
            torch.cuda.current_stream().wait_stream(side)
            d_map.record_stream(torch.cuda.current_stream())
            d_new = torch.zeros(orig_map.size, dtype=torch.float64, device=dev)
            st = _lib.current_stream()
            _lib.check(L.bfg_grid_regrid(ndim, N, _lib.ptr(d_map), _lib.ptr(d_off), _lib.ptr(d_new), lo, hi, st))
            del d_off
            d_map_sum = None
            if self.plane_range is not None:
                shared = self._slab_exchange(N, ndim, lo, hi, dev)
                if shared is not None:
                    return self._finish_sharded(shared, d_new, d_map, d_n, orig_map, N, ndim, lo, hi, dev)
                from .parallel import reduce_partial_map
                d_new, d_map_sum = reduce_partial_map(d_new, d_map)
            d_sums = torch.zeros(2, dtype=torch.float64, device=dev)
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_new), orig_map.size, _lib.ptr(d_sums), st))
            if d_map_sum is None:
                _lib.check(L.bfg_sum_f64(_lib.ptr(d_map), d_map.numel(), d_sums.data_ptr() + 8, st))
            else:
                d_sums[1] = d_map_sum
            out, out_np = _pinned_result(orig_map.size)
            out.copy_(d_new, non_blocking=True)
            sums = d_sums.cpu()
            n_up = int(d_n.cpu()[0])
            torch.cuda.current_stream().synchronize()
        new_sum, old_sum = float(sums[0]), float(sums[1])
        self.last_stats = dict(n_updates=n_up, new_sum=new_sum, old_sum=old_sum)
        assert np.isclose(new_sum, old_sum), \
            "ERROR in pixel regridding, sum(new_map) [%0.14e] != sum(oldmap) [%0.14e]" % (new_sum, old_sum)   # :616-619
        return out_np.reshape(orig_map.shape)


    def _slab_exchange(self, N, ndim, lo, hi, dev):
        """
        The SharedHostMaps of the slab-sharded end-to-end path, or None for the plain all-reduce path.  Taken when the ranks
        are the NCCL processes of ONE machine and the slabs are parallel.plane_ranges' equal, rank-ordered ones (same test on
        every rank: it depends on the group and the grid size only, plus a vote on the plane ranges).
        """
        from . import parallel
        dist = parallel._dist()
        if dist is None or os.environ.get("BFG_EXCHANGE", "p2p") == "allreduce" or not hasattr(os, 'memfd_create'):
            return None
        world, rank = dist.get_world_size(), dist.get_rank()
        if world < 2 or world > 8 or dist.get_backend() != 'nccl' or N % world or not parallel.single_node_group():
            return None
        torch = _torch()
        mine = tuple(parallel.plane_ranges(N, world)[rank]) == (int(lo), int(hi))
        vote = torch.tensor([1 if mine else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN)
        if int(vote.cpu()[0]) == 0:
            return None
        numel = N ** ndim
        key = (numel, world, rank, dev.index)
        if key not in _HOST_MAPS:
            el = N ** (ndim - 1)
            _HOST_MAPS[key] = parallel.SharedHostMaps(numel, rank, world, dev.index, own_range=(lo * el, hi * el))
        return _HOST_MAPS[key]

    def _finish_sharded(self, host, d_new, d_map, d_n, orig_map, N, ndim, lo, hi, dev):
        """
        Slab-sharded tail of process(): the CIC deposit of a slab reaches into the neighbouring slabs, so the per-rank partial
        maps are summed -- by ONE NCCL reduce-scatter that leaves every rank with its own slab of the new map (half the NVLink
        traffic of the all-reduce) -- and each rank copies only its slab into a page-locked host map that all processes of the
        box have mapped (parallel.SharedHostMaps): 1 / N of the full-map download per rank instead of N full copies through
        the one host link.  The small all-reduce of [sum new, sum old, n_updates] is enqueued behind the copy, so its
        completion also says that every rank's slab has arrived.
        """
        torch = _torch()
        import torch.distributed as dist
        from .parallel import SegmentsExhausted, reduce_partial_map
        L = _lib.lib()
        st = _lib.current_stream()
        el = N ** (ndim - 1)
        seg = addr = None
        try:
            seg, addr = host.acquire()
        except SegmentsExhausted:                # raised on all ranks: the caller holds MAX_SEGMENTS earlier results
            pass
        except OSError:                          # collective failure (agreed by all ranks)
            _HOST_MAPS[(N ** ndim, dist.get_world_size(), dist.get_rank(), dev.index)] = None
        d_acc = torch.zeros(3, dtype=torch.float64, device=dev)
        if seg is None:                          # private full copy on every rank (the all-reduce path)
            d_new, d_map_sum = reduce_partial_map(d_new, d_map)
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_new), d_new.numel(), _lib.ptr(d_acc), st))
            d_acc[0] /= dist.get_world_size()    # every rank holds the full sum; the all-reduce below adds them up again
            d_acc[1] = d_map_sum / dist.get_world_size()
            out, out_np = _pinned_result(orig_map.size)
            out.copy_(d_new, non_blocking=True)
        else:
            own = torch.empty((hi - lo) * el, dtype=torch.float64, device=dev)
            dist.reduce_scatter_tensor(own, d_new, op=dist.ReduceOp.SUM)
            _lib.check(L.bfg_copy_to_host_async(addr + 8 * lo * el, own.data_ptr(), 8 * own.numel(), st))
            _lib.check(L.bfg_sum_f64(own.data_ptr(), own.numel(), _lib.ptr(d_acc), st))
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_map), d_map.numel(), d_acc.data_ptr() + 8, st))
        d_acc[2] = d_n.reshape(()).to(torch.float64)                 # counts are < 2^53: exact as float64
        dist.all_reduce(d_acc)
        acc = d_acc.cpu()
        torch.cuda.current_stream().synchronize()
        new_sum, old_sum = float(acc[0]), float(acc[1])
        self.last_stats = dict(n_updates=int(acc[2]), new_sum=new_sum, old_sum=old_sum, sharded=True)
        assert np.isclose(new_sum, old_sum), \
            "ERROR in pixel regridding, sum(new_map) [%0.14e] != sum(oldmap) [%0.14e]" % (new_sum, old_sum)   # :616-619
        return host.export(seg, orig_map.shape) if seg is not None else out_np.reshape(orig_map.shape)


class PaintProfilesGrid(DefaultRunnerGrid):
    """BaryonForge/Runners/Map2DRunner.py:624-829."""

    def _device_records(self, dev):
        """Halo records (+ extras) of the owned planes on the device, box-cell ordered."""
        return self._records_on_device(True, dev)

    def paint_on_device(self):
        """The halo loop only: returns (painted owned planes on the device, flat; update-count tensor)."""
        torch = _torch()
        gm = self.GriddedMap
        ndim, N = (2 if gm.is2D else 3), gm.Npix
        lo, hi = self._planes(N)
        dev = self._device()
        L = _lib.lib()
        which = '2D' if gm.is2D else '3D'                                         # :763 projected / :792 real
        dV = float(np.power(gm.res, ndim)) if self.include_pixel_size else 1.0    # :723,825
        with torch.cuda.device(dev):
            d_rec, d_ext, n = self._device_records(dev)
            table = self._tables.get((_Ident(self.model), which, _Ident(getattr(self.model, 'interp' + which, None))),
                                     lambda: profile_table_of(self.model, which, dev.index))
            nloc = (hi - lo) * N ** (ndim - 1)
            d_new = torch.zeros(nloc, dtype=torch.float64, device=dev)
            d_n = torch.zeros(1, dtype=torch.int64, device=dev)
            n_cols = 0 if d_ext is None else d_ext.shape[1]
            _lib.check(L.bfg_grid_paint(table.handle, ndim, N, float(gm.res), dV, n, _lib.ptr(d_rec),
                                        _lib.ptr(d_ext), n_cols, 1 if self.use_ellipticity else 0, _lib.ptr(d_new), lo, hi,
                                        _lib.ptr(d_n), _lib.current_stream()))
        return d_new, d_n

    def process(self):
        torch = _torch()
        gm = self.GriddedMap
        ndim, N = (2 if gm.is2D else 3), gm.Npix
        dev = self._device()
        d_new, d_n = self.paint_on_device()
        with torch.cuda.device(dev):
            if self.plane_range is not None:
                from .parallel import gather_owned_ranges
                d_new = gather_owned_ranges(d_new, gm.map.size)
            out, out_np = _pinned_result(gm.map.size)
            out.copy_(d_new, non_blocking=True)
            n_up = int(d_n.cpu()[0])
            torch.cuda.current_stream().synchronize()
        self.last_stats = dict(n_updates=n_up)
        return out_np.reshape(gm.map.shape)


class PaintProfilesAnisGrid(PaintProfilesGrid):
    """BaryonForge/Runners/Map2DRunner.py:833-1015 (2-D maps only, :847)."""

    def __init__(self, HaloNDCatalog, GriddedMap, epsilon_max, model, Tracer_model, Mtot_model, background_val,
                 global_tracer_fraction, mass_def=None, include_pixel_size=True, use_ellipticity=False, verbose=True, *,
                 device=None, plane_range=None):
        self.Tracer_model = Tracer_model
        self.Mtot_model = Mtot_model
        self.background_val = background_val
        self.global_tracer_fraction = global_tracer_fraction
        super().__init__(HaloNDCatalog, GriddedMap, epsilon_max, model, use_ellipticity, mass_def, include_pixel_size,
                         verbose, device=device, plane_range=plane_range)
        self._tables2 = _TableCache()

    def __setstate__(self, d):
        super().__setstate__(d)
        self._tables2 = _TableCache()

    def __getstate__(self):
        d = super().__getstate__()
        d['_tables2'] = None
        return d

    def process(self):
        gm = self.GriddedMap
        assert gm.is2D == True, "Can only paint tSZ on 2D maps. You have passed a 3D Map"   # noqa: E712  (:847)
        torch = _torch()
        dev = self._device()
        L = _lib.lib()
        cosmo = cosmology.runner_cosmology(self.cosmo, with_w0=False)             # :849-853
        orig_map = gm.map
        N, res = gm.Npix, gm.res
        lo, hi = self._planes(N)
        nloc = (hi - lo) * N
        # total-mass map of the halos (:866-871): a PaintProfilesGrid pass of Mtot_model without the pixel size
        mrun = PaintProfilesGrid(self.HaloNDCatalog, gm, self.epsilon_max, self.Mtot_model, self.use_ellipticity,
                                 self.mass_def, False, self.verbose, device=self.device, plane_range=self.plane_range)
        d_mtot, _ = mrun.paint_on_device()
        dL = 2 * get_parameter(self.Mtot_model, 'proj_cutoff')                    # :877
        dV = np.power(res, 2) * dL                                                # :878
        with torch.cuda.device(dev):
            d_s = torch.zeros(1, dtype=torch.float64, device=dev)
            _lib.check(L.bfg_sum_f64(_lib.ptr(d_mtot), nloc, _lib.ptr(d_s), _lib.current_stream()))
            if self.plane_range is not None:
                from .parallel import all_reduce_sum
                all_reduce_sum(d_s)
            mtot_sum = float(d_s.cpu()[0])
        rho_halos = (mtot_sum / orig_map.size) / dL                               # :879 np.average(Mtot_map) / dL
        rho_m = float(cosmology.rho_matter(cosmo, 1 / (self.HaloNDCatalog.redshift + 1), is_comoving=True))   # :886
        drho_m = float(np.clip(rho_m - rho_halos, 0, None))                       # :887
        mtot_add = float(dV * drho_m)                                             # :888
        if self.verbose:
            print(f"Inputted halos contribute {100*(rho_halos/rho_m):0.2f}% of the total matter density.")
            print(f"Remaining density is assigned to a uniform background.")
        if rho_halos > rho_m:
            warnings.warn("Inputted halos contribute more mass than is available for this mean matter density."
                          "Your Mtot_model profiles are either too extended or you are using the wrong cosmology.")
        final = float(np.power(res, 2)) if self.include_pixel_size else 1.0       # :1012-1015
        with torch.cuda.device(dev):
            t_paint = self._tables.get((_Ident(self.model), '2D', _Ident(getattr(self.model, 'interp2D', None))),
                                       lambda: profile_table_of(self.model, '2D', dev.index))
            t_tracer = self._tables2.get((_Ident(self.Tracer_model), _Ident(getattr(self.Tracer_model, 'interp2D', None))),
                                         lambda: profile_table_of(self.Tracer_model, '2D', dev.index))
            d_rec, d_ext, n = self._device_records(dev)
            d_orig = _to_device(orig_map[lo:hi], dev, dtype=np.float64)
            d_new = torch.zeros(nloc, dtype=torch.float64, device=dev)
            d_n = torch.zeros(1, dtype=torch.int64, device=dev)
            n_cols = 0 if d_ext is None else d_ext.shape[1]
            st = _lib.current_stream()
            _lib.check(L.bfg_grid_paint_anis(t_paint.handle, t_tracer.handle, N, float(res), n, _lib.ptr(d_rec),
                                             _lib.ptr(d_ext), n_cols, 1 if self.use_ellipticity else 0, _lib.ptr(d_mtot),
                                             mtot_add, _lib.ptr(d_orig), _lib.ptr(d_new), lo, hi, _lib.ptr(d_n), st))
            _lib.check(L.bfg_anis_background(nloc, _lib.ptr(d_mtot), mtot_add, _lib.ptr(d_orig),
                                             float(self.background_val * self.global_tracer_fraction), final,
                                             _lib.ptr(d_new), st))
            if self.plane_range is not None:
                from .parallel import gather_owned_ranges
                d_new = gather_owned_ranges(d_new, orig_map.size)
            out, out_np = _pinned_result(orig_map.size)
            out.copy_(d_new, non_blocking=True)
            n_up = int(d_n.cpu()[0])
            torch.cuda.current_stream().synchronize()
        self.last_stats = dict(n_updates=n_up, rho_halos=rho_halos, rho_m=rho_m, dV=float(dV))
        return out_np.reshape(orig_map.shape)


# =====================================================================================================================
# particle snapshots
# =====================================================================================================================
class DefaultRunnerSnapshot(object):
    """Constructor contract of BaryonForge/Runners/SnapshotRunner.py:83-100.  The periodic KD-tree of :100 is replaced
    by a device cell list built inside process(); KDTree_kwargs is accepted and ignored."""

    def __init__(self, HaloNDCatalog, ParticleSnapshot, epsilon_max, model, mass_def=None, verbose=True,
                 KDTree_kwargs={}, *, device=None, ncell=None, keep_cells=False):
        self.HaloNDCatalog = HaloNDCatalog
        self.ParticleSnapshot = ParticleSnapshot
        self.epsilon_max = epsilon_max
        self.cosmo = HaloNDCatalog.cosmology
        self.model = model
        self.mass_def = mass_def
        self.verbose = verbose
        self.KDTree_kwargs = KDTree_kwargs
        self.device = device
        self.ncell = ncell
        # keep_cells: keep the device cell list (cell-sorted particle copies, 32 B per particle of HBM) between process() calls
        # on the same ParticleSnapshot, the way the reference builds its KD-tree once in __init__ (SnapshotRunner.py:95-100)
        # and re-uses it when only `Runner.model` changes (examples/10_...ipynb cell 15).  Default off (it holds 32 B of HBM per particle between calls).
        self.keep_cells = keep_cells
        self._cells = None
        self.last_stats = {}
        self._tables = _TableCache()

    __getstate__ = DefaultRunner.__getstate__
    __setstate__ = DefaultRunner.__setstate__
    _device = DefaultRunner._device

    def halo_records(self):
        """Per-halo scalars of SnapshotRunner.py:219-228 (+ BaryonCorrection.py:371,398-399,410), vectorised."""
        cat = self.HaloNDCatalog.cat
        n = cat.size
        Lbox = self.ParticleSnapshot.L
        ndim = 2 if self.ParticleSnapshot.is2D else 3
        cosmo = cosmology.runner_cosmology(self.cosmo, with_w0=False)             # :198-201 (no w0)
        M32 = cat['M'].astype('<f4')
        M = M32.astype(np.float64)
        a = 1 / (1 + self.HaloNDCatalog.redshift)                                 # :225
        rec = np.zeros((_lib.HALO_STRIDE, n), dtype=np.float64).T
        if n == 0:
            return rec, None
        R_phys = cosmology.radius_of_mass(cosmo, M, a, self.mass_def)             # :226
        rec[:, _lib.HB_RQ] = np.clip(self.epsilon_max * R_phys / a, 0, Lbox / 2)  # :227-228
        mcosmo = _model_cosmo(self.model, cosmo)
        R_mod = cosmology.radius_of_mass(mcosmo, M, a, getattr(self.model, 'mass_def', None)) / a
        rec[:, _lib.HB_RCUT] = self.model.epsilon_max * R_mod
        rec[:, _lib.HB_LNRCOM] = np.log(R_mod)
        rec[:, _lib.HB_LNZ] = np.log(1 / a)
        rec[:, _lib.HB_LNM] = np.log(M32).astype(np.float64)
        self.last_scalars = dict(R_phys=R_phys, R_model_com=R_mod)
        for k, name in enumerate(['x', 'y', 'z'][:ndim]):
            rec[:, _lib.HB_X + k] = cat[name].astype('<f4').astype(np.float64)
        keys = list(vars(self.model).get('p_keys', []))
        _check_keys(self.model, keys)
        return rec, _extras(cat, keys)

    # ---- small public helpers of the reference's DefaultRunnerSnapshot (host-side numpy; k_snap_halos does the same
    #      minimum-image arithmetic on the device)
    def enforce_periodicity(self, dx):
        """Minimum image of a coordinate difference, one wrap (SnapshotRunner.py:135-158)."""
        L = self.ParticleSnapshot.L
        dx = np.where(dx > L / 2, dx - L, dx)
        return np.where(dx < -L / 2, dx + L, dx)

    def compute_distance(self, *args):
        """Periodic Euclidean distance from per-axis differences (SnapshotRunner.py:103-132)."""
        d = 0
        for dx in args:
            d = d + self.enforce_periodicity(dx) ** 2
        return np.sqrt(d)

    def _pick_ncell(self, rq, n_part, ndim, Lbox):
        if self.ncell is not None:
            return int(self.ncell)
        cap = 256 if ndim == 3 else 4096
        by_count = int(max(1, (n_part / 8.0) ** (1.0 / ndim)))       # >= ~8 particles per cell
        med = float(np.median(rq)) if rq.size else Lbox
        by_radius = int(max(1, Lbox / max(0.5 * med, 1e-300)))       # cells about half a typical query radius
        return int(max(1, min(cap, by_count, by_radius)))


class BaryonifySnapshot(DefaultRunnerSnapshot):
    """BaryonForge/Runners/SnapshotRunner.py:161-274."""

    def process(self):
        ps = self.ParticleSnapshot
        names = ['x', 'y', 'z'][:(2 if ps.is2D else 3)]
        S = self._displace_sorted()
        if S.get('d_raw') is not None:
            # the catalogue is on the device as raw 32-byte records: replace x, y(, z) inside the records (one sector read, one
            # sector write per particle) and bring the records back as they are -- no per-field traffic on either side
            d_raw, slots, d_s = S['d_raw'], S['slots'], S['d_s']
            with _torch().cuda.device(S['dev']):
                _lib.check(_lib.lib().bfg_snap_apply_records(S['ndim'], S['n_part'], _lib.ptr(d_s[0]), _lib.ptr(d_s[1]),
                                                             _lib.ptr(d_s[2]), _lib.ptr(S['d_tot']), _lib.ptr(S['d_order']),
                                                             S['L'], _lib.ptr(d_raw), _lib.ptr(d_raw), slots['x'], slots['y'],
                                                             slots.get('z', 0), _lib.current_stream()))
                self._finish_stats(S)
                return _device_to_raw(d_raw).view(ps.cat.dtype)
        d_p = self._apply_to_columns(S)
        new_cat = np.empty_like(ps.cat)                                           # :263 (a copy with x, y, z replaced below)
        for name in ps.cat.dtype.names:
            if name not in names:
                src, dst = ps.cat[name], new_cat[name]
                _parallel_chunks(lambda sl: np.copyto(dst[sl], src[sl]), len(src), chunk=1 << 22)
        with _torch().cuda.device(self._device()):
            _device_to_fields(d_p, new_cat, names)
        return new_cat

    def _displace_sorted(self):
        """
        Cell list + halo loop (SnapshotRunner.py:217-260) on the device.  Returns a dict with the CELL-ORDERED particle
        coordinates `d_s`, their accumulated offsets `d_tot` [ndim][n_part], `d_order` (sorted slot -> caller's index), the
        caller-ordered device copies `d_p` (free to be overwritten) and the geometry.
        """
        torch = _torch()
        ps = self.ParticleSnapshot
        ndim = 2 if ps.is2D else 3
        n_part = len(ps.cat)
        Lbox = float(ps.L)
        dev = self._device()
        L = _lib.lib()
        _warn_table_range(self.model, self.HaloNDCatalog.redshift, self.HaloNDCatalog.cat['M'], self, self.HaloNDCatalog.cat)
        with torch.cuda.device(dev):
            table = self._tables.get((_Ident(self.model), _Ident(getattr(self.model, 'interp_d', None))),
                                     lambda: displacement_table_of(self.model, dev.index))
            if os.environ.get("BFG_DEVICE_RECORDS", "1") == "1":      # per-halo scalars on the device (bfg_box_records)
                cat = self.HaloNDCatalog.cat
                d_rec0, d_aux = _box_device_records(self, cat, self.HaloNDCatalog.redshift, ndim, False, False, Lbox / 2, dev)
                aux = d_aux.cpu().numpy()
                self.last_scalars = dict(R_phys=aux[0], R_model_com=aux[1])
                keys = list(vars(self.model).get('p_keys', []))
                _check_keys(self.model, keys)
                extras = _extras(cat, keys)
                rq = d_rec0[:, _lib.HB_RQ].cpu().numpy() if cat.size else np.zeros(0)
                n_rec = cat.size
            else:
                rec, extras = self.halo_records()
                rq = rec[:, _lib.HB_RQ]
                d_rec0 = _upload_records(rec, dev)
                n_rec = rec.shape[0]
        ncell = self._pick_ncell(rq, n_part, ndim, Lbox)
        with torch.cuda.device(dev):
            st = _lib.current_stream()
            names = ['x', 'y', 'z'][:ndim]
            d_raw = slots = None
            cells_key = (_Ident(ps.cat), ncell, ndim, str(dev))
            kept = getattr(self, '_cells', None) if getattr(self, 'keep_cells', False) else None
            if kept is not None and kept[0] == cells_key:
                # the cell list of this ParticleSnapshot is still on the device: no upload, no counting sort; fresh output buffers
                _, d_s, d_start, d_order = kept
                d_p = [torch.empty(n_part, dtype=torch.float64, device=dev) for _ in names] + ([None] if ndim == 2 else [])
            else:
                slots = None if getattr(self, 'keep_cells', False) or os.environ.get("BFG_SNAP_RAW", "1") != "1" \
                    else _record_layout(ps.cat, names)
                d_s = [torch.empty(n_part, dtype=torch.float64, device=dev) for _ in names] + ([None] if ndim == 2 else [])
                d_start = torch.empty(ncell ** ndim + 1, dtype=torch.int64, device=dev)
                d_order = torch.empty(n_part, dtype=torch.int64, device=dev)
                if slots is not None:
                    d_raw = _raw_to_device(ps.cat.view(np.float64), dev)
                    d_p, base = None, d_raw.data_ptr()
                    src, stride = [base + 8 * slots[nm] for nm in names] + [None], 4
                else:
                    d_p = _fields_to_device(ps.cat, names, dev) + ([None] if ndim == 2 else [])
                    src, stride = [_lib.ptr(t) for t in d_p], 1
                _lib.check(L.bfg_snap_build_cells_strided(ndim, n_part, src[0], src[1], src[2], stride, Lbox, ncell,
                                                          _lib.ptr(d_start), _lib.ptr(d_order), _lib.ptr(d_s[0]),
                                                          _lib.ptr(d_s[1]), _lib.ptr(d_s[2]), st))
                if getattr(self, 'keep_cells', False):
                    self._cells = (cells_key, d_s, d_start, d_order)
            d_ext = None if extras is None else _to_device(extras, dev)
            d_rec, d_ext = _sort_records(d_rec0, d_ext, 1, Lbox, 16, ndim)
            d_tot = torch.zeros((ndim, n_part), dtype=torch.float64, device=dev)
            d_n = torch.zeros(1, dtype=torch.int64, device=dev)
            _lib.check(L.bfg_snap_offsets(table.handle, ndim, n_part, _lib.ptr(d_s[0]), _lib.ptr(d_s[1]), _lib.ptr(d_s[2]),
                                          Lbox, ncell, _lib.ptr(d_start), n_rec, _lib.ptr(d_rec), _lib.ptr(d_ext),
                                          table.n_extra, _lib.ptr(d_tot), _lib.ptr(d_n), st))
        return dict(dev=dev, ndim=ndim, n_part=n_part, L=Lbox, d_p=d_p, d_s=d_s, d_tot=d_tot, d_order=d_order, d_n=d_n,
                    ncell=ncell, d_raw=d_raw, slots=slots)

    def _finish_stats(self, S):
        self.last_stats = dict(n_pairs=int(S['d_n'].cpu()[0]), ncell=S['ncell'])

    def process_on_device(self):
        """
        The work of process() without the final download: returns the displaced, wrapped coordinates as float64 device
        tensors [x, y(, z)] in the caller's particle order, so that what follows in the reference's workflow (NGP deposit,
        P(k): deposit_ngp / spectra.ShellPowerSpectrum) can run without the particles leaving HBM.
        """
        return self._apply_to_columns(self._displace_sorted())

    def _apply_to_columns(self, S):
        torch = _torch()
        d_p, d_s = S['d_p'], S['d_s']
        with torch.cuda.device(S['dev']):
            if d_p is None:   # the particles came up as raw records: fresh column outputs
                d_p = [torch.empty(S['n_part'], dtype=torch.float64, device=S['dev']) for _ in range(S['ndim'])] + \
                      ([None] if S['ndim'] == 2 else [])
            # displaced positions overwrite the (no longer needed) unsorted device copies
            _lib.check(_lib.lib().bfg_snap_apply(S['ndim'], S['n_part'], _lib.ptr(d_s[0]), _lib.ptr(d_s[1]), _lib.ptr(d_s[2]),
                                                 _lib.ptr(S['d_tot']), _lib.ptr(S['d_order']), S['L'], _lib.ptr(d_p[0]),
                                                 _lib.ptr(d_p[1]), _lib.ptr(d_p[2]), _lib.current_stream()))
            self._finish_stats(S)
        return d_p[:S['ndim']]

    def process_to_map_on_device(self, N_grid):
        """
        `ParticleSnapshot(<process() output>).make_map(N_grid)` (SnapshotRunner.py:263-273 + utils/io.py:629-677) in one pass
        over the cell-ordered particles: the displaced positions are deposited (NGP, mass-weighted) straight from the cell
        list's order, without the scatter back to the caller's order.  Returns a float64 device tensor (N_grid,)*ndim.
        """
        torch = _torch()
        ps = self.ParticleSnapshot
        M = ps.cat['M']
        msg = "If you want to make a map, provide a value for the particle mass"                           # io.py:659
        S = self._displace_sorted()
        d_s, ndim = S['d_s'], S['ndim']
        with torch.cuda.device(S['dev']):
            if S.get('d_raw') is not None and 'M' in S['slots'] and M.size:
                # the masses came up inside the raw records: the NaN and equal-mass scans run on the device copy instead of
                # two strided passes over the host array
                col = S['d_raw'].view(-1, 4)[:, S['slots']['M']]
                flags = torch.stack([torch.isnan(col).any(), (col == col[0]).all()]).cpu()
                assert not bool(flags[0]), msg
                equal = bool(flags[1])
                d_m = None if equal else col.contiguous()
            else:
                assert np.isnan(M).sum() == 0, msg
                equal = M.size == 0 or bool(np.all(M == M[0]))
                d_m = None if equal else _to_device(M, S['dev'], dtype=np.float64)
            d_grid = torch.zeros((int(N_grid),) * ndim, dtype=torch.float64, device=S['dev'])
            _lib.check(_lib.lib().bfg_snap_apply_deposit(ndim, S['n_part'], _lib.ptr(d_s[0]), _lib.ptr(d_s[1]), _lib.ptr(d_s[2]),
                                                         _lib.ptr(S['d_tot']), _lib.ptr(S['d_order']), _lib.ptr(d_m),
                                                         float(M[0]) if (equal and M.size) else 0.0, S['L'], int(N_grid),
                                                         d_grid.data_ptr(), _lib.current_stream()))
            self._finish_stats(S)
        return d_grid

    def process_to_map(self, N_grid):
        """process() followed by make_map(N_grid) of the displaced particles, as a numpy array (see process_to_map_on_device)."""
        d_grid = self.process_to_map_on_device(N_grid)
        with _torch().cuda.device(d_grid.device):
            return _device_to_raw(d_grid.reshape(-1)).reshape(d_grid.shape)       # page-locked, recycled result buffer




def deposit_ngp(coords, mass, L, N_grid, device=None):
    """ParticleSnapshot.make_map (BaryonForge/utils/io.py:629-677) on the GPU: NGP mass histogram, float64."""
    torch = _torch()
    ndim = len(coords)
    dev = torch.device('cuda', torch.cuda.current_device() if device is None else int(device))
    with torch.cuda.device(dev):
        d_p = [_to_device(c, dev, dtype=np.float64) for c in coords] + ([None] if ndim == 2 else [])
        d_m = _to_device(mass, dev, dtype=np.float64)
        d_grid = torch.zeros(int(N_grid) ** ndim, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().bfg_snap_deposit_ngp(ndim, d_m.numel(), _lib.ptr(d_p[0]), _lib.ptr(d_p[1]), _lib.ptr(d_p[2]),
                                                  _lib.ptr(d_m), float(L), int(N_grid), _lib.ptr(d_grid),
                                                  _lib.current_stream()))
        out = d_grid.cpu().numpy()
    return out.reshape((int(N_grid),) * ndim)
