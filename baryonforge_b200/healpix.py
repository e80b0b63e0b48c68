"""
Device HEALPix (RING) geometry, exposed for tests and shard assignment -- thin wrappers over the bfg_healpix_* entry
points.  These are the functions the reference takes from healpy (BaryonForge/Runners/HealpixRunner.py:327-361).
"""
import numpy as np

from . import _lib
from .runners import _torch, _to_device


def _dev():
    torch = _torch()
    return torch.device('cuda', torch.cuda.current_device())


def pix2vec(nside, pix_lo=0, pix_hi=None):
    torch = _torch()
    pix_hi = 12 * nside * nside if pix_hi is None else pix_hi
    out = torch.empty((3, pix_hi - pix_lo), dtype=torch.float64, device=_dev())
    _lib.check(_lib.lib().bfg_healpix_pix2vec(nside, pix_lo, pix_hi, _lib.ptr(out), _lib.current_stream()))
    return out.cpu().numpy()


def interp_weights(nside, theta, phi):
    torch = _torch()
    t, p = _to_device(theta, _dev(), np.float64), _to_device(phi, _dev(), np.float64)
    pix = torch.empty((4, t.numel()), dtype=torch.int64, device=_dev())
    w = torch.empty((4, t.numel()), dtype=torch.float64, device=_dev())
    _lib.check(_lib.lib().bfg_healpix_interp_weights(nside, t.numel(), _lib.ptr(t), _lib.ptr(p), _lib.ptr(pix), _lib.ptr(w),
                                                     _lib.current_stream()))
    return pix.cpu().numpy(), w.cpu().numpy()


def ang2pix(nside, theta, phi, nest=False):
    torch = _torch()
    t, p = _to_device(theta, _dev(), np.float64), _to_device(phi, _dev(), np.float64)
    pix = torch.empty(t.numel(), dtype=torch.int64, device=_dev())
    _lib.check(_lib.lib().bfg_healpix_ang2pix(nside, t.numel(), _lib.ptr(t), _lib.ptr(p), _lib.ptr(pix), _lib.current_stream()))
    if nest:
        out = torch.empty_like(pix)
        _lib.check(_lib.lib().bfg_healpix_reorder(nside, 1, pix.numel(), _lib.ptr(pix), _lib.ptr(out), _lib.current_stream()))
        pix = out
    return pix.cpu().numpy()


def reorder(nside, pix, to_nest):
    """RING -> NEST (to_nest=True) or NEST -> RING pixel indices on the device."""
    torch = _torch()
    d_in = _to_device(np.asarray(pix, dtype=np.int64), _dev())
    d_out = torch.empty_like(d_in)
    _lib.check(_lib.lib().bfg_healpix_reorder(nside, 1 if to_nest else 0, d_in.numel(), _lib.ptr(d_in), _lib.ptr(d_out),
                                              _lib.current_stream()))
    return d_out.cpu().numpy()


def halo_record(theta, phi, radius):
    """A shell halo record carrying only what query_disc needs (pointing + radius)."""
    rec = np.zeros(_lib.HALO_STRIDE)
    rec[_lib.HS_THETA], rec[_lib.HS_PHI], rec[_lib.HS_RADIUS] = theta, phi, radius
    return rec


def query_disc(nside, theta, phi, radius):
    """Ascending pixel list of the non-inclusive disc, computed on the device."""
    torch = _torch()
    rec = _to_device(halo_record(theta, phi, radius), _dev())
    cnt = torch.zeros(1, dtype=torch.int64, device=_dev())
    L = _lib.lib()
    _lib.check(L.bfg_healpix_query_disc(nside, _lib.ptr(rec), None, 0, _lib.ptr(cnt), _lib.current_stream()))
    n = int(cnt.cpu()[0])
    pix = torch.empty(max(n, 1), dtype=torch.int64, device=_dev())
    _lib.check(L.bfg_healpix_query_disc(nside, _lib.ptr(rec), _lib.ptr(pix), n, _lib.ptr(cnt), _lib.current_stream()))
    return np.sort(pix.cpu().numpy()[:n])


def disc_counts(nside, records):
    torch = _torch()
    rec = _to_device(records, _dev(), np.float64)
    out = torch.empty(records.shape[0], dtype=torch.int64, device=_dev())
    _lib.check(_lib.lib().bfg_healpix_disc_counts(nside, records.shape[0], _lib.ptr(rec), _lib.ptr(out), _lib.current_stream()))
    return out.cpu().numpy()


def table_readout(table, lnz, lnM, x, extras=None):
    """table(lnz, lnM, x_i[, extras]) on the device (exp() applied for log tables)."""
    torch = _torch()
    d_x = _to_device(x, _dev(), np.float64)
    out = torch.empty_like(d_x)
    ex = None if extras is None else np.ascontiguousarray(extras, dtype=np.float64)
    _lib.check(_lib.lib().bfg_table_readout(table.handle, float(lnz), float(lnM), _lib.ptr(ex), d_x.numel(), _lib.ptr(d_x),
                                            _lib.ptr(out), _lib.current_stream()))
    return out.cpu().numpy()
