"""ctypes binding of oracle/healpix_ring.c (test infrastructure; see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile the oracle's C files with gcc (outputs only under oracle/_build/)."""
    so = os.path.join(_HERE, "_build", "libhpo.so")
    srcs = [os.path.join(_HERE, f) for f in ("healpix_ring.c", "grid_deposit.c")]
    so2 = os.path.join(_HERE, "_build", "libgrid.so")
    stale = force or not (os.path.exists(so) and os.path.exists(so2)) or any(
        os.path.getmtime(s) > min(os.path.getmtime(so), os.path.getmtime(so2)) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libhpo.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        i64, dbl = C.c_int64, C.c_double
        pi64 = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        pdbl = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.hpo_pix2vec.argtypes = [i64, i64, pi64, pdbl, pdbl, pdbl]
        L.hpo_pix2vec_range.argtypes = [i64, i64, i64, pdbl, pdbl, pdbl]
        L.hpo_pix2ang.argtypes = [i64, i64, pi64, pdbl, pdbl]
        L.hpo_ang2pix.argtypes = [i64, i64, pdbl, pdbl, pi64]
        L.hpo_vec2pointing.argtypes = [pdbl, C.POINTER(dbl), C.POINTER(dbl)]
        L.hpo_query_disc.argtypes = [i64, dbl, dbl, dbl, C.c_void_p, i64]
        L.hpo_query_disc.restype = i64
        L.hpo_get_interpol.argtypes = [i64, i64, pdbl, pdbl, pi64, pdbl]
        L.hpo_ring_above.argtypes = [i64, dbl]
        L.hpo_ring_above.restype = i64
        L.hpo_ring2z.argtypes = [i64, i64]
        L.hpo_ring2z.restype = dbl
        L.hpo_ring_info.argtypes = [i64, i64, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int)]
        L.hpo_regrid_scatter.argtypes = [pdbl, i64, pdbl, pi64, pdbl]
        L.hpo_ring2nest.argtypes = [i64, i64, pi64, pi64]
        L.hpo_nest2ring.argtypes = [i64, i64, pi64, pi64]
        _LIB = L
    return _LIB


def pix2vec(nside, ipix):
    ipix = np.ascontiguousarray(np.atleast_1d(ipix), dtype=np.int64)
    x = np.empty(ipix.size); y = np.empty(ipix.size); z = np.empty(ipix.size)
    lib().hpo_pix2vec(int(nside), ipix.size, ipix.ravel(), x, y, z)
    return x.reshape(ipix.shape), y.reshape(ipix.shape), z.reshape(ipix.shape)


def pix2vec_range(nside, p0, n):
    x = np.empty(n); y = np.empty(n); z = np.empty(n)
    lib().hpo_pix2vec_range(int(nside), int(p0), int(n), x, y, z)
    return x, y, z


def pix2ang(nside, ipix):
    ipix = np.ascontiguousarray(np.atleast_1d(ipix), dtype=np.int64)
    t = np.empty(ipix.size); p = np.empty(ipix.size)
    lib().hpo_pix2ang(int(nside), ipix.size, ipix.ravel(), t, p)
    return t.reshape(ipix.shape), p.reshape(ipix.shape)


def ang2pix(nside, theta, phi):
    theta = np.ascontiguousarray(np.atleast_1d(theta), dtype=np.float64)
    phi = np.ascontiguousarray(np.atleast_1d(phi), dtype=np.float64)
    out = np.empty(theta.size, dtype=np.int64)
    lib().hpo_ang2pix(int(nside), theta.size, theta.ravel(), phi.ravel(), out)
    return out.reshape(theta.shape)


def vec2pointing(vec):
    v = np.ascontiguousarray(vec, dtype=np.float64)
    t = C.c_double(); p = C.c_double()
    lib().hpo_vec2pointing(v, C.byref(t), C.byref(p))
    return t.value, p.value


def query_disc(nside, theta, phi, radius):
    L = lib()
    n = L.hpo_query_disc(int(nside), float(theta), float(phi), float(radius), None, 0)
    out = np.empty(n, dtype=np.int64)
    if n:
        L.hpo_query_disc(int(nside), float(theta), float(phi), float(radius), out.ctypes.data, n)
    return out


def get_interpol(nside, theta, phi):
    theta = np.ascontiguousarray(np.atleast_1d(theta), dtype=np.float64).ravel()
    phi = np.ascontiguousarray(np.atleast_1d(phi), dtype=np.float64).ravel()
    pix = np.empty((4, theta.size), dtype=np.int64)
    wgt = np.empty((4, theta.size), dtype=np.float64)
    lib().hpo_get_interpol(int(nside), theta.size, theta, phi, pix, wgt)
    return pix, wgt


def regrid_scatter(hmap, parent_vals, child_pix, child_w):
    lib().hpo_regrid_scatter(hmap, parent_vals.size, np.ascontiguousarray(parent_vals),
                             np.ascontiguousarray(child_pix, dtype=np.int64), np.ascontiguousarray(child_w))
    return hmap


def ring2nest(nside, ipix):
    ipix = np.ascontiguousarray(np.atleast_1d(ipix), dtype=np.int64)
    out = np.empty(ipix.size, dtype=np.int64)
    lib().hpo_ring2nest(int(nside), ipix.size, ipix.ravel(), out)
    return out.reshape(ipix.shape)


def nest2ring(nside, ipix):
    ipix = np.ascontiguousarray(np.atleast_1d(ipix), dtype=np.int64)
    out = np.empty(ipix.size, dtype=np.int64)
    lib().hpo_nest2ring(int(nside), ipix.size, ipix.ravel(), out)
    return out.reshape(ipix.shape)
