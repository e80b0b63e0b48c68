"""
(z, ln M, ln r[, extras]) tables: from the reference's model objects to HBM, once.

The tables themselves are still built on the host by the reference's pyccl code
(Baryonification2D/3D.setup_interpolator, BaryonForge/Profiles/BaryonCorrection.py:142-328;
TabulatedProfile.setup_interpolator, BaryonForge/utils/Tabulate.py:193-276).  This module only reads the
public attributes those methods leave behind (`interp_d`, `interp2D`, `interp3D`, `Rdelta_sampling`,
`epsilon_max`, `p_keys`, `mass_def`, `cosmo`) and uploads them through bfg_table_create.

`DisplacementModel` / `ProfileModel` are minimal stand-ins carrying exactly those attributes, for use where the
reference package (pyccl) is not installed: synthetic benchmarks, tests, or tables loaded from disk.
"""
import ctypes as C

import numpy as np

from . import _lib

__all__ = ['DeviceTable', 'DisplacementModel', 'ProfileModel', 'displacement_table_of', 'profile_table_of',
           'get_parameter']


class _Grid(object):
    """Just enough of scipy's RegularGridInterpolator surface (.grid, .values) to be read back."""

    def __init__(self, axes, values):
        self.grid = tuple(np.ascontiguousarray(a, dtype=np.float64) for a in axes)
        self.values = np.ascontiguousarray(values, dtype=np.float64)


class DisplacementModel(object):
    """Carries what BaryonificationClass.setup_interpolator stores (BaryonCorrection.py:307-323)."""

    def __init__(self, axes, values, epsilon_max, cosmo, mass_def=None, Rdelta_sampling=False, p_keys=()):
        self.interp_d = _Grid(axes, values)
        self.raw_input_d = self.interp_d.values
        self.raw_input_z_range, self.raw_input_M_range, self.raw_input_r_range = self.interp_d.grid[:3]
        self.epsilon_max = epsilon_max
        self.cosmo = cosmo
        self.mass_def = mass_def
        self.Rdelta_sampling = Rdelta_sampling
        self.p_keys = list(p_keys)


class ProfileModel(object):
    """Carries what TabulatedProfile.setup_interpolator stores (Tabulate.py:261-271): interp2D/3D over log(table)."""

    def __init__(self, axes, raw3D=None, raw2D=None, mass_def=None, p_keys=(), proj_cutoff=None):
        with np.errstate(divide='ignore', invalid='ignore'):
            self.interp3D = None if raw3D is None else _Grid(axes, np.log(np.asarray(raw3D, dtype=np.float64)))
            self.interp2D = None if raw2D is None else _Grid(axes, np.log(np.asarray(raw2D, dtype=np.float64)))
        self.mass_def = mass_def
        self.p_keys = list(p_keys)
        if proj_cutoff is not None:     # the wrapped profile's line-of-sight half-length (Profiles/Base.py), read by the
            self.proj_cutoff = proj_cutoff   # anisotropic painters through _get_parameter


def get_parameter(obj, key, _depth=0):
    """
    utils/Tabulate.py:66-96 `_get_parameter`: the first attribute called `key` on `obj` or, recursively, on the halo
    profiles it wraps (TabulatedProfile.model, ...).  Without pyccl a "profile" is anything exposing real/projected.
    """
    def is_profile(v):
        if (hasattr(v, 'projected') and hasattr(v, 'real')) or hasattr(v, 'interp2D'):
            return True
        try:
            import pyccl as ccl
            return isinstance(v, ccl.halos.profiles.HaloProfile)
        except Exception:
            return False
    for k in dir(obj):
        if k == key:
            return getattr(obj, key)
        if k.startswith('__') or _depth > 8:
            continue
        try:
            v = getattr(obj, k)
        except Exception:
            continue
        if is_profile(v):
            return get_parameter(v, key, _depth + 1)
    return None


class DeviceTable(object):
    """Owns one bfg_table (HBM-resident).  flags: _lib.TABLE_LOG_VALUES | _lib.TABLE_RDELTA."""

    def __init__(self, axes, values, flags, device):
        axes = [np.ascontiguousarray(a, dtype=np.float64) for a in axes]
        values = np.ascontiguousarray(values, dtype=np.float64)
        if values.shape != tuple(a.size for a in axes):
            raise ValueError(f"table shape {values.shape} does not match its axes {[a.size for a in axes]}")
        if len(axes) < 3:
            raise ValueError("table needs (ln(1+z), ln M, ln r) axes")
        self.ndim = len(axes)
        self.shape = values.shape
        self.flags = flags
        self.device = int(device)
        self.n_extra = self.ndim - 3
        self._h = C.c_void_p()
        shape = (C.c_int64 * self.ndim)(*values.shape)
        ax = (C.c_void_p * self.ndim)(*[a.ctypes.data for a in axes])
        _lib.check(_lib.lib().bfg_table_create(C.byref(self._h), self.ndim, shape, ax, values.ctypes.data, flags,
                                               self.device))

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            _lib.lib().bfg_table_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _grid_of(interp, what):
    if interp is None or not hasattr(interp, 'grid') or not hasattr(interp, 'values'):
        raise TypeError(
            f"The GPU runners need a tabulated model: `{what}` (a RegularGridInterpolator) is missing. Run "
            "setup_interpolator() first; analytic (un-tabulated) profiles are not supported and there is no CPU fallback.")
    return [np.asarray(g, dtype=np.float64) for g in interp.grid], np.asarray(interp.values, dtype=np.float64)


def displacement_table_of(model, device):
    """Upload model.interp_d (BaryonCorrection.py:322).  Raises NameError like the reference when it was never built."""
    if not hasattr(model, 'interp_d'):
        if hasattr(model, 'displacement') and not hasattr(model, 'setup_interpolator'):
            raise TypeError(f"{type(model)} is not a tabulated displacement model (no `interp_d`)")
        raise NameError("No Table created. Run setup_interpolator() method first")   # BaryonCorrection.py:454-455
    axes, values = _grid_of(model.interp_d, 'interp_d')
    flags = _lib.TABLE_RDELTA if getattr(model, 'Rdelta_sampling', False) else 0
    return DeviceTable(axes, values, flags, device)


def profile_table_of(model, which, device):
    """Upload model.interp2D or .interp3D (already log, Tabulate.py:270-271,589-590)."""
    name = 'interp2D' if which == '2D' else 'interp3D'
    if not (hasattr(model, 'interp3D') and hasattr(model, 'interp2D')):
        if hasattr(model, 'setup_interpolator'):
            raise NameError("No Table created. Run setup_interpolator() method first")   # Tabulate.py:354-355
        raise TypeError(
            f"{type(model)} is not a TabulatedProfile/ParamTabulatedProfile: the GPU runners read profile tables and "
            "have no CPU fallback. Wrap the profile in TabulatedProfile(...).setup_interpolator(...).")
    axes, values = _grid_of(getattr(model, name), name)
    return DeviceTable(axes, values, _lib.TABLE_LOG_VALUES, device)
