"""
Data containers = the input/output layouts of the runner hot path.

Host-side mirror of BaryonForge/utils/io.py:9-677 (same class names, constructor signatures, attribute
names `.cat .cosmo .cosmology .map .NSIDE .redshift .bins .res .Npix .L .is2D .grid .inds` and error
behaviour), so user scripts that build these objects for the reference runners work unchanged.
Differences, all invisible to the runners:
  * GriddedMap.grid / .inds (io.py:463-470: d full-size float64 meshgrids + an int64 index cube, 34 GB at
    1024^3) are built lazily on first access -- the GPU path never touches them.
  * LightconeShell(path=...) reads the FITS file with healpy when it is installed, else with fits.read_map.
"""
import warnings

import numpy as np

__all__ = ['HaloLightConeCatalog', 'HaloNDCatalog', 'LightconeShell', 'GriddedMap', 'ParticleSnapshot']

_REQUIRED = ('Omega_m', 'sigma8', 'h', 'Omega_b', 'n_s', 'w0')


def _check_cosmo(cosmo):
    keys = cosmo.keys()
    if not all(k in keys for k in _REQUIRED):   # io.py:79-85
        raise ValueError("Not all cosmology parameters provided. I need Omega_m, sigma8, h, sigma8, Omega_b, n_s, w0")
    return cosmo


class _Container(object):
    @property
    def cosmology(self):
        return self.cosmo


class HaloLightConeCatalog(_Container):
    """Halos on the sky: float64 fields M, z, ra, dec (+ extra columns).  io.py:9-140."""

    def __init__(self, ra, dec, M, z, cosmo, **arrays):
        t = np.float64
        dtype = [('M', t), ('z', t), ('ra', t), ('dec', t)] + [(name, t) for name in arrays]
        cat = np.zeros(len(ra), dtype)
        if np.any(np.abs(dec) == 90):           # io.py:65-68
            dec = np.asarray(dec).astype(t)
            warnings.warn("Some halos found with declination exactly at the poles. Offsetting these by 4e-5 arcsec")
            dec = np.clip(dec, -90 + 1e-8, 90 - 1e-8)
        cat['ra'], cat['dec'], cat['z'], cat['M'] = ra, dec, z, M
        for name, arr in arrays.items():
            cat[name] = arr
        self.cat = cat
        self.cosmo = _check_cosmo(cosmo)

    @property
    def data(self):
        return self.cat

    def __getitem__(self, key):
        other = {k: self.cat[k][key] for k in self.cat.dtype.names if k not in ('ra', 'dec', 'M', 'z')}
        return HaloLightConeCatalog(ra=self.cat['ra'][key], dec=self.cat['dec'][key], M=self.cat['M'][key],
                                    z=self.cat['z'][key], cosmo=self.cosmo, **other)

    def __len__(self):
        return self.cat.size


class HaloNDCatalog(_Container):
    """Halos in a periodic box: big-endian float32 fields M, x, y, z (io.py:204-205) + extras."""

    def __init__(self, x, y, M, redshift, cosmo, z=None, **arrays):
        dtype = [('M', '>f'), ('x', '>f'), ('y', '>f'), ('z', '>f')]
        dtype = dtype + [(name, '>f', np.shape(arr)[1:] if np.ndim(arr) > 1 else '') for name, arr in arrays.items()]
        N = 1 if not isinstance(x, (list, np.ndarray, tuple)) else len(x)
        cat = np.zeros(N, dtype)
        cat['x'], cat['y'] = x, y
        cat['z'] = 0 if z is None else z
        cat['M'] = M
        for name, arr in arrays.items():
            cat[name] = arr
        self.cat = cat
        self.redshift = redshift
        self.cosmo = _check_cosmo(cosmo)

    @property
    def data(self):
        return self.cat

    def __getitem__(self, key):
        other = {k: self.cat[k][key] for k in self.cat.dtype.names if k not in ('x', 'y', 'z', 'M')}
        return HaloNDCatalog(x=self.cat['x'][key], y=self.cat['y'][key], z=self.cat['z'][key], M=self.cat['M'][key],
                             redshift=self.redshift, cosmo=self.cosmo, **other)

    def __len__(self):
        return self.cat.size


class LightconeShell(_Container):
    """A HEALPix RING map of one lightcone shell.  io.py:290-379."""

    def __init__(self, map=None, path=None, cosmo=None, redshift=None):
        if (path is None) & (map is None):
            raise ValueError("Need to provide either path to map, or provide map values in healpix ring configuration")
        elif isinstance(path, str):
            try:
                import healpy as hp
                self.map = hp.read_map(path)                              # io.py:347
            except ImportError:
                from .fits import read_map                                # same file layout, no healpy needed
                self.map = read_map(path)
        elif isinstance(map, np.ndarray):
            self.map = map
        nside = int(round(np.sqrt(self.map.size / 12.0)))
        if 12 * nside * nside != self.map.size:
            raise ValueError("Wrong pixel number (it is not 12*nside**2)")
        self.NSIDE = nside
        self.redshift = redshift
        self.cosmo = _check_cosmo(cosmo)

    @property
    def data(self):
        return self.map


class GriddedMap(_Container):
    """A square / cubic periodic grid.  io.py:382-494."""

    def __init__(self, map=None, redshift=None, bins=None, cosmo=None):
        self.map = map
        self.redshift = redshift
        self.Npix = self.map.shape[0]
        self.res = bins[1] - bins[0]
        self.bins = bins
        self.L = bins[-1] + self.res / 2
        self.is2D = True if len(self.map.shape) == 2 else False
        if self.is2D:
            assert self.map.shape[0] == self.map.shape[1]
        else:
            assert (self.map.shape[0] == self.map.shape[1]) & (self.map.shape[1] == self.map.shape[2])
        assert self.Npix == self.bins.size, f"Map has {self.Npix} pixels a side, but you passed {self.bins.size} bins"
        self._grid = None
        self._inds = None
        self.cosmo = _check_cosmo(cosmo)

    @property
    def grid(self):
        if self._grid is None:
            self._grid = np.meshgrid(*([self.bins] * (2 if self.is2D else 3)), indexing='xy')
        return self._grid

    @property
    def inds(self):
        if self._inds is None:
            self._inds = np.arange(self.map.size).reshape(self.map.shape)
        return self._inds

    @property
    def data(self):
        return self.map


class ParticleSnapshot(_Container):
    """Particles in a periodic box: float64 M, x, y, z.  io.py:497-677."""

    def __init__(self, x=None, y=None, z=None, M=None, L=None, redshift=None, cosmo=None):
        dtype = [('M', np.float64), ('x', np.float64), ('y', np.float64), ('z', np.float64)]
        cat = np.zeros(len(x), dtype)
        cat['x'], cat['y'] = x, y
        cat['z'] = 0 if z is None else z
        cat['M'] = M
        self.L = L
        self.cat = cat
        self.redshift = redshift
        self.is2D = True if z is None else False
        self.cosmo = _check_cosmo(cosmo)

    @property
    def data(self):
        return self.cat

    def make_map(self, N_grid):
        """NGP mass deposit (io.py:629-677), on the GPU."""
        from .runners import deposit_ngp
        assert np.isnan(self.cat['M']).sum() == 0, "If you want to make a map, provide a value for the particle mass"
        coords = [self.cat['x'], self.cat['y']] + ([] if self.is2D else [self.cat['z']])
        return deposit_ngp(coords, self.cat['M'], self.L, N_grid)
