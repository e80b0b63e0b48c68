class _FFTLogPrecision(dict):
    def to_dict(self):
        return dict(self)


class HaloProfile(object):
    """Import-time surface of ccl.halos.profiles.HaloProfile: real/projected dispatch to _real/_projected."""

    def __init__(self, mass_def=None, concentration=None, is_number_counts=False):
        self.mass_def = mass_def
        self.concentration = concentration
        self.precision_fftlog = _FFTLogPrecision(padding_lo_fftlog=0.1, padding_lo_extra=0.1, padding_hi_fftlog=10.,
                                                 padding_hi_extra=10., large_padding_2D=False, n_per_decade=100,
                                                 extrapol='linx_liny', plaw_fourier=-1.5, plaw_projected=-1.)

    def update_precision_fftlog(self, **kwargs):
        self.precision_fftlog.update(kwargs)

    def real(self, cosmo, r, M, a, **kw):
        if getattr(self, '_real', None) is None:
            raise NotImplementedError("pyccl shim: no FFTLog path")
        return self._real(cosmo, r, M, a, **kw)

    def projected(self, cosmo, r_t, M, a, **kw):
        if getattr(self, '_projected', None) is None:
            raise NotImplementedError("pyccl shim: no FFTLog path")
        return self._projected(cosmo, r_t, M, a, **kw)

    def fourier(self, cosmo, k, M, a, **kw):
        if getattr(self, '_fourier', None) is None:
            raise NotImplementedError("pyccl shim: no FFTLog path")
        return self._fourier(cosmo, k, M, a, **kw)
