def _fftlog_transform(*a, **k):
    raise NotImplementedError("pyccl shim: FFTLog is used only while building tables (host, out of scope)")


def resample_array(*a, **k):
    raise NotImplementedError("pyccl shim: resample_array is outside the runner hot path")
