class HMCalculator(object):
    def __init__(self, *a, **k):
        raise NotImplementedError("pyccl shim: halo model is outside the runner hot path")
