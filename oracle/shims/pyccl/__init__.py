"""
Stand-in for the third-party `pyccl` package (absent from this image, no network) --
TEST INFRASTRUCTURE.  It carries (i) the import-time surface /root/reference/BaryonForge
needs, and (ii) the only run-time arithmetic the runner hot path takes from pyccl:
  Cosmology(...), .compute_sigma() (no-op), angular_diameter_distance(cosmo, a),
  halos.massdef.MassDef(200, 'critical').get_radius(cosmo, M, a)
(/root/reference/BaryonForge/Runners/HealpixRunner.py:280-285,299,320; Map2DRunner.py:462-466,491;
SnapshotRunner.py:198-202,226; Profiles/BaryonCorrection.py:399).
Restates CCL's published background formulas (flat wCDM + photons + 3.044 massless neutrinos).
The per-halo scalars it yields are INPUTS shared by the oracle and the CUDA path, so hot-path
parity does not hinge on CCL's constants.
"""
import numpy as np
from . import physical_constants, pyutils, halos, core
from .core import Cosmology, angular_diameter_distance, comoving_radial_distance, h_over_h0, rho_x


def unlock_instance(func=None, **kw):
    if func is None:
        return lambda f: f
    return func


def sigmaM(cosmo, M, a):
    raise NotImplementedError("pyccl shim: sigma(M) is outside the runner hot path")


def correlation_3d(*a, **k):
    raise NotImplementedError("pyccl shim: correlation_3d is outside the runner hot path")
