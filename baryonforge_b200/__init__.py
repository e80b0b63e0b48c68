"""
baryonforge_b200 -- B200-native (sm_100a CUDA) back-end for BaryonForge's runner hot path.

Drop-in for `from BaryonForge.Runners import ...` / `from BaryonForge.utils.io import ...` on that path only:
    BaryonifyShell, PaintProfilesShell, BaryonifyGrid, PaintProfilesGrid, BaryonifySnapshot
    HaloLightConeCatalog, HaloNDCatalog, LightconeShell, GriddedMap, ParticleSnapshot
Tables are still built on the host by the reference's pyccl code; see tables.py.
"""
from .io import *            # noqa: F401,F403
from .runners import *       # noqa: F401,F403
from .tables import DeviceTable, DisplacementModel, ProfileModel   # noqa: F401
from .parallel import SimpleParallel, SplitJoinParallel             # noqa: F401
from .spectra import ShellPowerSpectrum                             # noqa: F401
from . import _lib, cosmology, harmonics, io, parallel, runners, spectra, synth, tables   # noqa: F401

__version__ = "0.1.0"
