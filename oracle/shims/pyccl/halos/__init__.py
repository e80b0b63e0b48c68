from . import massdef, profiles, concentration, halo_model, halo_model_base, mass_translator
from .massdef import MassDef, MassDef200c, MassDefVir
from .profiles import HaloProfile
from .halo_model import HMCalculator
from .halo_model_base import Concentration
