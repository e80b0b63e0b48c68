from . import hod
from .profile_base import HaloProfile
