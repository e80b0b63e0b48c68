from .profile_base import HaloProfile


class HaloProfileHOD(HaloProfile):
    pass
