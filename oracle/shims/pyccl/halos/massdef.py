import numpy as np
from .. import core


class MassDef(object):
    """MassDef(Delta, rho_type).get_radius -> physical Mpc, as CCL: R = (M / (4pi/3 Delta rho_x(a)))^(1/3)."""

    def __init__(self, Delta, rho_type):
        self.Delta, self.rho_type = Delta, rho_type
        self.name = "%s%s" % (Delta, rho_type[0])

    def get_Delta(self, cosmo, a):
        if self.Delta in ('vir', 'fof'):
            raise NotImplementedError("pyccl shim: only numeric Delta")
        return self.Delta

    def get_radius(self, cosmo, M, a):
        M_use = np.atleast_1d(M)
        Delta = self.get_Delta(cosmo, a)
        R = (M_use / (4.18879020479 * Delta * core.rho_x(cosmo, a, self.rho_type))) ** (1. / 3.)
        if np.ndim(M) == 0:
            return R[0]
        return R

    def get_mass(self, cosmo, R, a):
        R_use = np.atleast_1d(R)
        M = 4.18879020479 * core.rho_x(cosmo, a, self.rho_type) * self.get_Delta(cosmo, a) * R_use ** 3
        if np.ndim(R) == 0:
            return M[0]
        return M

    def __eq__(self, other):
        return isinstance(other, MassDef) and (self.Delta, self.rho_type) == (other.Delta, other.rho_type)

    def __hash__(self):
        return hash((self.Delta, self.rho_type))


MassDef200c = MassDef(200, 'critical')
MassDef200m = MassDef(200, 'matter')
MassDefVir = MassDef('vir', 'critical')
