"""ceviche_b200: B200-native FDTD time stepping behind ceviche's `fdtd` API.

Scope: the hot path of ceviche.fdtd.forward() (Yee-grid curl updates, sigma-PML, J injection),
its caller loop and its eps_r derivatives -- nothing else of ceviche."""
from .constants import C_0, EPSILON_0, ETA_0, MU_0
from .fdtd import fdtd
from .jacobians import jacobian
from . import modes, optimizers, parametrization, utils

__version__ = "0.1.0"
__all__ = ["fdtd", "jacobian", "utils", "modes", "optimizers", "parametrization", "C_0", "EPSILON_0", "MU_0", "ETA_0"]
