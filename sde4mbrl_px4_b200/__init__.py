"""B200-native drop-in for the gradient-based neural-SDE MPC solve of
wuwushrek/sde4mbrl_px4 (``sde4mbrl_px4/mpc_controller/sde_control.py``).

Public surface = what the reference node imports (sde_control.py:12-13):
``load_mpc_from_cfgfile`` and ``enu2ned`` (also reachable under the reference's module
layout as ``sde4mbrl_px4_b200.rotor_uav.sde_mpc_design`` / ``.rotor_uav.utils``).
Importing the package loads no CUDA context (fork rule, sde_control.py:723-728).
"""
__version__ = "0.2.0"

from .sde_mpc_design import load_mpc_from_cfgfile  # noqa: E402,F401
from .utils import enu2ned  # noqa: E402,F401

__all__ = ["load_mpc_from_cfgfile", "enu2ned", "__version__"]
