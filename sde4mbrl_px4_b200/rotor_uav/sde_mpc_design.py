"""``from sde4mbrl_px4_b200.rotor_uav.sde_mpc_design import load_mpc_from_cfgfile`` — the import the
reference node makes at sde_control.py:12, served by ``sde4mbrl_px4_b200.sde_mpc_design``."""
from ..sde_mpc_design import (HostArray, MPCController, OptState, PRNGKey, jit,  # noqa: F401
                              load_mpc_from_cfgfile, split)

__all__ = ["load_mpc_from_cfgfile", "MPCController", "OptState", "HostArray", "jit", "PRNGKey", "split"]
