"""Import-path shim: ``sde4mbrl_px4_b200.rotor_uav.sde_mpc_design`` / ``.utils`` mirror the two
modules the reference node imports from ``sde4mbrlExamples.rotor_uav`` (sde_control.py:12-13)."""
from .. import sde_mpc_design, utils  # noqa: F401
