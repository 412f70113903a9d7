"""``sde4mbrl_px4_b200.rotor_uav`` mirrors the package the reference node imports from
(``sde4mbrlExamples.rotor_uav``, sde_control.py:12-13): the submodules ``sde_mpc_design`` and
``utils`` are real modules, so the node's two import lines work with only the package name changed."""
from . import sde_mpc_design, utils  # noqa: F401
