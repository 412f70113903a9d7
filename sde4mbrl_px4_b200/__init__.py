"""B200-native drop-in for the gradient-based neural-SDE MPC solve of
wuwushrek/sde4mbrl_px4 (``sde4mbrl_px4/mpc_controller/sde_control.py``).

Public surface mirrors what the reference node imports
(sde_control.py:12-13): ``load_mpc_from_cfgfile`` and ``enu2ned``.
"""
__version__ = "0.1.0"
