"""``from sde4mbrl_px4_b200.rotor_uav.utils import enu2ned`` — the import the reference node makes at
sde_control.py:13, served by ``sde4mbrl_px4_b200.utils``."""
from ..utils import PRNGKey, enu2ned, ned2enu, split  # noqa: F401

__all__ = ["enu2ned", "ned2enu", "PRNGKey", "split"]
