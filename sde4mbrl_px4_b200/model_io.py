"""Learned-SDE model files (``learned_model_params`` of the YAML, launch/iris_sitl_traj_mpc.yaml:3).

The reference points at a Haiku pickle produced by the un-vendored ``sde4mbrl``
package; no pickle ships with it and unpickling needs JAX.  This module defines
the on-disk ``.npz`` format this framework loads, a seeded synthetic generator
(SURVEY.md section 8d "Synthetic inputs") and the packing into the flat blob of
``include/sdempc.h`` (``sdempc_model_header`` + weights).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os

import numpy as np

from . import _abi

NETS = ("drift", "diff")
LAYERS = ("W1", "b1", "W2", "b2", "W3", "b3")


@dataclasses.dataclass
class SDEModel:
    """Multirotor rigid-body prior + residual drift MLP + diffusion MLP."""

    nu: int
    width: int
    mass: float
    gravity: float
    k_thrust: float
    inertia: np.ndarray  # [3]
    mixer: np.ndarray  # [3, nu]  M_b = mixer @ T
    sigma_prior: np.ndarray  # [6]
    weights: dict  # "{net}_{layer}" -> float32 array

    @property
    def n_in(self) -> int:
        return 6 + self.nu

    def to_blob(self) -> bytes:
        h = _abi.ModelHeader()
        h.magic, h.version = _abi.MODEL_MAGIC, _abi.MODEL_VERSION
        h.nu, h.n_in, h.width, h.n_hidden, h.n_out = self.nu, self.n_in, self.width, 2, 6
        h.mass, h.gravity, h.k_thrust = self.mass, self.gravity, self.k_thrust
        for i in range(3):
            h.inertia[i] = float(self.inertia[i])
        mix = np.zeros((3, _abi.MAX_NU), np.float32)
        mix[:, : self.nu] = self.mixer
        for i, v in enumerate(mix.reshape(-1)):
            h.mixer[i] = float(v)
        for i in range(6):
            h.sigma_prior[i] = float(self.sigma_prior[i])
        parts = [bytes(h)]
        shapes = self.layer_shapes()
        for net in NETS:
            for lay in LAYERS:
                w = np.ascontiguousarray(self.weights[f"{net}_{lay}"], dtype=np.float32)
                assert w.shape == shapes[lay], (net, lay, w.shape, shapes[lay])
                parts.append(w.tobytes())
        return b"".join(parts)

    def layer_shapes(self) -> dict:
        W, n_in = self.width, self.n_in
        return {"W1": (W, n_in), "b1": (W,), "W2": (W, W), "b2": (W,), "W3": (6, W), "b3": (6,)}

    def save(self, path: str) -> None:
        np.savez(
            path, format=np.array("sdempc-model-v1"), nu=self.nu, width=self.width, mass=self.mass,
            gravity=self.gravity, k_thrust=self.k_thrust, inertia=self.inertia.astype(np.float32),
            mixer=self.mixer.astype(np.float32), sigma_prior=self.sigma_prior.astype(np.float32),
            **{k: v.astype(np.float32) for k, v in self.weights.items()},
        )

    @staticmethod
    def load(path: str) -> "SDEModel":
        path = os.path.expanduser(path)
        if path.endswith(".pkl"):
            raise RuntimeError(
                f"{path}: Haiku pickles of the upstream sde4mbrl package need JAX to load and are not supported; "
                "convert the parameters to the .npz format of sde4mbrl_px4_b200.model_io (see INTEGRATION.md)."
            )
        z = np.load(path, allow_pickle=False)
        if str(z["format"]) != "sdempc-model-v1":
            raise RuntimeError(f"{path}: unknown model format {z['format']}")
        weights = {f"{n}_{l}": np.asarray(z[f"{n}_{l}"], np.float32) for n in NETS for l in LAYERS}
        return SDEModel(
            nu=int(z["nu"]), width=int(z["width"]), mass=float(z["mass"]), gravity=float(z["gravity"]),
            k_thrust=float(z["k_thrust"]), inertia=np.asarray(z["inertia"], np.float32),
            mixer=np.asarray(z["mixer"], np.float32), sigma_prior=np.asarray(z["sigma_prior"], np.float32),
            weights=weights,
        )


def _mixer(arm: float, theta_deg, spin, c_m: float) -> np.ndarray:
    """FRD body, thrust along -z_b, motor i at arm*(cos th, sin th, 0):
    M_b = sum_i (-r_y T_i, +r_x T_i, s_i c_M T_i)   (SURVEY.md section 8a [SPEC] "Drift")."""
    th = np.deg2rad(np.asarray(theta_deg, np.float64))
    rx, ry = arm * np.cos(th), arm * np.sin(th)
    return np.stack([-ry, rx, np.asarray(spin, np.float64) * c_m]).astype(np.float32)


VEHICLES = {
    # PX4 quad-x motor order; hover at uref = 0.71 (launch/iris_sitl_traj_mpc.yaml:33)
    "iris": dict(nu=4, width=32, mass=1.5, inertia=(0.029, 0.029, 0.055), arm=0.255, hover=0.71,
                 theta=(45.0, 225.0, 315.0, 135.0), spin=(1, 1, -1, -1)),
    # PX4 hex-x motor order; hover at uref = 0.42 (launch/hexa_sitl_traj_mpc.yaml:14)
    "hexa": dict(nu=6, width=64, mass=2.0, inertia=(0.05, 0.05, 0.09), arm=0.30, hover=0.42,
                 theta=(90.0, 270.0, 330.0, 150.0, 30.0, 210.0), spin=(-1, 1, -1, 1, 1, -1)),
}


def synthetic_model(vehicle: str = "iris", seed: int = 0, width: int | None = None,
                    weight_scale: float = 0.1, bias_scale: float = 0.0) -> SDEModel:
    """Seeded synthetic learned model: hover-capable plant,
    MLP weights N(0, (weight_scale/sqrt(fan_in))^2), biases N(0, bias_scale^2) (zero by default)."""
    v = VEHICLES[vehicle]
    nu, W = v["nu"], int(width or v["width"])
    g = 9.81
    rng = np.random.default_rng(seed)
    n_in = 6 + nu
    weights = {}
    for net in NETS:
        for lay, (o, i) in (("1", (W, n_in)), ("2", (W, W)), ("3", (6, W))):
            weights[f"{net}_W{lay}"] = (rng.standard_normal((o, i)) * (weight_scale / np.sqrt(i))).astype(np.float32)
            weights[f"{net}_b{lay}"] = np.zeros((o,), np.float32)
    if bias_scale != 0.0:   # drawn after all weights so that the default models are unchanged
        for net in NETS:
            for lay in ("1", "2", "3"):
                b = weights[f"{net}_b{lay}"]
                weights[f"{net}_b{lay}"] = (rng.standard_normal(b.shape) * bias_scale).astype(np.float32)
    return SDEModel(
        nu=nu, width=W, mass=v["mass"], gravity=g, k_thrust=v["mass"] * g / (nu * v["hover"] ** 2),
        inertia=np.asarray(v["inertia"], np.float32), mixer=_mixer(v["arm"], v["theta"], v["spin"], 0.016),
        sigma_prior=np.asarray([0.05] * 3 + [0.1] * 3, np.float32), weights=weights,
    )


def blob_buffer(blob: bytes):
    """ctypes buffer that keeps ``blob`` alive for the duration of a C call."""
    return (C.c_char * len(blob)).from_buffer_copy(blob)
