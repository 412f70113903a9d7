#!/usr/bin/env python3
"""Converter STUB: upstream learned-SDE parameters -> this framework's ``.npz`` model format.

The reference's YAMLs point ``learned_model_params`` at a pickle written by the un-vendored ``sde4mbrl`` package
(launch/iris_sitl_traj_mpc.yaml:3).  Those pickles hold Haiku parameter pytrees of JAX arrays and need JAX to
unpickle, which this image does not have — so the conversion is split in two:

  1. on a machine WITH the upstream environment, dump the pytree to a flat ``.npz`` of NumPy arrays:

         import pickle, numpy as np, jax
         params = pickle.load(open("iris_sitl_sde.pkl", "rb"))
         flat = {"/".join(map(str, k)): np.asarray(v) for k, v in jax.tree_util.tree_flatten_with_path(params)[0]}
         np.savez("iris_sitl_sde_flat.npz", **flat)

  2. here, map the flat arrays onto the network layout this solver implements (SURVEY.md section 8a [SPEC]:
     two MLPs with two tanh hidden layers each, inputs [R(q)^T v, w, u], 6 outputs) plus the rigid-body
     constants, with an explicit ``--map`` file because the upstream parameter names are not pinned anywhere in
     the reference tree:

         python tools/convert_haiku_params.py iris_sitl_sde_flat.npz iris_sitl_sde.npz --vehicle iris --map map.json

     ``map.json``: {"drift_W1": "<flat key>", "drift_b1": ..., "drift_W2": ..., "drift_b2": ..., "drift_W3": ...,
     "drift_b3": ..., "diff_W1": ..., ..., "transpose": true|false, "mass": 1.5, "k_thrust": ..., "inertia": [...],
     "sigma_prior": [...]}.  Haiku stores Linear weights as [in, out]; this format is [out, in] (``transpose``).

Models whose architecture differs from the [SPEC] (other features, other activation, density-network diffusion)
cannot be represented and are rejected with a message rather than silently approximated.
"""
import argparse
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sde4mbrl_px4_b200 import model_io  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("flat_npz")
    ap.add_argument("out_npz")
    ap.add_argument("--vehicle", default="iris", choices=sorted(model_io.VEHICLES))
    ap.add_argument("--map", required=True, help="JSON file mapping this format's tensor names to keys of flat_npz")
    a = ap.parse_args()
    flat = np.load(a.flat_npz)
    mp = json.load(open(a.map))
    base = model_io.synthetic_model(a.vehicle)          # rigid-body defaults; overridden by the map where given
    tr = bool(mp.get("transpose", True))
    weights = {}
    for net in model_io.NETS:
        for lay in model_io.LAYERS:
            key = mp.get(f"{net}_{lay}")
            if key is None or key not in flat:
                raise SystemExit(f"map entry {net}_{lay} -> {key!r} not found in {a.flat_npz}; keys: {list(flat.keys())[:20]} ...")
            w = np.asarray(flat[key], np.float32)
            if lay.startswith("W") and tr:
                w = w.T
            weights[f"{net}_{lay}"] = np.ascontiguousarray(w)
    width = weights["drift_W1"].shape[0]
    m = model_io.SDEModel(
        nu=base.nu, width=width, mass=float(mp.get("mass", base.mass)), gravity=float(mp.get("gravity", base.gravity)),
        k_thrust=float(mp.get("k_thrust", base.k_thrust)), inertia=np.asarray(mp.get("inertia", base.inertia), np.float32),
        mixer=np.asarray(mp.get("mixer", base.mixer), np.float32),
        sigma_prior=np.asarray(mp.get("sigma_prior", base.sigma_prior), np.float32), weights=weights)
    shapes = m.layer_shapes()
    for k, v in weights.items():
        lay = k.split("_", 1)[1]
        if v.shape != shapes[lay]:
            raise SystemExit(f"{k}: shape {v.shape} does not fit the [SPEC] architecture {shapes[lay]} (inputs 6+nu, 2 hidden layers, 6 outputs)")
    m.to_blob()
    m.save(a.out_npz)
    print(f"wrote {a.out_npz}: nu={m.nu} width={m.width}")


if __name__ == "__main__":
    main()
