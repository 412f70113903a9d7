"""Generates tests/golden/golden_v1.npz: seeded inputs and the ORACLE's outputs for a set of
small cases (u*, x_evol, telemetry at it in {1, 5, 20, 200}; value_and_grad).  No reference
implementation of this path exists to generate vectors from (parity unpinned, see oracle header),
so the fixtures pin SPEC-ARITH itself: both the oracle and the CUDA path must reproduce them bit
for bit.  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.dirname(HERE)):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = [
    dict(name="iris_traj_it1", vehicle="iris", mode="traj", P=1, max_iter=1, B=4, seed=1),
    dict(name="iris_traj_it5", vehicle="iris", mode="traj", P=1, max_iter=5, B=4, seed=1),
    dict(name="iris_traj_it20", vehicle="iris", mode="traj", P=1, max_iter=20, B=4, seed=1),
    dict(name="iris_traj_it200", vehicle="iris", mode="traj", P=1, max_iter=200, B=4, seed=1),
    dict(name="iris_pos_it100", vehicle="iris", mode="pos", P=1, max_iter=100, B=3, seed=2),
    dict(name="iris_traj_p8_it20", vehicle="iris", mode="traj", P=8, max_iter=20, B=2, seed=3),
    dict(name="hexa_traj_it20", vehicle="hexa", mode="traj", P=1, max_iter=20, B=2, seed=4),
    dict(name="hexa_traj_p8_it10", vehicle="hexa", mode="traj", P=8, max_iter=10, B=2, seed=5),
]


def run_case(case, backend="oracle"):
    from conftest import make_setup, random_states
    from sde4mbrl_px4_b200 import synthetic

    cfg, blob, _ = make_setup(case["vehicle"], case["mode"], num_particles=case["P"], max_iter=case["max_iter"],
                              rtol=0.0, atol=0.0)
    B = case["B"]
    pr = synthetic.batched_problems(B, cfg.horizon, np.array(cfg.dt[: cfg.horizon]), seed=case["seed"])
    if backend == "oracle":
        from oracle import oracle as O

        s = O.Oracle(cfg, blob, "f32")
    else:
        from sde4mbrl_px4_b200 import solver

        s = solver.MPCSolver(cfg, blob)
    u0, i0 = s.reset(B)
    kw = dict(xref_win=pr["xref_win"]) if case["mode"] == "traj" else dict(xdes=pr["xref_win"][:, 0])
    rng = np.random.default_rng(case["seed"])
    uq = np.clip(u0 + 0.05 * rng.standard_normal(u0.shape), 1e-4, 1).astype(np.float32)
    J, g, _ = s.rollout(pr["x"], uq, u0[:, 0], rng=pr["rng"], **kw)
    u, xe, info, _ = s.solve(pr["x"], u0, i0, rng=pr["rng"], **kw)
    # second, warm-started tick from the predicted next state
    rng2 = pr["rng"].copy()
    rng2[:, 1] += 1
    u2, xe2, info2, _ = s.solve(xe[:, 1], u, info, rng=rng2, **kw)
    return dict(cost=J, grad=g, u=u, x_evol=xe, info=info[:, :7], u2=u2, x_evol2=xe2, info2=info2[:, :7])


if __name__ == "__main__":
    out = {"meta": np.array(json.dumps({"spec_arith": 1, "cases": CASES}))}
    for case in CASES:
        for k, v in run_case(case).items():
            out[f"{case['name']}/{k}"] = v
        print(case["name"], "opt_cost", out[f"{case['name']}/info"][:, 6])
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), os.path.getsize(os.path.join(HERE, "golden_v1.npz")), "bytes")
