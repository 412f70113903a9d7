"""world_size-2 gloo test of the N>1 path: static sharding + the final result gather."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT


def test_gloo_two_rank_shard_and_gather(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from sde4mbrl_px4_b200 import sharding
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        B = 37
        full = dict(u=np.arange(B * 6, dtype=np.float32).reshape(B, 2, 3), c=np.arange(B, dtype=np.float32) * 0.5)
        loc = sharding.shard_problem(full, r, w)
        lo, hi = sharding.shard_range(B, r, w)
        assert loc["u"].shape[0] == hi - lo
        out = sharding.gather_results({{k: v * 2 for k, v in loc.items()}}, B)
        if r == 0:
            assert np.array_equal(out["u"], full["u"] * 2) and np.array_equal(out["c"], full["c"] * 2)
            print("GATHER_OK")
        else:
            assert out is None
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER_OK" in out.stdout
