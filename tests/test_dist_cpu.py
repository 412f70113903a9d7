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


def test_gloo_two_rank_gather_to_rank0(tmp_path):
    """The collective inside sharding.solve_sharded (a true gather: rank 0 is the only receiver), on gloo."""
    script = tmp_path / "g.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        from sde4mbrl_px4_b200 import sharding
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        bufs = {{}}
        for rep in range(2):                       # second call reuses the cached receive buffers
            parts = {{"u": torch.arange(12, dtype=torch.float32) + 100 * r + rep, "info": torch.full((8,), float(r + rep))}}
            out = sharding.gather_to_rank0(parts, bufs)
            if r == 0:
                assert out["u"].shape == (w, 12) and out["info"].shape == (w, 8)
                for q in range(w):
                    assert torch.equal(out["u"][q], torch.arange(12, dtype=torch.float32) + 100 * q + rep)
                    assert torch.equal(out["info"][q], torch.full((8,), float(q + rep)))
            else:
                assert out is None
        if r == 0:
            print("GATHER0_OK")
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GATHER0_OK" in out.stdout


def test_gloo_two_rank_shared_memory_results(tmp_path):
    """sharding.SharedResults (the default final gather of solve_sharded): every rank writes its slice of a shared-memory
    array, rank 0 reads the concatenation; two generations alternate."""
    script = tmp_path / "s.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from sde4mbrl_px4_b200 import sharding
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()
        sr = sharding.SharedResults({{"u": (5, 3), "info": (5, 8)}})
        prev = None
        for rep in range(3):
            a = sr.next()
            a["u"][r] = 10 * r + rep
            a["info"][r] = r - rep
            dist.barrier()
            if r == 0:
                assert all(np.all(a["u"][q] == 10 * q + rep) and np.all(a["info"][q] == q - rep) for q in range(w))
                if prev is not None:      # the previous generation is still intact
                    assert all(np.all(prev["u"][q] == 10 * q + rep - 1) for q in range(w))
                prev = a
            dist.barrier()
        if r == 0:
            print("SHM_OK")
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29535", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "SHM_OK" in out.stdout


def test_gloo_two_rank_closed_loop_sharding(tmp_path):
    """BASELINE config 5 host logic at world size 2: contiguous blocks of rollouts per rank, one launch per rank,
    statistics gathered in rollout order on rank 0, device time = max over ranks.  The solver is a stand-in that
    records what it was asked to run (the CUDA path itself is covered by the -m gpu tests)."""
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from sde4mbrl_px4_b200 import sharding
        dist.init_process_group("gloo")
        r, w = dist.get_rank(), dist.get_world_size()

        class FakeSolver:
            def closed_loop(self, x0, t0, rng, ticks, want_hist=True):
                assert not want_hist and ticks == 7
                self.n = x0.shape[0]
                st = np.stack([x0[:, 0], t0, rng[:, 0].astype(np.float32), np.full(self.n, ticks, np.float32)], axis=1)
                return None, None, st.astype(np.float32)
            def last_launch_ms(self):
                return 100.0 * (r + 1)

        R = 11
        x0 = np.zeros((R, 13), np.float32); x0[:, 0] = np.arange(R)
        t0 = np.arange(R, dtype=np.float32) * 0.25
        rng = np.stack([1000 + np.arange(R), np.zeros(R)], axis=1).astype(np.uint64)
        fs = FakeSolver()
        out = sharding.closed_loop_sharded(fs, x0, t0, rng, 7)
        lo, hi = sharding.shard_range(R, r, w)
        assert fs.n == hi - lo
        if r == 0:
            st = out["stats"]
            assert st.shape == (R, 4)
            assert np.array_equal(st[:, 0], np.arange(R)) and np.array_equal(st[:, 1], t0)
            assert np.array_equal(st[:, 2], 1000 + np.arange(R)) and np.all(st[:, 3] == 7)
            assert abs(out["device_s"] - 0.1 * w) < 1e-6 and out["rollouts_per_rank"] == [6, 5]
            assert abs(out["ticks_per_s"] - R * 7 / (0.1 * w)) < 1e-3
            print("CLOSED_LOOP_OK")
        else:
            assert out is None
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "CLOSED_LOOP_OK" in out.stdout
