"""The C ABI: struct layouts match the header as compiled by gcc, the library loads without a GPU and
exports every symbol include/sdempc.h declares, host-only calls work, and compute calls fail loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT, make_setup
from sde4mbrl_px4_b200 import _abi, solver


def test_struct_sizes_match_header():
    src = r'''
    #include <stdio.h>
    #include "sdempc.h"
    int main(void){ printf("%zu %zu %zu %zu\n", sizeof(sdempc_config), sizeof(sdempc_model_header), sizeof(sdempc_info), sizeof(sdempc_solve_args)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        got = [int(v) for v in subprocess.check_output([exe]).split()]
    assert got == [C.sizeof(_abi.Config), C.sizeof(_abi.ModelHeader), C.sizeof(_abi.Info), C.sizeof(_abi.SolveArgs)]


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "sdempc.h")).read()
    declared = sorted(set(re.findall(r"\b(sdempc_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(_abi.HEADER_SYMBOLS)
    lib = _abi.load_library()
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.sdempc_version()


def test_host_only_calls_and_argument_errors(built):
    cfg, blob, _ = make_setup("iris", "traj")
    s = solver.MPCSolver(cfg, blob)          # host only: must not need a GPU
    u, info = s.reset(3)
    assert u.shape == (3, 20, 4) and np.allclose(u, 0.71) and np.allclose(info[:, 1], 0.01)
    with pytest.raises(RuntimeError, match="no trajectory"):
        s.state_from_traj(0.0)
    tab = np.zeros((3, 14), np.float32)
    tab[:, 0] = [0, 1, 1]
    tab[:, 7] = 1
    with pytest.raises(RuntimeError, match="strictly increasing"):
        s.set_trajectory(tab)
    tab[:, 0] = [0, 1, 2]
    tab[:, 1] = [0, 2, 4]
    s.set_trajectory(tab)
    assert np.allclose(s.state_from_traj(0.5)[0, 0], 1.0)
    # unsupported shape is rejected with a message, not a crash
    cfg2, _, _ = make_setup("iris", "traj", num_particles=3)
    with pytest.raises(RuntimeError, match="no compiled kernel"):
        solver.MPCSolver(cfg2, blob)
    bad = bytearray(blob)
    bad[0] ^= 0xFF
    with pytest.raises(RuntimeError, match="magic"):
        solver.MPCSolver(cfg, bytes(bad))


def test_empty_and_malformed_batches_are_rejected(built):
    """Argument errors are reported before any CUDA call (so this runs without a GPU)."""
    cfg, blob, _ = make_setup("iris", "traj")
    s = solver.MPCSolver(cfg, blob)
    e = np.zeros((0, 13), np.float32)
    with pytest.raises(RuntimeError, match="B must be >= 1"):
        s.solve(e, np.zeros((0, 20, 4), np.float32), np.zeros((0, 8), np.float32), xdes=e, rng=np.zeros((0, 2), np.uint64))
    x = np.zeros((2, 13), np.float32)
    x[:, 6] = 1
    u, info = s.reset(2)
    with pytest.raises(RuntimeError, match="one of xref_win, curr_t, xdes"):
        s.solve(x, u, info, rng=np.zeros((2, 2), np.uint64))
    with pytest.raises(RuntimeError, match="set_trajectory"):
        s.solve(x, u, info, curr_t=np.zeros(2, np.float32), rng=np.zeros((2, 2), np.uint64))
    with pytest.raises(RuntimeError, match="rng is required"):
        s.solve(x, u, info, xdes=x)


def test_compute_fails_loudly_without_gpu(built):
    """There is no CPU fallback: on a box without CUDA the solve raises instead of computing."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg, blob, _ = make_setup("iris", "pos")
    s = solver.MPCSolver(cfg, blob)
    u, info = s.reset(1)
    x = np.array([[0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]], np.float32)
    with pytest.raises(RuntimeError, match="no usable CUDA device|CUDA"):
        s.solve(x, u, info, xdes=x, rng=np.array([[1, 0]], np.uint64))


def test_product_does_not_import_the_oracle():
    """The product package never references oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "sde4mbrl_px4_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt and "oracle/" not in txt.replace("oracle/det_math.h", ""), f
