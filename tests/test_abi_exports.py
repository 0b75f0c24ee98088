"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/plen_b200.h declares; the ctypes
mirrors have the C struct sizes; without a GPU the product path fails loudly (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from plen_ml_walk_b200 import build
    build.build()
    from plen_ml_walk_b200 import _abi
    return _abi.load_library()


def test_every_declared_symbol_is_exported(lib):
    from plen_ml_walk_b200 import _abi
    hdr = open(os.path.join(ROOT, "include", "plen_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(plen_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.plen_version().decode().startswith("plen_b200")


def test_struct_mirrors_and_defaults(lib):
    from plen_ml_walk_b200 import _abi
    cfg = _abi.PlenConfigC()
    assert lib.plen_default_config(C.byref(cfg), 0) == 0
    assert abs(cfg.dt - 1 / 240) < 1e-9 and cfg.substeps == 4 and cfg.reset_ticks == 8          # plen_env.py:40-42, :569
    assert abs(cfg.motor_max_force - 0.15) < 1e-7 and cfg.max_episode_steps == 500               # :753, :15-19
    assert abs(cfg.mu_rolling - 0.08) < 1e-7 and cfg.linear_damping == 0.0                       # :439-481
    assert list(cfg.env_lo)[:3] == [-1.57, -0.15, -0.95] and list(cfg.env_hi)[-1] == 0.35        # :148-167
    cj = _abi.PlenConfigC()
    lib.plen_default_config(C.byref(cj), 1)
    assert cj.joint_act == 1 and abs(cj.mu_rolling - 0.008) < 1e-7 and abs(cj.linear_damping - 0.1) < 1e-7
    # the emulation harness compiles the same header: sizes must agree with the ctypes mirrors
    from emu_util import Emu
    e = Emu()
    assert e.L.emu_sizeof_config() == C.sizeof(_abi.PlenConfigC) and e.L.emu_sizeof_model() == C.sizeof(_abi.PlenModelC)


def test_sole_manifold_needs_hull_vertices(lib):
    """plen_create checks its arguments before it looks for a device: sole_manifold = 1 on a model without hull vertices (or
    with a negative tie tolerance) is refused with a message, not silently run as the four-corner path."""
    from plen_ml_walk_b200 import _abi
    from plen_ml_walk_b200.urdf_loader import packaged_model
    cfg = _abi.PlenConfigC()
    lib.plen_default_config(C.byref(cfg), 0)
    assert cfg.sole_manifold == 0 and abs(cfg.support_tie - 1e-7) < 1e-12
    model = _abi.model_to_c(packaged_model())
    assert model.n_hull[0] == 209 and model.n_hull[1] == 209          # plen.urdf feet: the STL hulls
    cfg.sole_manifold = 1
    model.n_hull[1] = 0
    assert not lib.plen_create(C.byref(cfg), C.byref(model), 4, 0)
    assert b"hull vertices" in lib.plen_last_error(None)
    model.n_hull[1] = 209
    cfg.support_tie = -1.0
    assert not lib.plen_create(C.byref(cfg), C.byref(model), 4, 0)
    assert b"support_tie" in lib.plen_last_error(None)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from plen_ml_walk_b200 import _abi
    from plen_ml_walk_b200.urdf_loader import packaged_model
    from plen_ml_walk_b200.vec_env import PlenVecEnv
    cfg = _abi.PlenConfigC()
    lib.plen_default_config(C.byref(cfg), 0)
    model = _abi.model_to_c(packaged_model())
    assert not lib.plen_create(C.byref(cfg), C.byref(model), 4, 0)          # NULL: no usable CUDA device
    assert b"no CPU fallback" in lib.plen_last_error(None)
    with pytest.raises(RuntimeError):
        PlenVecEnv(4)
    from plen_ml_walk_b200.gait import gait_trajectories
    with pytest.raises(RuntimeError):
        gait_trajectories([[30, 30, 10, 5, 10]])


def test_host_surface_matches_reference_tables():
    """action / observation spaces and joint ranges of PlenWalkEnv (plen_env.py:141-263) without touching the GPU."""
    import numpy as np
    from plen_ml_walk_b200 import vec_env
    a, o = vec_env._spaces()
    assert a.shape == (18,) and o.shape == (26,) and a.high[0] == 1.0 and a.low[0] == -1.0
    assert o.high[18] == 0.25 and o.low[24] == 0 and o.high[25] == 1 and np.isinf(o.high[19])
    assert len(vec_env.ENV_RANGES) == 18 and vec_env.REAL_RANGES[3] == [-1.0, 1.57] and vec_env.ENV_RANGES[3] == [-0.9, 0.3]
    s = a.sample()
    assert s.shape == (18,) and s.dtype == np.float32 and (abs(s) <= 1).all()
    ref = "/root/reference/plen_bullet/src/plen_bullet/plen_env.py"
    if os.path.exists(ref):        # authoring container only: the literal tables still match the reference source
        src = open(ref).read()
        nums = lambda blk: [float(x) for x in re.findall(r"-?\d+\.\d+", blk)]
        env_blk = src[src.index("self.env_ranges = ["):src.index("self.real_ranges = [")]
        real_blk = src[src.index("self.real_ranges = ["):src.index("self.real_ranges = [") + 1200]
        assert nums(env_blk)[:36] == [v for r in vec_env.ENV_RANGES for v in r]
        assert nums(real_blk)[:36] == [v for r in vec_env.REAL_RANGES for v in r]
