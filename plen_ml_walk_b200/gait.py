"""Batched sinewave gait + closed-form leg IK -- host mirror of the reference's TrajectoryGenerator
(plen_bullet/src/plen_bullet/trajectory_generator.py:9-277) and of the trajectory assembly in
plen_bullet/src/trajectory_eval.py:154-261, computed by the CUDA kernel behind `plen_gait_ik` (float64 on device).

    gen = TrajectoryGenerator(height=h[N], stride=s[N], body_sway=w[N], fwd_bias=b[N], device="cuda:0")
    gen.main()                       # one launch for all N parameter sets
    gen.cycle      [N,40,18]         # one gait cycle in action order (20 right-forward + 20 left-forward rows)
    gen.bend_legs  [N,18]            # trajectory_eval.py:251-261
    gen.foot_walk_rfwd / foot_walk_lfwd [N,20,12], gen.bend [N,3,12]   # the reference's attribute names / layouts
    gen.status     [N] uint8         # 1 where the reference would raise ValueError (math.acos domain, :187-189)

Unlike the reference this needs no env instance: the generator only ever read `env.real_ranges`
(trajectory_generator.py:52, :203-222).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi

DEFAULTS = dict(height=30.0, stride=30.0, bend_distance=10.0, body_sway=5.0, fwd_bias=10.0)   # :10-18


def gait_trajectories(params, device="cuda:0"):
    """params [N,5] float64 (height, stride, bend_distance, body_sway, fwd_bias; mm) -> (cycle, bend_legs, status)."""
    if not torch.cuda.is_available():
        raise RuntimeError("plen_gait_ik needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device(device)
    p = torch.as_tensor(params, dtype=torch.float64, device=dev).reshape(-1, 5).contiguous()
    n = p.shape[0]
    cycle = torch.empty((n, 40, 18), dtype=torch.float64, device=dev)
    bend = torch.empty((n, 18), dtype=torch.float64, device=dev)
    status = torch.empty(n, dtype=torch.uint8, device=dev)
    lib = _abi.load_library()
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with torch.cuda.device(dev):
        rc = lib.plen_gait_ik(idx, C.c_void_p(p.data_ptr()), n, C.c_void_p(cycle.data_ptr()), C.c_void_p(bend.data_ptr()),
                              C.c_void_p(status.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError("plen_gait_ik: %s" % lib.plen_last_error(None).decode())
    return cycle, bend, status


class TrajectoryGenerator:
    def __init__(self, num_DoubleSupport=5, num_SingleSupport=10, height=30.0, stride=30.0, bend_distance=10.0,
                 body_sway=5.0, fwd_bias=10.0, sway_steps=5, device="cuda:0"):
        if num_DoubleSupport != 5 or num_SingleSupport != 10:
            raise NotImplementedError("the kernel implements the reference's 5 + 10 + 5 step schedule (defaults, :10-11)")
        cols = [torch.as_tensor(v, dtype=torch.float64).reshape(-1) for v in (height, stride, bend_distance, body_sway, fwd_bias)]
        n = max(c.numel() for c in cols)
        self.params = torch.stack([c.expand(n) if c.numel() == 1 else c for c in cols], 1)
        self.device = device
        self.l_hip_knee, self.l_knee_foot = 25.0, 40.0                # :35-37

    def main(self):
        self.cycle, self.bend_legs, self.status = gait_trajectories(self.params, self.device)
        c = self.cycle
        # undo the sign map of trajectory_eval.py:180-205 to expose the reference's 12-column IK tables
        r = torch.cat([-c[..., 0:4], c[..., 4:6], c[..., 6:10], -c[..., 10:11], c[..., 11:12]], -1)
        self.foot_walk_rfwd, self.foot_walk_lfwd = r[:, :20], r[:, 20:]
        b = self.bend_legs
        row = torch.cat([-b[:, 0:4], b[:, 4:10], -b[:, 10:11], b[:, 11:12]], -1)
        self.bend = row[:, None, :].expand(-1, 3, -1)
        return self

    def full_trajectory(self, n_cycles=20):
        """[N, 40 n_cycles, 18]: the open-loop joint trajectory trajectory_eval.py replays (20 cycles = 800 rows)."""
        return self.cycle.repeat(1, n_cycles, 1)
