"""Sampling-MPC (MPPI) on the batched simulator: BASELINE.json configs[4] — `num_samples` action sequences of length
`horizon` rolled out from ONE shared robot state, costs reduced on the device (include/odg_mppi.h). The reference has
no MPC code; the cost is minus the reference's walk reward (WalkEnvironment.py:81-109, taken BEFORE its max(0, .)
clip, which would zero most of the signal) summed along the rollout.

Per plan(): broadcast the start state, ONE rollout launch (`odg_mppi_rollout`: every sample walks its whole horizon —
sample the action row, fused env step, accumulate the cost — without leaving the SM) and one reduction launch
(`odg_mppi_reduce`). The plan counter that keys the noise lives in device memory and is incremented on the stream, so
a captured CUDA graph draws fresh noise on every replay.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import lib as _lib
from .env import BatchedWalkEnv, _ptr


class MPPI:
    def __init__(self, num_samples: int = 1024, horizon: int = 64, sigma: float = 0.3, lam: float = 1.0,
                 termination_cost: float = 100.0, device=None, seed: int = 0, use_graph: bool = True, **sim_config):
        self.env = BatchedWalkEnv(num_samples, device=device, seed=seed, info_keys=None, auto_reset=0, **sim_config)
        self.L, self.dev = self.env.L, self.env.device
        self.N, self.T, self.A = num_samples, horizon, self.env.act_dim
        self.sigma, self.lam, self.term_cost, self.seed = sigma, lam, termination_cost, seed
        dev = self.dev
        self.mean = torch.zeros(horizon, self.A, device=dev)               # nominal sequence (actions in [-1, 1])
        self.new_mean = torch.zeros(horizon, self.A, device=dev)
        self.actions = torch.zeros(horizon, num_samples, self.A, device=dev)
        self.cost = torch.zeros(num_samples, device=dev)
        self.stats = torch.zeros(4, device=dev)
        # shared start state, broadcast to every sample at the start of a plan
        self.q0 = torch.zeros(num_samples, self.env.nq, device=dev)
        self.v0 = torch.zeros(num_samples, self.env.nv, device=dev)
        self.env_state0 = None
        self.iteration = 0
        self.iteration_dev = torch.zeros(1, dtype=torch.int32, device=dev)     # plan counter keying the noise (Philox)
        self.use_graph, self.graph = use_graph, None

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def set_start(self, qpos, qvel, last_action=None, desired_velocity=None):
        """The shared robot state: qpos [nq], qvel [nv] (+ the env-side state the walk reward reads)."""
        self.q0.copy_(torch.as_tensor(qpos, dtype=torch.float32).to(self.dev).reshape(1, -1).expand(self.N, -1))
        self.v0.copy_(torch.as_tensor(qvel, dtype=torch.float32).to(self.dev).reshape(1, -1).expand(self.N, -1))
        la = torch.zeros(self.A) if last_action is None else torch.as_tensor(last_action, dtype=torch.float32)
        dv = torch.tensor([0.75, 0.0, 0.0]) if desired_velocity is None else torch.as_tensor(desired_velocity, dtype=torch.float32)
        z = torch.zeros(self.N, dtype=torch.int32)
        self.env_state0 = dict(step=z.to(self.dev), gait_index=z.to(self.dev), gait_matches=z.to(self.dev),
                               last_action=la.reshape(1, -1).expand(self.N, -1).contiguous().to(self.dev),
                               desired_velocity=dv.reshape(1, -1).expand(self.N, -1).contiguous().to(self.dev),
                               fresh=torch.zeros(self.N, dtype=torch.uint8, device=self.dev))

    def _body(self):
        e, L, st = self.env, self.L, self._st()
        _lib.check(L.odg_set_state(e._h, _ptr(self.q0), _ptr(self.v0), None, st), "odg_set_state")
        s0 = self.env_state0
        _lib.check(L.odg_set_env_state(e._h, _ptr(s0["step"]), _ptr(s0["gait_index"]), _ptr(s0["gait_matches"]),
                                       _ptr(s0["last_action"]), _ptr(s0["desired_velocity"]), _ptr(s0["fresh"]), st),
                   "odg_set_env_state")
        _lib.check(L.odg_mppi_rollout(e._h, _ptr(self.mean), self.sigma, self.T, C.c_uint64(self.seed), 0, _ptr(self.iteration_dev),
                                      self.term_cost, _ptr(self.actions), _ptr(self.cost), st), "odg_mppi_rollout")
        _lib.check(L.odg_mppi_reduce(_ptr(self.cost), _ptr(self.actions), self.T, self.N, self.A, self.lam,
                                     _ptr(self.new_mean), _ptr(self.stats), st), "odg_mppi_reduce")
        self.iteration_dev.add_(1)

    def plan(self, update_mean: bool = True):
        """One MPPI iteration from the shared start state; returns the updated nominal sequence [T, A]."""
        assert self.env_state0 is not None, "call set_start first"
        if not self.use_graph:
            self._body()
        elif self.graph is not None:
            self.graph.replay()
        elif self.iteration == 0:
            self._body()                 # first plan runs eagerly: it is also the warm-up the capture needs
        else:
            torch.cuda.synchronize(self.dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()
            self.graph.replay()          # capture only records: run it once
        self.iteration += 1
        if update_mean:
            self.mean.copy_(self.new_mean)
        return self.new_mean

    @property
    def kernels_per_plan(self) -> int:
        return 9          # 2 + 5 state-broadcast transposes / copies, the rollout, the reduction, the counter
