"""On-device policy rollout: the loop of sim2real/train.py:537-549 (obs -> policy -> sample -> env.step -> store)
for all environments at once, with nothing leaving HBM.

Per step: one `k_mlp` launch (tcgen05 policy forward + sampling, writing action/logp/value straight into the
[T, N, .] rollout buffers) and one `k_step` launch (fused env step writing obs/reward/terminated/truncated straight
into the buffers). The whole horizon can be captured in one CUDA graph (`use_graph=True`): 2*T+2 kernels per replay,
no host work between them. GAE and advantage normalisation follow (`opendog_b200.policy.gae`), with the advantage
statistics all-reduced over NCCL when environments are sharded across GPUs.
"""
from __future__ import annotations

import torch

from .policy import gae


class Rollout:
    def __init__(self, env, policy, horizon: int = 24, gamma: float = 0.99, lam: float = 0.95, use_graph: bool = True,
                 first_row_id: int = 0):
        self.env, self.policy, self.T = env, policy, int(horizon)
        self.gamma, self.lam = gamma, lam
        N, S, A, dev = env.num_envs, env.obs_dim, env.act_dim, env.device
        assert policy.state_dim == S and policy.action_dim == A
        T = self.T
        self.first_row_id = first_row_id
        self.obs = torch.zeros(T + 1, N, S, device=dev)
        self.action = torch.zeros(T, N, A, device=dev)
        self.mean = torch.zeros(N, A, device=dev)
        self.logp = torch.zeros(T, N, device=dev)
        self.value = torch.zeros(T + 1, N, device=dev)
        self.reward = torch.zeros(T, N, device=dev)
        self.terminated = torch.zeros(T, N, dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros(T, N, dtype=torch.uint8, device=dev)
        self.done = torch.zeros(T, N, dtype=torch.uint8, device=dev)
        self.step_base = torch.zeros(1, dtype=torch.int32, device=dev)      # Philox step counter of the sampler
        self.graph = None
        self.use_graph = use_graph
        self.obs[0].copy_(env.reset())
        self._started = False              # row 0 holds the reset observation until the first horizon has been collected
        self._warm = False

    def _body(self):
        T, pol, env = self.T, self.policy, self.env
        for t in range(T):
            pol.act(self.obs[t], sample=True, step=t, step_base=self.step_base, first_row_id=self.first_row_id,
                    out=dict(mean=self.mean, value=self.value[t], action=self.action[t], logp=self.logp[t]))
            env.step_into(self.action[t], self.obs[t + 1], self.reward[t], self.terminated[t], self.truncated[t])
        pol.act(self.obs[T], sample=False, step=T, out=dict(mean=self.mean, value=self.value[T]))    # bootstrap value
        torch.bitwise_or(self.terminated, self.truncated, out=self.done)
        self.step_base.add_(T + 1)

    def collect(self):
        """One horizon of experience into the buffers. Returns self (buffers are [T(+1), N, ...]).

        The first call starts from the `env.reset()` observation stored by the constructor; every later call carries
        the last observation of the previous horizon over into row 0. The buffers a call returns are exactly what
        its own horizon produced: obs[t] is the observation action[t] / logp[t] / value[t] were computed from."""
        dev = self.env.device
        torch.cuda.nvtx.range_push("odg.rollout.collect")          # timeline ranges for nsys / ncu --nvtx (profiling)
        try:
            return self._collect(dev)
        finally:
            torch.cuda.nvtx.range_pop()

    def _collect(self, dev):
        if self._started:
            self.obs[0].copy_(self.obs[self.T])
        self._started = True
        if not self.use_graph:
            self._body()
        elif self.graph is not None:
            self.graph.replay()
        elif not self._warm:
            # first horizon: eager, on a side stream — real data, and at the same time the warm-up a capture needs
            s = torch.cuda.Stream(device=dev)
            s.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(s):
                self._body()
            torch.cuda.current_stream(dev).wait_stream(s)
            self._warm = True
        else:
            # second horizon: capture (records only, executes nothing), then run it as the first replay
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()
            self.graph.replay()
        return self

    def advantages(self, normalize: bool = True, group=None):
        """GAE over the collected horizon (sim2real/train.py:557-564). Returns (adv [T,N], returns [T,N], stats)."""
        with torch.cuda.nvtx.range("odg.rollout.gae"):
            return gae(self.reward, self.value, self.done, self.gamma, self.lam, normalize=normalize, group=group)

    @property
    def kernels_per_collect(self) -> int:
        return 2 * self.T + 1
