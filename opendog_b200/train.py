"""Multi-GPU training plumbing around the on-device rollout: how environments shard over ranks and the only two
collectives of the whole job (BASELINE.json north_star: "NCCL over NVLink is used only to allreduce rollout and
advantage statistics and policy gradients").

Environments are independent (the reference runs one per OS process: train/train.py:81-86), so rank r owns the
global env ids [r*N, (r+1)*N) and `step`/`reset` need no communication; RNG streams are keyed by GLOBAL env id, so
results do not depend on the number of ranks. Collectives:
  * advantage statistics [sum, sumsq, count] (3 x f64) — the reference normalises over the whole batch
    (sim2real/train.py:564), so every rank must use the job-wide mean/std;
  * the flat policy gradient (one buffer, ~0.3 M f32 for 33-512-256-8 actor+critic), averaged over ranks.
The host logic here is device-agnostic torch (so the world_size-2 `gloo` tests in tests/ exercise it on CPU); the
product path runs it on CUDA tensors over NCCL.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(rank: int, world: int, envs_per_rank: int):
    """Global env ids owned by `rank`: [first, first + envs_per_rank)."""
    first = rank * envs_per_rank
    return first, first + envs_per_rank


def _on(group):
    return dist.is_available() and dist.is_initialized() and (group is not False)


def allreduce_advantage_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of [sum(adv), sum(adv^2), count] over ranks."""
    if _on(group):
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group if group is not True else None)
    return stats


def stats_mean_std(stats: torch.Tensor):
    """(mean, unbiased std) from [sum, sumsq, count] — the quantities of `(adv - adv.mean()) / (adv.std() + 1e-8)`."""
    n = stats[2]
    mean = stats[0] / n
    var = torch.clamp((stats[1] - n * mean * mean) / torch.clamp(n - 1, min=1), min=0)
    return mean, torch.sqrt(var)


def allreduce_flat_grads(params, group=None, world: int | None = None):
    """Average the gradients of `params` over ranks with ONE all-reduce of a flat buffer. Returns the buffer."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return None
    flat = torch.cat([g.reshape(-1) for g in grads])
    if _on(group):
        g = group if group is not True else None
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=g)
        flat /= (world or dist.get_world_size(g))
    off = 0
    for gr in grads:
        n = gr.numel()
        gr.copy_(flat[off:off + n].view_as(gr))
        off += n
    return flat


def _loss_terms(policy, d, v, action, logp_old, adv, ret, clip, vf_coef, ent_coef):
    """Clipped-surrogate loss terms of one minibatch from the policy's (Normal, value) output: the fused kernel
    (policy.ppo_loss) when the policy exposes its state-independent `action_log_std` (the reference's ActorCritic,
    sim2real/train.py:132-149), the torch expression on the distribution otherwise."""
    ls = getattr(policy, "action_log_std", None)
    if ls is not None and d.loc.is_cuda and d.loc.dim() == 2:
        from .policy import ppo_loss
        return ppo_loss(d.loc, v, ls, action, logp_old, adv, ret, clip, vf_coef, ent_coef)
    logp = d.log_prob(action).float().sum(-1)
    ratio = torch.exp(logp - logp_old)
    pg = -torch.min(ratio * adv, torch.clamp(ratio, 1 - clip, 1 + clip) * adv).mean()
    vf = torch.nn.functional.mse_loss(v.float().squeeze(-1), ret)
    ent = d.entropy().float().sum(-1).mean()
    return pg + vf_coef * vf - ent_coef * ent, pg, vf, ent


def ppo_update(policy, optimizer, obs, action, logp_old, adv, ret, clip: float = 0.2, vf_coef: float = 0.5,
               ent_coef: float = 0.005, max_grad_norm: float = 0.5, epochs: int = 1, minibatches: int = 4, group=None,
               autocast: bool = True, timing: dict | None = None):
    """Clipped-surrogate update with the hyper-parameters of train/train.py:117-130 (clip .2, ent .005,
    max_grad_norm .5). obs [B,S], action [B,A], logp_old/adv/ret [B]; minibatches are fixed contiguous slices.
    Gradients are averaged over ranks (one flat all-reduce per minibatch); afterwards the tensor-core copy of the
    weights is refreshed. Returns the last minibatch's loss terms.

    `autocast` runs the differentiable forward / backward with bf16 operands on the tensor cores (fp32 master weights,
    fp32 loss and optimiser) — the same operand precision as the rollout-time kernel that produced `logp_old`, so the
    probability ratio compares like with like. `timing`, if given, receives CUDA events around every gradient
    all-reduce (`timing["allreduce"]` = list of (start, end) pairs) so a caller can report the collective's share."""
    # without autocast the GEMMs ([B, 33..512] x [512, 256] ...) run as TF32 (fp32 accumulate); the fp32 CUDA-core path is
    # ~5x slower on B200 and the reference's own update is plain fp32 torch on whatever device it finds
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    B = obs.shape[0]
    mb = (B + minibatches - 1) // minibatches
    params = [p for p in policy.parameters() if p.requires_grad]
    use_ac = bool(autocast) and obs.is_cuda
    padded = use_ac and hasattr(policy, "forward_padded")
    if padded:
        obs = policy.pad_obs(obs.to(torch.bfloat16))          # one cast + pad per update instead of one per minibatch
    out = {}
    for _ in range(epochs):
        for i in range(minibatches):
            sl = slice(i * mb, min(B, (i + 1) * mb))
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=use_ac):
                d, v = policy.forward_padded(obs[sl]) if padded else policy(obs[sl])
            loss, pg, vf, ent = _loss_terms(policy, d, v, action[sl], logp_old[sl], adv[sl], ret[sl], clip, vf_coef, ent_coef)
            optimizer.zero_grad(set_to_none=True)
            loss.backward()
            if timing is not None and obs.is_cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                allreduce_flat_grads(params, group=group)
                e1.record()
                timing.setdefault("allreduce", []).append((e0, e1))
            else:
                allreduce_flat_grads(params, group=group)
            torch.nn.utils.clip_grad_norm_(params, max_grad_norm)
            optimizer.step()
            out = dict(loss=loss.detach(), pg=pg.detach(), vf=vf.detach(), entropy=ent.detach())
    if hasattr(policy, "sync_weights"):
        policy.sync_weights()
    return out


class GraphedPPOUpdate:
    """`ppo_update` captured once as a CUDA graph per minibatch and replayed: same arithmetic, same hyper-parameters, no
    Python / launch / allocator time between the ~150 small kernels of a minibatch (forward, loss, backward, gradient
    all-reduce, clipping, fused Adam). The inputs are copied into static buffers of the captured shape
    (B = rows per update, fixed); the flat gradient all-reduce is captured with the rest (NCCL supports capture).
    Falls back to the eager `ppo_update` if capture is impossible (e.g. a CPU run of the gloo tests)."""

    def __init__(self, policy, optimizer, B: int, obs_dim: int, act_dim: int, minibatches: int = 4, clip: float = 0.2,
                 vf_coef: float = 0.5, ent_coef: float = 0.005, max_grad_norm: float = 0.5, group=None):
        if not all(g.get("capturable", False) for g in optimizer.param_groups):
            raise ValueError("GraphedPPOUpdate needs an optimizer created with capturable=True (its step is captured in a CUDA graph)")
        self.policy, self.opt, self.B, self.mbs, self.group = policy, optimizer, int(B), int(minibatches), group
        self.hp = dict(clip=clip, vf_coef=vf_coef, ent_coef=ent_coef, max_grad_norm=max_grad_norm)
        dev = next(policy.parameters()).device
        self.dev = dev
        self.params = [p for p in policy.parameters() if p.requires_grad]
        self.padded = hasattr(policy, "forward_padded")
        k = (-obs_dim) % 16 if self.padded else 0
        self.obs = torch.zeros(B, obs_dim + k, device=dev, dtype=torch.bfloat16 if self.padded else torch.float32)
        self.obs_dim = obs_dim
        self.action = torch.zeros(B, act_dim, device=dev)
        self.logp_old, self.adv, self.ret = (torch.zeros(B, device=dev) for _ in range(3))
        self.out = {k_: torch.zeros((), device=dev) for k_ in ("loss", "pg", "vf", "entropy")}
        self.graphs = None
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True

    def _minibatch(self, i):
        mb = (self.B + self.mbs - 1) // self.mbs
        sl = slice(i * mb, min(self.B, (i + 1) * mb))
        hp = self.hp
        with torch.autocast("cuda", dtype=torch.bfloat16):
            d, v = self.policy.forward_padded(self.obs[sl]) if self.padded else self.policy(self.obs[sl])
        loss, pg, vf, ent = _loss_terms(self.policy, d, v, self.action[sl], self.logp_old[sl], self.adv[sl], self.ret[sl],
                                        hp["clip"], hp["vf_coef"], hp["ent_coef"])
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        allreduce_flat_grads(self.params, group=self.group)
        torch.nn.utils.clip_grad_norm_(self.params, hp["max_grad_norm"], foreach=True)
        self.opt.step()
        for k_, t in (("loss", loss), ("pg", pg), ("vf", vf), ("entropy", ent)):
            self.out[k_].copy_(t.detach())

    def _capture(self):
        for p in self.params:                              # static gradient buffers
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):                         # warm-up outside capture (lazy inits, autotuning, optimizer state)
            for i in range(self.mbs):
                self._minibatch(i)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        graphs = []
        pool = None
        for i in range(self.mbs):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._minibatch(i)
            pool = g.pool()
            graphs.append(g)
        self.graphs = graphs

    def load(self, obs, action, logp_old, adv, ret):
        if self.padded:
            self.obs[:, :self.obs_dim].copy_(obs)          # fp32 -> bf16 cast + zero-padded columns stay zero
        else:
            self.obs.copy_(obs)
        self.action.copy_(action); self.logp_old.copy_(logp_old); self.adv.copy_(adv); self.ret.copy_(ret)

    def __call__(self, obs, action, logp_old, adv, ret, epochs: int = 1):
        torch.cuda.nvtx.range_push("odg.ppo_epoch")
        self.load(obs, action, logp_old, adv, ret)
        first = self.graphs is None
        if first:
            self._capture()                                # its eager warm-up pass is this call's first epoch
        for _ in range(epochs - 1 if first else epochs):
            for g in self.graphs:
                g.replay()
        if hasattr(self.policy, "sync_weights"):
            self.policy.sync_weights()
        torch.cuda.nvtx.range_pop()
        return self.out
