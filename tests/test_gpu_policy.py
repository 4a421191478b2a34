"""GPU tests of the rollout kernels (through the C ABI): tcgen05 policy MLP vs a torch reference of the same
ActorCritic (sim2real/train.py:132-149), sampling / log-prob vs a numpy restatement of the Philox stream, GAE vs
the reference's Python loop (sim2real/train.py:557-564).

Tolerances: the kernel computes with bf16 operands and fp32 accumulation. Against a torch model that rounds inputs,
weights and hidden activations to bf16 the outputs agree to 4e-3 (tanh.approx + accumulation order); against the
plain fp32 torch model to 3e-2 (bf16 quantisation of a 512-wide layer)."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _ref_forward(m, obs, emulate_bf16):
    def run(net, x, last_tanh):
        for li in (0, 2, 4):
            w, b = net[li].weight.detach(), net[li].bias.detach()
            if emulate_bf16:
                x = _bf16(x) @ _bf16(w).T + b
            else:
                x = x @ w.T + b
            if li != 4 or last_tanh:
                x = torch.tanh(x)
        return x
    return run(m.actor, obs, True), run(m.critic, obs, False)[:, 0]


@pytest.mark.parametrize("S,A,N", [(33, 8, 300), (22, 4, 128), (48, 12, 1000), (64, 16, 129)])
def test_policy_mlp_matches_torch(S, A, N):
    from opendog_b200.policy import ActorCriticB200
    torch.manual_seed(S * 100 + A)
    m = ActorCriticB200(S, A, 0.4, seed=7)
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 2 and p.shape[0] > 16:
                p.mul_(2.0)                      # push hidden units into the non-linear range of tanh
        m.action_log_std.copy_(torch.linspace(-1.5, -0.5, A)[None])
    m.sync_weights()
    obs = torch.randn(N, S, device="cuda") * 1.5
    action, logp, value, mean = m.act(obs, step=3, first_row_id=1000)
    torch.cuda.synchronize()
    r_mean, r_val = _ref_forward(m, obs, True)
    assert (mean - r_mean).abs().max().item() < 4e-3
    assert (value - r_val).abs().max().item() < 4e-3 * max(1.0, r_val.abs().max().item())
    f_mean, f_val = _ref_forward(m, obs, False)
    assert (mean - f_mean).abs().max().item() < 3e-2
    assert (value - f_val).abs().max().item() < 3e-2 * max(1.0, f_val.abs().max().item())
    # sampling: action = mean + exp(log_std) * eps with eps from Philox (seed, first_row + row, step, block)
    from oracle.oracle import philox
    ls = m.action_log_std.detach().cpu().numpy().reshape(-1)
    act = action.cpu().numpy(); mu = mean.cpu().numpy(); lp = logp.cpu().numpy()
    for row in (0, 1, N // 2, N - 1):
        eps = []
        for blk in range((A + 3) // 4):
            r = philox(7, 1000 + row, 3, blk, 0x53414d50)
            for pr in range(2):
                u1 = (np.float32(r[2 * pr] >> 8) + np.float32(1.0)) * np.float32(2.0 ** -24)
                u2 = np.float32(r[2 * pr + 1] >> 8) * np.float32(2.0 ** -24)
                rad = math.sqrt(-2.0 * math.log(u1))
                eps += [rad * math.cos(2 * math.pi * u2), rad * math.sin(2 * math.pi * u2)]
        eps = np.array(eps[:A])
        assert np.abs(act[row] - (mu[row] + np.exp(ls) * eps)).max() < 1e-4
        assert abs(lp[row] - np.sum(-0.5 * eps ** 2 - ls - 0.5 * math.log(2 * math.pi))) < 1e-3
    # the eps of a whole batch look standard normal
    z = ((action - mean) / torch.exp(m.action_log_std)).flatten()
    assert abs(z.mean().item()) < 0.15 and abs(z.std().item() - 1.0) < 0.15
    # deterministic path
    a2, lp2, v2, mean2 = m.act(obs, sample=False)
    assert lp2 is None and torch.equal(a2, mean2) and torch.equal(mean2, mean)


def test_gae_matches_reference_loop():
    from opendog_b200.policy import gae
    T, N = 24, 777
    g = torch.Generator().manual_seed(0)
    rew = torch.randn(T, N, generator=g); val = torch.randn(T + 1, N, generator=g)
    done = torch.rand(T, N, generator=g) < 0.1
    adv, ret, stats = gae(rew.cuda(), val.cuda(), done.cuda(), 0.99, 0.95, normalize=False, group=False)
    # sim2real/train.py:557-561 per environment
    r64, v64, m64 = rew.double().numpy(), val.double().numpy(), (~done).double().numpy()
    ref = np.zeros((T, N)); a = np.zeros(N)
    for t in reversed(range(T)):
        delta = r64[t] + 0.99 * v64[t + 1] * m64[t] - v64[t]
        a = delta + 0.99 * 0.95 * m64[t] * a
        ref[t] = a
    assert np.abs(adv.cpu().numpy() - ref).max() < 1e-4
    assert np.abs(ret.cpu().numpy() - (ref + v64[:T])).max() < 1e-4
    s = stats.cpu().numpy()
    a32 = adv.cpu().double().numpy()
    assert abs(s[0] - a32.sum()) < 1e-6 * max(1, abs(a32.sum())) + 1e-6 and abs(s[1] - (a32 ** 2).sum()) < 1e-6 * (a32 ** 2).sum()
    assert s[2] == T * N
    adv_n, _, _ = gae(rew.cuda(), val.cuda(), done.cuda(), 0.99, 0.95, normalize=True, group=False)
    ref_n = (torch.from_numpy(ref).float() - torch.from_numpy(ref).float().mean()) / (torch.from_numpy(ref).float().std() + 1e-8)
    assert (adv_n.cpu() - ref_n).abs().max().item() < 1e-4


def test_rollout_graph_equals_eager_and_feeds_gae():
    """One horizon collected through the CUDA graph equals the same horizon collected launch by launch."""
    from opendog_b200.env import BatchedWalkEnv
    from opendog_b200.policy import ActorCriticB200
    from opendog_b200.rollout import Rollout
    outs = []
    for use_graph in (False, True):
        torch.manual_seed(0)
        env = BatchedWalkEnv(96, seed=5, info_keys=None, max_episode_steps=30)
        pol = ActorCriticB200(env.obs_dim, env.act_dim, 0.4, seed=11)
        ro = Rollout(env, pol, horizon=12, use_graph=use_graph)
        for _ in range(3):                       # the graph path warms up once, captures once, then replays
            ro.collect()
        torch.cuda.synchronize()
        outs.append([x.clone() for x in (ro.obs, ro.action, ro.logp, ro.value, ro.reward, ro.done)])
        adv, ret, stats = ro.advantages(normalize=True, group=False)
        assert torch.isfinite(adv).all() and abs(adv.mean().item()) < 1e-3 and abs(adv.std().item() - 1) < 1e-3
    # the eager run did 3 collects; the graph run did warm-up + capture + 1 replay = 3 collects as well
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert outs[0][5].any(), "episodes should end (truncation at 30) inside the last collected horizon (steps 25-36)"


@pytest.mark.parametrize("use_graph", [False, True])
def test_first_horizon_starts_from_the_reset_observation(use_graph):
    """the first collect() must act on the `env.reset()` observation (not on a zero row), and in every horizon
    it returns — eager, the eager warm-up of the graph path, the capture + first replay, later replays — obs[t] must be the
    observation that action[t] / logp[t] were computed from."""
    from opendog_b200.env import BatchedWalkEnv
    from opendog_b200.policy import ActorCriticB200
    from opendog_b200.rollout import Rollout
    env = BatchedWalkEnv(160, seed=8, info_keys=None)
    pol = ActorCriticB200(env.obs_dim, env.act_dim, 0.4, seed=3)
    ro = Rollout(env, pol, horizon=6, use_graph=use_graph)
    reset_obs = ro.obs[0].clone()
    assert reset_obs.abs().sum().item() > 0
    std = torch.exp(pol.action_log_std.detach()).reshape(-1)
    last = None
    for k in range(4):
        ro.collect()
        torch.cuda.synchronize()
        if k == 0:
            assert torch.equal(ro.obs[0], reset_obs), "row 0 of the first horizon is the reset observation"
        else:
            assert torch.equal(ro.obs[0], last), "row 0 carries the previous horizon's last observation"
        for t in (0, ro.T - 1):
            mean = pol.act(ro.obs[t].clone(), sample=False)[3]
            z = (ro.action[t] - mean) / std
            logp = (-0.5 * z * z - torch.log(std) - 0.5 * math.log(2 * math.pi)).sum(-1)
            assert (logp - ro.logp[t]).abs().max().item() < 2e-3, (k, t)
        last = ro.obs[ro.T].clone()


def test_reference_checkpoint_loads_and_runs_through_the_kernel(tmp_path):
    """SURVEY section 5 (checkpoint/resume): a `.pth` written by the reference's own trainer
    (sim2real/train.py:587-588, `torch.save(agent.state_dict())`; fixture = its shipped
    sim2real/output/pth/quadruped_ac_sym_ep3700.pth, 22 -> 512 -> 256 -> 4) loads unchanged into ActorCriticB200, the
    tcgen05 forward reproduces the reference module's torch forward at the stated bf16 tolerances, and the deterministic
    closed-loop rollout + JSON export of sim2real/train.py:600-636 runs on it with the structural invariants of the
    symmetric-trot mapping (train.py:243-259) and the ctrlrange clip (:276)."""
    import json
    import os
    import torch.nn as nn
    from opendog_b200.compat import BatchedQuadrupedEnv
    from opendog_b200.gait import export_walk_json, TRAIN_REAL_HOME_DEG
    from opendog_b200.policy import ActorCriticB200
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_quadruped_ac_sym_ep3700.pth")
    sd = torch.load(path, map_location="cpu", weights_only=True)

    class RefActorCritic(nn.Module):                    # the reference module, restated (sim2real/train.py:132-149)
        def __init__(self, s, a):
            super().__init__()
            self.actor = nn.Sequential(nn.Linear(s, 512), nn.Tanh(), nn.Linear(512, 256), nn.Tanh(), nn.Linear(256, a), nn.Tanh())
            self.critic = nn.Sequential(nn.Linear(s, 512), nn.Tanh(), nn.Linear(512, 256), nn.Tanh(), nn.Linear(256, 1))
            self.action_log_std = nn.Parameter(torch.zeros(1, a))
    ref = RefActorCritic(22, 4)
    ref.load_state_dict(sd)                             # strict: same keys, same shapes
    pol = ActorCriticB200(22, 4, 0.4)
    missing, unexpected = pol.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    pol.sync_weights()
    env = BatchedQuadrupedEnv(256, auto_reset=True)
    obs = env.reset().clone()
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(6):                                  # observations the policy actually sees, plus random ones
        x = obs if t % 2 == 0 else torch.randn(256, 22, device="cuda", generator=g)
        mean, _, value, _ = pol.act(x, sample=False)
        with torch.no_grad():
            rm = ref.actor(x.cpu()); rv = ref.critic(x.cpu())[:, 0]
            bm, bv = _ref_forward(ref, x.cpu(), True)
        # (a trained critic's values reach ~20 here, not O(1) as for a fresh network: bf16 rounding errors scale with the
        #  magnitude of the partial sums, so the value tolerances are relative to the batch's largest |value|)
        vs = max(1.0, rv.abs().max().item())
        assert (mean.cpu() - rm).abs().max().item() < 3e-2 and (value.cpu() - rv).abs().max().item() < 3e-2 * vs
        assert (mean.cpu() - bm).abs().max().item() < 4e-3 and (value.cpu() - bv).abs().max().item() < 4e-3 * vs
        obs, _, _, _ = env.step(mean)
        obs = obs.clone()
    out = tmp_path / "walk.json"
    env2 = BatchedQuadrupedEnv(4, auto_reset=False)
    export_walk_json(pol, env2, str(out), num_steps=50)
    steps = json.load(open(out))
    assert len(steps) >= 5 and all(s["duration"] == 0.1 and len(s["targets_deg"]) == 8 for s in steps)
    H = TRAIN_REAL_HOME_DEG
    tol = 0.011                                         # targets are rounded to 2 decimals (train.py:628)
    for k, s in enumerate(steps):
        d = {n: s["targets_deg"][n] - H[n] for n in H}
        # BL mirrors FR and BR mirrors FL on the thighs (train.py:243-247); thigh targets are clipped to ctrlrange
        # [2.36, 2.8] around home 2.35619: delta in [+0.22, +25.43] degrees (our_robot.xml:14-15)
        assert abs(d["BL_tigh_actuator"] - d["FR_tigh_actuator"]) <= tol and abs(d["BR_tigh_actuator"] - d["FL_tigh_actuator"]) <= tol
        for n in ("FR_tigh_actuator", "FL_tigh_actuator"):
            assert 0.21 <= d[n] <= 25.44
        # the stance pair's knees stay home, the swinging pair's knees are antisymmetric (:249-259) up to the ctrlrange clip
        # [-1.8, -1.2] around home -1.5708: delta in [-13.13, +21.25] degrees (our_robot.xml:19-20); phases alternate
        swing, stance = (("FR_knee_actuator", "BL_knee_actuator"), ("FL_knee_actuator", "BR_knee_actuator"))[::1 if k % 2 == 0 else -1]
        assert abs(d[stance[0]]) <= tol and abs(d[stance[1]]) <= tol
        a, b = d[swing[0]], d[swing[1]]
        assert -13.14 <= a <= 21.26 and -13.14 <= b <= 21.26
        clipped = min(a, b) <= -13.12 or max(a, b) >= 21.24
        assert abs(a + b) <= 2 * tol or clipped


def test_graphed_ppo_update_equals_the_eager_update():
    """train.GraphedPPOUpdate (the PPO epoch of BASELINE configs[3] replayed as CUDA graphs) against train.ppo_update on
    the same batch from the same initial weights: same losses, same updated parameters (both run the bf16-autocast
    forward / backward; the graph only removes launch and allocator time)."""
    from opendog_b200.policy import ActorCriticB200
    from opendog_b200.train import GraphedPPOUpdate, ppo_update
    B, S, A = 8192, 33, 8
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn(B, S, device="cuda", generator=g); act = torch.randn(B, A, device="cuda", generator=g).clamp(-1, 1)
    logp = torch.randn(B, device="cuda", generator=g) * 0.1 - 8
    adv = torch.randn(B, device="cuda", generator=g); ret = torch.randn(B, device="cuda", generator=g)
    pols, outs = [], []
    for graphed in (False, True):
        pol = ActorCriticB200(S, A, 0.4, seed=0)
        torch.manual_seed(0)
        for p in pol.parameters():
            torch.nn.init.normal_(p, std=0.05)
        opt = torch.optim.Adam(pol.parameters(), lr=1e-3, fused=True, capturable=True)
        if graphed:
            upd = GraphedPPOUpdate(pol, opt, B, S, A, minibatches=4)
            for _ in range(3):                       # first call = capture (its warm-up pass is a real update), then replays
                out = upd(obs, act, logp, adv, ret)
        else:
            for _ in range(3):
                out = ppo_update(pol, opt, obs, act, logp, adv, ret, epochs=1, minibatches=4)
        pols.append(pol); outs.append({k: float(v) for k, v in out.items()})
    for k in outs[0]:
        assert abs(outs[0][k] - outs[1][k]) <= 2e-3 * max(1.0, abs(outs[0][k])), (k, outs)
    for (n0, p0), (_, p1) in zip(pols[0].named_parameters(), pols[1].named_parameters()):
        assert torch.allclose(p0, p1, rtol=0, atol=3e-3), (n0, (p0 - p1).abs().max().item())
    # the tensor-core copy of the weights was refreshed: the rollout-time forward sees the updated parameters
    m0 = pols[0].act(obs[:256], sample=False)[3]; m1 = pols[1].act(obs[:256], sample=False)[3]
    assert torch.allclose(m0, m1, atol=2e-2)


@pytest.mark.parametrize("rows,cols", [(1000, 512), (4097, 256), (37, 1024), (600, 8), (5000, 2048)])
def test_tanh_backward_with_bias_gradient_matches_torch(rows, cols):
    """`odg_tanh_backward_bias` (include/odg_policy.h) against torch's two passes: gy * (1 - y*y) in fp32 rounded once to
    bf16 (≤ 1 bf16 ulp: FMA contraction may differ; torch's own bf16 kernel, which rounds three times, within a few ulp) and
    the column sums of the kernel's OWN rounded gradient in float64 (fp32 accumulation in a
    fixed order: 1e-5 relative to the column's absolute sum); ragged row counts; deterministic."""
    import ctypes as C
    from opendog_b200 import lib as _lib
    L = _lib.load()
    g_ = torch.Generator(device="cuda").manual_seed(rows + cols)
    gy = torch.randn(rows, cols, device="cuda", generator=g_).to(torch.bfloat16)
    y = torch.tanh(torch.randn(rows, cols, device="cuda", generator=g_) * 2).to(torch.bfloat16)
    p = lambda t: C.c_void_p(t.data_ptr())
    outs = []
    for _ in range(2):
        g = torch.zeros_like(y); gb = torch.zeros(cols, device="cuda")
        scratch = torch.empty(L.odg_tanh_backward_bias_scratch_floats(cols), device="cuda")
        _lib.check(L.odg_tanh_backward_bias(p(gy), p(y), p(g), p(gb), p(scratch), rows, cols,
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "odg_tanh_backward_bias")
        outs.append((g, gb))
    (g, gb), (g2, gb2) = outs
    assert torch.equal(g, g2) and torch.equal(gb, gb2)                          # deterministic
    ref = (gy.float() * (1.0 - y.float() * y.float())).to(torch.bfloat16)     # fp32 arithmetic, ONE rounding
    d = (g.float() - ref.float()).abs()
    assert bool((d <= 2 ** -7 * ref.float().abs() + 1e-30).all())               # within one bf16 ulp everywhere (FMA contraction)
    assert float((g == ref).float().mean()) > 0.99
    # torch's own bf16 kernel rounds y*y, 1 - y*y and the product separately: it sits within a few ulp of the above
    tref = torch.ops.aten.tanh_backward(gy, y).float()
    assert bool(((g.float() - tref).abs() <= 2 ** -5 * tref.abs() + 2 ** -9 * gy.float().abs() + 1e-30).all())
    s64 = g.double().sum(0)
    assert torch.allclose(gb.double(), s64, rtol=0, atol=1e-5 * float(g.double().abs().sum(0).max()) + 1e-12)
    assert L.odg_tanh_backward_bias(p(gy), p(y), p(g), p(gb), p(scratch), rows, 24, None) == -1           # ODG_ERR_INVALID: 24 / 8 = 3
    assert b"odg_tanh_backward_bias" in L.odg_last_error()


def test_fused_hidden_layer_backward_matches_autocast_torch():
    """`policy.linear_tanh` (custom backward: one pass for tanh' and the bias gradient) against the plain autocast
    expression it replaces in the PPO update's forward: outputs within one bf16 ulp (hardware tanh), gradients within bf16
    rounding of each other."""
    from opendog_b200.policy import linear_tanh
    torch.manual_seed(5)
    B, K, N = 3000, 48, 512
    x0 = torch.randn(B, K, device="cuda"); w0 = torch.randn(N, K, device="cuda") * 0.1; b0 = torch.randn(N, device="cuda") * 0.1
    go = torch.randn(B, N, device="cuda")
    res = []
    for fused in (False, True):
        x, w, b = (t.clone().requires_grad_(True) for t in (x0, w0, b0))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = linear_tanh(x, w, b) if fused else torch.tanh(torch.nn.functional.linear(x, w, b))
        (y.float() * go).sum().backward()
        res.append((y.detach(), x.grad, w.grad, b.grad))
    (y0, gx0, gw0, gb0), (y1, gx1, gw1, gb1) = res
    # forward: the hardware tanh of the rollout kernel (tanh.approx, rel. error 2^-11) rounded to bf16: within one bf16 ulp
    # of torch's tanhf-based result, the same bits for most elements
    assert y1.dtype == torch.bfloat16
    assert bool(((y1.float() - y0.float()).abs() <= 2 ** -7 * y0.float().abs() + 1e-6).all())
    assert float((y0 == y1).float().mean()) > 0.85
    assert gx1.dtype == torch.float32 and gw1.dtype == torch.float32 and gb1.dtype == torch.float32
    rel = lambda a, b_: float((a - b_).norm() / b_.norm())
    assert rel(gx1, gx0) < 1e-2 and rel(gw1, gw0) < 1e-2 and rel(gb1, gb0) < 1e-2
    # without autocast the plain fp32 / TF32 expression is what runs
    x = x0.clone().requires_grad_(True)
    assert linear_tanh(x, w0, b0).dtype == torch.float32


@pytest.mark.parametrize("B,A", [(5000, 8), (333, 12), (70000, 4), (1, 16)])
def test_fused_ppo_loss_matches_the_torch_expression(B, A):
    """`policy.ppo_loss` on CUDA (odg_ppo_loss: loss terms and gradients in one pass) against the torch expression it
    replaces (the objective of train/train.py:117-130 on the ActorCritic of sim2real/train.py:132-149), autograd in float64:
    loss terms to 1e-5 relative, gradients to 1e-4 of their largest entry; ratios inside and outside the clip range, zero
    advantages, and exact ties (ratio = 1)."""
    from opendog_b200.policy import ppo_loss
    g = torch.Generator(device="cuda").manual_seed(B + A)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    mean0, value0, ls0 = torch.tanh(r(B, A)), r(B, 1), (r(1, A) * 0.3 - 0.9)
    action = mean0 + torch.exp(ls0) * r(B, A)
    with torch.no_grad():
        lp = torch.distributions.Normal(mean0, torch.exp(ls0).expand_as(mean0)).log_prob(action).sum(-1)
    logp_old = lp + r(B) * 0.25                     # ratios spread over ~[0.5, 2]
    logp_old[::7] = lp[::7]                         # exact ties: ratio = 1
    adv, ret = r(B), r(B)
    adv[::11] = 0.0
    clip, vf_coef, ent_coef = 0.2, 0.5, 0.005
    # fused
    m1, v1, l1 = (t.clone().requires_grad_(True) for t in (mean0, value0, ls0))
    loss1, pg1, vf1, ent1 = ppo_loss(m1, v1, l1, action, logp_old, adv, ret, clip, vf_coef, ent_coef)
    loss1.backward()
    # torch, float64
    m2, v2, l2 = (t.double().clone().requires_grad_(True) for t in (mean0, value0, ls0))
    d = torch.distributions.Normal(m2, torch.exp(l2.expand_as(m2)), validate_args=False)
    logp = d.log_prob(action.double()).sum(-1)
    ratio = torch.exp(logp - logp_old.double())
    a = adv.double()
    pg2 = -torch.min(ratio * a, torch.clamp(ratio, 1 - clip, 1 + clip) * a).mean()
    vf2 = torch.nn.functional.mse_loss(v2.squeeze(-1), ret.double())
    ent2 = d.entropy().sum(-1).mean()
    loss2 = pg2 + vf_coef * vf2 - ent_coef * ent2
    loss2.backward()
    for x, y in ((loss1, loss2), (pg1, pg2), (vf1, vf2), (ent1, ent2)):
        x, y = float(x.detach()), float(y.detach())
        assert abs(x - y) <= 1e-5 * (1 + abs(y)), (x, y)
    # a sample whose float32 ratio lands on the other side of a clip boundary than the float64 one flips its gradient: leave
    # those (ratio within 1e-5 of a boundary) out of the comparison
    rt = ratio.detach()
    near = ((rt - (1 - clip)).abs() < 1e-5) | ((rt - (1 + clip)).abs() < 1e-5)
    keep = ~near
    gm1, gm2 = m1.grad.double()[keep], m2.grad[keep]
    assert float((gm1 - gm2).abs().max()) <= 1e-4 * float(gm2.abs().max()) + 1e-12
    assert float((v1.grad.double() - v2.grad).abs().max()) <= 1e-5 * float(v2.grad.abs().max()) + 1e-12
    if not bool(near.any()):
        assert float((l1.grad.double() - l2.grad).abs().max()) <= 1e-4 * float(l2.grad.abs().max()) + 1e-9
    assert m1.grad.shape == mean0.shape and v1.grad.shape == value0.shape and l1.grad.shape == ls0.shape


def test_fused_ppo_loss_with_infinite_clip_is_the_sim2real_update():
    """sim2real/train.py:566-569: actor_loss = -(log_prob * adv).mean(), critic MSE, entropy — the clipped objective with
    clip = +inf and logp_old = logp (ratio 1): same loss gradients from `policy.ppo_loss`."""
    from opendog_b200.policy import ppo_loss
    g = torch.Generator(device="cuda").manual_seed(77)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    B, A = 4000, 8
    mean0, value0, ls0 = torch.tanh(r(B, A)), r(B, 1), (r(1, A) * 0.3 - 0.9)
    action = mean0 + torch.exp(ls0) * r(B, A)
    adv, ret = r(B), r(B)
    VALUE_LOSS_COEF, ENT = 0.5, 0.01
    m2, v2, l2 = (t.double().clone().requires_grad_(True) for t in (mean0, value0, ls0))
    d = torch.distributions.Normal(m2, torch.exp(l2.expand_as(m2)), validate_args=False)
    logp = d.log_prob(action.double()).sum(-1)
    # (the reference takes entropy().mean() over [B, A]; per-dimension mean = the summed entropy / A)
    loss2 = -(logp * adv.double()).mean() + VALUE_LOSS_COEF * torch.nn.functional.mse_loss(v2.squeeze(-1), ret.double()) \
        - ENT * d.entropy().mean()
    loss2.backward()
    m1, v1, l1 = (t.clone().requires_grad_(True) for t in (mean0, value0, ls0))
    loss1, _, _, _ = ppo_loss(m1, v1, l1, action, logp.detach().float(), adv, ret, float("inf"), VALUE_LOSS_COEF, ENT / A)
    loss1.backward()
    for a_, b_ in ((m1.grad, m2.grad), (v1.grad, v2.grad), (l1.grad, l2.grad)):
        assert float((a_.double() - b_).abs().max()) <= 2e-4 * float(b_.abs().max()) + 1e-12
