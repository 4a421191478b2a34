"""GPU parity of the Go1 fidelity rows: the full collision set of unitree_go1/go1.xml:26-64 (trunk
and leg primitives, condim-6 feet), data.cfrc_ext, and the Jump task, through the C ABI against the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_emu_parity import FLIP_M  # noqa: E402


def _tumble_states(n_trials=8, steps=260, every=3, seed=5):
    from oracle.oracle import Sim
    sim = Sim("go1")
    rng = np.random.default_rng(seed)
    lo = np.array([r[0] for r in sim.desc["act_ctrlrange"]]); hi = np.array([r[1] for r in sim.desc["act_ctrlrange"]])
    out = []
    for trial in range(n_trials):
        sim.reset_keyframe()
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        sim.qpos[3:7] = q; sim.qpos[2] = 0.45
        sim.qpos[7:] = rng.uniform(lo, hi)
        sim.qvel[:] = rng.normal(size=18) * 0.5
        for k in range(steps):
            if k % 50 == 0:
                sim.ctrl[:] = rng.uniform(lo, hi)
            if k % every == 0:
                out.append((sim.qpos.copy(), sim.qvel.copy(), sim.qacc_warmstart.copy(), sim.ctrl.copy()))
            sim.step()
    return sim, out


def test_go1_colliders_condim6_and_cfrc_ext_tumble_parity():
    from opendog_b200.env import BatchedWalkEnv
    sim, states = _tumble_states()
    N = len(states)
    f32 = lambda i: np.stack([st[i] for st in states]).astype(np.float32)
    env = BatchedWalkEnv(N, model="go1", frame_skip=1, scale_actions=0, auto_reset=0, solver_iterations=100,
                         info_keys=("ncon", "contact_normal_force", "cfrc_ext"))
    env.set_state(f32(0), f32(1), f32(2))
    _, _, _, info = env.step(torch.from_numpy(f32(3)).cuda())
    gq, gv = [t.cpu().numpy() for t in env.get_state()]
    ncon = info["ncon"].cpu().numpy(); gcf = info["cfrc_ext"].cpu().numpy()
    geoms = sim.desc["geoms"]
    seen, bad, flips, worst_cf = set(), [], 0, 0.0
    for i in range(N):
        sim.reset_keyframe()
        sim.qpos[:] = f32(0)[i]; sim.qvel[:] = f32(1)[i]; sim.qacc_warmstart[:] = f32(2)[i]; sim.ctrl[:] = f32(3)[i]
        sim.step()
        for c in sim.contacts():
            seen.add((geoms[c["geom"]]["type"], geoms[c["geom"]]["leg"] < 0, c["dim"]))
        if ncon[i] != sim.ncon:
            assert sim.decision_gaps[0] < FLIP_M, (i, ncon[i], sim.ncon, sim.decision_gaps)
            flips += 1
            continue
        eq = (np.abs(gq[i] - sim.qpos) - (1e-6 + 1e-5 * np.abs(sim.qpos))).max()
        ev = (np.abs(gv[i] - sim.qvel) - (1e-4 + 1e-3 * np.abs(sim.qvel))).max()
        if eq > 0 or ev > 0:
            bad.append((i, float(eq), float(ev)))
        cf = sim.cfrc_ext()[1:]
        worst_cf = max(worst_cf, np.abs(gcf[i] - cf).max() / (1.0 + np.abs(cf).max()))
    assert {(1, False, 6), (2, False, 3), (2, True, 3), (3, False, 3), (3, True, 3), (4, True, 3)} <= seen, seen
    assert flips <= N // 100 + 1
    assert len(bad) <= 0.005 * N + 1 and all(b[1] < 1e-5 and b[2] < 1e-3 for b in bad), bad
    assert worst_cf < 2e-2, worst_cf


def test_jump_task_matches_the_oracle_and_auto_resets():
    from opendog_b200.env import BatchedWalkEnv, JUMP_TERMS
    from oracle.go1_tasks import JumpEnv
    N, seed = 8, 11
    env = BatchedWalkEnv(N, model="go1", seed=seed, task="jump", auto_reset=0, solver_iterations=100,
                         info_keys=("task_terms", "reward_unclipped", "x_position", "y_position", "distance_from_origin", "cfrc_ext"))
    assert env.obs_dim == 21 and env.act_dim == 12 and len(JUMP_TERMS) == 10
    ws = [JumpEnv(seed=seed, env_id=i) for i in range(N)]
    obs = env.reset().cpu().numpy()
    assert np.abs(obs - np.stack([w.reset() for w in ws])).max() < 1e-6
    assert np.array_equal(env.get_state()[0].cpu().numpy(), np.stack([w.qpos for w in ws]).astype(np.float32))
    assert np.array_equal(env.get_env_state()["desired_velocity"].cpu().numpy(), np.stack([w.desired_velocity for w in ws]))
    rng = np.random.default_rng(3)
    desc = ws[0].sim.desc
    lo = np.array([r[0] for r in desc["act_ctrlrange"]]); hi = np.array([r[1] for r in desc["act_ctrlrange"]])
    key = np.array(desc["key_ctrl"])
    n_cc = n_air = n_term = 0
    for t in range(40):
        ctrl = np.clip(key + rng.uniform(-0.5, 0.5, (N, 12)), lo, hi).astype(np.float32)
        for i, w in enumerate(ws):
            w.step(ctrl[i])
            if i % 4 == 1 and t % 5 == 2:
                w.qpos[:3] = [rng.uniform(0.2, 1.4), rng.uniform(-0.4, 0.4), rng.uniform(0.48, 0.7)]
            if i % 4 == 2 and t % 6 == 3:
                w.qpos[2] = 0.13; w.qpos[3:7] = [np.cos(0.725), np.sin(0.725), 0.0, 0.0]
            if i % 4 == 3 and t % 9 == 4:
                a = 0.36 if t % 2 else -0.34
                w.qpos[3:7] = [np.cos(a / 2), 0.0, 0.0, np.sin(a / 2)]
            w.qpos[:] = w.qpos.astype(np.float32); w.qvel[:] = w.qvel.astype(np.float32)
        env.set_state(np.stack([w.qpos for w in ws]), np.stack([w.qvel for w in ws]), np.stack([w.sim.qacc_warmstart for w in ws]))
        env.set_env_state(step=torch.tensor([w.step_count for w in ws], dtype=torch.int32))
        obs, rew, term, trunc, info = env.evaluate(torch.from_numpy(ctrl))
        obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        inf = {k: v.cpu().numpy() for k, v in info.items()}
        for i, w in enumerate(ws):
            w.sim.ctrl[:] = ctrl[i]
            oo, orew, oterm, otrunc, oinfo = w.evaluate()
            assert np.abs(obs[i] - oo).max() < 2e-6, (t, i)
            assert term[i] == oterm and trunc[i] == otrunc, (t, i)                 # bit-exact flags
            assert inf["x_position"][i] == np.float32(oinfo["x_position"]) and inf["y_position"][i] == np.float32(oinfo["z_position"])
            if abs(oinfo["collision_norm"] - 0.1) >= 1e-3:
                assert np.allclose(inf["task_terms"][i], oinfo["terms"], rtol=3e-6, atol=1e-6), (t, i)
                assert abs(rew[i] - orew) <= 3e-6 * max(1.0, abs(orew)), (t, i)
            n_cc += oinfo["terms"][9] > 0; n_air += oinfo["terms"][0] > 0; n_term += oterm
    assert n_cc >= 4 and n_air >= 4 and n_term >= 4
    # stepping with auto-reset: episodes end (static_stability / 750-step truncation shortened), reset obs follow
    e = BatchedWalkEnv(256, model="go1", seed=2, task="jump", max_episode_steps=20, info_keys=("terminal_obs",))
    o = e.reset().clone()
    dones = 0
    for t in range(45):
        a = torch.from_numpy(np.clip(key + rng.uniform(-0.8, 0.8, (256, 12)), lo, hi).astype(np.float32)).cuda()
        o, r, d, inf = e.step(a)
        assert torch.isfinite(o).all() and torch.isfinite(r).all() and (r >= 0).all()
        dones += int(d.sum())
        if d.any():
            st = e.get_env_state()["step"]
            assert (st[d] == 0).all()
            assert (o[d][:, 2:6] == 0).all()                                      # reset obs: zero velocities
    assert dones >= 256 * 2


@pytest.mark.parametrize("n", [1, 3, 65, 1000])
def test_ragged_batch_sizes_on_every_model_and_task(n):
    """Batch sizes that do not fill a warp, a block or a wave (the state arrays are padded to whole blocks and the padding
    environments never touch caller buffers): every model / task steps, stays finite, and the first environment of a
    ragged handle is bit-identical to the same environment in a 64-env handle."""
    from opendog_b200.env import BatchedWalkEnv
    g = torch.Generator(device="cuda").manual_seed(n)
    for kw in (dict(), dict(model="go1"), dict(model="go1", task="jump")):
        e = BatchedWalkEnv(n, seed=1, info_keys=None, **kw)
        ref = BatchedWalkEnv(64, seed=1, info_keys=None, **kw)
        o, o64 = e.reset(), ref.reset()
        assert torch.equal(o[0], o64[0])
        for t in range(4):
            a64 = (torch.rand(64, e.act_dim, device="cuda", generator=g) * 2 - 1) * (0.3 if kw else 1.0)
            if kw.get("task") == "jump":
                a64 = torch.tensor(e.desc["key_ctrl"], device="cuda").expand(64, -1) + a64
            a = a64[:n] if n <= 64 else torch.cat([a64, a64[:1].expand(n - 64, -1)]).contiguous()
            o, r, d, _ = e.step(a.contiguous())
            o64, r64, d64, _ = ref.step(a64.contiguous())
            assert torch.isfinite(o).all() and torch.isfinite(r).all(), (n, kw, t)
            assert torch.equal(o[0], o64[0]) and torch.equal(r[0], r64[0]) and torch.equal(d[0], d64[0]), (n, kw, t)
