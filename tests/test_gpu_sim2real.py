"""GPU parity of the QuadrupedEnv surface (include/odg_sim2real.h, opendog_b200/compat.py) against
oracle/sim2real_oracle.py — itself pinned to the reference's own class by tests/test_golden_sim2real.py.

Tolerances (fp32 physics over 50 substeps per policy step, chaotic contacts): reset obs 2e-4 (100 settle
substeps), per-step obs 2e-3 and reward 2e-3 * max(1, |r|) with the oracle re-synchronised to the GPU state
before every step; an env-step outside must have had a discrete collision decision within FLIP_M of flipping in one of
its 50 substeps (oracle-confirmed; amplified by the substeps that follow) and at most 4 % of env-steps may be such;
median obs error < 5e-5; commanded targets (`sim_target_rad`) exact, done flags and termination reasons exact on in-tolerance
steps."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_emu_parity import FLIP_M  # noqa: E402

REASONS = ["max_steps", "mj_error", "orientation_limit", "too_much_backward"]


def test_quadruped_env_matches_oracle():
    from opendog_b200.compat import BatchedQuadrupedEnv
    from oracle.sim2real_oracle import QuadrupedEnvOracle
    N, T = 16, 14
    env = BatchedQuadrupedEnv(N, auto_reset=False)
    obs = env.reset().cpu().numpy()
    orcs = [QuadrupedEnvOracle() for _ in range(N)]
    oobs = np.stack([o.reset() for o in orcs])
    assert np.abs(obs - oobs).max() < 2e-4, "settled reset observation"
    rng = np.random.default_rng(3)
    bad, errs = 0, []
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 4)).astype(np.float32)
        a[1, 0] = a[1, 2] = -abs(a[1, 0])                         # env 1 walks backwards
        # identical pre-step state on both sides
        gq, gv = [x.cpu().numpy() for x in env.sim.get_state()]
        for i, o in enumerate(orcs):
            o.sim.qpos[:] = gq[i]; o.sim.qvel[:] = gv[i]
        obs, rew, done, info = env.step(torch.from_numpy(a))
        obs = obs.cpu().numpy(); rew = rew.cpu().numpy(); done = done.cpu().numpy()
        tgt = info["sim_target_rad"].cpu().numpy(); reason = info["termination_reason"].cpu().numpy()
        for i, o in enumerate(orcs):
            o.prev_x = float(gq[i][0]) if t == 0 else o.prev_x
            eo, er, ed, ei = o.step(a[i])
            assert np.abs(tgt[i] - ei["sim_target_rad"]).max() < 1e-6
            ok = np.abs(obs[i] - eo).max() < 2e-3 and abs(rew[i] - er) < 2e-3 * max(1.0, abs(er))
            bad += (not ok)
            errs.append(float(np.abs(obs[i] - eo).max()))
            if not ok:
                assert min(o.min_gap[0], o.min_gap[1]) < FLIP_M, f"t={t} env={i}: outside tolerance without a decision flip {o.min_gap}"
                print(f"outlier t={t} env={i}: obs err {np.abs(obs[i] - eo).max():.3e} (idx {np.abs(obs[i] - eo).argmax()}) "
                      f"reward {rew[i]:.5f} vs {er:.5f}")
            if ok:
                assert bool(done[i]) == ed and REASONS[int(reason[i])] == ei["termination_reason"], (t, i)
    # One policy step is 50 substeps of impact-rich contact dynamics from an identical state: when a hull vertex sits
    # within fp32 rounding of the 1 mm contact margin in any of them, the two paths disagree on one contact for one
    # substep and the difference is amplified by the remaining substeps (measured: median 6e-6, 99th percentile 2e-5,
    # a few percent of env-steps at 1e-2..1e-1 in a joint velocity). Bound the rate, and the bulk tightly.
    assert bad <= (N * T) * 4 // 100, f"{bad}/{N * T} env-steps outside tolerance"
    assert np.median(errs) < 5e-5 and np.percentile(errs, 90) < 2e-4


def test_single_env_facade_and_termination_paths():
    """The numpy 4-tuple façade (class name, constructor, info keys of sim2real/train.py) and auto-reset."""
    from opendog_b200.compat import BatchedQuadrupedEnv, QuadrupedEnv
    env = QuadrupedEnv("our_robot/walking_scene.xml")
    assert (env.state_dim, env.action_dim) == (22, 4)
    s = env.reset()
    assert s.dtype == np.float32 and s.shape == (22,) and s[20] == 0.0 and s[21] == 1.0     # phase 0: sin 0, cos 1
    s2, r, d, info = env.step(np.zeros(4, np.float32))
    assert isinstance(r, float) and isinstance(d, bool) and info["sim_target_rad"].shape == (8,)
    assert info["termination_reason"] == "max_steps" and abs(s2[21] + 1.0) < 1e-6         # phase 1: cos(pi) = -1
    # zero action = home targets clipped to ctrlrange: thigh 2.35619 -> 2.36 (our_robot.xml:15), knee -1.5708
    assert np.allclose(info["sim_target_rad"], [2.36, -1.5708] * 4, atol=1e-6)
    env.close()
    # batched + auto-reset: roll two envs past 25 degrees -> done, reason orientation_limit, obs = settled reset obs
    b = BatchedQuadrupedEnv(4, auto_reset=True)
    first = b.reset().clone()
    qpos, qvel = b.sim.get_state()
    qpos[1, 3:7] = torch.tensor([np.cos(0.6), np.sin(0.6), 0.0, 0.0]); qpos[3, 3:7] = qpos[1, 3:7]
    b.sim.set_state(qpos, qvel)
    obs, rew, done, info = b.step(torch.zeros(4, 4))
    assert done.cpu().tolist() == [False, True, False, True]
    assert info["termination_reason"].cpu().tolist() == [0, 2, 0, 2]
    assert torch.equal(obs[1], first[1]) and torch.equal(obs[3], first[3])
    assert (info["terminal_obs"][1] - obs[1]).abs().max() > 0.1
    assert rew[1] < rew[0] - 4.0                                                           # the -5 termination penalty


def test_auto_reset_enforces_the_training_loops_episode_cap():
    """sim2real/train.py:68,539: the reference's loop ends an episode after MAX_STEPS_PER_EPISODE policy steps. With
    auto_reset the batched env plays that role: done = 1, reason "max_steps" (no penalty), bookkeeping and state reset."""
    from opendog_b200.compat import BatchedQuadrupedEnv, REASONS
    b = BatchedQuadrupedEnv(8, auto_reset=True, max_steps=5)
    first = b.reset().clone()
    a = torch.zeros(8, 4)
    for t in range(1, 12):
        obs, rew, done, info = b.step(a)
        if t % 5 == 0:
            assert done.all() and (info["termination_reason"] == 0).all() and REASONS[0] == "max_steps"
            assert torch.equal(obs, first), "a capped episode restarts from the settled reset state"
            assert (rew > -1.0).all()                                # no termination penalty
        else:
            assert not done.any()
    nocap = BatchedQuadrupedEnv(8, auto_reset=True, max_steps=0)
    nocap.reset()
    for t in range(7):
        _, _, done, _ = nocap.step(a)
        assert not done.any()
