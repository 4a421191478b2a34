"""GPU parity of the terrain trainer's environment (sim2real/train2.py) against
oracle/sim2real_oracle.py — itself pinned to the reference's own class and terrain generator by
tests/test_golden_terrain.py. Same tolerances as tests/test_gpu_sim2real.py (40 substeps per policy step here)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_emu_parity import FLIP_M  # noqa: E402

REASONS = ["max_steps", "mj_error", "orientation_limit", "too_much_backward"]


def test_terrain_env_matches_oracle():
    from opendog_b200.compat import BatchedTerrainQuadrupedEnv
    from oracle.sim2real_oracle import QuadrupedEnvV2Oracle
    N, T = 16, 14
    env = BatchedTerrainQuadrupedEnv(N, auto_reset=False)
    assert (env.state_dim, env.action_dim) == (12, 8)
    obs = env.reset().cpu().numpy()
    orcs = [QuadrupedEnvV2Oracle(terrain=False) for _ in range(N)]
    oobs = np.stack([o.reset() for o in orcs])
    assert np.abs(obs - oobs).max() < 2e-4, "settled reset observation"
    rng = np.random.default_rng(3)
    bad, errs, seen = 0, [], set()
    for t in range(T):
        a = rng.uniform(-1, 1, (N, 8)).astype(np.float32)
        a[1, [0, 2]] = -abs(a[1, 0]); a[1, [4, 6]] = abs(a[1, 0])          # env 1 shuffles backwards
        gq, gv = [x.cpu().numpy() for x in env.sim.get_state()]
        if t == 5:                                                       # env 2: rolled past 35 degrees
            gq[2, 3:7] = [np.cos(0.4), np.sin(0.4), 0.0, 0.0]
            env.sim.set_state(gq, gv)
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
            else:
                assert bool(done[i]) == ed and REASONS[int(reason[i])] == ei["termination_reason"], (t, i)
                seen.add(ei["termination_reason"])
    assert bad <= (N * T) * 4 // 100, f"{bad}/{N * T} env-steps outside tolerance"
    assert np.median(errs) < 5e-5 and np.percentile(errs, 90) < 2e-4
    assert "orientation_limit" in seen


def test_terrain_generator_tail_height_lookup_and_statistics():
    from opendog_b200.compat import BatchedTerrainQuadrupedEnv
    from oracle.sim2real_oracle import smooth_and_normalise, terrain_height
    N = 256
    env = BatchedTerrainQuadrupedEnv(N, seed=5, auto_reset=True, max_steps=3)
    # (1) the deterministic tail (4 blend passes outside the flat disc, min-max normalisation, transpose) on given raw heights
    rng = np.random.default_rng(0)
    raw = rng.uniform(-2, 2, (N, 100, 100)).astype(np.float32)
    rad = rng.uniform(0.1, 0.4, N).astype(np.float32)
    out = env.terrain_from_raw(torch.from_numpy(raw), torch.from_numpy(rad)).cpu().numpy()
    for i in (0, 1, 17, 255):
        ref = smooth_and_normalise(raw[i], float(rad[i]))
        assert np.abs(out[i] - ref).max() < 2e-6, i
    flat = np.zeros((N, 100, 100), np.float32)                            # max <= min + 1e-4 -> 0.5 everywhere (:264)
    assert (env.terrain_from_raw(torch.from_numpy(flat), torch.from_numpy(rad)).cpu().numpy() == 0.5).all()
    # (2) the generator: half of the resets are flat, the others normalised to [0, 1] with a flat disc around the start
    env.reset()
    h = env.hfield_data.cpu().numpy()
    is_flat = (h == 0.5).all(1)
    assert 0.35 < is_flat.mean() < 0.65
    rough = h[~is_flat]
    assert np.all(rough.min(1) == 0.0) and np.all(rough.max(1) == 1.0)
    grid = rough.reshape(-1, 100, 100).transpose(0, 2, 1)                 # [env, row, col]
    cs = 5.0 / 99
    rr, cc = np.meshgrid(np.arange(100), np.arange(100), indexing="ij")
    disc = np.hypot(-2.5 + cc * cs, -2.5 + rr * cs) < 0.1                 # inside every possible flat disc (radius >= 0.1)
    assert disc.sum() >= 4 and np.all(np.ptp(grid[:, disc], axis=1) == 0)
    far = np.hypot(-2.5 + cc * cs, -2.5 + rr * cs) > 1.6
    assert np.all(grid[:, far].std(1) > 0.05)                             # and rough away from it
    # (3) get_terrain_height on the generated fields, exactly
    pts = rng.uniform(-2.7, 2.7, (N, 2)).astype(np.float32)
    got = env.terrain_height(torch.from_numpy(pts)).cpu().numpy()
    want = np.array([terrain_height(h[i], float(pts[i, 0]), float(pts[i, 1])) for i in range(N)])
    assert np.array_equal(got, want.astype(np.float32))
    # (4) episodic regeneration: only environments whose episode ended get a new field; streams are keyed by global env id
    before = h.copy()
    for t in range(2):
        _, _, done, _ = env.step(torch.zeros(N, 8))
        assert not done.any() and np.array_equal(env.hfield_data.cpu().numpy(), before)
    _, _, done, _ = env.step(torch.zeros(N, 8))
    assert done.all()                                                    # max_steps = 3
    after = env.hfield_data.cpu().numpy()
    assert (after != before).any(1).mean() > 0.6                          # (flat -> flat redraws look unchanged)
    lo = BatchedTerrainQuadrupedEnv(128, seed=5); hi = BatchedTerrainQuadrupedEnv(128, seed=5, first_env_id=128)
    lo.reset(); hi.reset()
    assert np.array_equal(np.concatenate([lo.hfield_data.cpu().numpy(), hi.hfield_data.cpu().numpy()]), before)


def test_single_env_terrain_facade():
    from opendog_b200.compat import TerrainQuadrupedEnv
    env = TerrainQuadrupedEnv("our_robot/walking_scene.xml")
    assert (env.state_dim, env.action_dim, env.sim_steps_per_policy_step) == (12, 8, 40)
    s = env.reset()
    assert s.dtype == np.float32 and s.shape == (12,)
    s2, r, d, info = env.step(np.zeros(8, np.float32))
    assert isinstance(r, float) and isinstance(d, bool) and info["sim_target_rad"].shape == (8,)
    assert np.allclose(info["sim_target_rad"], [2.36, -1.5708] * 4, atol=1e-6)      # home targets clipped to ctrlrange
    assert env.hfield_data.shape == (10000,) and 0.0 <= env.get_terrain_height(0.0, 0.0) <= 0.302
    env.close()


def test_terrain_policy_network_and_rollout():
    """The terrain trainer's ActorCritic (train2.py:149-157: 12 -> 1024 -> 512 -> 8, std 0.3) in the reference's
    state_dict layout, driving the batched terrain environment through the on-device rollout."""
    import torch.nn as nn
    from opendog_b200.compat import BatchedTerrainQuadrupedEnv
    from opendog_b200.policy import ActorCriticB200
    pol = ActorCriticB200(12, 8, 0.3, hidden=(1024, 512), seed=2)
    sd = pol.state_dict()
    assert tuple(sd["actor.0.weight"].shape) == (1024, 12) and tuple(sd["actor.2.weight"].shape) == (512, 1024)
    assert tuple(sd["actor.4.weight"].shape) == (8, 512) and tuple(sd["critic.4.weight"].shape) == (1, 512)
    assert tuple(sd["action_log_std"].shape) == (1, 8) and abs(float(sd["action_log_std"][0, 0]) - np.log(0.3)) < 1e-6
    env = BatchedTerrainQuadrupedEnv(512, auto_reset=True, seed=1)
    obs = env.reset()
    mean, _, value, _ = pol.act(obs, sample=False)
    with torch.no_grad():
        d, v = pol(obs)                                        # the reference module's fp32 forward
    assert (mean - d.mean).abs().max().item() < 3e-2 and (value - v[:, 0]).abs().max().item() < 3e-2
    ret = 0.0
    for t in range(20):
        a, logp, value, mean = pol.act(obs, sample=True)
        z = (a - mean) / 0.3
        assert torch.allclose(logp, (-0.5 * z * z - np.log(0.3) - 0.5 * np.log(2 * np.pi)).sum(-1), atol=1e-3)
        obs, rew, done, info = env.step(a.clamp(-1, 1))
        ret += float(rew.mean())
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert np.isfinite(ret)
