"""oracle/sim2real_oracle.py (train2.py part) against golden vectors produced by the REFERENCE's own terrain-trainer
environment (tools/make_golden_terrain.py: sim2real/train2.py imported unmodified, `mujoco` stubbed onto the oracle
physics): env steps, the terrain generator under the reference's own seeding, and `get_terrain_height`."""
import os
import random

import numpy as np

from oracle.sim2real_oracle import (QuadrupedEnvV2Oracle, generate_terrain, smooth_and_normalise, terrain_height)

GOLD = os.path.join(os.path.dirname(__file__), "golden", "terrain_env_golden.npz")
REASONS = ["max_steps", "mj_error", "orientation_limit", "too_much_backward"]


def test_oracle_reproduces_reference_terrain_env():
    g = np.load(GOLD)
    seen = set()
    env = QuadrupedEnvV2Oracle()
    assert (env.state_dim, env.action_dim, env.n_sub) == (12, 8, 40)
    for ep in range(len(g["length"])):
        obs = env.reset(seed=int(g["seed"][ep]))
        assert np.array_equal(obs, g["reset_obs"][ep])
        assert np.array_equal(env.hfield_data, g["hfield"][ep])                        # the generated terrain, bit-exact
        for t in range(int(g["length"][ep])):
            inj = g["inject"][ep, t]
            if not np.isnan(inj[0]):
                env.sim.qpos[3:7] = inj
            obs, r, done, info = env.step(g["action"][ep, t])
            assert np.array_equal(obs, g["obs"][ep, t]), (ep, t)                      # float32 obs, bit-exact
            assert abs(r - g["reward"][ep, t]) <= 1e-12 * max(1.0, abs(r)), (ep, t)
            assert done == bool(g["done"][ep, t])
            assert info["termination_reason"] == REASONS[int(g["reason"][ep, t])]
            assert np.array_equal(info["sim_target_rad"], g["sim_target_rad"][ep, t])
            seen.add(info["termination_reason"])
        assert done or int(g["length"][ep]) == g["action"].shape[1]
    assert {"orientation_limit", "too_much_backward"} <= seen
    assert sum(float(np.ptp(h)) > 0.5 for h in g["hfield"]) >= 2 and any(float(np.ptp(h)) == 0 for h in g["hfield"])


def test_terrain_generator_structure_and_height_lookup():
    g = np.load(GOLD)
    k = int(g["height_terrain"])
    hts = np.array([terrain_height(g["hfield"][k], x, y) for x, y in g["height_points"]])
    assert np.array_equal(hts, g["heights"])
    # structure of a rough terrain (train2.py:203-270): normalised to [0, 1], flat disc of radius 0.1-0.4 m around the start
    # (raw height 0 there), stored TRANSPOSED: hfield_data[c * nrow + r]
    random.seed(int(g["seed"][k]))
    h = generate_terrain(random)
    assert np.array_equal(h, g["hfield"][k]) and h.min() == 0.0 and h.max() == 1.0
    grid = h.reshape(100, 100).T                                                       # [row, col]
    cs = 5.0 / 99
    centre = [(r, c) for r in range(100) for c in range(100) if np.hypot(-2.5 + c * cs, -2.5 + r * cs) < 0.1]
    assert len(centre) >= 4 and len({grid[r, c] for r, c in centre}) == 1              # the disc is one flat level
    # the deterministic tail alone: blur outside the disc, normalise, transpose
    raw = np.random.default_rng(0).uniform(-1, 1, (100, 100)).astype(np.float32)
    out = smooth_and_normalise(raw, 0.3).reshape(100, 100).T
    assert out.min() == 0.0 and out.max() == 1.0
    # border cells are never smoothed (loops run 1..R-2): they keep their raw value up to the one affine normalisation map
    ratio = (out[0, 1:] - out[0, :-1]) / np.where(raw[0, 1:] - raw[0, :-1] == 0, 1, raw[0, 1:] - raw[0, :-1])
    assert np.allclose(ratio, ratio[0], rtol=1e-3)
