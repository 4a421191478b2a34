"""The Jump task oracle (oracle/go1_tasks.py) against golden vectors produced by the REFERENCE's own
JumpEnvironmentV0 + JumpEnvironmentRewardCalc code (tools/make_golden_jump.py; tests/golden/jump_env_golden.npz)."""
import os

import numpy as np

from oracle.go1_tasks import JUMP_TERMS, JumpEnv

GOLD = os.path.join(os.path.dirname(__file__), "golden", "jump_env_golden.npz")


def test_oracle_reproduces_reference_jump_env():
    g = np.load(GOLD)
    n_envs, n_steps = g["action"].shape[:2]
    seen_cc = seen_air = 0
    for i in range(n_envs):
        w = JumpEnv(seed=int(g["seed"]), env_id=i, max_steps=int(g["max_steps"]))
        assert np.array_equal(w.desired_velocity, g["desired_velocity"][i])
        w.reset()
        for t in range(n_steps):
            iq, iv = g["inject_qpos"][i, t], g["inject_qvel"][i, t]
            w.qpos[~np.isnan(iq)] = iq[~np.isnan(iq)]
            w.qvel[~np.isnan(iv)] = iv[~np.isnan(iv)]
            obs, rew, term, trunc, info = w.step(g["action"][i, t])
            assert np.array_equal(obs, g["obs"][i, t]), (i, t)                      # float64, bit-exact
            assert rew == g["reward"][i, t] and term == g["terminated"][i, t] and trunc == g["truncated"][i, t], (i, t)
            assert np.array_equal(info["terms"], g["terms"][i, t]), (i, t, dict(zip(JUMP_TERMS, info["terms"])))
            for k in ("x_position", "y_position", "z_position", "distance_from_origin"):
                assert info[k] == g[k][i, t]
            assert info["z_position"] == w.qpos[1]                                  # quirk C17 (JumpEnvironment.py:81)
            seen_cc += info["terms"][9] > 0; seen_air += info["terms"][0] > 0
            if g["did_reset"][i, t]:
                assert np.array_equal(w.reset(), g["reset_obs"][i, t])
    assert seen_cc > 5 and seen_air > 5 and g["terminated"].sum() > 3 and g["truncated"].sum() > 1
