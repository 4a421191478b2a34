"""The oracle's env logic against golden vectors produced by the REFERENCE's own Python code
(tools/make_golden_walk.py: WalkEnvironment.py + ScaleActionEnvironment.py +
walk_environment_reward_calc.py imported unmodified, physics supplied by the oracle)."""
import os

import numpy as np

from oracle.oracle import WalkEnv

GOLD = os.path.join(os.path.dirname(__file__), "golden", "walk_env_golden.npz")


def test_oracle_reproduces_reference_env_outputs():
    g = np.load(GOLD)
    n_envs, n_steps = g["action"].shape[:2]
    assert g["terminated"].sum() >= 4 and g["truncated"].sum() >= 4 and (g["gait_first"] > 0).sum() >= 4
    for i in range(n_envs):
        w = WalkEnv(seed=int(g["seed"]), env_id=i)
        w.e.max_steps = int(g["max_steps"])
        w.reset()
        assert np.array_equal(w.qpos, g["init_qpos"][i])
        assert np.array_equal(w.desired_velocity, g["desired_velocity"][i])
        for t in range(n_steps):
            inj = g["inject"][i, t]
            if not np.isnan(inj[0]):
                w.qpos[3:7] = inj[:4]
            if not np.isnan(inj[4]):
                w.qvel[0] = inj[4]
            obs, r, term, trunc, info = w.step(g["action"][i, t])
            assert np.array_equal(obs, g["obs"][i, t]), (i, t)                  # float64 obs, bit-exact
            assert abs(r - g["reward"][i, t]) <= 4e-16 * max(1.0, abs(r)), (i, t)   # exp() may differ by 1 ulp
            assert term == bool(g["terminated"][i, t]) and trunc == bool(g["truncated"][i, t]), (i, t)
            assert info["patterns_matches"] == g["patterns_matches"][i, t]
            assert info["gait_first_call"] == g["gait_first"][i, t]
            for k in ("x_position", "y_position", "distance_from_origin", "reward_ctrl"):
                assert abs(info[k] - g[k][i, t]) <= 1e-15 * max(1.0, abs(g[k][i, t])), (k, i, t)
            assert np.allclose(info["paw_contact_forces"], g["paw_contact_forces"][i, t], rtol=0, atol=1e-12)
            if term or trunc:
                robs = w.reset()
                assert g["did_reset"][i, t]
                assert np.array_equal(robs, g["reset_obs"][i, t])
