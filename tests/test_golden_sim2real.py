"""oracle/sim2real_oracle.py against golden vectors produced by the REFERENCE's own QuadrupedEnv code
(tools/make_golden_sim2real.py: sim2real/train.py imported unmodified, `mujoco` stubbed onto the oracle physics)."""
import os

import numpy as np

from oracle.sim2real_oracle import QuadrupedEnvOracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sim2real_env_golden.npz")
REASONS = ["max_steps", "mj_error", "orientation_limit", "too_much_backward"]


def test_oracle_reproduces_reference_quadruped_env():
    g = np.load(GOLD)
    seen = set()
    env = QuadrupedEnvOracle()
    for ep in range(len(g["length"])):
        obs = env.reset()
        assert np.array_equal(obs, g["reset_obs"][ep])
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
