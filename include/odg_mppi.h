/* odg_mppi.h — sampling-MPC (MPPI) on the fused step kernel: BASELINE.json configs[4] "1024 action-sequence samples
 * x horizon 64 rolled out from a shared robot state, cost reduction on-device". The reference has no MPC code
 * (SURVEY §8d config 5): this is a new capability on the same hot path, expressed through the same env semantics —
 * a sample's cost is minus the sum of the walk-environment rewards (WalkEnvironment.py:81-109; the caller passes
 * OdgInfoPtrs.reward_unclipped, i.e. before the max(0, .) clip) along its rollout,
 * plus a penalty when it terminates (is_healthy, reward_calc:117-135).
 *
 * Per planning call: the caller broadcasts one robot state to N = num_samples environments of an OdgSim
 * (odg_set_state / odg_set_env_state), then for t in [0, T): odg_mppi_sample -> odg_step -> odg_mppi_accumulate,
 * and finally odg_mppi_reduce. Same conventions as odg.h.
 */
#ifndef ODG_MPPI_H
#define ODG_MPPI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* action[n][a] = clamp(mean[a] + sigma * eps, -1, 1), eps ~ N(0,1) from Philox4x32-10 keyed by
 * (seed, n, iteration, t, block). mean_dev [A]; action_dev [N][A] (one row block of the [T][N][A] action tensor). */
int odg_mppi_sample(const float* mean_dev, float sigma, int n_samples, int act_dim, uint64_t seed, uint32_t iteration,
                    uint32_t t, float* action_dev, void* stream);

/* cost[n] += alive[n] ? -reward[n] : 0;  a sample that terminates pays `termination_cost` once and stops
 * accumulating (alive[n] = 0). cost_dev [N] f32, alive_dev [N] u8. */
int odg_mppi_accumulate(const float* reward_dev, const uint8_t* terminated_dev, int n_samples, float termination_cost,
                        float* cost_dev, uint8_t* alive_dev, void* stream);

/* The whole rollout of a plan in ONE launch (what opendog_b200.mppi.MPPI runs): for every sample n of `sim`
 * (created with auto_reset = 0; the caller has broadcast the shared start state with odg_set_state / odg_set_env_state)
 * and t in [0, horizon): draw action[t][n][:] exactly as odg_mppi_sample does, take the fused environment step of
 * odg_step, and accumulate cost[n] as odg_mppi_accumulate does (a terminated sample pays `termination_cost` once, stops
 * stepping and only keeps drawing its action rows). State stays on the SM between the steps of a sample's horizon; no
 * launch or grid-wide barrier separates them. mean_dev [T][A]; actions_dev [T][N][A]; cost_dev [N] (overwritten).
 * `iteration_dev` (nullable) points to a device-side plan counter that overrides `iteration`: a CUDA graph that
 * increments it replays with fresh noise every plan. */
struct OdgSim;
int odg_mppi_rollout(struct OdgSim* sim, const float* mean_dev, float sigma, int horizon, uint64_t seed, uint32_t iteration,
                     const uint32_t* iteration_dev, float termination_cost, float* actions_dev, float* cost_dev, void* stream);

/* Information-theoretic MPPI update, one block, fixed summation order (deterministic):
 *   w[n] = exp(-(cost[n] - min cost) / lambda);  mean_out[t][a] = sum_n w[n] action[t][n][a] / sum_n w[n]
 * actions_dev [T][N][A]; mean_out_dev [T][A]; stats_dev [4] f32 = (min cost, mean cost, sum w, argmin) nullable. */
int odg_mppi_reduce(const float* cost_dev, const float* actions_dev, int horizon, int n_samples, int act_dim,
                    float lambda, float* mean_out_dev, float* stats_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ODG_MPPI_H */
