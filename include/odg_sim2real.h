/* odg_sim2real.h — C ABI of the second environment surface of the reference: `QuadrupedEnv`
 * (/root/reference/Code/mujoco/sim2real/train.py:151-411), the 4-tuple `reset()` / `step(action[4]) ->
 * (obs[22], reward, done, info)` environment the hand-rolled actor-critic trains on. Batched: one handle steps
 * `num_envs` environments; state never leaves HBM. It rides on an OdgSim created with frame_skip =
 * int(0.10 / timestep) = 50, scale_actions = 0, auto_reset = 0 (this layer owns resets).
 *
 * Per policy step: k_s2r_pre (symmetric-trot mapping of 4 actions to 8 clipped ctrl targets, :235-285) ->
 * the fused physics kernel (50 x mj_step) -> k_s2r_post (obs :184-207, the nine reward terms :313-392,
 * termination :393-402, bookkeeping, optional auto-reset to the settled keyframe state :209-233).
 * Same conventions as odg.h (caller-owned device pointers, stream as void*, 0 / negative status).
 */
#ifndef ODG_SIM2REAL_H
#define ODG_SIM2REAL_H

#include <stdint.h>
#include "odg.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OdgS2R OdgS2R;

enum OdgS2RReason {                 /* info["termination_reason"], sim2real/train.py:393-402 */
  ODG_S2R_RUNNING = 0,              /* "max_steps" is what the reference reports while not done */
  ODG_S2R_MJ_ERROR = 1,
  ODG_S2R_ORIENTATION_LIMIT = 2,
  ODG_S2R_TOO_MUCH_BACKWARD = 3
};

typedef struct OdgS2RConfig {
  double action_amplitude_rad;      /* ACTION_AMPLITUDE_RAD = radians(40)                 :75-76 */
  int settle_steps;                 /* NUM_SETTLE_STEPS = 100 (multiple of the sim's frame_skip) :91 */
  int auto_reset;                   /* 1: a done env is reset inside odg_s2r_step (obs = reset obs) */
  double real_home_deg[8];          /* real_robot_home_deg_map in ACTUATOR_NAMES_ORDERED order  :95-102 */
  double joint_scale[8];            /* joint_scale_factors (all 1)                               :103 */
  int max_steps;                    /* MAX_STEPS_PER_EPISODE = 250 (:68): the reference's training loop ends an episode
                                       after this many policy steps (:539) and resets the env. With auto_reset = 1 this
                                       layer plays that loop's role: an env whose counter reaches max_steps reports done = 1
                                       with reason ODG_S2R_RUNNING ("max_steps", no penalty) and is reset. 0 = no cap;
                                       ignored when auto_reset = 0 (the caller's own loop owns the cap, as in the reference) */
  int variant;                      /* OdgS2RVariant: which of the reference's two `QuadrupedEnv` classes this handle is */
} OdgS2RConfig;

enum OdgS2RVariant {
  ODG_S2R_TRAIN = 0,                /* sim2real/train.py:151-411 — 4 actions through the symmetric-trot mapping, obs 22,
                                       0.10 s per policy step (50 x mj_step), nine reward terms, limits 25 deg */
  ODG_S2R_TERRAIN = 1               /* sim2real/train2.py:159-411 — the terrain trainer: 8 direct joint targets (amplitude
                                       50 deg, :93), obs 12 (yaw pitch roll, 8 joint offsets, v_x; :197-201), 0.08 s per
                                       policy step (40 x mj_step, :106,178), twelve reward terms (:352-407), termination at
                                       35 deg roll/pitch, 52.5 deg yaw, backward ratio 0.85 (:409-414), MAX_STEPS 1000 (:88).
                                       As shipped that trainer loads walking_scene.xml (:66), whose height field is an asset
                                       without a geom: the generated terrain (odg_terrain_*) never touches the physics */
};

void odg_s2r_default_config(OdgS2RConfig* cfg);
/* the same for a given variant (amplitude, max_steps and variant filled in accordingly) */
void odg_s2r_default_config_for(OdgS2RConfig* cfg, int variant);
int odg_s2r_obs_dim(const OdgS2R* e);     /* 22 / 12 */
int odg_s2r_act_dim(const OdgS2R* e);     /* 4 / 8 */

/* Replaces `QuadrupedEnv(xml_path)` (:152-182). Computes the settled reset state once (keyframe + settle_steps
 * x mj_step with ctrl = home; it is deterministic, so every reset restores the same state). */
int odg_s2r_create(OdgSim* sim, const OdgModel* model, const OdgS2RConfig* cfg, OdgS2R** out);
void odg_s2r_destroy(OdgS2R* e);

/* Replaces `env.reset()` (:227-233). mask_dev [N] u8 or NULL = all; obs_dev [N][obs_dim] f32 or NULL. */
int odg_s2r_reset(OdgS2R* e, const uint8_t* mask_dev, float* obs_dev, void* stream);

/* Replaces `env.step(action)` (:287-411). action_dev [N][act_dim] f32 in [-1,1]; obs_dev [N][obs_dim]; reward_dev [N];
 * done_dev [N] u8; reason_dev [N] u8 (OdgS2RReason, nullable); sim_target_rad_dev [N][8] f32 in ctrl order
 * (info["sim_target_rad"], nullable); terminal_obs_dev [N][obs_dim] (obs before auto-reset, nullable). */
int odg_s2r_step(OdgS2R* e, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                 uint8_t* reason_dev, float* sim_target_rad_dev, float* terminal_obs_dev, void* stream);

/* Test hook: overwrite the per-episode bookkeeping (step counter, previous x, cumulative +/- displacement,
 * previous net displacement, last commanded targets [N][8]); NULL pointers are skipped. */
int odg_s2r_set_bookkeeping(OdgS2R* e, const int32_t* counter, const double* prev_x, const double* cum_pos,
                            const double* cum_neg, const double* prev_net, const float* last_cmd, void* stream);

/* ---- the terrain trainer's height field (sim2real/train2.py:203-304), one 100 x 100 field per environment ---------
 * `_generate_random_terrain`: with probability 1/2 a flat field (0.5 everywhere); else per cell outside a flat disc of
 * random radius 0.1-0.4 m around the robot's start: U(-1.5, 1.5) + 1.05 (sin(x fx) cos(y fy) + sin(2 x fx) cos(2 y fy))
 * with fx, fy ~ U(0.2, 0.6) per cell + (20 %) a spike U(-1.2, 1.2), x 1.5 within 1 m of the disc's rim; 4 passes of a
 * 3 x 3 blend (factor 0.3) over interior cells outside the disc; min-max normalisation; stored transposed
 * (`hfield_data[c * nrow + r]`). Random numbers are Philox4x32 keyed by (seed, global env id, generation, cell) — the
 * reference seeds Python's `random` from the wall clock, so only the distribution can be matched, not the stream; the
 * deterministic tail (blend / normalise / transpose) is held to the reference within 2e-6 (float32 like numpy; only the
 * order of the nine additions of a blend differs) and the lookup exactly. */
typedef struct OdgTerrain OdgTerrain;
int odg_terrain_create(int num_envs, int device, uint64_t seed, int first_env_id, OdgTerrain** out);
void odg_terrain_destroy(OdgTerrain* t);
/* new terrain for masked envs (NULL = all); bumps each env's generation counter */
int odg_terrain_generate(OdgTerrain* t, const uint8_t* mask_dev, void* stream);
/* test hook: run only the deterministic tail on caller-provided raw heights [N][100][100] (row-major [r][c]) and flat-disc
 * radii [N] */
int odg_terrain_from_raw(OdgTerrain* t, const float* raw_dev, const float* radius_dev, void* stream);
/* `get_terrain_height(x, y)` (:295-304) for one query point per env: xy_dev [N][2] -> height_dev [N] */
int odg_terrain_height(const OdgTerrain* t, const float* xy_dev, float* height_dev, void* stream);
/* device pointer to the fields, [N][ncol * nrow] f32 (MuJoCo's hfield_data layout per env) */
const float* odg_terrain_data(const OdgTerrain* t);

#ifdef __cplusplus
}
#endif
#endif /* ODG_SIM2REAL_H */
