/* odg_oracle.h — CPU fp64 ORACLE for the OpenDOG hot path. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (opendog_b200/) never does.
 *
 * PARITY: UNPINNED at the single-step level, LOOSELY PINNED at the trajectory level. The arithmetic being restated
 * lives in the un-vendored third-party wheel mujoco==3.2.3 (/root/reference/Code/mujoco/install.sh:27), which cannot
 * be imported or built in this environment, and the reference ships no tests / golden vectors of mj_step itself
 * (SURVEY.md §4): no qpos / qvel / contact force of this file has ever been compared with MuJoCo's. What the reference
 * does ship are ten (policy checkpoint, 50-step closed-loop walk recorded in real MuJoCo) pairs written by the revision
 * of sim2real/train.py in the tree; under teacher forcing this file tracks them to 0.1 degree of policy output (of +-40)
 * over the first 150 engine steps and to a few degrees afterwards, and resolves the reference's 100 settling steps to the
 * step (tools/pin_walk_json.py, tests/test_mujoco_pin_walk_json.py, profiles/r02g_mujoco_pin.txt).
 * This file restates MuJoCo's published algorithm ("Computation" chapter of its documentation) for
 * the model class the reference uses, anchored on the reference's own call sites
 * (environments/WalkEnvironment.py:56-79, rewards/walk_environment_reward_calc.py, sim2real/train.py).
 */
#ifndef ODG_ORACLE_H
#define ODG_ORACLE_H

#include <stdint.h>
#include "../include/odg_model.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ODGO_MAX_CON (ODG_MAX_GEOM * ODG_MAX_CON_PER_GEOM)
#define ODGO_MAX_EFC (ODG_MAX_NV + ODG_MAX_NV + 6 * ODGO_MAX_CON)

enum { ODGO_ROW_FRICTION = 0, ODGO_ROW_LIMIT = 1, ODGO_ROW_CONTACT = 2 };

typedef struct OdgoContact {
  int geom;            /* index into OdgModel.geom */
  int vert;            /* hull vertex index (global) or -1 */
  int efc;             /* first row in the efc arrays */
  int dim;
  double dist;
  double pos[3];       /* contact point (midway between surfaces) */
  double frame[9];     /* rows: normal, tangent1, tangent2 */
  double force[6];     /* contact-frame force after the solve: normal, 2 sliding, torsional, 2 rolling */
} OdgoContact;

typedef struct OdgoData {
  /* state */
  double qpos[ODG_MAX_NQ], qvel[ODG_MAX_NV], time;
  double qacc_warmstart[ODG_MAX_NV];
  double ctrl[ODG_MAX_NU];
  /* kinematics */
  double xpos[1 + ODG_MAX_LEG * ODG_MAX_JL][3], xmat[1 + ODG_MAX_LEG * ODG_MAX_JL][9];
  double xquat[1 + ODG_MAX_LEG * ODG_MAX_JL][4];
  double xipos[1 + ODG_MAX_LEG * ODG_MAX_JL][3];
  double anchor[ODG_MAX_LEG][ODG_MAX_JL][3], axis[ODG_MAX_LEG][ODG_MAX_JL][3];
  /* dynamics */
  double M[ODG_MAX_NV * ODG_MAX_NV];
  double qfrc_bias[ODG_MAX_NV], qfrc_passive[ODG_MAX_NV], qfrc_actuator[ODG_MAX_NV];
  double qfrc_smooth[ODG_MAX_NV], qacc_smooth[ODG_MAX_NV], qacc[ODG_MAX_NV], qfrc_constraint[ODG_MAX_NV];
  double actuator_force[ODG_MAX_NU];
  /* constraints */
  int ncon, nefc, solver_iter;
  OdgoContact contact[ODGO_MAX_CON];
  int efc_type[ODGO_MAX_EFC], efc_id[ODGO_MAX_EFC];
  double efc_J[ODGO_MAX_EFC * ODG_MAX_NV];
  double efc_pos[ODGO_MAX_EFC], efc_margin[ODGO_MAX_EFC], efc_vel[ODGO_MAX_EFC];
  double efc_aref[ODGO_MAX_EFC], efc_R[ODGO_MAX_EFC], efc_D[ODGO_MAX_EFC];
  double efc_frictionloss[ODGO_MAX_EFC], efc_force[ODGO_MAX_EFC];
  double solver_cost, solver_gradnorm;
  /* How close the last forward pass came to taking a DIFFERENT discrete decision (diagnostics for the parity tests:
   * an fp32 implementation may legitimately decide the other way when one of these is at rounding level):
   *   gap_contact  min |dist - margin| over every support point tested for inclusion (m)
   *   gap_support  min lead of a winning support vertex over the runner-up of the same search (m)
   *   gap_limit    min |q - bound| over limited joints (rad) */
  double gap_contact, gap_support, gap_limit;
  /* broad-phase cache: bounding sphere of every hull in its link frame (filled on first use; model constants) */
  double bs_center[ODG_MAX_GEOM][3], bs_radius[ODG_MAX_GEOM];
  int bs_ready;
  /* mj_rnePostConstraint (odgo_cfrc_ext): [torque, force] per body, body 0 = trunk (MuJoCo body id - 1) */
  double cfrc_ext[1 + ODG_MAX_LEG * ODG_MAX_JL][6];
} OdgoData;

/* Walk-environment state that lives outside mjData in the reference (WalkEnvironmentV0 +
 * WalkEnvironmentRewardCalc attributes). */
typedef struct OdgoWalkEnv {
  const OdgModel* m;
  OdgoData d;
  uint64_t seed;
  uint32_t env_id;
  uint32_t episode;          /* number of resets so far: Philox counter word */
  int step;                  /* WalkEnvironmentV0._step */
  int frame_skip;
  int max_steps;             /* 750 */
  float last_action[8];      /* utils.last_action (scaled ctrl target, float32) */
  int last_action_is_reset;  /* 1 right after reset_model (np.zeros(8), float64) */
  int gait_index, gait_matches; /* utils.current_pattern_index / consecutive_matches (never reset) */
  double desired_velocity[3];
} OdgoWalkEnv;

typedef struct OdgoWalkInfo {
  double x_position, y_position, distance_from_origin;
  double paw_contact_forces[4][6];
  double patterns_matches;
  double linear_vel_tracking_reward, reward_ctrl;
  int paws_in_ground[4];
  int gait_first_call;       /* value used in the reward */
  double reward_terms[6];    /* lin_track, safe_range, gait, joint_cost, action_rate, y_cost (unweighted) */
  double min_gap[3];         /* min over the substeps of this env-step of OdgoData.gap_contact / gap_support / gap_limit */
} OdgoWalkInfo;

int  odgo_sizeof_model(void);
int  odgo_sizeof_data(void);
int  odgo_sizeof_walkenv(void);

/* physics: mj_resetData / mj_forward / mj_step for this model class */
void odgo_reset_data(const OdgModel* m, OdgoData* d);
void odgo_forward(const OdgModel* m, OdgoData* d);
void odgo_step(const OdgModel* m, OdgoData* d);
/* pieces, exposed for tests */
void odgo_kinematics(const OdgModel* m, OdgoData* d);
void odgo_mass_matrix(const OdgModel* m, OdgoData* d);
void odgo_bias(const OdgModel* m, OdgoData* d);
void odgo_collision(const OdgModel* m, OdgoData* d);
void odgo_cfrc_ext(const OdgModel* m, OdgoData* d);
double odgo_constraint_cost(const OdgModel* m, const OdgoData* d, const double* qacc);

/* Newton stopping rules: oracle-tight by default; bench.py's CPU baseline legs switch to MuJoCo's defaults
 * (tolerance 1e-8 on scaled gradient or improvement, ls_tolerance 0.01, ls_iterations 50). Process-wide. */
void odgo_set_solver(double tolerance, double ls_tolerance, int ls_iterations, int improvement_exit);

/* counter-based RNG shared bit-for-bit with the CUDA library */
void odgo_philox4x32(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]);
float odgo_u01(uint32_t x);

/* WalkEnvironmentV0 wrapped in ScaleActionWrapper */
void odgo_walk_init(OdgoWalkEnv* e, const OdgModel* m, uint64_t seed, uint32_t env_id);
void odgo_walk_reset(OdgoWalkEnv* e, double* obs33);
void odgo_walk_scale_action(const float* action, float* scaled);
void odgo_walk_step(OdgoWalkEnv* e, const float* action, double* obs33, double* reward,
                    int* terminated, int* truncated, OdgoWalkInfo* info);
/* evaluate obs/reward/termination on the CURRENT state after one mj_forward (no integration) */
void odgo_walk_evaluate(OdgoWalkEnv* e, const float* scaled_action, double* obs33, double* reward,
                        int* terminated, int* truncated, OdgoWalkInfo* info);
/* SB3 VecEnv worker semantics: step, and on done store the terminal obs and reset */
void odgo_walk_step_autoreset(OdgoWalkEnv* e, const float* action, double* obs33, double* reward,
                              int* done, int* truncated, double* terminal_obs33, OdgoWalkInfo* info);

void odgo_walk_step_autoreset_batch(OdgoWalkEnv** envs, int n, const float* actions, double* obs, double* reward, int* done);

#ifdef __cplusplus
}
#endif
#endif
