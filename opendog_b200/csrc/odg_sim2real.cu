// odg_sim2real.cu — the QuadrupedEnv surface (include/odg_sim2real.h) on top of the fused physics kernel.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <new>
#include <string>

#include "../../include/odg_sim2real.h"
#include "odg_sim_internal.h"

namespace {

using odg_internal::set_error;
#define CUDA_TRY(expr)                                                                         \
  do { cudaError_t e_ = (expr);                                                                \
       if (e_ != cudaSuccess) return set_error(ODG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

constexpr int kObs = 22;          // largest observation (train.py: 22, train2.py: 12)

struct S2RConst {
  int N, stride, nq, nv, auto_reset;   // N environments, SoA stride of the simulator's state arrays
  int max_steps;            // episode cap of the training loop (train.py:68,539), enforced when auto_reset
  int variant, obs_dim, act_dim;   // OdgS2RVariant
  double init_z_flat, settled_z;   // train2.py: initial_body_z_pos_on_flat (keyframe z), current_initial_body_z_pos (z after the settle)
  int act_id[8];            // ctrl index of ACTUATOR_NAMES_ORDERED[o]  (FR, FL, BR, BL) x (tigh, knee)
  int qidx[8], vidx[8];     // qpos / qvel index of that actuator's joint
  double home[8];           // sim_keyframe_home_qpos_map
  double clo[8], chi[8];    // actuator ctrlrange
  double real_home_deg[8], scale[8];
  double amp;
  double init_y;            // initial_body_y_pos
  double settled_x;
};
struct S2RState {
  int* counter; double* prev_x; double* cum_pos; double* cum_neg; double* prev_net;
  float* last_cmd;          // [N][8] by ctrl index
  float* ctrl;              // [N][8] scratch: the targets handed to the physics kernel
  const float* settled;     // [nq + nv + nv] state after the settle steps
};

// quat_to_ypr (sim2real/train.py:110-118) in double
__device__ void quat_to_ypr(double q0, double q1, double q2, double q3, double& yaw, double& pitch, double& roll) {
  const double sinr = 2 * (q0 * q1 + q2 * q3), cosr = 1 - 2 * (q1 * q1 + q2 * q2);
  roll = atan2(sinr, cosr);
  const double sinp = 2 * (q0 * q2 - q3 * q1);
  pitch = fabs(sinp) < 1 ? asin(sinp) : copysign(1.5707963267948966, sinp);
  const double siny = 2 * (q0 * q3 + q1 * q2), cosy = 1 - 2 * (q2 * q2 + q3 * q3);
  yaw = atan2(siny, cosy);
}

// _get_observation (:184-207)
__device__ void write_obs(const S2RConst& C, const float* qpos, const float* qvel, int env, int counter, float* o) {
  const int N = C.stride;
  double yaw, pitch, roll;
  quat_to_ypr(qpos[3 * N + env], qpos[4 * N + env], qpos[5 * N + env], qpos[6 * N + env], yaw, pitch, roll);
  o[0] = (float)yaw; o[1] = (float)pitch; o[2] = (float)roll;
  if (C.variant == ODG_S2R_TERRAIN) {              // train2.py:197-201: [yaw pitch roll, q_j - home (8), v_x]
    for (int k = 0; k < 8; k++) o[3 + k] = (float)((double)qpos[C.qidx[k] * N + env] - C.home[k]);
    o[11] = qvel[0 * N + env];
    return;
  }
  for (int k = 0; k < 8; k++) {
    o[3 + k] = (float)((double)qpos[C.qidx[k] * N + env] - C.home[k]);
    o[11 + k] = qvel[C.vidx[k] * N + env];
  }
  o[19] = qvel[0 * N + env];
  const double phase = (double)(counter % 2);               // PHASE_CYCLE_DURATION_POLICY_STEPS = 2
  o[20] = (float)sin(phase * 3.141592653589793); o[21] = (float)cos(phase * 3.141592653589793);
}

// _apply_actions_and_step, command part (:235-279) + the scaling of step() (:292-295)
__global__ void k_s2r_pre(const S2RConst C, const S2RState S, const float* __restrict__ action) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= C.N) return;
  if (C.variant == ODG_S2R_TERRAIN) {              // train2.py:348-354: home + amplitude * a, clipped to ctrlrange
    for (int o = 0; o < 8; o++) {
      const double t = fmin(fmax(C.home[o] + (double)action[env * 8 + o] * C.amp, C.clo[o]), C.chi[o]);
      S.ctrl[env * 8 + C.act_id[o]] = (float)t;
    }
    return;
  }
  const int phase = S.counter[env] % 2;
  const double fr_t = (double)action[env * 4 + 0] * C.amp, k1 = (double)action[env * 4 + 1] * C.amp;
  const double fl_t = (double)action[env * 4 + 2] * C.amp, k2 = (double)action[env * 4 + 3] * C.amp;
  // ordered: FR_t FR_k FL_t FL_k BR_t BR_k BL_t BL_k
  double d[8];
  d[0] = fr_t; d[2] = fl_t; d[6] = fr_t; d[4] = fl_t;       // BL mirrors FR, BR mirrors FL
  d[1] = d[3] = d[5] = d[7] = 0.0;
  if (phase == 0) { d[1] = k1; d[7] = -k1; } else { d[3] = k2; d[5] = -k2; }
  for (int o = 0; o < 8; o++) {
    const double t = fmin(fmax(C.home[o] + d[o], C.clo[o]), C.chi[o]);
    S.ctrl[env * 8 + C.act_id[o]] = (float)t;
  }
}

__device__ void restore_settled(const S2RConst& C, const S2RState& S, const odg::SimPtrs& P, int env) {
  const int N = C.stride;
  for (int i = 0; i < C.nq; i++) P.qpos[i * N + env] = S.settled[i];
  for (int i = 0; i < C.nv; i++) { P.qvel[i * N + env] = S.settled[C.nq + i]; P.warm[i * N + env] = S.settled[C.nq + C.nv + i]; }
  S.counter[env] = 0; S.prev_x[env] = (double)S.settled[0];
  S.cum_pos[env] = 0.0; S.cum_neg[env] = 0.0; S.prev_net[env] = 0.0;
}

// everything of step() after the physics (:299-411)
__global__ void k_s2r_post(const S2RConst C, const S2RState S, const odg::SimPtrs P, float* __restrict__ obs,
                           float* __restrict__ reward, unsigned char* __restrict__ done_out, unsigned char* __restrict__ reason_out,
                           float* __restrict__ sim_target, float* __restrict__ terminal_obs, const float* __restrict__ home_ctrl) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = C.stride;
  if (env >= C.N) return;
  const int phase_action = S.counter[env] % 2;
  const int counter = S.counter[env] + 1;
  bool finite = true;
  for (int i = 0; i < C.nq; i++) finite = finite && isfinite(P.qpos[i * N + env]);
  for (int i = 0; i < C.nv; i++) finite = finite && isfinite(P.qvel[i * N + env]);
  const double x = P.qpos[0 * N + env], y = P.qpos[1 * N + env];
  const double dx = x - S.prev_x[env];
  double cpos = S.cum_pos[env], cneg = S.cum_neg[env];
  if (dx > 0) cpos += dx; else if (dx < 0) cneg += fabs(dx);
  float o[kObs];
  write_obs(C, P.qpos, P.qvel, env, counter, o);
  const double vx = P.qvel[0 * N + env], vy = P.qvel[1 * N + env];
  if (C.variant == ODG_S2R_TERRAIN) {
    // train2.py:357-414
    const double net = cpos - cneg, dnet = net - S.prev_net[env];
    double r = 450.0 * vx;
    if (dnet > 0.0005) r += 20.0 * dnet;
    if (vx < -0.005) r += -9.0 * fabs(vx);
    r += 0.005; r += 0.01;
    r += -0.3 * fabs(vy);
    r += -0.5 * fabs(vy);
    r += -0.15 * fabs(y - C.init_y);
    const double z = P.qpos[2 * N + env];
    const double zs = z - C.settled_z, zi = z - C.init_z_flat;
    double pz = 0.0;
    if (zs < -0.03) pz -= (0.25 * 0.5) * (fabs(zs) - 0.03) * (fabs(zs) - 0.03);
    if (fabs(zi) > 0.05) pz -= (0.25 * 0.25) * (fabs(zi) - 0.05) * (fabs(zi) - 0.05);
    r += pz;
    double yaw, pitch, roll;
    quat_to_ypr(P.qpos[3 * N + env], P.qpos[4 * N + env], P.qpos[5 * N + env], P.qpos[6 * N + env], yaw, pitch, roll);
    const double th = 15.0 * 0.017453292519943295, thy = 35.0 * 0.017453292519943295, lim = 35.0 * 0.017453292519943295;
    double po = 0.0;
    if (fabs(roll) > th) po += -0.08 * (fabs(roll) - th) * (fabs(roll) - th);
    if (fabs(pitch) > th) po += -0.08 * (fabs(pitch) - th) * (fabs(pitch) - th);
    if (fabs(yaw) > thy) po += -0.08 * (fabs(yaw) - thy) * (fabs(yaw) - thy);
    r += po;
    double dsq = 0.0;
    for (int k = 0; k < 8; k++) {
      const double d = (double)S.ctrl[env * 8 + C.act_id[k]] - (double)S.last_cmd[env * 8 + C.act_id[k]];
      dsq += d * d;
    }
    r += -0.005 * dsq;
    r += dx > 0 ? 70.0 * dx : (dx < 0.0005 ? -1.0 : 0.0);                     // reward_step_displacement (:363-366)
    double jvm = 0.0;
    for (int i = 7; i < 15 && i < C.nv; i++) jvm += fabs((double)P.qvel[i * N + env]);   // np.abs(qvel[7:15]): 7 of the 8 hinges (:398)
    r += -0.05 * exp(-jvm * 5.0);
    bool done = false; int reason = ODG_S2R_RUNNING;
    if (!finite) { r -= 50.0; done = true; reason = ODG_S2R_MJ_ERROR; }
    if (!done && (fabs(roll) > lim || fabs(pitch) > lim || fabs(yaw) > lim * 1.5)) { r -= 150.0; done = true; reason = ODG_S2R_ORIENTATION_LIMIT; }
    if (!done && cpos > 0.05 && cneg > 0.85 * cpos) { r -= 50.0; done = true; reason = ODG_S2R_TOO_MUCH_BACKWARD; }
    if (!done && C.auto_reset && C.max_steps > 0 && counter >= C.max_steps) done = true;
    if (reward) reward[env] = (float)r;
    if (done_out) done_out[env] = done ? 1 : 0;
    if (reason_out) reason_out[env] = (unsigned char)reason;
    if (sim_target) for (int u = 0; u < 8; u++) sim_target[env * 8 + u] = S.ctrl[env * 8 + u];
    if (terminal_obs) for (int k = 0; k < C.obs_dim; k++) terminal_obs[env * C.obs_dim + k] = o[k];
    S.counter[env] = counter; S.prev_x[env] = x; S.cum_pos[env] = cpos; S.cum_neg[env] = cneg; S.prev_net[env] = net;
    for (int u = 0; u < 8; u++) S.last_cmd[env * 8 + u] = S.ctrl[env * 8 + u];
    if (done && C.auto_reset) {
      restore_settled(C, S, P, env);
      for (int u = 0; u < 8; u++) S.last_cmd[env * 8 + u] = home_ctrl[u];
      write_obs(C, P.qpos, P.qvel, env, 0, o);
    }
    if (obs) for (int k = 0; k < C.obs_dim; k++) obs[env * C.obs_dim + k] = o[k];
    return;
  }
  double r = 150.0 * vx;
  const double net = cpos - cneg, dnet = net - S.prev_net[env];
  if (dnet > 0.0005) r += 15.0 * dnet;
  if (vx < -0.005) r += -5.0 * fabs(vx);
  r += 0.05;
  r += -0.2 * fabs(vy);
  r += -0.1 * fabs(y - C.init_y);
  double yaw, pitch, roll;
  quat_to_ypr(P.qpos[3 * N + env], P.qpos[4 * N + env], P.qpos[5 * N + env], P.qpos[6 * N + env], yaw, pitch, roll);
  const double th = 5.0 * 0.017453292519943295, thy = 10.0 * 0.017453292519943295, lim = 25.0 * 0.017453292519943295;
  double ro = 0.0;
  if (fabs(roll) > th) ro += -0.05 * (fabs(roll) - th) * (fabs(roll) - th);
  if (fabs(pitch) > th) ro += -0.05 * (fabs(pitch) - th) * (fabs(pitch) - th);
  if (fabs(yaw) > thy) ro += -0.05 * (fabs(yaw) - thy) * (fabs(yaw) - thy);
  r += ro;
  double dsq = 0.0;
  for (int k = 0; k < 8; k++) {
    const double d = (double)S.ctrl[env * 8 + C.act_id[k]] - (double)S.last_cmd[env * 8 + C.act_id[k]];
    dsq += d * d;
  }
  r += -0.01 * dsq;
  // leg positioning (:340-383): legs in ordered pairs FR(0,1) FL(2,3) BR(4,5) BL(6,7); FR/BL swing in phase 0
  int too_far = 0, not_home = 0;
  for (int leg = 0; leg < 4; leg++) {
    const bool swinging = (phase_action == 0) ? (leg == 0 || leg == 3) : (leg == 1 || leg == 2);
    double maxdev = 0.0; bool at_home = true;
    for (int j = 0; j < 2; j++) {
      const int k = leg * 2 + j;
      const double cmd = (double)S.ctrl[env * 8 + C.act_id[k]];
      const double real_target = C.real_home_deg[k] + C.scale[k] * ((cmd - C.home[k]) * 57.29577951308232);
      const double dev = fabs(real_target - C.real_home_deg[k]);
      maxdev = fmax(maxdev, dev);
      if (dev > 15.0) at_home = false;
    }
    if (swinging) { if (maxdev > 40.0) too_far++; } else if (!at_home) not_home++;
  }
  if (too_far > 0 || not_home > 0) r += -(double)(too_far + not_home) * 0.5;
  bool done = false; int reason = ODG_S2R_RUNNING;
  if (!finite) { r -= 20.0; done = true; reason = ODG_S2R_MJ_ERROR; }
  if (fabs(roll) > lim || fabs(pitch) > lim || fabs(yaw) > lim) { r -= 5.0; done = true; reason = ODG_S2R_ORIENTATION_LIMIT; }
  if (!done && cpos > 0.05 && cneg > 0.75 * cpos) { r -= 5.0; done = true; reason = ODG_S2R_TOO_MUCH_BACKWARD; }
  // the training loop's `for step_num in range(MAX_STEPS_PER_EPISODE)` (:539): episode over, no penalty, reason "max_steps"
  if (!done && C.auto_reset && C.max_steps > 0 && counter >= C.max_steps) done = true;
  if (reward) reward[env] = (float)r;
  if (done_out) done_out[env] = done ? 1 : 0;
  if (reason_out) reason_out[env] = (unsigned char)reason;
  if (sim_target) for (int u = 0; u < 8; u++) sim_target[env * 8 + u] = S.ctrl[env * 8 + u];
  if (terminal_obs) for (int k = 0; k < C.obs_dim; k++) terminal_obs[env * C.obs_dim + k] = o[k];
  // bookkeeping for the next step
  S.counter[env] = counter; S.prev_x[env] = x; S.cum_pos[env] = cpos; S.cum_neg[env] = cneg; S.prev_net[env] = net;
  for (int u = 0; u < 8; u++) S.last_cmd[env * 8 + u] = S.ctrl[env * 8 + u];
  if (done && C.auto_reset) {
    restore_settled(C, S, P, env);
    for (int u = 0; u < 8; u++) S.last_cmd[env * 8 + u] = home_ctrl[u];
    write_obs(C, P.qpos, P.qvel, env, 0, o);
  }
  if (obs) for (int k = 0; k < C.obs_dim; k++) obs[env * C.obs_dim + k] = o[k];
}

__global__ void k_s2r_reset(const S2RConst C, const S2RState S, const odg::SimPtrs P, const unsigned char* __restrict__ mask,
                            float* __restrict__ obs, const float* __restrict__ home_ctrl) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= C.N) return;
  if (mask && !mask[env]) return;
  restore_settled(C, S, P, env);
  for (int u = 0; u < 8; u++) S.last_cmd[env * 8 + u] = home_ctrl[u];
  if (obs) { float o[kObs]; write_obs(C, P.qpos, P.qvel, env, 0, o); for (int k = 0; k < C.obs_dim; k++) obs[env * C.obs_dim + k] = o[k]; }
}
__global__ void k_s2r_fill_ctrl(float* ctrl, const float* home_ctrl, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * 8) ctrl[i] = home_ctrl[i & 7];
}
__global__ void k_s2r_snapshot(const odg::SimPtrs P, int nq, int nv, float* settled) {
  const int i = threadIdx.x;
  const int N = P.N;
  if (i < nq) settled[i] = P.qpos[i * N];
  if (i < nv) { settled[nq + i] = P.qvel[i * N]; settled[nq + nv + i] = P.warm[i * N]; }
}

}  // namespace

struct DevScope {          // run an entry point on the handle's device, whatever the caller's current device is
  int prev = -1;
  explicit DevScope(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DevScope() { if (prev >= 0) cudaSetDevice(prev); }
};

struct OdgS2R {
  OdgSim* sim = nullptr;
  S2RConst C{};
  S2RState S{};
  void* d_slab = nullptr;
  float* d_settled = nullptr; float* d_home_ctrl = nullptr;
};

extern "C" {

void odg_s2r_default_config_for(OdgS2RConfig* c, int variant) {
  if (!c) return;
  odg_s2r_default_config(c);
  if (variant == ODG_S2R_TERRAIN) {
    c->variant = ODG_S2R_TERRAIN;
    c->action_amplitude_rad = 50.0 * 3.14159265358979323846 / 180.0;     // ACTION_AMPLITUDE_DEG (train2.py:93)
    c->max_steps = 1000;                                                  // MAX_STEPS_PER_EPISODE (train2.py:88)
  }
}
int odg_s2r_obs_dim(const OdgS2R* e) { return e ? e->C.obs_dim : 0; }
int odg_s2r_act_dim(const OdgS2R* e) { return e ? e->C.act_dim : 0; }

void odg_s2r_default_config(OdgS2RConfig* c) {
  if (!c) return;
  c->action_amplitude_rad = 40.0 * 3.14159265358979323846 / 180.0;
  c->settle_steps = 100; c->auto_reset = 0; c->max_steps = 250; c->variant = ODG_S2R_TRAIN;
  const double home[8] = { -45.0, 45.0, 45.0, 45.0, 45.0, -45.0, 45.0, -45.0 };   // FR_t FR_k FL_t FL_k BR_t BR_k BL_t BL_k
  for (int i = 0; i < 8; i++) { c->real_home_deg[i] = home[i]; c->joint_scale[i] = 1.0; }
}

int odg_s2r_create(OdgSim* sim, const OdgModel* m, const OdgS2RConfig* cfg_in, OdgS2R** out) {
  if (!sim || !m || !out) return set_error(ODG_ERR_INVALID, "odg_s2r_create: null argument");
  *out = nullptr;
  OdgS2RConfig cfg;
  if (cfg_in) cfg = *cfg_in; else odg_s2r_default_config(&cfg);
  const odg::DevConst& DC = sim->prep.C;
  if (m->nu != 8 || m->njl != 2) return set_error(ODG_ERR_INVALID, "QuadrupedEnv needs the 8-actuator OpenDOG model");
  if (DC.scale_actions || DC.auto_reset) return set_error(ODG_ERR_INVALID, "create the OdgSim with scale_actions = 0 and auto_reset = 0");
  if (cfg.settle_steps < 1) return set_error(ODG_ERR_INVALID, "settle_steps must be >= 1");
  if (cfg.variant != ODG_S2R_TRAIN && cfg.variant != ODG_S2R_TERRAIN) return set_error(ODG_ERR_INVALID, "unknown OdgS2RVariant");
  if (cfg.max_steps < 0) return set_error(ODG_ERR_INVALID, "max_steps must be >= 0");
  DevScope scope(sim->device);
  OdgS2R* e = new (std::nothrow) OdgS2R();
  if (!e) return set_error(ODG_ERR_ALLOC, "out of host memory");
  e->sim = sim;
  S2RConst& C = e->C;
  C.N = sim->N; C.stride = sim->P.N; C.nq = DC.nq; C.nv = DC.nv; C.auto_reset = cfg.auto_reset; C.max_steps = cfg.max_steps; C.amp = cfg.action_amplitude_rad;
  C.variant = cfg.variant; C.obs_dim = cfg.variant == ODG_S2R_TERRAIN ? 12 : 22; C.act_dim = cfg.variant == ODG_S2R_TERRAIN ? 8 : 4;
  C.init_z_flat = m->key_qpos[2]; C.settled_z = m->key_qpos[2];
  // ACTUATOR_NAMES_ORDERED = FR FL BR BL; model legs are in body order FL FR BL BR
  const int leg_of[4] = { 1, 0, 3, 2 };
  for (int o = 0; o < 8; o++) {
    const int leg = leg_of[o / 2], j = o % 2;
    int u = -1;
    for (int k = 0; k < m->nu; k++) if (m->act_leg[k] == leg && m->act_joint[k] == j) u = k;
    if (u < 0) { delete e; return set_error(ODG_ERR_INVALID, "actuator missing"); }
    C.act_id[o] = u; C.qidx[o] = 7 + leg * m->njl + j; C.vidx[o] = 6 + leg * m->njl + j;
    C.home[o] = m->key_qpos[C.qidx[o]]; C.clo[o] = m->act_ctrlrange[u][0]; C.chi[o] = m->act_ctrlrange[u][1];
    C.real_home_deg[o] = cfg.real_home_deg[o]; C.scale[o] = cfg.joint_scale[o];
  }
  C.init_y = m->key_qpos[1];
  const size_t N = (size_t)sim->N;
  const size_t bytes = N * (4 * sizeof(double) + sizeof(int) + 16 * sizeof(float));
  if (cudaMalloc(&e->d_slab, bytes) != cudaSuccess || cudaMalloc(&e->d_settled, (C.nq + 2 * C.nv) * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&e->d_home_ctrl, 8 * sizeof(float)) != cudaSuccess) { odg_s2r_destroy(e); return set_error(ODG_ERR_ALLOC, "cudaMalloc(s2r) failed"); }
  double* dp = static_cast<double*>(e->d_slab);
  e->S.prev_x = dp; e->S.cum_pos = dp + N; e->S.cum_neg = dp + 2 * N; e->S.prev_net = dp + 3 * N;
  float* fp = reinterpret_cast<float*>(dp + 4 * N);
  e->S.last_cmd = fp; e->S.ctrl = fp + 8 * N;
  e->S.counter = reinterpret_cast<int*>(fp + 16 * N);
  e->S.settled = e->d_settled;
  CUDA_TRY(cudaMemset(e->d_slab, 0, bytes));
  float hc[8];
  for (int u = 0; u < 8; u++) hc[u] = (float)m->key_ctrl[u];
  CUDA_TRY(cudaMemcpy(e->d_home_ctrl, hc, sizeof(hc), cudaMemcpyHostToDevice));
  // settled reset state: keyframe (odg_create left every env there) + settle_steps x mj_step with ctrl = home
  k_s2r_fill_ctrl<<<(unsigned)((N * 8 + 255) / 256), 256>>>(e->S.ctrl, e->d_home_ctrl, sim->N);
  {
    // one launch with frame_skip = settle_steps (the policy-step length need not divide the settle: 100 vs 40 in train2.py)
    const int fs = DC.frame_skip;
    odg_set_frame_skip(sim, cfg.settle_steps);
    int rc = odg_step(sim, e->S.ctrl, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    odg_set_frame_skip(sim, fs);
    if (rc != ODG_OK) { odg_s2r_destroy(e); return rc; }
  }
  k_s2r_snapshot<<<1, 32>>>(sim->P, C.nq, C.nv, e->d_settled);
  CUDA_TRY(cudaDeviceSynchronize());
  {
    float z = 0.f;                                   // current_initial_body_z_pos (train2.py:318): trunk height after the settle
    CUDA_TRY(cudaMemcpy(&z, e->d_settled + 2, sizeof(float), cudaMemcpyDeviceToHost));
    C.settled_z = (double)z;
  }
  *out = e;
  return odg_s2r_reset(e, nullptr, nullptr, nullptr);
}

void odg_s2r_destroy(OdgS2R* e) {
  if (!e) return;
  DevScope scope(e->sim->device);
  cudaFree(e->d_slab); cudaFree(e->d_settled); cudaFree(e->d_home_ctrl);
  delete e;
}

int odg_s2r_reset(OdgS2R* e, const uint8_t* mask_dev, float* obs_dev, void* stream) {
  if (!e) return set_error(ODG_ERR_INVALID, "odg_s2r_reset: null handle");
  DevScope scope(e->sim->device);
  const int N = e->C.N;
  k_s2r_reset<<<(N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(e->C, e->S, e->sim->P, mask_dev, obs_dev, e->d_home_ctrl);
  e->sim->launches++;
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

int odg_s2r_step(OdgS2R* e, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                 uint8_t* reason_dev, float* sim_target_rad_dev, float* terminal_obs_dev, void* stream) {
  if (!e || !action_dev) return set_error(ODG_ERR_INVALID, "odg_s2r_step: null handle or action");
  DevScope scope(e->sim->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = e->C.N;
  k_s2r_pre<<<(N + 127) / 128, 128, 0, st>>>(e->C, e->S, action_dev);
  int rc = odg_step(e->sim, e->S.ctrl, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
  if (rc != ODG_OK) return rc;
  k_s2r_post<<<(N + 127) / 128, 128, 0, st>>>(e->C, e->S, e->sim->P, obs_dev, reward_dev, done_dev, reason_dev,
                                               sim_target_rad_dev, terminal_obs_dev, e->d_home_ctrl);
  e->sim->launches += 2;
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

int odg_s2r_set_bookkeeping(OdgS2R* e, const int32_t* counter, const double* prev_x, const double* cum_pos,
                            const double* cum_neg, const double* prev_net, const float* last_cmd, void* stream) {
  if (!e) return set_error(ODG_ERR_INVALID, "odg_s2r_set_bookkeeping: null handle");
  DevScope scope(e->sim->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t N = (size_t)e->C.N;
  if (counter) CUDA_TRY(cudaMemcpyAsync(e->S.counter, counter, N * 4, cudaMemcpyDeviceToDevice, st));
  if (prev_x) CUDA_TRY(cudaMemcpyAsync(e->S.prev_x, prev_x, N * 8, cudaMemcpyDeviceToDevice, st));
  if (cum_pos) CUDA_TRY(cudaMemcpyAsync(e->S.cum_pos, cum_pos, N * 8, cudaMemcpyDeviceToDevice, st));
  if (cum_neg) CUDA_TRY(cudaMemcpyAsync(e->S.cum_neg, cum_neg, N * 8, cudaMemcpyDeviceToDevice, st));
  if (prev_net) CUDA_TRY(cudaMemcpyAsync(e->S.prev_net, prev_net, N * 8, cudaMemcpyDeviceToDevice, st));
  if (last_cmd) CUDA_TRY(cudaMemcpyAsync(e->S.last_cmd, last_cmd, N * 8 * 4, cudaMemcpyDeviceToDevice, st));
  return ODG_OK;
}

}  // extern "C"
