/* odg_oracle.c — CPU fp64 ORACLE (test infrastructure, see odg_oracle.h: PARITY UNPINNED at the single-step level,
 * loosely pinned against the reference's shipped MuJoCo walk files).
 *
 * Restates, for "free trunk + hinge leg chains over a floor plane" models, what the reference's
 * hot path executes inside the third-party `mujoco==3.2.3` wheel and its own Python reward code:
 *
 *   mj_step            <- WalkEnvironment.py:58 (do_simulation, frame_skip 10), sim2real/train.py:281-284
 *   obs/reward/done    <- WalkEnvironment.py:56-151, rewards/walk_environment_reward_calc.py:117-390
 *   action scaling     <- ScaleActionEnvironment.py:8-23
 *
 * The physics pipeline follows MuJoCo's documented computation order [3P-recalled]:
 * kinematics -> composite-rigid-body mass matrix -> RNE bias -> passive -> actuation -> smooth
 * acceleration -> collision (plane vs convex hull) -> constraint rows (dof friction loss, joint
 * limits, elliptic-cone contacts; solref/solimp impedance) -> primal Newton solve of the convex
 * constraint problem -> semi-implicit Euler.
 *
 * Written with dense nv x nv algebra on purpose: it shares no structure with the CUDA kernel's
 * block-arrow / leg-per-lane formulation, so agreement between the two is meaningful.
 *
 * Build: see oracle/Makefile (-ffp-contract=off: the float32 pieces must round like numpy).
 */
#include "odg_oracle.h"

#include <math.h>
#include <string.h>

#define NV_MAX ODG_MAX_NV
#define MINVAL 1e-15
#define ODG_PI 3.14159265358979323846
#define MINIMP 0.0001
#define MAXIMP 0.9999

/* ------------------------------------------------------------------------------------------ */
/* small vector helpers                                                                        */
static void v3_set(double* r, double x, double y, double z) { r[0] = x; r[1] = y; r[2] = z; }
static void v3_copy(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static void v3_add(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static void v3_sub(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static void v3_addscl(double* r, const double* a, const double* b, double s) { r[0] = a[0] + s * b[0]; r[1] = a[1] + s * b[1]; r[2] = a[2] + s * b[2]; }
static double v3_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void v3_cross(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static void m3_mulv(double* r, const double* M, const double* v) {
  double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
  double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
  double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void m3_tmulv(double* r, const double* M, const double* v) {
  double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
  double y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
  double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void m3_col(double* r, const double* M, int k) { r[0] = M[k]; r[1] = M[3 + k]; r[2] = M[6 + k]; }
static void quat_norm(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }   /* mju_normalize4 */
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
static void quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat2mat(double* M, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  M[0] = w * w + x * x - y * y - z * z; M[1] = 2 * (x * y - w * z); M[2] = 2 * (x * z + w * y);
  M[3] = 2 * (x * y + w * z); M[4] = w * w - x * x + y * y - z * z; M[5] = 2 * (y * z - w * x);
  M[6] = 2 * (x * z - w * y); M[7] = 2 * (y * z + w * x); M[8] = w * w - x * x - y * y + z * z;
}
static void axisangle2quat(double* q, const double* axis, double angle) {
  double s = sin(0.5 * angle);
  q[0] = cos(0.5 * angle); q[1] = s * axis[0]; q[2] = s * axis[1]; q[3] = s * axis[2];
}

/* dense symmetric positive-definite solve (lower Cholesky, in place) */
static int chol_factor(double* A, int n, int ld) {
  for (int j = 0; j < n; j++) {
    double s = A[j * ld + j];
    for (int k = 0; k < j; k++) s -= A[j * ld + k] * A[j * ld + k];
    if (s <= 0) return -1;
    s = sqrt(s);
    A[j * ld + j] = s;
    for (int i = j + 1; i < n; i++) {
      double t = A[i * ld + j];
      for (int k = 0; k < j; k++) t -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = t / s;
    }
  }
  return 0;
}
static void chol_solve(const double* L, int n, int ld, double* x) {
  for (int i = 0; i < n; i++) {
    double t = x[i];
    for (int k = 0; k < i; k++) t -= L[i * ld + k] * x[k];
    x[i] = t / L[i * ld + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double t = x[i];
    for (int k = i + 1; k < n; k++) t -= L[k * ld + i] * x[k];
    x[i] = t / L[i * ld + i];
  }
}

int odgo_sizeof_model(void) { return (int)sizeof(OdgModel); }
int odgo_sizeof_data(void) { return (int)sizeof(OdgoData); }
int odgo_sizeof_walkenv(void) { return (int)sizeof(OdgoWalkEnv); }

/* body index: 0 = trunk, 1 + leg*njl + j = leg link; dof index: 0..5 trunk, 6 + leg*njl + j */
static int body_of(const OdgModel* m, int leg, int link) { return leg < 0 ? 0 : 1 + leg * m->njl + link; }
static int dof_of(const OdgModel* m, int leg, int j) { return 6 + leg * m->njl + j; }

/* ------------------------------------------------------------------------------------------ */
void odgo_reset_data(const OdgModel* m, OdgoData* d) {
  /* mj_resetData: qpos = qpos0, everything else zero. (The env overwrites qpos right after.) */
  memset(d, 0, sizeof(*d));
  d->qpos[2] = 0.0; d->qpos[3] = 1.0;
  (void)m;
}

/* mj_kinematics (+ the part of mj_comPos that places the inertial frames) */
void odgo_kinematics(const OdgModel* m, OdgoData* d) {
  quat_norm(d->qpos + 3);                       /* mj_kinematics normalises qpos quaternions in place */
  v3_copy(d->xpos[0], d->qpos);
  memcpy(d->xquat[0], d->qpos + 3, 4 * sizeof(double));
  quat2mat(d->xmat[0], d->xquat[0]);
  double t[3];
  m3_mulv(t, d->xmat[0], m->base_ipos);
  v3_add(d->xipos[0], d->xpos[0], t);
  for (int l = 0; l < m->nleg; l++)
    for (int j = 0; j < m->njl; j++) {
      int b = body_of(m, l, j), p = j == 0 ? 0 : b - 1;
      double pos[3], quat[4], R[9], qloc[4], q2[4];
      m3_mulv(t, d->xmat[p], m->body_pos[l][j]);
      v3_add(pos, d->xpos[p], t);
      quat_mul(quat, d->xquat[p], m->body_quat[l][j]);
      quat2mat(R, quat);
      m3_mulv(t, R, m->jnt_pos[l][j]);
      v3_add(d->anchor[l][j], pos, t);
      m3_mulv(d->axis[l][j], R, m->jnt_axis[l][j]);
      axisangle2quat(qloc, m->jnt_axis[l][j], d->qpos[7 + l * m->njl + j]);   /* ref = 0 */
      quat_mul(q2, quat, qloc);
      quat_norm(q2);
      memcpy(d->xquat[b], q2, sizeof(q2));
      quat2mat(d->xmat[b], q2);
      m3_mulv(t, d->xmat[b], m->jnt_pos[l][j]);                               /* off-centre rotation */
      v3_sub(d->xpos[b], d->anchor[l][j], t);
      m3_mulv(t, d->xmat[b], m->ipos[l][j]);
      v3_add(d->xipos[b], d->xpos[b], t);
    }
}

/* spatial inertia about a world-aligned frame at point O: mass, h = m*c, IO (3x3) */
typedef struct { double m, h[3], I[9]; } SpI;
static void spi_body(SpI* s, double mass, const double* inertia_local, const double* R, const double* c) {
  double RI[9], Iw[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double a = 0; for (int k = 0; k < 3; k++) a += R[i * 3 + k] * inertia_local[k * 3 + j];
    RI[i * 3 + j] = a;
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double a = 0; for (int k = 0; k < 3; k++) a += RI[i * 3 + k] * R[j * 3 + k];
    Iw[i * 3 + j] = a;
  }
  double cc = v3_dot(c, c);
  s->m = mass;
  for (int i = 0; i < 3; i++) {
    s->h[i] = mass * c[i];
    for (int j = 0; j < 3; j++) s->I[i * 3 + j] = Iw[i * 3 + j] + mass * ((i == j ? cc : 0.0) - c[i] * c[j]);
  }
}
static void spi_add(SpI* a, const SpI* b) {
  a->m += b->m;
  for (int i = 0; i < 3; i++) a->h[i] += b->h[i];
  for (int i = 0; i < 9; i++) a->I[i] += b->I[i];
}
/* momentum (L about O, p) of spatial inertia s moving with twist (w, vO) */
static void spi_apply(const SpI* s, const double* w, const double* v, double* L, double* p) {
  double t[3];
  v3_cross(t, w, s->h);
  for (int i = 0; i < 3; i++) p[i] = s->m * v[i] + t[i];
  m3_mulv(L, s->I, w);
  v3_cross(t, s->h, v);
  v3_add(L, L, t);
}

/* motion axis of dof i as a twist (w, vO) about O = trunk origin */
static void dof_twist(const OdgModel* m, const OdgoData* d, int dof, double* w, double* v) {
  if (dof < 3) { v3_set(w, 0, 0, 0); v3_set(v, 0, 0, 0); v[dof] = 1; return; }
  if (dof < 6) { m3_col(w, d->xmat[0], dof - 3); v3_set(v, 0, 0, 0); return; }
  int l = (dof - 6) / m->njl, j = (dof - 6) % m->njl;
  double r[3];
  v3_copy(w, d->axis[l][j]);
  v3_sub(r, d->anchor[l][j], d->xpos[0]);
  v3_cross(v, r, w);                               /* velocity at O of a rotation about the anchor */
}

/* mj_crb: composite-rigid-body mass matrix (dense, symmetric) + armature */
void odgo_mass_matrix(const OdgModel* m, OdgoData* d) {
  int nv = m->nv;
  SpI comp[1 + ODG_MAX_LEG * ODG_MAX_JL];
  double c[3];
  v3_sub(c, d->xipos[0], d->xpos[0]);
  spi_body(&comp[0], m->base_mass, m->base_inertia, d->xmat[0], c);
  for (int l = 0; l < m->nleg; l++) {
    for (int j = 0; j < m->njl; j++) {
      int b = body_of(m, l, j);
      v3_sub(c, d->xipos[b], d->xpos[0]);
      spi_body(&comp[b], m->mass[l][j], m->inertia[l][j], d->xmat[b], c);
    }
    for (int j = m->njl - 2; j >= 0; j--) spi_add(&comp[body_of(m, l, j)], &comp[body_of(m, l, j + 1)]);
    spi_add(&comp[0], &comp[body_of(m, l, 0)]);
  }
  memset(d->M, 0, sizeof(d->M));
  for (int i = 0; i < nv; i++) {
    int bi = i < 6 ? 0 : 1 + (i - 6);
    double w[3], v[3], L[3], p[3];
    dof_twist(m, d, i, w, v);
    spi_apply(&comp[bi], w, v, L, p);
    /* ancestors of dof i (including itself): trunk dofs, and for a leg dof the dofs above it in its leg */
    for (int j = 0; j <= i; j++) {
      int anc = j < 6 && i < 6 ? 1 : 0;
      if (i >= 6) {
        int li = (i - 6) / m->njl;
        anc = j < 6 || ((j - 6) / m->njl == li);
      }
      if (!anc) continue;
      double wj[3], vj[3];
      dof_twist(m, d, j, wj, vj);
      double val = v3_dot(wj, L) + v3_dot(vj, p);
      d->M[i * nv + j] = val; d->M[j * nv + i] = val;
    }
  }
  for (int i = 0; i < 6; i++) d->M[i * nv + i] += m->base_armature[i];
  for (int l = 0; l < m->nleg; l++)
    for (int j = 0; j < m->njl; j++) { int k = dof_of(m, l, j); d->M[k * nv + k] += m->armature[l][j]; }
}

/* translational Jacobian (3 x nv) of world point p attached to body (leg,link); leg<0 = trunk */
static void jac_point(const OdgModel* m, const OdgoData* d, int leg, int link, const double* p, double* Jp) {
  int nv = m->nv;
  memset(Jp, 0, 3 * nv * sizeof(double));
  double r[3], col[3], ax[3];
  v3_sub(r, p, d->xpos[0]);
  for (int i = 0; i < 3; i++) Jp[i * nv + i] = 1;
  for (int i = 0; i < 3; i++) {
    m3_col(ax, d->xmat[0], i);
    v3_cross(col, ax, r);
    for (int k = 0; k < 3; k++) Jp[k * nv + 3 + i] = col[k];
  }
  if (leg >= 0)
    for (int j = 0; j <= link; j++) {
      v3_sub(r, p, d->anchor[leg][j]);
      v3_cross(col, d->axis[leg][j], r);
      for (int k = 0; k < 3; k++) Jp[k * nv + dof_of(m, leg, j)] = col[k];
    }
}

/* rotational Jacobian of the body (leg, link): mj_jacBody's jacr — trunk rotational DoFs are in the trunk frame */
static void jac_rot(const OdgModel* m, const OdgoData* d, int leg, int link, double* Jr) {
  int nv = m->nv;
  memset(Jr, 0, 3 * nv * sizeof(double));
  double ax[3];
  for (int i = 0; i < 3; i++) {
    m3_col(ax, d->xmat[0], i);
    for (int k = 0; k < 3; k++) Jr[k * nv + 3 + i] = ax[k];
  }
  if (leg >= 0)
    for (int j = 0; j <= link; j++)
      for (int k = 0; k < 3; k++) Jr[k * nv + dof_of(m, leg, j)] = d->axis[leg][j][k];
}

/* mj_rne(flg_acc=0): Coriolis + centrifugal + gravity, by per-body Newton-Euler in the world frame */
void odgo_bias(const OdgModel* m, OdgoData* d) {
  int nv = m->nv;
  memset(d->qfrc_bias, 0, sizeof(d->qfrc_bias));
  double w0[3], t[3], t2[3];
  m3_mulv(w0, d->xmat[0], d->qvel + 3);                 /* trunk angular velocity, world */
  for (int b = 0; b < 1 + m->nleg * m->njl; b++) {
    int leg = b == 0 ? -1 : (b - 1) / m->njl, link = b == 0 ? -1 : (b - 1) % m->njl;
    /* walk down from the trunk: body-fixed reference point r with acceleration aref, w, alpha */
    double w[3], al[3] = {0, 0, 0}, r[3], aref[3] = {0, 0, 0};
    v3_copy(w, w0); v3_copy(r, d->xpos[0]);
    for (int j = 0; j <= link; j++) {
      /* acceleration of the anchor as a point of the parent */
      double rho[3], acc[3];
      v3_sub(rho, d->anchor[leg][j], r);
      v3_cross(t, al, rho); v3_cross(t2, w, rho); v3_cross(t2, w, t2);
      v3_add(acc, aref, t); v3_add(acc, acc, t2);
      v3_copy(aref, acc); v3_copy(r, d->anchor[leg][j]);
      double qd = d->qvel[dof_of(m, leg, j)];
      v3_cross(t, w, d->axis[leg][j]);                  /* d/dt(axis) = w_parent x axis */
      v3_addscl(al, al, t, qd);
      v3_addscl(w, w, d->axis[leg][j], qd);
    }
    double mass = b == 0 ? m->base_mass : m->mass[leg][link];
    const double* Iloc = b == 0 ? m->base_inertia : m->inertia[leg][link];
    double rho[3], acom[3], F[3], N[3], Iw_w[3], Iw_al[3];
    v3_sub(rho, d->xipos[b], r);
    v3_cross(t, al, rho); v3_cross(t2, w, rho); v3_cross(t2, w, t2);
    v3_add(acom, aref, t); v3_add(acom, acom, t2);
    for (int k = 0; k < 3; k++) F[k] = mass * (acom[k] - m->gravity[k]);
    /* world inertia applied to a vector: R I R^T v */
    m3_tmulv(t, d->xmat[b], w); m3_mulv(t2, Iloc, t); m3_mulv(Iw_w, d->xmat[b], t2);
    m3_tmulv(t, d->xmat[b], al); m3_mulv(t2, Iloc, t); m3_mulv(Iw_al, d->xmat[b], t2);
    v3_cross(t, w, Iw_w);
    v3_add(N, Iw_al, t);
    /* project the wrench (F at COM, N) onto the dofs supporting this body */
    double Jp[3 * NV_MAX];
    jac_point(m, d, leg, link, d->xipos[b], Jp);
    for (int i = 0; i < nv; i++) d->qfrc_bias[i] += Jp[i] * F[0] + Jp[nv + i] * F[1] + Jp[2 * nv + i] * F[2];
    for (int i = 0; i < 3; i++) { m3_col(t, d->xmat[0], i); d->qfrc_bias[3 + i] += v3_dot(t, N); }
    for (int j = 0; j <= link; j++) d->qfrc_bias[dof_of(m, leg, j)] += v3_dot(d->axis[leg][j], N);
  }
}

/* mj_passive (joint damping only) and mj_fwdActuation (<position> actuators) */
static void passive_and_actuation(const OdgModel* m, OdgoData* d) {
  memset(d->qfrc_passive, 0, sizeof(d->qfrc_passive));
  memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
  for (int i = 0; i < 6; i++) d->qfrc_passive[i] = -m->base_damping[i] * d->qvel[i];
  for (int l = 0; l < m->nleg; l++)
    for (int j = 0; j < m->njl; j++) { int k = dof_of(m, l, j); d->qfrc_passive[k] = -m->damping[l][j] * d->qvel[k]; }
  for (int u = 0; u < m->nu; u++) {
    double c = d->ctrl[u];
    if (m->act_ctrllimited[u]) c = fmax(m->act_ctrlrange[u][0], fmin(m->act_ctrlrange[u][1], c));
    int l = m->act_leg[u], j = m->act_joint[u], k = dof_of(m, l, j);
    /* gain*ctrl + bias, gainprm = kp, biasprm = (0, -kp, -kv) */
    double f = m->act_kp[u] * c - m->act_kp[u] * d->qpos[7 + l * m->njl + j] - m->act_kv[u] * d->qvel[k];
    if (m->act_forcelimited[u]) f = fmax(m->act_forcerange[u][0], fmin(m->act_forcerange[u][1], f));
    d->actuator_force[u] = f;
    d->qfrc_actuator[k] += f;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* collision: floor plane z=0 (normal +z) against hulls / spheres — mjc_PlaneConvex / mjc_PlaneSphere */
static void add_contact(const OdgModel* m, OdgoData* d, int g, int vert, const double* point, double dist) {
  if (d->ncon >= ODGO_MAX_CON) return;
  OdgoContact* c = &d->contact[d->ncon++];
  memset(c, 0, sizeof(*c));
  c->geom = g; c->vert = vert; c->dist = dist; c->dim = m->geom[g].condim;
  v3_copy(c->pos, point);
  c->pos[2] -= 0.5 * dist;                                  /* midway between the two surfaces */
  /* frame: x = plane normal; mju_makeFrame picks y=(0,1,0), z = x cross y = (-1,0,0) */
  c->frame[2] = 1; c->frame[4] = 1; c->frame[6] = -1;
}

static __thread double g_world[ODG_MAX_VERT][3];          /* scratch: one hull in world coordinates */
static void note_gap(double* slot, double v) { v = fabs(v); if (v < *slot) *slot = v; }

static void bounding_spheres(const OdgModel* m, OdgoData* d) {
  for (int g = 0; g < m->ngeom; g++) {
    const OdgGeom* G = &m->geom[g];
    double c[3] = {0, 0, 0}, r2 = 0;
    for (int k = 0; k < G->vert_count; k++) for (int i = 0; i < 3; i++) c[i] += m->vert[G->vert_start + k][i] / G->vert_count;
    for (int k = 0; k < G->vert_count; k++) {
      double e[3]; v3_sub(e, m->vert[G->vert_start + k], c);
      r2 = fmax(r2, v3_dot(e, e));
    }
    v3_copy(d->bs_center[g], c); d->bs_radius[g] = sqrt(r2) * (1 + 1e-12) + 1e-12;
  }
  d->bs_ready = 0x5ca1ab1e;
}

void odgo_collision(const OdgModel* m, OdgoData* d) {
  d->ncon = 0;
  d->gap_contact = d->gap_support = 1e30;
  if (d->bs_ready != 0x5ca1ab1e) bounding_spheres(m, d);
  const double tilt = m->multicontact_tilt;
  for (int g = 0; g < m->ngeom; g++) {
    const OdgGeom* G = &m->geom[g];
    int b = body_of(m, G->leg, G->link);
    const double* R = d->xmat[b];
    const double* x = d->xpos[b];
    if (G->type == ODG_GEOM_SPHERE) {
      double c[3], p[3];
      m3_mulv(c, R, G->center); v3_add(c, c, x);
      double dist = c[2] - G->radius;
      note_gap(&d->gap_contact, dist - G->margin);
      if (dist > G->margin) continue;
      v3_set(p, c[0], c[1], c[2] - G->radius);
      add_contact(m, d, g, -1, p, dist);
      continue;
    }
    if (G->type == ODG_GEOM_CAPSULE || G->type == ODG_GEOM_CYLINDER || G->type == ODG_GEOM_BOX) {
      /* geom frame in the world: centre c, axes = columns of M = R * rot */
      double c[3], M[9];
      m3_mulv(c, R, G->center); v3_add(c, c, x);
      for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += R[i * 3 + k] * G->rot[k * 3 + j];
        M[i * 3 + j] = s;
      }
      const double n[3] = { 0, 0, 1 };
      if (G->type == ODG_GEOM_CAPSULE) {
        /* mjc_PlaneCapsule: sphere tests at the two ends of the segment, +axis first */
        double ax[3]; m3_col(ax, M, 2);
        for (int sgn = 1; sgn >= -1; sgn -= 2) {
          double e[3], p[3];
          v3_addscl(e, c, ax, sgn * G->size[1]);
          double dist = e[2] - G->size[0];
          note_gap(&d->gap_contact, dist - G->margin);
          if (dist > G->margin) continue;
          v3_set(p, e[0], e[1], e[2] - G->size[0]);
          add_contact(m, d, g, -1, p, dist);
        }
      } else if (G->type == ODG_GEOM_BOX) {
        /* mjc_PlaneBox: the 8 corners in index order, skipping corners above the centre or out of the margin, at most 4 */
        int cnt = 0;
        for (int i = 0; i < 8 && cnt < 4; i++) {
          double loc[3] = { (i & 1) ? G->size[0] : -G->size[0], (i & 2) ? G->size[1] : -G->size[1], (i & 4) ? G->size[2] : -G->size[2] };
          double corner[3], p[3];
          m3_mulv(corner, M, loc);
          double ldist = corner[2];
          if (ldist <= 0) note_gap(&d->gap_contact, c[2] + ldist - G->margin);
          if (c[2] + ldist > G->margin || ldist > 0) continue;
          v3_add(p, corner, c);
          add_contact(m, d, g, -1, p, c[2] + ldist);
          cnt++;
        }
      } else {
        /* mjc_PlaneCylinder: the lowest rim point of the disc nearer the plane, the rim point below it on the other disc,
         * and two more points on the near disc's rim 120 degrees to either side */
        double ax[3]; m3_col(ax, M, 2);
        double prjaxis = v3_dot(n, ax);
        if (prjaxis > 0) { for (int k = 0; k < 3; k++) ax[k] = -ax[k]; prjaxis = -prjaxis; }
        double dist0 = c[2];
        double vec[3];
        for (int k = 0; k < 3; k++) vec[k] = ax[k] * prjaxis - n[k];
        double len2 = v3_dot(vec, vec);
        if (len2 >= 1e-30) { double s = G->size[0] / sqrt(len2); for (int k = 0; k < 3; k++) vec[k] *= s; }
        else { double xa[3]; m3_col(xa, M, 0); for (int k = 0; k < 3; k++) vec[k] = xa[k] * G->size[0]; }
        double prjvec = v3_dot(vec, n);
        double axs[3]; for (int k = 0; k < 3; k++) axs[k] = ax[k] * G->size[1];
        double prjaxs = prjaxis * G->size[1];
        note_gap(&d->gap_contact, dist0 + prjaxs + prjvec - G->margin);
        if (dist0 + prjaxs + prjvec > G->margin) continue;
        {
          double dist = dist0 + prjaxs + prjvec, p[3];
          for (int k = 0; k < 3; k++) p[k] = c[k] + vec[k] + axs[k];
          add_contact(m, d, g, -1, p, dist);
        }
        note_gap(&d->gap_contact, dist0 - prjaxs + prjvec - G->margin);
        if (dist0 - prjaxs + prjvec <= G->margin) {
          double dist = dist0 - prjaxs + prjvec, p[3];
          for (int k = 0; k < 3; k++) p[k] = c[k] + vec[k] - axs[k];
          add_contact(m, d, g, -1, p, dist);
        }
        double prjvec1 = -prjvec * 0.5;
        note_gap(&d->gap_contact, dist0 + prjaxs + prjvec1 - G->margin);
        if (dist0 + prjaxs + prjvec1 <= G->margin) {
          double vec1[3];
          v3_cross(vec1, vec, ax);
          double l1 = sqrt(v3_dot(vec1, vec1));
          if (l1 > 0) for (int k = 0; k < 3; k++) vec1[k] *= G->size[0] * sqrt(3.0) / 2 / l1;
          double dist = dist0 + prjaxs + prjvec1;
          for (int sgn = 1; sgn >= -1; sgn -= 2) {
            double p[3];
            for (int k = 0; k < 3; k++) p[k] = c[k] + sgn * vec1[k] + axs[k] - vec[k] * 0.5;
            add_contact(m, d, g, -1, p, dist);
          }
        }
      }
      continue;
    }
    {
      /* broad phase (what mj_collision's bounding-volume test does for a geom far from the plane): the hull lies inside
       * the sphere (center, radius) of its geom frame, so it cannot come within the margin if the sphere does not.
       * Exact: never changes which contacts exist. */
      double c[3];
      m3_mulv(c, R, d->bs_center[g]);
      if (c[2] + x[2] - d->bs_radius[g] > G->margin) continue;
    }
    /* world coordinates of the hull once; support vertex in direction -normal */
    double (*W)[3] = g_world;
    for (int k = 0; k < G->vert_count; k++) { m3_mulv(W[k], R, m->vert[G->vert_start + k]); v3_add(W[k], W[k], x); }
    int best = -1; double zmin = 0, z2 = 1e30, pw[3], bestp[3] = {0, 0, 0};
    for (int k = 0; k < G->vert_count; k++) {
      v3_copy(pw, W[k]);
      if (best < 0 || pw[2] < zmin) { if (best >= 0) z2 = zmin; best = k; zmin = pw[2]; v3_copy(bestp, pw); }
      else if (pw[2] < z2) z2 = pw[2];
    }
    if (best >= 0) note_gap(&d->gap_contact, zmin - G->margin);
    if (best < 0 || zmin > G->margin) continue;
    note_gap(&d->gap_support, z2 - zmin);
    add_contact(m, d, g, G->vert_start + best, bestp, zmin);
    /* up to three more support points, from directions tilted off -normal by `tilt`, 120 deg apart */
    int found[ODG_MAX_CON_PER_GEOM]; int nf = 1; found[0] = best;
    for (int i = 0; i < 3 && tilt > 0; i++) {
      double ang = 2.0 * ODG_PI * i / 3.0;
      /* tangent basis of the contact frame: t1 = (0,1,0), t2 = (-1,0,0) */
      double dir[3] = { -sin(tilt) * sin(ang), sin(tilt) * cos(ang), -cos(tilt) };
      int bi = -1; double smax = 0, s2 = -1e30, bp[3] = {0, 0, 0};
      for (int k = 0; k < G->vert_count; k++) {
        v3_copy(pw, W[k]);
        double s = v3_dot(dir, pw);
        if (bi < 0 || s > smax) { if (bi >= 0) s2 = smax; bi = k; smax = s; v3_copy(bp, pw); }
        else if (s > s2) s2 = s;
      }
      note_gap(&d->gap_support, smax - s2);
      int dup = 0;
      for (int k = 0; k < nf; k++) dup |= (found[k] == bi);
      if (!dup) note_gap(&d->gap_contact, bp[2] - G->margin);
      if (dup || bp[2] > G->margin) continue;
      found[nf++] = bi;
      add_contact(m, d, g, G->vert_start + bi, bp, bp[2]);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* constraint rows: mj_makeConstraint + mj_makeImpedance                                       */
static double impedance(const double* solimp, double pos, double margin) {
  double d0 = fmin(MAXIMP, fmax(MINIMP, solimp[0])), d1 = fmin(MAXIMP, fmax(MINIMP, solimp[1]));
  double width = fmax(MINVAL, solimp[2]), mid = fmin(MAXIMP, fmax(MINIMP, solimp[3])), power = fmax(1.0, solimp[4]);
  if (d0 == d1 || width <= MINVAL) return 0.5 * (d0 + d1);
  double x = fabs((pos - margin) / width);
  if (x >= 1) return d1;
  if (x <= 0) return d0;
  double y;
  if (power == 1) y = x;
  else if (x <= mid) y = pow(x, power) / pow(mid, power - 1);
  else y = 1 - pow(1 - x, power) / pow(1 - mid, power - 1);
  return d0 + y * (d1 - d0);
}
static void solref_kb(const OdgModel* m, const double* solref, const double* solimp, double* K, double* B) {
  double dmax = fmin(MAXIMP, fmax(MINIMP, solimp[1]));
  if (solref[0] > 0) {
    double tc = fmax(solref[0], 2 * m->timestep);           /* refsafe */
    double dr = solref[1];
    *K = 1 / fmax(MINVAL, dmax * dmax * tc * tc * dr * dr);
    *B = 2 / fmax(MINVAL, dmax * tc);
  } else {
    *K = -solref[0] / fmax(MINVAL, dmax * dmax);
    *B = -solref[1] / fmax(MINVAL, dmax);
  }
}

static int add_row(const OdgModel* m, OdgoData* d, int type, int id, const double* J, double pos, double margin,
                   double frictionloss, const double* solref, const double* solimp, double diagApprox, int kzero) {
  int nv = m->nv, r = d->nefc++;
  memcpy(d->efc_J + r * nv, J, nv * sizeof(double));
  d->efc_type[r] = type; d->efc_id[r] = id; d->efc_pos[r] = pos; d->efc_margin[r] = margin;
  d->efc_frictionloss[r] = frictionloss;
  double vel = 0;
  for (int i = 0; i < nv; i++) vel += J[i] * d->qvel[i];
  d->efc_vel[r] = vel;
  double K, B, imp = impedance(solimp, pos, margin);
  solref_kb(m, solref, solimp, &K, &B);
  if (kzero) K = 0;
  d->efc_R[r] = fmax(MINVAL, (1 - imp) / imp * diagApprox);
  d->efc_D[r] = 1 / d->efc_R[r];
  d->efc_aref[r] = -B * vel - K * imp * (pos - margin);
  return r;
}

static void make_constraints(const OdgModel* m, OdgoData* d) {
  int nv = m->nv;
  double J[NV_MAX];
  d->nefc = 0;
  d->gap_limit = 1e30;
  /* 1. dof friction loss */
  for (int i = 0; i < nv; i++) {
    double fl = i < 6 ? m->base_frictionloss[i] : m->frictionloss[(i - 6) / m->njl][(i - 6) % m->njl];
    double w = i < 6 ? m->base_invweight0[i] : m->dof_invweight0[(i - 6) / m->njl][(i - 6) % m->njl];
    if (fl <= 0) continue;
    memset(J, 0, sizeof(J)); J[i] = 1;
    add_row(m, d, ODGO_ROW_FRICTION, i, J, 0, 0, fl, m->dof_solref, m->dof_solimp, w, 1);
  }
  /* 2. joint limits (hinge): lower then upper, active when dist < margin (= 0) */
  for (int l = 0; l < m->nleg; l++)
    for (int j = 0; j < m->njl; j++) {
      if (!m->jnt_limited[l][j]) continue;
      int k = dof_of(m, l, j);
      double q = d->qpos[7 + l * m->njl + j];
      for (int side = -1; side <= 1; side += 2) {
        double dist = side * (m->jnt_range[l][j][(side + 1) / 2] - q);
        note_gap(&d->gap_limit, dist);
        if (dist < 0) {
          memset(J, 0, sizeof(J)); J[k] = -side;
          add_row(m, d, ODGO_ROW_LIMIT, k, J, dist, 0, 0, m->lim_solref, m->lim_solimp, m->dof_invweight0[l][j], 0);
        }
      }
    }
  /* 3. contacts: elliptic cone, condim 1, 3 or 6. Rows: normal, two sliding rows (linear Jacobian of the contact point
   * along the frame axes), then for condim 6 the torsional row (relative angular velocity about the normal) and two
   * rolling rows (about the tangents) — mj_makeConstraint's mj_jacDifPair jacp / jacr blocks. The plane is static,
   * so the relative Jacobians are the geom body's own. */
  for (int c = 0; c < d->ncon; c++) {
    OdgoContact* con = &d->contact[c];
    const OdgGeom* G = &m->geom[con->geom];
    double Jp[3 * NV_MAX], Jr[3 * NV_MAX];
    jac_point(m, d, G->leg, G->link, con->pos, Jp);
    jac_rot(m, d, G->leg, G->link, Jr);
    con->efc = d->nefc;
    if (con->dist >= G->margin) { con->efc = -1; continue; }      /* excluded (in the gap) */
    const double fri5[5] = { G->friction, G->friction, G->friction_torsion, G->friction_roll, G->friction_roll };
    for (int k = 0; k < con->dim; k++) {
      const double* Jsrc = k < 3 ? Jp : Jr;
      const int ax = k < 3 ? k : k - 3;
      for (int i = 0; i < nv; i++)
        J[i] = con->frame[3 * ax] * Jsrc[i] + con->frame[3 * ax + 1] * Jsrc[nv + i] + con->frame[3 * ax + 2] * Jsrc[2 * nv + i];
      int r = add_row(m, d, ODGO_ROW_CONTACT, c, J, k == 0 ? con->dist : 0, k == 0 ? G->margin : 0, 0,
                      G->solref, G->solimp, G->invweight0, k > 0);
      if (k > 0) {
        /* elliptic cone (mj_makeImpedance): friction rows reuse the normal row's impedance; R_1 = R_n / impratio,
         * R_j = R_1 * mu_1^2 / mu_j^2 for the further friction dimensions */
        int r0 = con->efc;
        double B = 0, K = 0;
        solref_kb(m, G->solref, G->solimp, &K, &B);
        double R1 = fmax(MINVAL, d->efc_R[r0] / fmax(MINVAL, m->impratio));
        double fj = fmax(1e-5, fri5[k - 1]), f0 = fmax(1e-5, fri5[0]);          /* mjMINMU */
        d->efc_R[r] = fmax(MINVAL, R1 * ((f0 * f0) / (fj * fj)));     /* (ratio first: exactly R1 for the second sliding row) */
        d->efc_D[r] = 1 / d->efc_R[r];
        d->efc_aref[r] = -B * d->efc_vel[r];
      }
    }
  }
}

/* per-block cost of z = J*qacc - aref:  s(z), ds/dz, d2s/dz2 (dim x dim).  force = -ds/dz. */
static double block_cost(const OdgModel* m, const OdgoData* d, int r, int dim, const double* z, double* g, double* H) {
  for (int i = 0; i < dim * dim; i++) H[i] = 0;
  for (int i = 0; i < dim; i++) g[i] = 0;
  int type = d->efc_type[r];
  if (type == ODGO_ROW_FRICTION) {
    double f = d->efc_frictionloss[r], R = d->efc_R[r], D = d->efc_D[r];
    if (z[0] <= -R * f) { g[0] = -f; return f * (-0.5 * R * f - z[0]); }
    if (z[0] >= R * f) { g[0] = f; return f * (-0.5 * R * f + z[0]); }
    g[0] = D * z[0]; H[0] = D; return 0.5 * D * z[0] * z[0];
  }
  if (type == ODGO_ROW_LIMIT || dim == 1) {
    double D = d->efc_D[r];
    if (z[0] >= 0) return 0;
    g[0] = D * z[0]; H[0] = D; return 0.5 * D * z[0] * z[0];
  }
  /* elliptic contact, dim 3 or 6: map to the regular cone (mj_constraintUpdate) */
  const OdgGeom* G = &m->geom[d->contact[d->efc_id[r]].geom];
  const double fri5[5] = { G->friction, G->friction, G->friction_torsion, G->friction_roll, G->friction_roll };
  double mu = fmax(1e-5, fri5[0]) * sqrt(d->efc_R[r + 1] / d->efc_R[r]);
  double sc[6] = { 0 }, U[6] = { 0 };
  sc[0] = mu;
  for (int i = 1; i < dim; i++) sc[i] = fmax(1e-5, fri5[i - 1]);
  for (int i = 0; i < dim; i++) U[i] = z[i] * sc[i];
  double N = U[0], T = 0;
  for (int i = 1; i < dim; i++) T += U[i] * U[i];
  T = sqrt(T);
  if ((T <= 0 && N >= 0) || (T > 0 && N >= mu * T)) return 0;                    /* top zone: separated */
  if ((T <= 0 && N < 0) || (T > 0 && mu * N + T <= 0)) {                         /* bottom zone: sticking */
    double c = 0;
    for (int i = 0; i < dim; i++) { double D = d->efc_D[r + i]; g[i] = D * z[i]; H[i * dim + i] = D; c += 0.5 * D * z[i] * z[i]; }
    return c;
  }
  /* middle zone: on the cone surface (sliding) */
  double Dm = d->efc_D[r] / (mu * mu * (1 + mu * mu));
  double NmT = N - mu * T;
  double gU[6], HU[36];
  gU[0] = Dm * NmT;
  for (int i = 1; i < dim; i++) gU[i] = -Dm * NmT * mu * U[i] / T;
  double a = mu * N / (T * T * T), b = mu * mu - mu * N / T;
  HU[0] = 1;
  for (int i = 1; i < dim; i++) {
    HU[i] = HU[i * dim] = -mu * U[i] / T;
    for (int j = i; j < dim; j++) HU[i * dim + j] = HU[j * dim + i] = a * U[i] * U[j] + (i == j ? b : 0);   /* symmetric by construction */
  }
  for (int i = 0; i < dim; i++) {
    g[i] = gU[i] * sc[i];
    for (int j = 0; j < dim; j++) H[i * dim + j] = Dm * HU[i * dim + j] * sc[i] * sc[j];
  }
  return 0.5 * Dm * NmT * NmT;
}

static int block_dim(const OdgoData* d, int r) {
  return d->efc_type[r] == ODGO_ROW_CONTACT ? d->contact[d->efc_id[r]].dim : 1;
}

/* total cost, optionally gradient (nv) and Hessian (nv x nv) at qacc */
static double total_cost(const OdgModel* m, const OdgoData* d, const double* a, double* grad, double* H) {
  int nv = m->nv;
  double da[NV_MAX], Mda[NV_MAX], cost = 0;
  for (int i = 0; i < nv; i++) da[i] = a[i] - d->qacc_smooth[i];
  for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->M[i * nv + j] * da[j]; Mda[i] = s; cost += 0.5 * s * da[i]; }
  if (grad) memcpy(grad, Mda, nv * sizeof(double));
  if (H) memcpy(H, d->M, nv * nv * sizeof(double));
  for (int r = 0; r < d->nefc;) {
    int dim = block_dim(d, r);
    double z[6], g[6], Hb[36];
    for (int k = 0; k < dim; k++) {
      double s = -d->efc_aref[r + k];
      for (int i = 0; i < nv; i++) s += d->efc_J[(r + k) * nv + i] * a[i];
      z[k] = s;
    }
    cost += block_cost(m, d, r, dim, z, g, Hb);
    if (grad)
      for (int k = 0; k < dim; k++) for (int i = 0; i < nv; i++) grad[i] += d->efc_J[(r + k) * nv + i] * g[k];
    if (H)
      for (int k = 0; k < dim; k++) for (int l = 0; l < dim; l++) {
        double h = Hb[k * dim + l];
        if (h == 0) continue;
        const double* Jk = d->efc_J + (r + k) * nv; const double* Jl = d->efc_J + (r + l) * nv;
        for (int i = 0; i < nv; i++) { double t = h * Jk[i]; if (t != 0) for (int j = 0; j < nv; j++) H[i * nv + j] += t * Jl[j]; }
      }
    r += dim;
  }
  return cost;
}
double odgo_constraint_cost(const OdgModel* m, const OdgoData* d, const double* qacc) { return total_cost(m, d, qacc, 0, 0); }

/* phi'(alpha), phi''(alpha) along a + alpha*p */
static void line_derivs(const OdgModel* m, const OdgoData* d, const double* a, const double* p, const double* Jp_, const double* jar0,
                        double g0, double h0, double alpha, double* d1, double* d2) {
  (void)a; (void)p;
  double s1 = g0 + alpha * h0, s2 = h0;
  for (int r = 0; r < d->nefc;) {
    int dim = block_dim(d, r);
    double z[6], g[6], Hb[36];
    for (int k = 0; k < dim; k++) z[k] = jar0[r + k] + alpha * Jp_[r + k];
    block_cost(m, d, r, dim, z, g, Hb);
    for (int k = 0; k < dim; k++) {
      s1 += g[k] * Jp_[r + k];
      for (int l = 0; l < dim; l++) s2 += Hb[k * dim + l] * Jp_[r + k] * Jp_[r + l];
    }
    r += dim;
  }
  *d1 = s1; *d2 = s2;
}

/* Stopping rules of the Newton solver. As an ORACLE the solve runs far past MuJoCo's defaults (scaled gradient < 1e-11,
 * line search to 1e-14 of phi'(0)) so that the fp32 kernel is compared with the minimiser itself. As a CPU BASELINE
 * (bench.py's cpu_baseline / --impl reference legs call odgo_set_solver) it stops where MuJoCo's defaults stop:
 * mjOption.tolerance = 1e-8 on the scaled gradient OR the scaled cost improvement of the last iteration,
 * ls_tolerance = 0.01, ls_iterations = 50 — otherwise the timed work would be several times what mj_step does. */
static double g_tol = 1e-11, g_ls_tol = 1e-14;
static int g_ls_iter = 100, g_improvement_exit = 0;
void odgo_set_solver(double tolerance, double ls_tolerance, int ls_iterations, int improvement_exit) {
  g_tol = tolerance; g_ls_tol = ls_tolerance; g_ls_iter = ls_iterations; g_improvement_exit = improvement_exit;
}

/* mj_fwdConstraint with the Newton solver */
static void solve_constraints(const OdgModel* m, OdgoData* d) {
  int nv = m->nv;
  memcpy(d->qacc, d->qacc_smooth, nv * sizeof(double));
  memset(d->qfrc_constraint, 0, sizeof(d->qfrc_constraint));
  d->solver_iter = 0;
  if (d->nefc == 0) return;
  double a[NV_MAX];
  /* warm start: previous qacc if it has lower cost than qacc_smooth */
  double cw = total_cost(m, d, d->qacc_warmstart, 0, 0), cs = total_cost(m, d, d->qacc_smooth, 0, 0);
  memcpy(a, cw < cs ? d->qacc_warmstart : d->qacc_smooth, nv * sizeof(double));
  double grad[NV_MAX], H[NV_MAX * NV_MAX], p[NV_MAX], Mp[NV_MAX];
  double jar0[ODGO_MAX_EFC], Jp_[ODGO_MAX_EFC];
  double scale = 0;
  for (int i = 0; i < nv; i++) scale += d->M[i * nv + i];
  scale = 1.0 / fmax(MINVAL, scale);                       /* 1 / (meaninertia * nv) */
  double prev_cost = 0;
  for (int it = 0; it < 100; it++) {
    double cost = total_cost(m, d, a, grad, H);
    double gn = 0;
    for (int i = 0; i < nv; i++) gn += grad[i] * grad[i];
    gn = sqrt(gn);
    d->solver_cost = cost; d->solver_gradnorm = gn * scale; d->solver_iter = it;
    if (gn * scale < g_tol) break;
    if (g_improvement_exit && it > 0 && scale * (prev_cost - cost) < g_tol) break;
    prev_cost = cost;
    if (chol_factor(H, nv, nv) != 0) break;
    for (int i = 0; i < nv; i++) p[i] = -grad[i];
    chol_solve(H, nv, nv, p);
    /* exact line search on the convex, C1, piecewise-quadratic phi(alpha) */
    double g0 = 0, h0 = 0;
    for (int i = 0; i < nv; i++) { double s = 0; for (int j = 0; j < nv; j++) s += d->M[i * nv + j] * p[j]; Mp[i] = s; }
    for (int i = 0; i < nv; i++) {
      double s = 0; for (int j = 0; j < nv; j++) s += d->M[i * nv + j] * (a[j] - d->qacc_smooth[j]);
      g0 += p[i] * s; h0 += p[i] * Mp[i];
    }
    for (int r = 0; r < d->nefc; r++) {
      double s = -d->efc_aref[r], t = 0;
      for (int i = 0; i < nv; i++) { s += d->efc_J[r * nv + i] * a[i]; t += d->efc_J[r * nv + i] * p[i]; }
      jar0[r] = s; Jp_[r] = t;
    }
    double d1, d2, d10;
    line_derivs(m, d, a, p, Jp_, jar0, g0, h0, 0.0, &d10, &d2);
    if (d10 >= 0) break;                                    /* not a descent direction: converged */
    double lo = 0, hi = -1, alpha = 1.0;
    for (int ls = 0; ls < g_ls_iter; ls++) {
      line_derivs(m, d, a, p, Jp_, jar0, g0, h0, alpha, &d1, &d2);
      if (fabs(d1) <= g_ls_tol * fabs(d10)) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      double next = alpha - d1 / d2;
      if (hi < 0) { if (!(next > lo)) next = 2 * alpha; }
      else if (!(next > lo && next < hi)) next = 0.5 * (lo + hi);
      if (hi > 0 && (hi - lo) <= 1e-15 * hi) break;
      alpha = next;
    }
    for (int i = 0; i < nv; i++) a[i] += alpha * p[i];
  }
  memcpy(d->qacc, a, nv * sizeof(double));
  /* forces */
  for (int r = 0; r < d->nefc;) {
    int dim = block_dim(d, r);
    double z[6], g[6], Hb[36];
    for (int k = 0; k < dim; k++) {
      double s = -d->efc_aref[r + k];
      for (int i = 0; i < nv; i++) s += d->efc_J[(r + k) * nv + i] * a[i];
      z[k] = s;
    }
    block_cost(m, d, r, dim, z, g, Hb);
    for (int k = 0; k < dim; k++) {
      d->efc_force[r + k] = -g[k];
      for (int i = 0; i < nv; i++) d->qfrc_constraint[i] += d->efc_J[(r + k) * nv + i] * (-g[k]);
    }
    r += dim;
  }
  for (int c = 0; c < d->ncon; c++) {
    OdgoContact* con = &d->contact[c];
    for (int k = 0; k < 6; k++) con->force[k] = (con->efc >= 0 && k < con->dim) ? d->efc_force[con->efc + k] : 0;
  }
}

/* mj_rnePostConstraint's cfrc_ext: per body, the contact wrenches acting on it as [torque, force] in world axes about
 * the subtree centre of mass of the kinematic tree's root (body 1 = trunk: the whole robot's COM). gymnasium's
 * do_simulation calls it after the last mj_step, so it belongs to the post-integration state's kinematics but the
 * contact forces of the last substep's forward pass; here it is evaluated with the forward pass's own kinematics
 * (positions differ by one substep of motion: the reward code only thresholds norms of it). */
void odgo_cfrc_ext(const OdgModel* m, OdgoData* d) {
  int nb = 1 + m->nleg * m->njl;
  double com[3] = { 0, 0, 0 }, mt = m->base_mass;
  for (int k = 0; k < 3; k++) com[k] = m->base_mass * d->xipos[0][k];
  for (int l = 0; l < m->nleg; l++)
    for (int j = 0; j < m->njl; j++) {
      int b = body_of(m, l, j);
      for (int k = 0; k < 3; k++) com[k] += m->mass[l][j] * d->xipos[b][k];
      mt += m->mass[l][j];
    }
  for (int k = 0; k < 3; k++) com[k] /= mt;
  memset(d->cfrc_ext, 0, sizeof(d->cfrc_ext));
  for (int c = 0; c < d->ncon; c++) {
    const OdgoContact* con = &d->contact[c];
    if (con->efc < 0) continue;
    const OdgGeom* G = &m->geom[con->geom];
    int b = body_of(m, G->leg, G->link);
    if (b >= nb) continue;
    double f[3] = { 0, 0, 0 }, t[3] = { 0, 0, 0 }, r[3], rxf[3];
    for (int k = 0; k < 3; k++)
      for (int a = 0; a < 3; a++) {
        f[k] += con->frame[3 * a + k] * con->force[a];                       /* frame^T * force */
        if (con->dim > 3) t[k] += con->frame[3 * a + k] * con->force[3 + a];
      }
    v3_sub(r, con->pos, com);
    v3_cross(rxf, r, f);
    for (int k = 0; k < 3; k++) { d->cfrc_ext[b][k] += t[k] + rxf[k]; d->cfrc_ext[b][3 + k] += f[k]; }
  }
}

void odgo_forward(const OdgModel* m, OdgoData* d) {
  int nv = m->nv;
  odgo_kinematics(m, d);
  odgo_mass_matrix(m, d);
  odgo_bias(m, d);
  passive_and_actuation(m, d);
  double L[NV_MAX * NV_MAX];
  memcpy(L, d->M, sizeof(L));
  for (int i = 0; i < nv; i++) d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
  memcpy(d->qacc_smooth, d->qfrc_smooth, nv * sizeof(double));
  chol_factor(L, nv, nv);
  chol_solve(L, nv, nv, d->qacc_smooth);
  odgo_collision(m, d);
  make_constraints(m, d);
  solve_constraints(m, d);
}

/* mj_Euler (implicit in joint damping when any is present) + mj_advance */
static void euler(const OdgModel* m, OdgoData* d) {
  int nv = m->nv, njl = m->njl;
  double h = m->timestep, qacc[NV_MAX], damp[NV_MAX];
  int any = 0;
  for (int i = 0; i < nv; i++) { damp[i] = i < 6 ? m->base_damping[i] : m->damping[(i - 6) / njl][(i - 6) % njl]; any |= damp[i] > 0; }
  if (!any) memcpy(qacc, d->qacc, nv * sizeof(double));
  else {
    double A[NV_MAX * NV_MAX];
    memcpy(A, d->M, sizeof(A));
    for (int i = 0; i < nv; i++) { A[i * nv + i] += h * damp[i]; qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    chol_factor(A, nv, nv);
    chol_solve(A, nv, nv, qacc);
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int i = 0; i < 3; i++) d->qpos[i] += h * d->qvel[i];
  double w[3] = { d->qvel[3], d->qvel[4], d->qvel[5] };
  double n = sqrt(v3_dot(w, w));
  if (n > MINVAL) {                                          /* mju_quatIntegrate, body-frame angular velocity */
    double ax[3] = { w[0] / n, w[1] / n, w[2] / n }, dq[4], q[4];
    axisangle2quat(dq, ax, h * n);
    quat_norm(d->qpos + 3);
    quat_mul(q, d->qpos + 3, dq);
    quat_norm(q);
    memcpy(d->qpos + 3, q, sizeof(q));
  } else quat_norm(d->qpos + 3);
  for (int i = 6; i < nv; i++) d->qpos[i + 1] += h * d->qvel[i];
  d->time += h;
  memcpy(d->qacc_warmstart, d->qacc, nv * sizeof(double));
}

void odgo_step(const OdgModel* m, OdgoData* d) {
  odgo_forward(m, d);
  euler(m, d);
}

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10, shared bit-for-bit with opendog_b200/csrc (reset noise, desired velocity)     */
void odgo_philox4x32(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
float odgo_u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }  /* [0,1), 24 bits */

/* ------------------------------------------------------------------------------------------ */
/* WalkEnvironmentV0 (+ ScaleActionWrapper)                                                     */
static const int kPattern[8][4] = {                       /* reward_calc:54-63, order FL FR BL BR */
  {1, 1, 1, 1}, {1, 1, 0, 1}, {1, 0, 0, 1}, {1, 0, 1, 1}, {1, 1, 1, 1}, {1, 1, 1, 0}, {0, 1, 1, 0}, {1, 1, 1, 1} };
static const int kPawBody[4] = { 4, 7, 10, 13 };          /* reward_calc:93 */
#define STREAM_RESET 0x52534554u                          /* 'RSET' */
#define STREAM_DESVEL 0x44564c00u                         /* 'DVL'  */

void odgo_walk_init(OdgoWalkEnv* e, const OdgModel* m, uint64_t seed, uint32_t env_id) {
  memset(e, 0, sizeof(*e));
  e->m = m; e->seed = seed; e->env_id = env_id;
  e->frame_skip = 10;                                     /* WalkEnvironment.py:36 */
  e->max_steps = 750;                                     /* 15.0 / (0.002*10), WalkEnvironment.py:52,63 */
  /* reward_calc:75,301-305: desired velocity sampled ONCE, x in [0.5, 1.0], y = z = 0 */
  uint32_t r[4];
  odgo_philox4x32(seed, env_id, 0, 0, STREAM_DESVEL, r);
  float u = odgo_u01(r[0]);
  e->desired_velocity[0] = (double)(0.5f + 0.5f * u);
  e->desired_velocity[1] = 0; e->desired_velocity[2] = 0;
  odgo_reset_data(m, &e->d);
}

/* reward_calc:372-390 */
static void euler_from_quat(const double* q, double* roll, double* pitch, double* yaw) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  double t0 = 2.0 * (w * x + y * z), t1 = 1.0 - 2.0 * (x * x + y * y);
  *roll = atan2(t0, t1);
  double t2 = 2.0 * (w * y - z * x);
  t2 = t2 > 1.0 ? 1.0 : t2; t2 = t2 < -1.0 ? -1.0 : t2;
  *pitch = asin(t2);
  double t3 = 2.0 * (w * z + x * y), t4 = 1.0 - 2.0 * (y * y + z * z);
  *yaw = atan2(t3, t4);
}

/* WalkEnvironment.py:115-136 */
static void walk_obs(const OdgoWalkEnv* e, double* obs) {
  const OdgModel* m = e->m; const OdgoData* d = &e->d;
  double off = m->key_ctrl[7];                             /* key_ctrl[0, 7:] -> one value, broadcast (quirk C1) */
  int k = 0;
  for (int i = 0; i < 3; i++) obs[k++] = d->qvel[i] * 2.0;
  for (int i = 0; i < 3; i++) obs[k++] = d->qvel[3 + i] * 0.25;
  for (int i = 0; i < 3; i++) obs[k++] = e->desired_velocity[i] * 2.0;
  for (int i = 0; i < 8; i++) obs[k++] = (d->qpos[7 + i] - off) * 1.0;
  for (int i = 0; i < 8; i++) obs[k++] = d->qvel[6 + i] * 0.05;
  for (int i = 0; i < 8; i++) obs[k++] = (double)e->last_action[i];
  for (int i = 0; i < 33; i++) obs[i] = obs[i] < -100.0 ? -100.0 : (obs[i] > 100.0 ? 100.0 : obs[i]);
}

void odgo_walk_reset(OdgoWalkEnv* e, double* obs) {
  const OdgModel* m = e->m; OdgoData* d = &e->d;
  odgo_reset_data(m, d);                                   /* MujocoEnv.reset -> mj_resetData */
  /* reset_model (WalkEnvironment.py:138-151): qpos = key_qpos + U(-0.02, 0.02)^nq, qvel = 0.
   * data.ctrl noise is unobservable (ctrl is overwritten before the next mj_step) and not drawn. */
  for (int blk = 0; blk * 4 < m->nq; blk++) {
    uint32_t r[4];
    odgo_philox4x32(e->seed, e->env_id, e->episode, (uint32_t)blk, STREAM_RESET, r);
    for (int k = 0; k < 4 && blk * 4 + k < m->nq; k++) {
      float u = odgo_u01(r[k]);
      float noise = -0.02f + 0.04f * u;
      d->qpos[blk * 4 + k] = (double)((float)m->key_qpos[blk * 4 + k] + noise);
    }
  }
  e->episode++;
  e->step = 0;
  memset(e->last_action, 0, sizeof(e->last_action));
  e->last_action_is_reset = 1;
  /* NOT reset: gait_index / gait_matches / desired_velocity (quirks C3, C4) */
  if (obs) walk_obs(e, obs);
}

/* ScaleActionEnvironment.py:21-23, float32 arithmetic exactly as numpy evaluates it */
void odgo_walk_scale_action(const float* action, float* scaled) {
  static const float lo[8] = { 2.36f, -1.8f, 2.36f, -1.8f, 2.36f, -1.8f, 2.36f, -1.8f };
  static const float hi[8] = { 2.8f, -1.20f, 2.8f, -1.20f, 2.8f, -1.20f, 2.8f, -1.20f };
  for (int i = 0; i < 8; i++) {
    float t = action[i] + 1.0f;
    float w = hi[i] - lo[i];
    float p = t * w;
    scaled[i] = lo[i] + p / 2.0f;
  }
}

/* reward_calc:318-337: last contact index per paw between the floor (geom 0) and a paw body */
static void paws_in_ground(const OdgoWalkEnv* e, int* contact_of_paw) {
  const OdgModel* m = e->m; const OdgoData* d = &e->d;
  for (int p = 0; p < 4; p++) contact_of_paw[p] = -1;
  for (int c = 0; c < d->ncon; c++)
    for (int p = 0; p < 4; p++)
      if (m->geom[d->contact[c].geom].mj_body_id == kPawBody[p]) { contact_of_paw[p] = c; break; }
}

/* reward_calc:203-234 (stateful; called twice per step by the reference, quirk C2) */
static int diagonal_gait_reward(OdgoWalkEnv* e) {
  int cp[4];
  paws_in_ground(e, cp);
  int match = 1;
  for (int p = 0; p < 4; p++) match &= ((cp[p] >= 0) == kPattern[e->gait_index][p]);
  match &= (e->d.qvel[0] >= 0.5);                          /* desired_velocity_min[0] */
  if (match) {
    e->gait_matches += 8;
    e->gait_index = (e->gait_index + 1) % 8;
    return e->gait_matches;
  }
  e->gait_matches = 0; e->gait_index = 0;
  return 0;
}

/* reward_calc:351-370 + 339-349 (quirk C9) */
static void paw_contact_forces(const OdgoWalkEnv* e, double out[4][6]) {
  const OdgoData* d = &e->d;
  int cp[4];
  paws_in_ground(e, cp);
  for (int p = 0; p < 4; p++) {
    for (int k = 0; k < 6; k++) out[p][k] = 0;
    if (cp[p] < 0) continue;
    const OdgoContact* c = &d->contact[cp[p]];
    double fg[3], fb[3];
    m3_mulv(fg, c->frame, c->force);                       /* R_c @ f (rows used as columns) */
    /* xquat[paw_body - 1]: the calf link = the fused last link of leg p */
    m3_tmulv(fb, d->xmat[body_of(e->m, p, e->m->njl - 1)], fg);
    for (int k = 0; k < 3; k++) out[p][k] = fb[k];
  }
}

static int all_finite(const OdgoWalkEnv* e) {
  for (int i = 0; i < e->m->nq; i++) if (!isfinite(e->d.qpos[i])) return 0;
  for (int i = 0; i < e->m->nv; i++) if (!isfinite(e->d.qvel[i])) return 0;
  return 1;
}

static float sum8f(const float* r) { return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7])); }
static double sum8d(const double* r) { return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7])); }

/* everything WalkEnvironmentV0.step does after do_simulation (WalkEnvironment.py:59-79) */
static void walk_post(OdgoWalkEnv* e, const float* scaled, double* obs, double* reward, int* terminated,
                      int* truncated, OdgoWalkInfo* info) {
  const OdgModel* m = e->m; OdgoData* d = &e->d;
  const double lim = 15.0 * ODG_PI / 180.0;                  /* np.deg2rad(15) */
  walk_obs(e, obs);
  /* --- _calculate_positive_rewards */
  double lin = 0;
  if (d->qpos[0] > 0) {
    double ex = e->desired_velocity[0] - d->qvel[0], ey = e->desired_velocity[1] - d->qvel[1];
    lin = exp(-(ex * ex + ey * ey) / 0.25);
  }
  double roll = 0, pitch = 0, yaw = 0, safe = 0;
  int finite = all_finite(e);
  if (finite) {
    euler_from_quat(d->qpos + 3, &roll, &pitch, &yaw);
    double dr = !(fabs(roll) > lim) ? lim - fabs(roll) : 0;
    double dp = !(fabs(pitch) > lim) ? lim - fabs(pitch) : 0;
    double dy = !(fabs(yaw) > lim) ? lim - fabs(yaw) : 0;
    safe = (dr + dp + dy) / (0.110 + lim + lim + lim);
  }
  int gait = diagonal_gait_reward(e);
  double rewards = lin * 1.5 + safe * .015 + (double)gait * 3;
  /* --- _calculate_negative_costs */
  double dq[8];
  for (int i = 0; i < 8; i++) { double t = d->qpos[7 + i] - m->key_ctrl[i]; dq[i] = t * t; }
  double joint_cost = sum8d(dq);
  double rate;
  if (e->last_action_is_reset) {                           /* float64 zeros minus float32 action -> float64 */
    double s[8];
    for (int i = 0; i < 8; i++) { double t = 0.0 - (double)scaled[i]; s[i] = t * t; }
    rate = sum8d(s);
  } else {                                                 /* float32 - float32 stays float32 in numpy */
    float s[8];
    for (int i = 0; i < 8; i++) { float t = e->last_action[i] - scaled[i]; s[i] = t * t; }
    rate = (double)sum8f(s);
  }
  /* numpy 1.26 (reference pin): np.float32 scalar * python float -> float64 */
  double rate_w = rate * 0.01;
  double ycost = fabs(d->qpos[1]);
  double costs = joint_cost * 0.1 + rate_w + ycost;
  double r = rewards - costs;
  *reward = r > 0.0 ? r : 0.0;
  /* --- termination */
  int healthy = finite && (-lim < roll && roll < lim) && (-lim < pitch && pitch < lim) && (-lim < yaw && yaw < lim);
  *terminated = !healthy;
  *truncated = e->step >= e->max_steps;
  if (info) {
    info->x_position = d->qpos[0]; info->y_position = d->qpos[1];
    info->distance_from_origin = sqrt(d->qpos[0] * d->qpos[0] + d->qpos[1] * d->qpos[1]);
    paw_contact_forces(e, info->paw_contact_forces);
    int cp[4]; paws_in_ground(e, cp);
    for (int p = 0; p < 4; p++) info->paws_in_ground[p] = cp[p] >= 0;
    info->gait_first_call = gait;
    info->linear_vel_tracking_reward = lin;
    double tq = 0;
    for (int i = 0; i < 8; i++) tq += d->qfrc_actuator[m->nv - 8 + i] * d->qfrc_actuator[m->nv - 8 + i];
    info->reward_ctrl = tq;
    info->reward_terms[0] = lin; info->reward_terms[1] = safe; info->reward_terms[2] = gait;
    info->reward_terms[3] = joint_cost; info->reward_terms[4] = rate; info->reward_terms[5] = ycost;
  }
  int second = diagonal_gait_reward(e);                    /* info["patterns_matches"], WalkEnvironment.py:70 */
  if (info) info->patterns_matches = second;
  memcpy(e->last_action, scaled, 8 * sizeof(float));       /* set_last_action, :78 */
  e->last_action_is_reset = 0;
}

void odgo_walk_step(OdgoWalkEnv* e, const float* action, double* obs, double* reward, int* terminated,
                    int* truncated, OdgoWalkInfo* info) {
  float scaled[8];
  odgo_walk_scale_action(action, scaled);
  e->step += 1;
  for (int i = 0; i < 8; i++) e->d.ctrl[i] = (double)scaled[i];
  double gap[3] = { 1e30, 1e30, 1e30 };
  for (int s = 0; s < e->frame_skip; s++) {
    odgo_step(e->m, &e->d);
    gap[0] = fmin(gap[0], e->d.gap_contact); gap[1] = fmin(gap[1], e->d.gap_support); gap[2] = fmin(gap[2], e->d.gap_limit);
  }
  walk_post(e, scaled, obs, reward, terminated, truncated, info);
  if (info) for (int k = 0; k < 3; k++) info->min_gap[k] = gap[k];
}

void odgo_walk_evaluate(OdgoWalkEnv* e, const float* scaled, double* obs, double* reward, int* terminated,
                        int* truncated, OdgoWalkInfo* info) {
  for (int i = 0; i < 8; i++) e->d.ctrl[i] = (double)scaled[i];
  odgo_forward(e->m, &e->d);
  walk_post(e, scaled, obs, reward, terminated, truncated, info);
  if (info) { info->min_gap[0] = e->d.gap_contact; info->min_gap[1] = e->d.gap_support; info->min_gap[2] = e->d.gap_limit; }
}

/* a worker's whole env set in one call (no per-env foreign-function overhead): actions [n][8], obs [n][33] */
void odgo_walk_step_autoreset_batch(OdgoWalkEnv** envs, int n, const float* actions, double* obs, double* reward, int* done) {
  for (int i = 0; i < n; i++) {
    int trunc = 0;
    odgo_walk_step_autoreset(envs[i], actions + 8 * i, obs + 33 * i, reward + i, done + i, &trunc, 0, 0);
  }
}

void odgo_walk_step_autoreset(OdgoWalkEnv* e, const float* action, double* obs, double* reward, int* done,
                              int* truncated, double* terminal_obs, OdgoWalkInfo* info) {
  int term = 0, trunc = 0;
  odgo_walk_step(e, action, obs, reward, &term, &trunc, info);
  *done = term || trunc;
  *truncated = trunc && !term;                             /* SB3: TimeLimit.truncated = truncated and not terminated */
  if (*done) {
    if (terminal_obs) memcpy(terminal_obs, obs, 33 * sizeof(double));
    odgo_walk_reset(e, obs);
  }
}
