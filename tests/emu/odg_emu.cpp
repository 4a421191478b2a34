// odg_emu.cpp — TEST-ONLY lane emulator for opendog_b200/csrc/odg_core.cuh.
//
// Compiles the exact source of the CUDA step (same float32 arithmetic, same 4-lane cooperative
// algorithm) for the host, with 4 std::threads standing in for the 4 lanes of an environment and a
// spin barrier standing in for __shfl_xor_sync. It exists so the kernel's numerics can be checked
// against oracle/ in the CPU-only container (`-m "not gpu"` tests). It is NOT a CPU fallback: the
// package never builds, loads or links it; opendog_b200 fails loudly without the CUDA library.
#define ODG_HOST_EMU 1
#include <atomic>
#include <thread>
#include <vector>
#include <cstdio>

#include "../../opendog_b200/csrc/odg_prep.h"

namespace {
thread_local int t_lane = 0;
double g_buf[4];
std::atomic<int> g_count{0};
std::atomic<int> g_sense{0};
void barrier() {
  static thread_local int local_sense = 0;
  local_sense ^= 1;
  if (g_count.fetch_add(1, std::memory_order_acq_rel) == 3) {
    g_count.store(0, std::memory_order_relaxed);
    g_sense.store(local_sense, std::memory_order_release);
  } else {
    while (g_sense.load(std::memory_order_acquire) != local_sense) { }
  }
}
}  // namespace

float odg_emu_shfl_xor(float v, int m) {
  g_buf[t_lane] = (double)v; barrier(); float r = (float)g_buf[t_lane ^ m]; barrier(); return r;
}
double odg_emu_shfl_xor_d(double v, int m) {
  g_buf[t_lane] = v; barrier(); double r = g_buf[t_lane ^ m]; barrier(); return r;
}

#ifdef ODG_EMU_STATS
namespace { std::vector<float> g_stats; }
void odg_emu_stat(int it, float rel_step, float alpha, int ls_passes, int nc_leg, float d10) {
  g_stats.insert(g_stats.end(), { (float)it, rel_step, alpha, (float)ls_passes, (float)nc_leg, d10 });
}
extern "C" int emu_stats(float* out, int max_floats) {          // drains the record buffer (6 floats per record)
  const int n = (int)g_stats.size() < max_floats ? (int)g_stats.size() : max_floats;
  for (int i = 0; i < n; i++) out[i] = g_stats[i];
  g_stats.clear();
  return n;
}
#endif

struct Emu {
  odg::Prepared prep;
  int N;
  std::vector<float> qpos, qvel, warm, last_action, desvel;
  std::vector<int> step, gidx, gcnt, work; std::vector<unsigned> episode; std::vector<unsigned char> fresh;
  odg::SimPtrs P;
};

template <class F> static void run4(F f) {
  std::thread th[4];
  for (int l = 0; l < 4; l++) th[l] = std::thread([=]() { t_lane = l; f(l); });
  for (int l = 0; l < 4; l++) th[l].join();
}

extern "C" {

// Test hook for the support-vertex candidate lists (odg_prep.h): copies the hull of (slot, leg) as the kernel sees it
// (float32, padded rows excluded) and the candidate indices of one direction cell. Returns the number of hull vertices;
// *n_cand receives the list length.
int emu_support_candidates(const Emu* e, int slot, int leg, int cell, float* verts_xyz, int* cand, int* n_cand) {
  const odg::DevConst& C = e->prep.C;
  const int nv = C.slot_nvert[slot], vs = C.slot_vstart[slot];
  for (int k = 0; k < nv; k++)
    for (int c = 0; c < 3; c++) verts_xyz[k * 3 + c] = e->prep.vert[((size_t)(vs + k) * 4 + leg) * 4 + c];
  const float* gc = e->prep.gc.data();
  const int ent = reinterpret_cast<const int*>(gc + C.oct_off)[(slot * odg::kSupportCells + cell) * 4 + leg];
  const unsigned char* idx = reinterpret_cast<const unsigned char*>(gc + C.idx_off) + (ent & 0xFFFF);
  *n_cand = ent >> 16;
  for (int i = 0; i < *n_cand; i++) cand[i] = idx[i];
  return nv;
}
int emu_num_slots(const Emu* e) { return e->prep.C.nslot; }
float emu_tilt_dir(const Emu* e, int i, int c) { return e->prep.C.tilt_dir[i][c]; }

const char* emu_last_error() { static thread_local std::string e; return e.c_str(); }

Emu* emu_create(const OdgModel* m, const OdgEnvConfig* cfg, int N, uint64_t seed) {
  Emu* e = new Emu();
  std::string err = odg::prepare(*m, *cfg, seed, &e->prep);
  if (!err.empty()) { fprintf(stderr, "emu_create: %s\n", err.c_str()); delete e; return nullptr; }
  const odg::DevConst& C = e->prep.C;
  e->N = N;
  e->qpos.assign((size_t)C.nq * N, 0.f); e->qvel.assign((size_t)C.nv * N, 0.f); e->warm.assign((size_t)C.nv * N, 0.f);
  e->last_action.assign((size_t)C.nu * N, 0.f); e->desvel.assign((size_t)3 * N, 0.f);
  e->work.assign(N, 0); e->step.assign(N, 0); e->gidx.assign(N, 0); e->gcnt.assign(N, 0); e->episode.assign(N, 0); e->fresh.assign(N, 1);
  e->P = odg::SimPtrs{ N, N, e->qpos.data(), e->qvel.data(), e->warm.data(), e->last_action.data(), e->desvel.data(),
                       e->step.data(), e->gidx.data(), e->gcnt.data(), e->episode.data(), e->fresh.data(), e->work.data() };
  for (int i = 0; i < N; i++) odg::env_init(C, e->P, i);
  return e;
}
void emu_destroy(Emu* e) { delete e; }

void emu_reset(Emu* e, const unsigned char* mask, float* obs) {
  run4([=](int l) {
    for (int i = 0; i < e->N; i++) {
      if (mask && !mask[i]) continue;
      if (e->prep.C.njl == 2) odg::env_reset<2>(e->prep.C, e->prep.lc.data(), e->P, obs, i, l, 0xFu);
      else odg::env_reset<3>(e->prep.C, e->prep.lc.data(), e->P, obs, i, l, 0xFu);
    }
  });
}

void emu_step(Emu* e, const odg::StepArgs* A) {
  odg::StepArgs a = *A;
  run4([=](int l) {
    const float4* sv = reinterpret_cast<const float4*>(e->prep.vert.data());
    alignas(16) static float s_red[odg::kRedGroup];     // the group's shared-memory reduction rows
    for (int i = 0; i < e->N; i++) {
      if (e->prep.C.njl == 2) odg::env_step<2, true>(e->prep.C, e->prep.lc.data(), e->prep.gc.data(), sv, e->P, a, a.action, i, l, 0xFu, s_red);
      else odg::env_step<3, true>(e->prep.C, e->prep.lc.data(), e->prep.gc.data(), sv, e->P, a, a.action, i, l, 0xFu, s_red);
    }
  });
}

// state access in the C-ABI layouts ([N][nq] etc.)
void emu_get_state(Emu* e, float* qpos, float* qvel, float* warm) {
  const odg::DevConst& C = e->prep.C; int N = e->N;
  for (int i = 0; i < N; i++) {
    if (qpos) for (int k = 0; k < C.nq; k++) qpos[i * C.nq + k] = e->qpos[(size_t)k * N + i];
    if (qvel) for (int k = 0; k < C.nv; k++) qvel[i * C.nv + k] = e->qvel[(size_t)k * N + i];
    if (warm) for (int k = 0; k < C.nv; k++) warm[i * C.nv + k] = e->warm[(size_t)k * N + i];
  }
}
void emu_set_state(Emu* e, const float* qpos, const float* qvel, const float* warm) {
  const odg::DevConst& C = e->prep.C; int N = e->N;
  for (int i = 0; i < N; i++) {
    if (qpos) for (int k = 0; k < C.nq; k++) e->qpos[(size_t)k * N + i] = qpos[i * C.nq + k];
    if (qvel) for (int k = 0; k < C.nv; k++) e->qvel[(size_t)k * N + i] = qvel[i * C.nv + k];
    for (int k = 0; k < C.nv; k++) e->warm[(size_t)k * N + i] = warm ? warm[i * C.nv + k] : 0.f;
  }
}
void emu_get_env_state(Emu* e, int* step, int* gidx, int* gcnt, float* last_action, float* desvel, unsigned char* fresh) {
  const odg::DevConst& C = e->prep.C; int N = e->N;
  for (int i = 0; i < N; i++) {
    if (step) step[i] = e->step[i];
    if (gidx) gidx[i] = e->gidx[i];
    if (gcnt) gcnt[i] = e->gcnt[i];
    if (fresh) fresh[i] = e->fresh[i];
    if (last_action) for (int u = 0; u < C.nu; u++) last_action[i * C.nu + u] = e->last_action[(size_t)u * N + i];
    if (desvel) for (int k = 0; k < 3; k++) desvel[i * 3 + k] = e->desvel[(size_t)k * N + i];
  }
}
void emu_set_env_state(Emu* e, const int* step, const int* gidx, const int* gcnt, const float* last_action,
                       const float* desvel, const unsigned char* fresh) {
  const odg::DevConst& C = e->prep.C; int N = e->N;
  for (int i = 0; i < N; i++) {
    if (step) e->step[i] = step[i];
    if (gidx) e->gidx[i] = gidx[i];
    if (gcnt) e->gcnt[i] = gcnt[i];
    if (fresh) e->fresh[i] = fresh[i];
    if (last_action) for (int u = 0; u < C.nu; u++) e->last_action[(size_t)u * N + i] = last_action[i * C.nu + u];
    if (desvel) for (int k = 0; k < 3; k++) e->desvel[(size_t)k * N + i] = desvel[i * 3 + k];
  }
}
void emu_get_work(Emu* e, int* out) { for (int i = 0; i < e->N; i++) out[i] = e->work[i]; }   // Newton iterations + line-search passes of the last env-step
int emu_sizeof_stepargs() { return (int)sizeof(odg::StepArgs); }
void emu_default_config(OdgEnvConfig* c) { odg::default_config(c); }
}
