// odg_mppi.cu — MPPI sampling, cost accumulation and the on-device softmin reduction (include/odg_mppi.h).
#include <cuda_runtime.h>

#include <string>

#include "../../include/odg.h"
#include "../../include/odg_mppi.h"
#include "odg_sim_internal.h"

namespace {
using odg_internal::set_error;
#define CUDA_TRY(expr)                                                                         \
  do { cudaError_t e_ = (expr);                                                                \
       if (e_ != cudaSuccess) return set_error(ODG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

constexpr uint32_t kStreamMppi = 0x4d505049u;      // Philox stream id "MPPI"

// action[a] = clamp(mean[a] + sigma * eps, -1, 1) for one sample: Philox4x32-10 keyed by (seed, sample, iteration, t, block),
// Box-Muller pairs. One definition for the stand-alone sampler and for the fused rollout kernel.
__device__ __forceinline__ void mppi_sample_row(const float* __restrict__ mean, float sigma, int n, int N, int A, uint32_t seed_lo,
                                                uint32_t seed_hi, uint32_t iteration, uint32_t t, float* __restrict__ action) {
  for (int blk = 0; blk * 4 < A; blk++) {
    uint32_t r[4];
    odg::philox4x32(seed_lo, seed_hi, (uint32_t)n, iteration, (t << 8) | (uint32_t)blk, kStreamMppi, r);
    for (int pr = 0; pr < 2; pr++) {
      const float u1 = ((float)(r[2 * pr] >> 8) + 1.0f) * 5.9604644775390625e-08f;
      const float u2 = odg::u01(r[2 * pr + 1]);
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs; sincosf(6.283185307179586f * u2, &sn, &cs);
      const float e[2] = { rad * cs, rad * sn };
      for (int q = 0; q < 2; q++) {
        const int a = blk * 4 + pr * 2 + q;
        if (a < A) action[(size_t)n * A + a] = fminf(1.f, fmaxf(-1.f, mean[a] + sigma * e[q]));
      }
    }
  }
  (void)N;
}

__global__ void k_mppi_sample(const float* __restrict__ mean, float sigma, int N, int A, uint32_t seed_lo, uint32_t seed_hi,
                              uint32_t iteration, uint32_t t, float* __restrict__ action) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < N) mppi_sample_row(mean, sigma, n, N, A, seed_lo, seed_hi, iteration, t, action);
}

// The whole rollout of one plan in ONE launch: every 4-lane group owns one sample and walks the horizon — draw the
// step's action row (the stream of k_mppi_sample), take the fused environment step (odg_core.cuh: env_step, the code
// k_step runs), add minus the unclipped reward to the sample's cost. Between the steps of a sample nothing waits on any
// other sample: no launch boundaries, no grid-wide barrier per horizon step; a sample that terminates pays once and
// only keeps drawing its (cheap) action rows, which the weighted mean of the reduction still needs.
struct MppiArgs {
  const float* mean;              // [T][A]
  float sigma, term_cost;
  int T;
  uint32_t seed_lo, seed_hi, iteration;
  const uint32_t* iteration_dev;  // nullable: device-side plan counter (CUDA-graph replays draw fresh noise)
  float* actions;                 // [T][n][A]
  float* cost;                    // [n]
};

template <int NJL>
__global__ void __launch_bounds__(128, 1) k_mppi_rollout(const __grid_constant__ odg::DevConst C, const odg::SimPtrs P, const MppiArgs M,
                                                         const float* __restrict__ g_lc, const float* __restrict__ g_gc,
                                                         const float* __restrict__ g_vert, SmemLayout L, int lanes) {
  extern __shared__ __align__(16) float smem[];
  const float4* s_vert; const float* s_lc; const float* s_gc;
  stage_constants(smem, g_lc, g_gc, g_vert, L, &s_vert, &s_lc, &s_gc);
  const int leg = threadIdx.x & 3, lane = threadIdx.x & 31;
  float* s_red = smem + L.vert_floats + L.lc_floats + L.gc_floats + (threadIdx.x >> 2) * (4 * odg::kRedStride);
  if (lane >= lanes) return;
  const unsigned gm = 0xFu << (lane & 28);
  const int envs_per_warp = lanes >> 2;
  const int env = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * envs_per_warp + (lane >> 2);
  if (env >= P.N) return;
  const bool real = env < P.n;
  const uint32_t iteration = M.iteration_dev ? *M.iteration_dev : M.iteration;
  const int A = C.nu;
  odg::StepArgs S;
  memset(&S, 0, sizeof(S));
  float cost = 0.f;
  bool alive = true;
  for (int t = 0; t < M.T; t++) {
    float* row = M.actions + (size_t)t * P.n * A;
    if (leg == 0 && real) mppi_sample_row(M.mean + (size_t)t * A, M.sigma, env, P.n, A, M.seed_lo, M.seed_hi, iteration, (uint32_t)t, row);
    __syncwarp(gm);                                // the group's other lanes read the row lane 0 just wrote
    if (alive) {
      S.action = row;
      const odg::StepResult r = odg::env_step<NJL>(C, s_lc, s_gc, s_vert, P, S, env, leg, gm, s_red);
      cost -= r.reward_unclipped;
      if (r.terminated) { cost += M.term_cost; alive = false; }
    }
  }
  if (leg == 0 && real) M.cost[env] = cost;
}

__global__ void k_mppi_accum(const float* __restrict__ reward, const unsigned char* __restrict__ term, int N, float term_cost,
                             float* __restrict__ cost, unsigned char* __restrict__ alive) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N || !alive[n]) return;
  float c = cost[n] - reward[n];
  if (term[n]) { c += term_cost; alive[n] = 0; }
  cost[n] = c;
}

constexpr int kRedThreads = 1024;
__global__ void __launch_bounds__(kRedThreads) k_mppi_reduce(const float* __restrict__ cost, const float* __restrict__ actions, int T,
                                                             int N, int A, float lambda, float* __restrict__ mean_out,
                                                             float* __restrict__ stats) {
  extern __shared__ float s_w[];                 // [N] weights
  __shared__ float s_red[kRedThreads]; __shared__ int s_idx[kRedThreads];
  const int tid = threadIdx.x;
  // min cost (and its index), fixed tree
  float m = 3.4e38f; int mi = 0; float sum_c = 0.f;
  for (int n = tid; n < N; n += kRedThreads) { const float c = cost[n]; sum_c += c; if (c < m) { m = c; mi = n; } }
  s_red[tid] = m; s_idx[tid] = mi;
  __syncthreads();
  for (int d = kRedThreads / 2; d > 0; d >>= 1) {
    if (tid < d) {
      const float o = s_red[tid + d]; const int oi = s_idx[tid + d];
      if (o < s_red[tid] || (o == s_red[tid] && oi < s_idx[tid])) { s_red[tid] = o; s_idx[tid] = oi; }
    }
    __syncthreads();
  }
  const float cmin = s_red[0]; const int imin = s_idx[0];
  __syncthreads();
  // weights and their sum
  float ws = 0.f;
  for (int n = tid; n < N; n += kRedThreads) { const float w = expf(-(cost[n] - cmin) / lambda); s_w[n] = w; ws += w; }
  s_red[tid] = ws;
  __syncthreads();
  for (int d = kRedThreads / 2; d > 0; d >>= 1) { if (tid < d) s_red[tid] += s_red[tid + d]; __syncthreads(); }
  const float wsum = s_red[0];
  __syncthreads();
  s_red[tid] = sum_c;
  __syncthreads();
  for (int d = kRedThreads / 2; d > 0; d >>= 1) { if (tid < d) s_red[tid] += s_red[tid + d]; __syncthreads(); }
  if (tid == 0 && stats) { stats[0] = cmin; stats[1] = s_red[0] / (float)N; stats[2] = wsum; stats[3] = (float)imin; }
  // weighted mean per (t, a): thread per output element, samples summed in index order
  for (int o = tid; o < T * A; o += kRedThreads) {
    const int t = o / A, a = o % A;
    const float* p = actions + (size_t)t * N * A + a;
    float acc = 0.f;
    for (int n = 0; n < N; n++) acc += s_w[n] * p[(size_t)n * A];
    mean_out[o] = acc / wsum;
  }
}
}  // namespace

extern "C" {

int odg_mppi_sample(const float* mean_dev, float sigma, int n_samples, int act_dim, uint64_t seed, uint32_t iteration,
                    uint32_t t, float* action_dev, void* stream) {
  if (!mean_dev || !action_dev || n_samples < 1 || act_dim < 1 || act_dim > 16)
    return set_error(ODG_ERR_INVALID, "odg_mppi_sample: bad arguments");
  k_mppi_sample<<<(n_samples + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      mean_dev, sigma, n_samples, act_dim, (uint32_t)seed, (uint32_t)(seed >> 32), iteration, t, action_dev);
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

int odg_mppi_accumulate(const float* reward_dev, const uint8_t* terminated_dev, int n_samples, float termination_cost,
                        float* cost_dev, uint8_t* alive_dev, void* stream) {
  if (!reward_dev || !terminated_dev || !cost_dev || !alive_dev || n_samples < 1)
    return set_error(ODG_ERR_INVALID, "odg_mppi_accumulate: bad arguments");
  k_mppi_accum<<<(n_samples + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(reward_dev, terminated_dev, n_samples,
                                                                                        termination_cost, cost_dev, alive_dev);
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

int odg_mppi_rollout(OdgSim* s, const float* mean_dev, float sigma, int horizon, uint64_t seed, uint32_t iteration,
                     const uint32_t* iteration_dev, float termination_cost, float* actions_dev, float* cost_dev, void* stream) {
  if (!s || !mean_dev || !actions_dev || !cost_dev || horizon < 1 || !(sigma >= 0.f))
    return set_error(ODG_ERR_INVALID, "odg_mppi_rollout: bad arguments");
  if (s->prep.C.auto_reset) return set_error(ODG_ERR_INVALID, "odg_mppi_rollout: the handle must be created with auto_reset = 0");
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != s->device) cudaSetDevice(s->device);
  odg::DevConst C = s->prep.C;
  C.lockstep = 0;                                  // samples leave the horizon loop at different times: no block barriers
  MppiArgs M;
  M.mean = mean_dev; M.sigma = sigma; M.term_cost = termination_cost; M.T = horizon;
  M.seed_lo = (uint32_t)seed; M.seed_hi = (uint32_t)(seed >> 32); M.iteration = iteration; M.iteration_dev = iteration_dev;
  M.actions = actions_dev; M.cost = cost_dev;
  // one wave: as few environments per warp as the batch allows (the rollout is one long latency chain per sample, and
  // samples that share a warp wait for each other's Newton iterations)
  const int lanes = s->cfg_lanes ? s->cfg_lanes : (s->P.N <= 8 * s->num_sms * 2 ? 8 : s->step_lanes);
  const int block = s->cfg_block ? s->cfg_block : 64;
  const int envs_per_block = (block / 32) * (lanes / 4);
  const int grid = (s->P.N + envs_per_block - 1) / envs_per_block;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (C.njl == 2) {
    e = cudaFuncSetAttribute(k_mppi_rollout<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_step);
    if (e == cudaSuccess) k_mppi_rollout<2><<<grid, block, s->smem_step, st>>>(C, s->P, M, s->d_lc, s->d_gc, s->d_vert, s->L, lanes);
  } else {
    e = cudaFuncSetAttribute(k_mppi_rollout<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_step);
    if (e == cudaSuccess) k_mppi_rollout<3><<<grid, block, s->smem_step, st>>>(C, s->P, M, s->d_lc, s->d_gc, s->d_vert, s->L, lanes);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  s->launches++;
  if (prev != s->device && prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) return set_error(ODG_ERR_CUDA, std::string("odg_mppi_rollout: ") + cudaGetErrorString(e));
  return ODG_OK;
}

int odg_mppi_reduce(const float* cost_dev, const float* actions_dev, int horizon, int n_samples, int act_dim, float lambda,
                    float* mean_out_dev, float* stats_dev, void* stream) {
  if (!cost_dev || !actions_dev || !mean_out_dev || horizon < 1 || n_samples < 1 || n_samples > 12000 || act_dim < 1 || !(lambda > 0.f))
    return set_error(ODG_ERR_INVALID, "odg_mppi_reduce: bad arguments (n_samples <= 12000)");
  // [n_samples] weights in dynamic shared memory next to 8 KB of static reduction rows: above 48 KB in total the kernel
  // needs the opt-in limit (n_samples > ~10000)
  CUDA_TRY(cudaFuncSetAttribute(k_mppi_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)n_samples * sizeof(float))));
  k_mppi_reduce<<<1, kRedThreads, (size_t)n_samples * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      cost_dev, actions_dev, horizon, n_samples, act_dim, lambda, mean_out_dev, stats_dev);
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

}  // extern "C"
