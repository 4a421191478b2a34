// odg_terrain.cu — the terrain trainer's height fields (include/odg_sim2real.h: odg_terrain_*), one 100 x 100 field per
// environment: `_generate_random_terrain` and `get_terrain_height` of sim2real/train2.py:203-304 as kernels.
// One block per environment; the field lives in shared memory through generation, the four blend passes and the
// normalisation, and is written to HBM once (transposed, MuJoCo's hfield_data layout).
#include <cuda_runtime.h>

#include <new>
#include <string>

#include "../../include/odg_sim2real.h"
#include "odg_sim_internal.h"

namespace {
using odg_internal::set_error;
#define CUDA_TRY(expr)                                                                         \
  do { cudaError_t e_ = (expr);                                                                \
       if (e_ != cudaSuccess) return set_error(ODG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

constexpr int kR = 100, kC = 100, kCells = kR * kC;         // TERRAIN_ROWS / TERRAIN_COLS (train2.py:111-112)
constexpr float kMaxH = 1.5f, kSmooth = 0.3f;               // TERRAIN_MAX_ABS_HEIGHT, TERRAIN_SMOOTHNESS_FACTOR (:113-114)
constexpr int kPasses = 4;                                  // TERRAIN_NUM_SMOOTH_PASSES (:115)
constexpr double kSizeX = 5.0, kSizeY = 5.0, kSizeZ = 0.3, kBase = 0.001;   // hfield size (walking_scene.xml:19)
constexpr uint32_t kStreamTerrain = 0x54455252u;            // Philox stream id "TERR"
constexpr int kThreads = 256;

struct TerrainArgs {
  int N; uint32_t seed_lo, seed_hi; int first_env_id;
  float* data; unsigned* gen;
  const unsigned char* mask; const float* raw; const float* radius;
  double start_x, start_y;
};

__device__ __forceinline__ double cell_x(int c) { return 0.0 - (kSizeX / 2.0) + c * (kSizeX / (kC - 1)); }
__device__ __forceinline__ double cell_y(int r) { return 0.0 - (kSizeY / 2.0) + r * (kSizeY / (kR - 1)); }

__global__ void __launch_bounds__(kThreads) k_terrain(const TerrainArgs A) {
  extern __shared__ float sm[];                    // two fields: current and scratch
  float* cur = sm; float* nxt = sm + kCells;
  __shared__ float s_lo[kThreads], s_hi[kThreads];
  const int env = blockIdx.x, tid = threadIdx.x;
  if (A.mask && !A.mask[env]) return;
  float* out = A.data + (size_t)env * kCells;
  float radius;
  if (A.raw) {                                     // test hook: the deterministic tail on given raw heights
    radius = A.radius[env];
    for (int i = tid; i < kCells; i += kThreads) cur[i] = A.raw[(size_t)env * kCells + i];
  } else {
    const unsigned gen = A.gen[env];
    const uint32_t gid = (uint32_t)(A.first_env_id + env);
    uint32_t h[4];
    odg::philox4x32(A.seed_lo, A.seed_hi, gid, gen, 0xFFFFFFFFu, kStreamTerrain, h);
    if (odg::u01(h[0]) < 0.5f) {                   // 50 %: flat terrain, normalised height 0.5 (:206-210)
      for (int i = tid; i < kCells; i += kThreads) out[i] = 0.5f;
      __syncthreads();
      if (tid == 0) A.gen[env] = gen + 1;
      return;
    }
    radius = 0.1f + 0.3f * odg::u01(h[1]);         // flat_circle_radius ~ U(0.1, 0.4) (:224)
    for (int i = tid; i < kCells; i += kThreads) {
      const int r = i / kC, c = i % kC;
      const double wx = cell_x(c), wy = cell_y(r);
      const double dist = sqrt((wx - A.start_x) * (wx - A.start_x) + (wy - A.start_y) * (wy - A.start_y));
      float v = 0.f;
      if (dist >= (double)radius) {
        uint32_t a[4], b[4];
        odg::philox4x32(A.seed_lo, A.seed_hi, gid, gen, (uint32_t)(2 * i), kStreamTerrain, a);
        odg::philox4x32(A.seed_lo, A.seed_hi, gid, gen, (uint32_t)(2 * i + 1), kStreamTerrain, b);
        const double H = (double)kMaxH;
        const double base = -H + 2.0 * H * (double)odg::u01(a[0]);
        const double fx = 0.2 + 0.4 * (double)odg::u01(a[1]), fy = 0.2 + 0.4 * (double)odg::u01(a[2]);
        const double noise = (sin(wx * fx) * cos(wy * fy) + sin(wx * fx * 2) * cos(wy * fy * 2)) * H * 0.7;
        double spike = 0.0;
        if (odg::u01(a[3]) < 0.2f) spike = -H * 0.8 + 1.6 * H * (double)odg::u01(b[0]);
        v = (float)(base + noise + spike);
        if (fabs(dist - (double)radius) < 1.0) v = v * 1.5f;
      }
      cur[i] = v;
    }
  }
  __syncthreads();
  // ---- 4 passes of the 3 x 3 blend over interior cells outside the flat disc (:250-262), float32 like numpy
  for (int pass = 0; pass < kPasses; pass++) {
    for (int i = tid; i < kCells; i += kThreads) {
      const int r = i / kC, c = i % kC;
      float v = cur[i];
      if (r >= 1 && r < kR - 1 && c >= 1 && c < kC - 1) {
        const double wx = cell_x(c), wy = cell_y(r);
        const double dist = sqrt((wx - A.start_x) * (wx - A.start_x) + (wy - A.start_y) * (wy - A.start_y));
        if (dist >= (double)radius) {
          float s = 0.f;
          for (int dr = -1; dr <= 1; dr++) for (int dc = -1; dc <= 1; dc++) s = __fadd_rn(s, cur[(r + dr) * kC + c + dc]);
          const float avg = __fdiv_rn(s, 9.0f);
          v = __fadd_rn(__fmul_rn(v, 1.0f - kSmooth), __fmul_rn(avg, kSmooth));
        }
      }
      nxt[i] = v;
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  // ---- min-max normalisation (:263-265), transpose, store (:270)
  float lo = 3.4e38f, hi = -3.4e38f;
  for (int i = tid; i < kCells; i += kThreads) { lo = fminf(lo, cur[i]); hi = fmaxf(hi, cur[i]); }
  s_lo[tid] = lo; s_hi[tid] = hi;
  __syncthreads();
  for (int d = kThreads / 2; d > 0; d >>= 1) {
    if (tid < d) { s_lo[tid] = fminf(s_lo[tid], s_lo[tid + d]); s_hi[tid] = fmaxf(s_hi[tid], s_hi[tid + d]); }
    __syncthreads();
  }
  lo = s_lo[0]; hi = s_hi[0];
  const bool degenerate = hi <= lo + 1e-4f;
  for (int i = tid; i < kCells; i += kThreads) {
    const int c = i / kR, r = i % kR;              // output index i = c * nrow + r  (norm.T.flatten())
    out[i] = degenerate ? 0.5f : __fdiv_rn(__fsub_rn(cur[r * kC + c], lo), __fsub_rn(hi, lo));
  }
  if (!A.raw && tid == 0) A.gen[env] += 1;
}

// get_terrain_height (train2.py:295-304), one query per env
__global__ void k_terrain_height(const float* __restrict__ data, const float* __restrict__ xy, float* __restrict__ h, int N) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= N) return;
  const double lx = (double)xy[2 * env], ly = (double)xy[2 * env + 1];
  const double cf = (lx + kSizeX / 2.0) / kSizeX * (kC - 1), rf = (ly + kSizeY / 2.0) / kSizeY * (kR - 1);
  const int c = (int)fmin(fmax(cf, 0.0), (double)(kC - 1)), r = (int)fmin(fmax(rf, 0.0), (double)(kR - 1));
  h[env] = (float)(0.0 + (kBase + (double)data[(size_t)env * kCells + c * kR + r] * kSizeZ));
}

struct Scope {
  int prev = -1;
  explicit Scope(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~Scope() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

struct OdgTerrain {
  int N = 0, device = 0, first_env_id = 0;
  uint64_t seed = 0;
  float* d_data = nullptr; unsigned* d_gen = nullptr;
};

extern "C" {

int odg_terrain_create(int num_envs, int device, uint64_t seed, int first_env_id, OdgTerrain** out) {
  if (!out || num_envs < 1) return set_error(ODG_ERR_INVALID, "odg_terrain_create: bad arguments");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_error(ODG_ERR_NO_DEVICE, "no CUDA device: libodgsim has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(ODG_ERR_INVALID, "odg_terrain_create: device out of range");
  Scope scope(device);
  OdgTerrain* t = new (std::nothrow) OdgTerrain();
  if (!t) return set_error(ODG_ERR_ALLOC, "out of host memory");
  t->N = num_envs; t->device = device; t->seed = seed; t->first_env_id = first_env_id;
  if (cudaMalloc(&t->d_data, (size_t)num_envs * kCells * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&t->d_gen, (size_t)num_envs * sizeof(unsigned)) != cudaSuccess) {
    odg_terrain_destroy(t); return set_error(ODG_ERR_ALLOC, "cudaMalloc(terrain) failed");
  }
  CUDA_TRY(cudaMemset(t->d_gen, 0, (size_t)num_envs * sizeof(unsigned)));
  CUDA_TRY(cudaMemset(t->d_data, 0, (size_t)num_envs * kCells * sizeof(float)));
  CUDA_TRY(cudaFuncSetAttribute(k_terrain, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kCells * (int)sizeof(float)));
  *out = t;
  return ODG_OK;
}

void odg_terrain_destroy(OdgTerrain* t) {
  if (!t) return;
  Scope scope(t->device);
  cudaFree(t->d_data); cudaFree(t->d_gen);
  delete t;
}

static int launch(OdgTerrain* t, const uint8_t* mask, const float* raw, const float* radius, void* stream) {
  Scope scope(t->device);
  TerrainArgs A;
  A.N = t->N; A.seed_lo = (uint32_t)t->seed; A.seed_hi = (uint32_t)(t->seed >> 32); A.first_env_id = t->first_env_id;
  A.data = t->d_data; A.gen = t->d_gen; A.mask = mask; A.raw = raw; A.radius = radius;
  A.start_x = 0.0; A.start_y = 0.0;                // initial_qpos_home[0:2] (train2.py:222-223; the keyframe starts at the origin)
  k_terrain<<<t->N, kThreads, 2 * kCells * sizeof(float), static_cast<cudaStream_t>(stream)>>>(A);
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

int odg_terrain_generate(OdgTerrain* t, const uint8_t* mask_dev, void* stream) {
  if (!t) return set_error(ODG_ERR_INVALID, "odg_terrain_generate: null handle");
  return launch(t, mask_dev, nullptr, nullptr, stream);
}

int odg_terrain_from_raw(OdgTerrain* t, const float* raw_dev, const float* radius_dev, void* stream) {
  if (!t || !raw_dev || !radius_dev) return set_error(ODG_ERR_INVALID, "odg_terrain_from_raw: null argument");
  return launch(t, nullptr, raw_dev, radius_dev, stream);
}

int odg_terrain_height(const OdgTerrain* t, const float* xy_dev, float* height_dev, void* stream) {
  if (!t || !xy_dev || !height_dev) return set_error(ODG_ERR_INVALID, "odg_terrain_height: null argument");
  Scope scope(t->device);
  k_terrain_height<<<(t->N + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(t->d_data, xy_dev, height_dev, t->N);
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

const float* odg_terrain_data(const OdgTerrain* t) { return t ? t->d_data : nullptr; }

}  // extern "C"
