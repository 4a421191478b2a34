// odg_policy.cu — the rollout side of libodgsim (C ABI: include/odg_policy.h).  sm_100a only.
//
//   k_mlp        ActorCritic forward for 128 environments per CTA on the 5th-generation tensor cores:
//                obs -> Linear(S,512)-tanh -> Linear(512,256)-tanh -> Linear(256,A)-tanh   (actor)
//                    -> Linear(S,512)-tanh -> Linear(512,256)-tanh -> Linear(256,1)        (critic)
//                (reference: sim2real/train.py:132-149) fused with Normal sampling and log-prob (:542-543).
//                bf16 operands in shared memory (canonical K-major core-matrix layout, no swizzle), fp32
//                accumulators in TMEM (tcgen05.mma issued by one thread), weights streamed chunk by chunk with
//                1-D bulk async copies (TMA engine, mbarrier complete_tx) double-buffered against the MMAs,
//                activations never leave the SM: the 512- and 256-wide hidden layers go TMEM -> registers
//                (tcgen05.ld) -> bias+tanh -> bf16 -> shared memory as the next layer's A operand.
//   k_pack       fp32 torch state_dict weights -> bf16 operand chunks in that canonical layout.
//   k_gae*       GAE scan + deterministic advantage statistics; k_adv_norm: normalisation.
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/odg.h"
#include "../../include/odg_policy.h"
#include "odg_core.cuh"      // philox4x32 / u01 (same counter-based RNG as the step kernel and the oracle)

namespace odg_internal { int set_error(int code, const std::string& msg); }

namespace {

using odg_internal::set_error;
#define CUDA_TRY(expr)                                                                         \
  do { cudaError_t e_ = (expr);                                                                \
       if (e_ != cudaSuccess) return set_error(ODG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); } while (0)

constexpr int kH1 = ODG_POLICY_H1, kH2 = ODG_POLICY_H2;
constexpr int kM = 128;                  // environments (rows) per CTA = UMMA M
constexpr int kNOut = 16;                // output layer padded to the smallest UMMA N
constexpr int kK2Chunk = 64;             // columns of W2 streamed per chunk
constexpr int kNumChunks = 2 + kH1 / kK2Chunk + 1;   // W1 halves, W2 k-chunks, W3  (= 11 per network)
constexpr uint32_t kStreamSample = 0x53414d50u;      // Philox stream id "SAMP"

// ---- shared memory map (bytes)
constexpr int kA0Bytes = kM * ODG_POLICY_MAX_STATE * 2;       // 16 KB  obs tile
constexpr int kA1Bytes = kM * kH1 * 2;                         // 128 KB hidden 1 (hidden 2 aliases its first half)
constexpr int kWBufBytes = 256 * kK2Chunk * 2;                 // 32 KB  one weight chunk
constexpr int kOffA0 = 0, kOffA1 = kOffA0 + kA0Bytes, kOffW = kOffA1 + kA1Bytes, kOffBar = kOffW + 2 * kWBufBytes;
constexpr int kBiasFloats = kH1 + kH2 + kNOut;                  // per network: b1 | b2 | b3
constexpr int kOffBias = kOffBar + 64;
constexpr int kSmemBytes = kOffBias + 2 * kBiasFloats * 4;

// byte offset of element (r, k) of an operand tile with kc columns in the canonical K-major no-swizzle layout:
// 8x8 core matrices (8 rows x 16 bytes, contiguous 128 B); K-adjacent core matrices are 128 B apart (LBO),
// 8-row groups are (kc/8)*128 B apart (SBO).
__host__ __device__ inline int canon_off(int r, int k, int kc) {
  return (r >> 3) * ((kc >> 3) * 128) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2;
}

struct ChunkDesc { int rows, kc; int layer, row0, k0; };      // layer 0..2; rows x kc block of W_layer at (row0, k0)
__host__ __device__ inline ChunkDesc chunk_desc(int c, int K0p) {
  ChunkDesc d;
  if (c < 2) { d.rows = 256; d.kc = K0p; d.layer = 0; d.row0 = c * 256; d.k0 = 0; }
  else if (c < 2 + kH1 / kK2Chunk) { d.rows = 256; d.kc = kK2Chunk; d.layer = 1; d.row0 = 0; d.k0 = (c - 2) * kK2Chunk; }
  else { d.rows = kNOut; d.kc = kH2; d.layer = 2; d.row0 = 0; d.k0 = 0; }
  return d;
}
__host__ __device__ inline int chunk_bytes(int c, int K0p) { ChunkDesc d = chunk_desc(c, K0p); return d.rows * d.kc * 2; }
__host__ __device__ inline int chunk_offset(int c, int K0p) { int o = 0; for (int i = 0; i < c; i++) o += chunk_bytes(i, K0p); return o; }

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a barrier that never completes is a bug; trap (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(b, parity)) if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // tcgen05 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
  // base_offset 0, layout_type SWIZZLE_NONE (0) [61,64)
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  // kind::f16 instruction descriptor: D = F32 [4,6)=1, A = BF16 [7,10)=1, B = BF16 [10,13)=1, A and B K-major,
  // N>>3 [17,23), M>>4 [24,29)
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

#ifdef ODG_MLP_TIMING
__device__ long long g_mlp_t[64];
#define MLP_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_mlp_t[i] = clock64(); } while (0)
#else
#define MLP_STAMP(i) do { } while (0)
#endif

struct MlpParams {
  const float* obs; int n, S, A, K0p;
  const uint8_t* wpack[2];     // actor / critic operand chunks (bf16, canonical layout, chunk after chunk)
  const float* bias[2];        // b1[512] | b2[256] | b3[16]
  const float* log_std;
  float* mean; float* value; float* action; float* logp;
  uint32_t seed_lo, seed_hi, step; const uint32_t* step_base; int first_row;
};

// hidden-layer epilogue: D[row][0..ncols) (TMEM, fp32) + bias -> tanh -> bf16 -> A operand tile with `kc_out` columns,
// written at columns [kofs, kofs + ncols)
__device__ __forceinline__ void epilogue_hidden(uint32_t taddr_row, int ncols, const float* __restrict__ bias,
                                                uint8_t* a_out, int kc_out, int kofs, int row) {
  for (int cb = 0; cb < ncols; cb += 32) {
    uint32_t v[32];
    tmem_ld32(taddr_row + cb, v);
#pragma unroll
    for (int g = 0; g < 4; g++) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int col = cb + g * 8 + i * 2;
        float x0 = tanh_fast(__uint_as_float(v[g * 8 + i * 2]) + bias[col]);
        float x1 = tanh_fast(__uint_as_float(v[g * 8 + i * 2 + 1]) + bias[col + 1]);
        w[i] = pack_bf16(x0, x1);
      }
      *reinterpret_cast<uint4*>(a_out + canon_off(row, kofs + cb + g * 8, kc_out)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__global__ void __launch_bounds__(kM, 1) k_mlp(const MlpParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA0 = smem + kOffA0; uint8_t* sA1 = smem + kOffA1;
  uint8_t* sW[2] = { smem + kOffW, smem + kOffW + kWBufBytes };
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + kOffBar);       // [2] weight chunk landed in buffer b
  uint64_t* bar_free = bar_w + 2;                                       // [2] the MMAs that read buffer b have retired
  uint64_t* bar_acc = bar_w + 4;                                        // the accumulator of a layer is complete
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_w + 6);
  float* s_bias = reinterpret_cast<float*>(smem + kOffBias);            // both networks' biases (epilogue broadcasts)
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row = tid;                                                  // TMEM lane = row of the tile
  const int grow = blockIdx.x * kM + row;                               // row in the batch
  const int K0p = P.K0p;
  const int total_chunks = 2 * kNumChunks;

  if (tid == 0) {
    mbar_init(&bar_w[0], 1); mbar_init(&bar_w[1], 1); mbar_init(&bar_free[0], 1); mbar_init(&bar_free[1], 1); mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < 2 * kBiasFloats; i += kM) s_bias[i] = P.bias[i / kBiasFloats][i % kBiasFloats];
  // obs tile -> bf16 A0 (rows past the batch and columns past S are zero)
  {
    const float* o = P.obs + (size_t)grow * P.S;
    for (int k8 = 0; k8 < K0p; k8 += 8) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int k = k8 + 2 * i;
        float x0 = (grow < P.n && k < P.S) ? o[k] : 0.f;
        float x1 = (grow < P.n && k + 1 < P.S) ? o[k + 1] : 0.f;
        w[i] = pack_bf16(x0, x1);
      }
      *reinterpret_cast<uint4*>(sA0 + canon_off(row, k8, K0p)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  MLP_STAMP(0);
  const uint32_t tmem = *s_tmem;
  const uint32_t tmem_row = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's lane quarter
  constexpr uint32_t kD1 = 0, kD2 = 256;                                // accumulator columns: layer 1 / 3 at 0, layer 2 at 256

  // weight chunk c of the flat sequence (actor chunks, then critic chunks) -> buffer c & 1
  auto prefetch = [&](int c) {
    const int net = c / kNumChunks, cc = c % kNumChunks;
    const uint32_t bytes = (uint32_t)chunk_bytes(cc, K0p);
    mbar_expect_tx(&bar_w[c & 1], bytes);
    bulk_g2s(sW[c & 1], P.wpack[net] + chunk_offset(cc, K0p), bytes, &bar_w[c & 1]);
  };
  if (tid == 0) prefetch(0);

  int c = 0;                                                            // flat chunk counter (uniform across threads)
  int n_acc = 0;                                                        // accumulators completed so far (bar_acc phase)
  // One chunk, issued by thread 0 only: wait for its weights, issue `ksteps` MMAs (K = 16 each)
  // D[.., ncols] (+)= A[128 x 16k] * W_chunk^T, commit them to the buffer's "free" barrier (and, for the last chunk of
  // a layer, to the accumulator barrier), then prefetch the next chunk into the other buffer as soon as the MMAs that
  // read it (chunk c-1) have retired. MMAs of consecutive chunks queue back to back in the tensor pipe; nobody waits for
  // an individual chunk. The other 127 threads only wait for the accumulator (`wait_acc`).
  auto run_chunk = [&](uint32_t a_saddr, uint32_t a_sbo, int ksteps, uint32_t dcol, int ncols, bool accumulate, bool last) {
    if (tid == 0) {
      mbar_wait(&bar_w[c & 1], (uint32_t)((c >> 1) & 1));
      tc_fence_after();
      const ChunkDesc d = chunk_desc(c % kNumChunks, K0p);
      const uint32_t b_saddr = smem_u32(sW[c & 1]), b_sbo = (uint32_t)(d.kc >> 3) * 128u;
      const uint32_t idesc = make_idesc(kM, ncols);
      for (int k = 0; k < ksteps; k++)
        umma_bf16(tmem + dcol, make_desc(a_saddr + k * 256, 128, a_sbo), make_desc(b_saddr + k * 256, 128, b_sbo), idesc,
                  (accumulate || k > 0) ? 1u : 0u);
      umma_commit(&bar_free[c & 1]);
      if (last) umma_commit(bar_acc);
      if (c + 1 < total_chunks) {
        if (c >= 1) mbar_wait(&bar_free[(c + 1) & 1], (uint32_t)(((c - 1) >> 1) & 1));   // chunk c-1 read that buffer
        prefetch(c + 1);
      }
    }
    c++;
  };
  auto wait_acc = [&]() { mbar_wait(bar_acc, (uint32_t)(n_acc & 1)); tc_fence_after(); n_acc++; };
  // all tcgen05.ld of an accumulator region done + A operand writes visible to the async proxy, before the next MMAs
  auto sync_after_epilogue = [&]() { fence_async_smem(); tc_fence_before(); __syncthreads(); tc_fence_after(); };

  float outv[kNOut];
  for (int net = 0; net < 2; net++) {
    const float* b1 = s_bias + net * kBiasFloats; const float* b2 = b1 + kH1; const float* b3 = b2 + kH2;
    for (int h = 0; h < 2; h++) {                                        // layer 1, two 256-wide halves
      run_chunk(smem_u32(sA0), (uint32_t)(K0p >> 3) * 128u, K0p / 16, kD1, 256, false, true);
      wait_acc();
      MLP_STAMP(1 + net * 12 + h * 2);
      epilogue_hidden(tmem_row + kD1, 256, b1 + h * 256, sA1, kH1, h * 256, row);
      sync_after_epilogue();
      MLP_STAMP(2 + net * 12 + h * 2);
    }
    for (int kc = 0; kc < kH1 / kK2Chunk; kc++)                          // layer 2, K streamed in 64-column chunks
      run_chunk(smem_u32(sA1) + kc * (kK2Chunk / 8) * 128, (uint32_t)(kH1 >> 3) * 128u, kK2Chunk / 16, kD2, 256, kc > 0,
                kc == kH1 / kK2Chunk - 1);
    wait_acc();
    MLP_STAMP(5 + net * 12);
    epilogue_hidden(tmem_row + kD2, 256, b2, sA1, kH2, 0, row);          // hidden 2 overwrites hidden 1 (all its MMAs retired)
    sync_after_epilogue();
    MLP_STAMP(6 + net * 12);
    run_chunk(smem_u32(sA1), (uint32_t)(kH2 >> 3) * 128u, kH2 / 16, kD1, kNOut, false, true);   // output layer
    wait_acc();
    MLP_STAMP(7 + net * 12);
    {
      uint32_t v[16];
      tmem_ld16(tmem_row + kD1, v);
      if (net == 0) {
#pragma unroll
        for (int a = 0; a < kNOut; a++) outv[a] = tanhf(__uint_as_float(v[a]) + b3[a]);   // actor: Tanh head
      } else if (grow < P.n && P.value) {
        P.value[grow] = __uint_as_float(v[0]) + b3[0];
      }
    }
    sync_after_epilogue();
  }
  MLP_STAMP(30);
  // ---- Normal(mean, exp(log_std)): sample + log-prob  (sim2real/train.py:542-543)
  if (grow < P.n) {
    const uint32_t step_ctr = P.step + (P.step_base ? __ldg(P.step_base) : 0u);
    float lp = 0.f;
#pragma unroll
    for (int blk = 0; blk < kNOut / 4; blk++) {
      if (blk * 4 >= P.A) break;
      uint32_t r[4];
      odg::philox4x32(P.seed_lo, P.seed_hi, (uint32_t)(P.first_row + grow), step_ctr, (uint32_t)blk, kStreamSample, r);
#pragma unroll
      for (int pr = 0; pr < 2; pr++) {                                  // Box-Muller: 2 uniforms -> 2 normals
        const float u1 = ((float)(r[2 * pr] >> 8) + 1.0f) * 5.9604644775390625e-08f;     // (0, 1]
        const float u2 = odg::u01(r[2 * pr + 1]);
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs; sincosf(6.283185307179586f * u2, &sn, &cs);
        const float e[2] = { rad * cs, rad * sn };
#pragma unroll
        for (int q = 0; q < 2; q++) {
          const int a = blk * 4 + pr * 2 + q;
          if (a < P.A) {
            const float ls = __ldg(P.log_std + a);
            if (P.mean) P.mean[(size_t)grow * P.A + a] = outv[a];
            if (P.action) P.action[(size_t)grow * P.A + a] = outv[a] + expf(ls) * e[q];
            lp += -0.5f * e[q] * e[q] - ls - 0.9189385332046727f;
          }
        }
      }
    }
    if (P.logp) P.logp[grow] = lp;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

// fp32 [out][in] weights -> bf16 chunks. One thread per destination element pair.
__global__ void k_pack(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                       int S, int nout, int K0p, uint8_t* __restrict__ dst) {
  const int c = blockIdx.y;
  const ChunkDesc d = chunk_desc(c, K0p);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;                  // element index r * kc + k
  if (e >= d.rows * d.kc) return;
  const int r = e / d.kc, k = e % d.kc;
  float v = 0.f;
  if (d.layer == 0) { if (k < S) v = w0[(size_t)(d.row0 + r) * S + k]; }
  else if (d.layer == 1) v = w1[(size_t)r * kH1 + d.k0 + k];
  else { if (r < nout) v = w2[(size_t)r * kH2 + k]; }
  *reinterpret_cast<__nv_bfloat16*>(dst + chunk_offset(c, K0p) + canon_off(r, k, d.kc)) = __float2bfloat16_rn(v);
}
__global__ void k_pack_bias(const float* b0, const float* b1, const float* b2, int nout, float* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kH1) dst[i] = b0[i];
  else if (i < kH1 + kH2) dst[i] = b1[i - kH1];
  else if (i < kH1 + kH2 + kNOut) { const int a = i - kH1 - kH2; dst[i] = a < nout ? b2[a] : 0.f; }
}

// ---------------------------------------------------------------------------------------------- GAE
constexpr int kGaeBlock = 256;
// one thread per environment: reverse scan over T steps; block-level fixed-order sum of adv and adv^2
__global__ void k_gae(const float* __restrict__ rew, const float* __restrict__ val, const uint8_t* __restrict__ done,
                      int T, int n, float gamma, float lam, float* __restrict__ adv, float* __restrict__ ret,
                      double* __restrict__ partial) {
  __shared__ double s_sum[kGaeBlock], s_sq[kGaeBlock];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  double sum = 0.0, sq = 0.0;
  if (e < n) {
    float a = 0.f, vnext = val[(size_t)T * n + e];
    for (int t = T - 1; t >= 0; t--) {
      const size_t i = (size_t)t * n + e;
      const float m = done[i] ? 0.f : 1.f, v = val[i];
      const float delta = rew[i] + gamma * vnext * m - v;
      a = delta + gamma * lam * m * a;
      adv[i] = a; if (ret) ret[i] = a + v;
      sum += (double)a; sq += (double)a * (double)a;
      vnext = v;
    }
  }
  s_sum[threadIdx.x] = sum; s_sq[threadIdx.x] = sq;
  __syncthreads();
  for (int d = kGaeBlock / 2; d > 0; d >>= 1) {
    if (threadIdx.x < d) { s_sum[threadIdx.x] += s_sum[threadIdx.x + d]; s_sq[threadIdx.x] += s_sq[threadIdx.x + d]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[2 * blockIdx.x] = s_sum[0]; partial[2 * blockIdx.x + 1] = s_sq[0]; }
}
__global__ void k_gae_finish(const double* __restrict__ partial, int nblocks, long long count, double* __restrict__ stats) {
  __shared__ double s_sum[kGaeBlock], s_sq[kGaeBlock];
  double sum = 0.0, sq = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += kGaeBlock) { sum += partial[2 * i]; sq += partial[2 * i + 1]; }
  s_sum[threadIdx.x] = sum; s_sq[threadIdx.x] = sq;
  __syncthreads();
  for (int d = kGaeBlock / 2; d > 0; d >>= 1) {
    if (threadIdx.x < d) { s_sum[threadIdx.x] += s_sum[threadIdx.x + d]; s_sq[threadIdx.x] += s_sq[threadIdx.x + d]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { stats[0] = s_sum[0]; stats[1] = s_sq[0]; stats[2] = (double)count; }
}
__global__ void k_adv_norm(float* __restrict__ adv, long long count, const double* __restrict__ stats) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const double n = stats[2], mean = stats[0] / n;
  const double var = n > 1.0 ? fmax(0.0, (stats[1] - n * mean * mean) / (n - 1.0)) : 0.0;   // torch.std: unbiased
  adv[i] = (float)(((double)adv[i] - mean) / (sqrt(var) + 1e-8));
}

}  // namespace

struct OdgPolicy {
  int device = 0, S = 0, A = 0, K0p = 0;
  uint8_t* d_wpack[2] = { nullptr, nullptr };
  float* d_bias[2] = { nullptr, nullptr };
  float* d_log_std = nullptr;
  double* d_partial = nullptr; int partial_cap = 0;
  size_t wpack_bytes = 0;
  long long launches = 0;
};

namespace {
struct DevGuard {
  int prev = -1;
  explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
// device that owns a caller buffer (entry points without a handle run where their tensors live)
int device_of(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess || a.type != cudaMemoryTypeDevice) { cudaGetLastError(); return -1; }
  return a.device;
}
}  // namespace

extern "C" {

int odg_policy_create(int state_dim, int action_dim, int device, OdgPolicy** out) {
  if (!out || state_dim < 1 || state_dim > ODG_POLICY_MAX_STATE || action_dim < 1 || action_dim > ODG_POLICY_MAX_ACTION)
    return set_error(ODG_ERR_INVALID, "odg_policy_create: state_dim must be in [1,64], action_dim in [1,16]");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return set_error(ODG_ERR_NO_DEVICE, "no CUDA device: libodgsim has no CPU fallback");
  if (device < 0 || device >= ndev) return set_error(ODG_ERR_INVALID, "odg_policy_create: device out of range");
  DevGuard guard(device);
  OdgPolicy* p = new (std::nothrow) OdgPolicy();
  if (!p) return set_error(ODG_ERR_ALLOC, "out of host memory");
  p->device = device; p->S = state_dim; p->A = action_dim; p->K0p = (state_dim + 15) / 16 * 16;
  p->wpack_bytes = (size_t)chunk_offset(kNumChunks, p->K0p);
  for (int net = 0; net < 2; net++) {
    if (cudaMalloc(&p->d_wpack[net], p->wpack_bytes) != cudaSuccess ||
        cudaMalloc(&p->d_bias[net], (kH1 + kH2 + kNOut) * sizeof(float)) != cudaSuccess) {
      odg_policy_destroy(p); return set_error(ODG_ERR_ALLOC, "cudaMalloc(policy) failed");
    }
  }
  if (cudaMalloc(&p->d_log_std, ODG_POLICY_MAX_ACTION * sizeof(float)) != cudaSuccess) {
    odg_policy_destroy(p); return set_error(ODG_ERR_ALLOC, "cudaMalloc(policy) failed");
  }
  CUDA_TRY(cudaMemset(p->d_log_std, 0, ODG_POLICY_MAX_ACTION * sizeof(float)));
  CUDA_TRY(cudaFuncSetAttribute(k_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  *out = p;
  return ODG_OK;
}

void odg_policy_destroy(OdgPolicy* p) {
  if (!p) return;
  DevGuard guard(p->device);
  for (int net = 0; net < 2; net++) { cudaFree(p->d_wpack[net]); cudaFree(p->d_bias[net]); }
  cudaFree(p->d_log_std);
  delete p;
}

int odg_policy_load(OdgPolicy* p, const OdgPolicyWeights* w, void* stream) {
  if (!p || !w) return set_error(ODG_ERR_INVALID, "odg_policy_load: null argument");
  DevGuard guard(p->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int net = 0; net < 2; net++) {
    const float* const* W = net == 0 ? w->actor_w : w->critic_w;
    const float* const* B = net == 0 ? w->actor_b : w->critic_b;
    for (int i = 0; i < 3; i++) if (!W[i] || !B[i]) return set_error(ODG_ERR_INVALID, "odg_policy_load: missing weight pointer");
    const int nout = net == 0 ? p->A : 1;
    const int max_elems = 256 * (p->K0p > kK2Chunk ? p->K0p : kK2Chunk);
    dim3 grid((max_elems + 255) / 256, kNumChunks);
    k_pack<<<grid, 256, 0, st>>>(W[0], W[1], W[2], p->S, nout, p->K0p, p->d_wpack[net]);
    k_pack_bias<<<(kH1 + kH2 + kNOut + 255) / 256, 256, 0, st>>>(B[0], B[1], B[2], nout, p->d_bias[net]);
    p->launches += 2;
  }
  if (w->action_log_std)
    CUDA_TRY(cudaMemcpyAsync(p->d_log_std, w->action_log_std, p->A * sizeof(float), cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

int odg_policy_forward(OdgPolicy* p, const float* obs_dev, int n, float* mean_dev, float* value_dev, float* action_dev,
                       float* logp_dev, uint64_t seed, uint32_t step, const uint32_t* step_base_dev,
                       int first_row_id, void* stream) {
  if (!p || !obs_dev || n < 1) return set_error(ODG_ERR_INVALID, "odg_policy_forward: bad arguments");
  DevGuard guard(p->device);
  MlpParams P;
  P.obs = obs_dev; P.n = n; P.S = p->S; P.A = p->A; P.K0p = p->K0p;
  P.wpack[0] = p->d_wpack[0]; P.wpack[1] = p->d_wpack[1]; P.bias[0] = p->d_bias[0]; P.bias[1] = p->d_bias[1];
  P.log_std = p->d_log_std; P.mean = mean_dev; P.value = value_dev; P.action = action_dev; P.logp = logp_dev;
  P.seed_lo = (uint32_t)seed; P.seed_hi = (uint32_t)(seed >> 32); P.step = step; P.step_base = step_base_dev; P.first_row = first_row_id;
  k_mlp<<<(n + kM - 1) / kM, kM, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(P);
  p->launches++;
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

int odg_gae(const float* reward_dev, const float* value_dev, const uint8_t* done_dev, int T, int n, float gamma,
            float lambda, float* adv_dev, float* ret_dev, double* stats_dev, void* stream) {
  if (!reward_dev || !value_dev || !done_dev || !adv_dev || T < 1 || n < 1)
    return set_error(ODG_ERR_INVALID, "odg_gae: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblocks = (n + kGaeBlock - 1) / kGaeBlock;
  const int dev = device_of(reward_dev);
  if (dev < 0) return set_error(ODG_ERR_INVALID, "odg_gae: reward_dev is not device memory");
  DevGuard guard(dev);
  // per-block partial sums: a stream-ordered allocation that lives for this call only (no state shared between
  // streams, threads or devices)
  double* partial = nullptr;
  CUDA_TRY(cudaMallocAsync(&partial, (size_t)2 * nblocks * sizeof(double), st));
  k_gae<<<nblocks, kGaeBlock, 0, st>>>(reward_dev, value_dev, done_dev, T, n, gamma, lambda, adv_dev, ret_dev, partial);
  if (stats_dev) k_gae_finish<<<1, kGaeBlock, 0, st>>>(partial, nblocks, (long long)T * n, stats_dev);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(partial, st);
  if (e != cudaSuccess) return set_error(ODG_ERR_CUDA, std::string("odg_gae: ") + cudaGetErrorString(e));
  return ODG_OK;
}

int odg_normalize_advantages(float* adv_dev, long long count, const double* stats_dev, void* stream) {
  if (!adv_dev || !stats_dev || count < 1) return set_error(ODG_ERR_INVALID, "odg_normalize_advantages: bad arguments");
  const int dev = device_of(adv_dev);
  if (dev < 0) return set_error(ODG_ERR_INVALID, "odg_normalize_advantages: adv_dev is not device memory");
  DevGuard guard(dev);
  k_adv_norm<<<(unsigned)((count + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(adv_dev, count, stats_dev);
  CUDA_TRY(cudaGetLastError());
  return ODG_OK;
}

long long odg_policy_launch_count(const OdgPolicy* p) { return p ? p->launches : 0; }
#ifdef ODG_MLP_TIMING
int odg_mlp_timing(long long* out) { return cudaMemcpyFromSymbol(out, g_mlp_t, sizeof(long long) * 64) == cudaSuccess ? 0 : -1; }
#endif

}  // extern "C"
