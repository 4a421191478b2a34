// odg_sim_internal.h — the handle behind `OdgSim*`, shared by the translation units of libodgsim.
#pragma once
#include <string>

#include "odg_prep.h"

struct SmemLayout { int lc_floats, gc_floats, vert_floats; };

struct OdgSim {
  int device = 0, N = 0, num_sms = 0;
  odg::Prepared prep;
  odg::SimPtrs P{};
  float* d_lc = nullptr; float* d_gc = nullptr; float* d_vert = nullptr;
  void* d_state = nullptr;           // one allocation behind all SoA arrays
  SmemLayout L{};
  size_t smem_step = 0, smem_const = 0;
  int step_block = 128, step_grid = 1, step_lanes = 32;
  int cfg_lanes = 0, cfg_block = 0, cfg_lockstep = -1, cfg_fat = -1;   // OdgEnvConfig::launch_*
  int step_fat = 1;                                       // which instantiation of the step kernel launch_step uses
  long long launches = 0;
};



namespace odg_internal { int set_error(int code, const std::string& msg); }

#ifdef __CUDACC__
// Stage the model constants of a persistent block into shared memory with 16-byte asynchronous copies (LDGSTS): every
// thread issues all of its copies back to back and waits once, instead of a load -> store round trip per element (at
// 4096 environments a block lives for a single tile, so this prologue is on the critical path of the step).
__device__ __forceinline__ void stage_constants(float* smem, const float* __restrict__ g_lc, const float* __restrict__ g_gc,
                                                const float* __restrict__ g_vert, SmemLayout L,
                                                const float4** s_vert, const float** s_lc, const float** s_gc) {
  // layout: [vert (16B aligned)][lc][gc]; all three sizes are multiples of 4 floats (host: SmemLayout)
  float* slc = smem + L.vert_floats;
  float* sgc = slc + L.lc_floats;
  auto copy16 = [](float* dst, const float* src, int n_floats) {
    for (int i = threadIdx.x * 4; i < n_floats; i += blockDim.x * 4) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
    }
  };
  copy16(smem, g_vert, L.vert_floats);
  copy16(slc, g_lc, L.lc_floats);
  copy16(sgc, g_gc, L.gc_floats);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  *s_vert = reinterpret_cast<const float4*>(smem); *s_lc = slc; *s_gc = sgc;
}
#endif

