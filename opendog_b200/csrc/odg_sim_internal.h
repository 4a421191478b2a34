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
  int* d_order = nullptr; int* d_hist = nullptr; int regroup = 0;
  long long launches = 0;
};

namespace odg_internal { int set_error(int code, const std::string& msg); }
