// env_kernels.cuh - K1 kernels (step / reset) and their templated launchers. Included by one
// translation unit per team width G (env_inst_g*.cu) so the instantiations compile in parallel.
//
// CTA layout: block_threads(G) threads = four teams (more for narrow teams), one environment each. Dynamic shared memory:
//   [ lookup tables shared by the CTA | per-team double scratch ... | per-team word scratch ... ]
#pragma once
#include "lib_common.h"
#include "spec_build.h"

namespace marlsc {

// Threads per CTA: 128 for teams inside a warp, four teams for multi-warp teams.
template <int G>
struct Block {
  static constexpr int threads = G <= 32 ? 128 : 4 * G;
  static constexpr int teams = threads / G;
};
inline int block_teams(int G) { return G <= 32 ? 128 / G : 4; }

struct LaunchArgs {
  DevSpec ds;
  marlsc_env_state_t st;
  int max_smem_optin;
  bool lean;   // every capability this launch needs is in kCapsLean
};

__device__ __forceinline__ Tables stage_tables(const DevSpec& sp, unsigned char* smem) {
  // cooperative copy of the per-CTA lookup tables (a few KB, L2 resident) into shared memory
  const void* src[7] = {sp.skw, sp.pen_rate, sp.hold_rate, sp.prio, sp.prio_static, sp.home_mask, sp.lead_u8};
  const int at[7] = {sp.t_skw, sp.t_pen, sp.t_hold, sp.t_prio, sp.t_pstat, sp.t_hmask, sp.t_lead};
  const int bytes[7] = {sp.S * 8, sp.S * 8, sp.S * 8, sp.R * ((sp.W + 3) & ~3), sp.R, sp.home_mask ? sp.R * 4 : 0, sp.W * sp.S};
#pragma unroll 1
  for (int k = 0; k < 7; ++k) {
    // every table starts 16-byte aligned in the device blob and in shared memory and is padded to 16
    const uint32_t* s4 = static_cast<const uint32_t*>(src[k]);
    uint32_t* d4 = reinterpret_cast<uint32_t*>(smem + at[k]);
    for (int i = threadIdx.x; i < ((bytes[k] + 3) >> 2); i += blockDim.x) d4[i] = s4[i];
  }
  Tables tb;
  tb.skw = reinterpret_cast<const double*>(smem + sp.t_skw);
  tb.pen = reinterpret_cast<const double*>(smem + sp.t_pen);
  tb.hold = reinterpret_cast<const double*>(smem + sp.t_hold);
  tb.prio = smem + sp.t_prio;
  tb.pstat = smem + sp.t_pstat;
  tb.hmask = sp.home_mask ? reinterpret_cast<const uint32_t*>(smem + sp.t_hmask) : nullptr;
  tb.lead = smem + sp.t_lead;
  return tb;
}

// Geometries for which the lean instantiation is built (the automatic choices of auto_team_size()).
constexpr bool has_lean(int G, int SPL) {
  return (G == 1 && (SPL == 2 || SPL == 4 || SPL == 8)) || (G == 4 && SPL == 4) || (G == 8 && SPL == 4) ||
         (G == 16 && SPL == 4) || (G == 32 && (SPL == 4 || SPL == 8 || SPL == 16)) || G > 32;
}
// Multi-warp teams exist for the lean instantiation only (the generic allocation votes across the team per
// warehouse visit, which does not pay across warps); the host falls back to 32 lanes for generic launches.
constexpr bool has_generic(int G) { return G <= 32; }
#ifndef MARLSC_G1_MIN_BLOCKS
#define MARLSC_G1_MIN_BLOCKS 6
#endif
constexpr int min_blocks(int G, uint32_t CAPS) {
  return CAPS != kCapsLean ? 1 : (G == 32 ? 6 : (G > 32 ? 4 : (G == 1 ? MARLSC_G1_MIN_BLOCKS : 1)));   // two-warp teams: 64 registers beat a fifth CTA
}

template <int G, int SPL, uint32_t CAPS>
__global__ void __launch_bounds__(Block<G>::threads, min_blocks(G, CAPS))
env_step_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ marlsc_env_state_t st,
                const __grid_constant__ marlsc_step_io_t io, int t) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double xchg[Block<G>::threads / 32];
  constexpr int TEAMS = Block<G>::teams;
  const Tables tb = stage_tables(sp, smem);
  __syncthreads();
  const int team = threadIdx.x / G;
  const int64_t e = (int64_t)blockIdx.x * TEAMS + team;
  const unsigned live = __ballot_sync(0xffffffffu, e < st.num_envs);
  if (e >= st.num_envs) return;   // whole teams leave together; everything below is team-local
  Team<G> tm;
  tm.init(xchg, live);
  Scratch sc;
  unsigned char* base = smem + sp.t_bytes;
  sc.d = reinterpret_cast<double*>(base) + (size_t)team * sp.d_words;
  sc.w = reinterpret_cast<int32_t*>(base + (size_t)TEAMS * sp.d_words * sizeof(double)) + (size_t)team * sp.w_words;
  step_env<G, SPL, CAPS>(sp, tb, tm, sc, st, io, e, t);
}

template <int G, int SPL, uint32_t CAPS>
__global__ void __launch_bounds__(Block<G>::threads)
env_reset_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ marlsc_env_state_t st,
                 const int32_t* __restrict__ init_inventory, int per_env, float* __restrict__ obs) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ double xchg[Block<G>::threads / 32];
  constexpr int TEAMS = Block<G>::teams;
  const Tables tb = stage_tables(sp, smem);
  __syncthreads();
  const int team = threadIdx.x / G;
  const int64_t e = (int64_t)blockIdx.x * TEAMS + team;
  if (e >= st.num_envs) return;
  Team<G> tm;
  tm.init(xchg);
  reset_env<G, SPL, CAPS>(sp, tb, tm, st, init_inventory, per_env, obs, e);
}

inline size_t step_smem_bytes(const DevSpec& ds, int G) {
  const int teams = block_teams(G);
  return (size_t)ds.t_bytes + (size_t)teams * ((size_t)ds.d_words * sizeof(double) + (size_t)ds.w_words * sizeof(int32_t));
}

template <int G, int SPL, uint32_t CAPS>
int launch_step_caps(const LaunchArgs& a, const marlsc_step_io_t& io, int t, cudaStream_t s) {
  const int teams = Block<G>::teams;
  const size_t smem = step_smem_bytes(a.ds, G);
  if ((int)smem > a.max_smem_optin)
    return set_error(MARLSC_EUNSUPPORTED, "shared-memory scratch of " + std::to_string(smem) + " bytes per CTA does not fit; use a wider team");
  // kernel attributes belong to the device (context) they were set on: cache per device, not per thread
  static std::atomic<size_t> configured[kMaxDevices];
  static std::atomic<bool> carveout[kMaxDevices];
  int dev = 0;
  MARLSC_CUDA(cudaGetDevice(&dev));
  const int di = dev < kMaxDevices ? dev : kMaxDevices - 1;
  if (dev >= kMaxDevices || !carveout[di].load()) {   // the scratch is what bounds residency: ask for the largest shared-memory carveout
    MARLSC_CUDA(cudaFuncSetAttribute((const void*)env_step_kernel<G, SPL, CAPS>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
    carveout[di].store(true);
  }
  if (smem > 48 * 1024 && (dev >= kMaxDevices || smem > configured[di].load())) {
    MARLSC_CUDA(cudaFuncSetAttribute((const void*)env_step_kernel<G, SPL, CAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[di].store(smem);
  }
  const unsigned grid = (unsigned)((a.st.num_envs + teams - 1) / teams);
  env_step_kernel<G, SPL, CAPS><<<grid, Block<G>::threads, smem, s>>>(a.ds, a.st, io, t);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

template <int G, int SPL>
int launch_step_t(const LaunchArgs& a, const marlsc_step_io_t& io, int t, cudaStream_t s) {
  if constexpr (has_lean(G, SPL)) {
    if (a.lean) return launch_step_caps<G, SPL, kCapsLean>(a, io, t, s);
  }
  if constexpr (has_generic(G)) return launch_step_caps<G, SPL, kCapsAll>(a, io, t, s);
  return set_error(MARLSC_EUNSUPPORTED, "multi-warp teams run the lean instantiation only");
}

template <int G, int SPL, uint32_t CAPS>
int launch_reset_caps(const LaunchArgs& a, const int32_t* init, int per_env, float* obs, cudaStream_t s) {
  const int teams = Block<G>::teams;
  const size_t smem = (size_t)a.ds.t_bytes;
  const unsigned grid = (unsigned)((a.st.num_envs + teams - 1) / teams);
  env_reset_kernel<G, SPL, CAPS><<<grid, Block<G>::threads, smem, s>>>(a.ds, a.st, init, per_env, obs);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

template <int G, int SPL>
int launch_reset_t(const LaunchArgs& a, const int32_t* init, int per_env, float* obs, cudaStream_t s) {
  return launch_reset_caps<G, SPL, kCapsAll>(a, init, per_env, obs, s);   // reset is not a hot kernel
}

// K5 launcher (samplers.cu)
int launch_base_stock(const DevSpec& ds, const marlsc_env_state_t& st, const float* level, int level_per_env, int t,
                      float* actions, cudaStream_t s);

// One pair of entry points per team width, defined in env_inst_g*.cu; spl selects the instantiation.
#define MARLSC_DECLARE_G(G)                                                                                   \
  int launch_step_g##G(int spl, const LaunchArgs& a, const marlsc_step_io_t& io, int t, cudaStream_t s);      \
  int launch_reset_g##G(int spl, const LaunchArgs& a, const int32_t* init, int per_env, float* obs, cudaStream_t s);
MARLSC_DECLARE_G(1)
MARLSC_DECLARE_G(2)
MARLSC_DECLARE_G(4)
MARLSC_DECLARE_G(8)
MARLSC_DECLARE_G(16)
MARLSC_DECLARE_G(32)
MARLSC_DECLARE_G(64)

#define MARLSC_SPL_CASE(G, SPL, FN, ...) case SPL: return FN<G, SPL>(__VA_ARGS__);
#define MARLSC_DEFINE_G(G, CASES_STEP, CASES_RESET)                                                           \
  namespace marlsc {                                                                                          \
  int launch_step_g##G(int spl, const LaunchArgs& a, const marlsc_step_io_t& io, int t, cudaStream_t s) {     \
    switch (spl) { CASES_STEP default: break; }                                                               \
    return set_error(MARLSC_EUNSUPPORTED, "no kernel instantiated for this team size / SKU count");           \
  }                                                                                                           \
  int launch_reset_g##G(int spl, const LaunchArgs& a, const int32_t* init, int per_env, float* obs,           \
                        cudaStream_t s) {                                                                     \
    switch (spl) { CASES_RESET default: break; }                                                              \
    return set_error(MARLSC_EUNSUPPORTED, "no kernel instantiated for this team size / SKU count");           \
  }                                                                                                           \
  }

}  // namespace marlsc
