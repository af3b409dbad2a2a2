// env_step.cu - K1: fused environment step / reset kernels for sm_100a and their C-ABI launchers.
//
// Mapping: one team of TPE threads per environment (env_core.cuh), BLOCK/TPE teams per CTA, each
// team with a private shared-memory scratch. Large shapes (S > 64) use a 4-warp team so the
// streaming phases move 128 elements per instruction while warp 0 walks the order list; tiny shapes
// use one thread per environment.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <new>
#include <string>

#include "lib_common.h"
#include "spec_build.h"

namespace marlsc {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

int set_error(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

}  // namespace marlsc

using namespace marlsc;

struct marlsc_env {
  DevSpec ds;
  HostTables tb;
  int device = 0;
  int team = 1;        // threads per environment in use
  int team_auto = 1;
  void* d_blob = nullptr;  // one allocation holding every device table
  int max_smem_optin = 0;
};

namespace {

template <int TPE>
struct Block {
  static constexpr int kThreads = TPE > 128 ? TPE : 128;
  static constexpr int kTeams = kThreads / TPE;
};

template <int TPE>
__global__ void __launch_bounds__(Block<TPE>::kThreads)
env_step_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ marlsc_env_state_t st,
                const __grid_constant__ marlsc_step_io_t io, int t, int d_stride, int w_stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int TEAMS = Block<TPE>::kTeams;
  const int team_in_block = threadIdx.x / TPE;
  const int64_t e = (int64_t)blockIdx.x * TEAMS + team_in_block;
  if (e >= st.num_envs) return;   // whole teams leave together; barriers below are per team
  Team<TPE> tm;
  tm.init(team_in_block);
  Scratch sc;
  sc.d = reinterpret_cast<double*>(smem) + (size_t)team_in_block * d_stride;
  sc.w = reinterpret_cast<int32_t*>(smem + (size_t)TEAMS * d_stride * sizeof(double)) + (size_t)team_in_block * w_stride;
  step_env<TPE>(sp, tm, sc, st, io, e, t);
}

template <int TPE>
__global__ void __launch_bounds__(Block<TPE>::kThreads)
env_reset_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ marlsc_env_state_t st,
                 const int32_t* __restrict__ init_inventory, int per_env, float* __restrict__ obs, int d_stride,
                 int w_stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int TEAMS = Block<TPE>::kTeams;
  const int team_in_block = threadIdx.x / TPE;
  const int64_t e = (int64_t)blockIdx.x * TEAMS + team_in_block;
  if (e >= st.num_envs) return;
  Team<TPE> tm;
  tm.init(team_in_block);
  Scratch sc;
  sc.d = reinterpret_cast<double*>(smem) + (size_t)team_in_block * d_stride;
  sc.w = reinterpret_cast<int32_t*>(smem + (size_t)TEAMS * d_stride * sizeof(double)) + (size_t)team_in_block * w_stride;
  reset_env<TPE>(sp, tm, sc, st, init_inventory, per_env, obs, e);
}

struct LaunchGeom {
  int block, teams, d_stride, w_stride;
  size_t smem;
  unsigned grid;
};

template <int TPE>
LaunchGeom geom(const DevSpec& ds, int64_t num_envs) {
  LaunchGeom g;
  g.block = Block<TPE>::kThreads;
  g.teams = g.block / TPE;
  g.d_stride = ds.d_words;   // odd strides keep same-offset accesses of neighbouring teams on distinct banks
  g.w_stride = ds.w_words;
  g.smem = (size_t)g.teams * ((size_t)g.d_stride * sizeof(double) + (size_t)g.w_stride * sizeof(int32_t));
  g.grid = (unsigned)((num_envs + g.teams - 1) / g.teams);
  return g;
}

template <int TPE>
int prepare(marlsc_env* env, const void* kernel, const LaunchGeom& g) {
  if ((int)g.smem > env->max_smem_optin)
    return set_error(MARLSC_EUNSUPPORTED, "team scratch of " + std::to_string(g.smem) + " bytes exceeds shared memory; use a larger team size");
  if (g.smem > 48 * 1024) MARLSC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  return MARLSC_OK;
}

template <int TPE>
int launch_step(marlsc_env* env, const marlsc_env_state_t& st, const marlsc_step_io_t& io, int t, cudaStream_t s) {
  const LaunchGeom g = geom<TPE>(env->ds, st.num_envs);
  int rc = prepare<TPE>(env, (const void*)env_step_kernel<TPE>, g);
  if (rc) return rc;
  env_step_kernel<TPE><<<g.grid, g.block, g.smem, s>>>(env->ds, st, io, t, g.d_stride, g.w_stride);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

template <int TPE>
int launch_reset(marlsc_env* env, const marlsc_env_state_t& st, const int32_t* init, int per_env, float* obs,
                 cudaStream_t s) {
  const LaunchGeom g = geom<TPE>(env->ds, st.num_envs);
  int rc = prepare<TPE>(env, (const void*)env_reset_kernel<TPE>, g);
  if (rc) return rc;
  env_reset_kernel<TPE><<<g.grid, g.block, g.smem, s>>>(env->ds, st, init, per_env, obs, g.d_stride, g.w_stride);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

#define MARLSC_DISPATCH_TEAM(team, CALL)                                                   \
  switch (team) {                                                                          \
    case 1: return CALL(1);                                                                \
    case 2: return CALL(2);                                                                \
    case 4: return CALL(4);                                                                \
    case 8: return CALL(8);                                                                \
    case 16: return CALL(16);                                                              \
    case 32: return CALL(32);                                                              \
    case 64: return CALL(64);                                                              \
    case 128: return CALL(128);                                                            \
    case 256: return CALL(256);                                                            \
    default: return set_error(MARLSC_EINVAL, "team size must be a power of two in [1,256]"); \
  }

int check_state(const marlsc_env* env, const marlsc_env_state_t* st) {
  if (!env || !st) return set_error(MARLSC_EINVAL, "null handle or state");
  if (st->num_envs < 1) return set_error(MARLSC_EINVAL, "num_envs must be positive");
  if (st->num_envs > (int64_t)0x7fffffff) return set_error(MARLSC_EINVAL, "num_envs too large");
  if (!st->inventory || !st->ring_qty) return set_error(MARLSC_EINVAL, "state.inventory / state.ring_qty are NULL");
  if (env->ds.lead_mode == MARLSC_LEAD_STOCHASTIC && !st->ring_lead)
    return set_error(MARLSC_EINVAL, "state.ring_lead is required with a stochastic lead-time sampler");
  if (env->ds.need_hist && !st->demand_hist) return set_error(MARLSC_EINVAL, "state.demand_hist is required by this configuration");
  if (env->ds.need_fcst && !st->forecast) return set_error(MARLSC_EINVAL, "state.forecast is required by this configuration");
  return MARLSC_OK;
}

template <typename T>
size_t blob_add(size_t& off, const std::vector<T>& v) {
  off = (off + 15) & ~size_t(15);
  const size_t at = off;
  off += v.size() * sizeof(T);
  return at;
}

}  // namespace

extern "C" {

int marlsc_env_create(const marlsc_env_spec_t* spec, int device, marlsc_env_t** out) {
  if (!spec || !out) return set_error(MARLSC_EINVAL, "null spec or out");
  marlsc_env* env = new (std::nothrow) marlsc_env();
  if (!env) return set_error(MARLSC_ENOMEM, "out of host memory");
  const std::string err = build_devspec(*spec, env->ds, env->tb);
  if (!err.empty()) {
    delete env;
    return set_error(MARLSC_EINVAL, err);
  }
  env->device = device;
  env->team = env->team_auto = auto_team_size(env->ds.S);
  cudaError_t ce = cudaSetDevice(device);
  if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&env->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (ce != cudaSuccess) {
    delete env;
    return set_error(MARLSC_ECUDA, std::string("no usable CUDA device: ") + cudaGetErrorString(ce));
  }
  // pack every table into one device allocation
  const HostTables& t = env->tb;
  size_t off = 0;
  const size_t o_amax = blob_add(off, t.action_max), o_of = blob_add(off, t.out_fixed), o_ov = blob_add(off, t.out_var),
               o_if = blob_add(off, t.in_fixed), o_iv = blob_add(off, t.in_var), o_hr = blob_add(off, t.hold_rate),
               o_pr = blob_add(off, t.pen_rate), o_sw = blob_add(off, t.skw), o_le = blob_add(off, t.lead_exp),
               o_hm = blob_add(off, t.home), o_cl = blob_add(off, t.closest), o_rm = blob_add(off, t.region_map),
               o_pp = blob_add(off, t.prio), o_ps = blob_add(off, t.prio_static), o_om = blob_add(off, t.obs_mean),
               o_os = blob_add(off, t.obs_std);
  std::vector<unsigned char> host(off + 16, 0);
  auto put = [&](size_t at, const void* src, size_t n) { if (n) std::memcpy(host.data() + at, src, n); };
  put(o_amax, t.action_max.data(), t.action_max.size() * 8); put(o_of, t.out_fixed.data(), t.out_fixed.size() * 8);
  put(o_ov, t.out_var.data(), t.out_var.size() * 8); put(o_if, t.in_fixed.data(), t.in_fixed.size() * 8);
  put(o_iv, t.in_var.data(), t.in_var.size() * 8); put(o_hr, t.hold_rate.data(), t.hold_rate.size() * 8);
  put(o_pr, t.pen_rate.data(), t.pen_rate.size() * 8); put(o_sw, t.skw.data(), t.skw.size() * 8);
  put(o_le, t.lead_exp.data(), t.lead_exp.size() * 4); put(o_hm, t.home.data(), t.home.size() * 4);
  put(o_cl, t.closest.data(), t.closest.size() * 4); put(o_rm, t.region_map.data(), t.region_map.size() * 4);
  put(o_pp, t.prio.data(), t.prio.size()); put(o_ps, t.prio_static.data(), t.prio_static.size());
  put(o_om, t.obs_mean.data(), t.obs_mean.size() * 4); put(o_os, t.obs_std.data(), t.obs_std.size() * 4);
  ce = cudaMalloc(&env->d_blob, host.size());
  if (ce == cudaSuccess) ce = cudaMemcpy(env->d_blob, host.data(), host.size(), cudaMemcpyHostToDevice);
  if (ce != cudaSuccess) {
    if (env->d_blob) cudaFree(env->d_blob);
    delete env;
    return set_error(MARLSC_ECUDA, std::string("uploading tables: ") + cudaGetErrorString(ce));
  }
  unsigned char* b = static_cast<unsigned char*>(env->d_blob);
  auto D = [&](size_t at) { return reinterpret_cast<const double*>(b + at); };
  auto I = [&](size_t at) { return reinterpret_cast<const int32_t*>(b + at); };
  bind_tables(env->ds, D(o_amax), D(o_of), D(o_ov), D(o_if), D(o_iv), D(o_hr), D(o_pr), D(o_sw), I(o_le), I(o_hm),
              I(o_cl), t.region_map.empty() ? nullptr : I(o_rm), b + o_pp, b + o_ps,
              t.obs_mean.empty() ? nullptr : reinterpret_cast<const float*>(b + o_om),
              t.obs_std.empty() ? nullptr : reinterpret_cast<const float*>(b + o_os));
  *out = env;
  return MARLSC_OK;
}

void marlsc_env_destroy(marlsc_env_t* env) {
  if (!env) return;
  if (env->d_blob) cudaFree(env->d_blob);
  delete env;
}

int32_t marlsc_env_obs_dim(const marlsc_env_t* env) { return env ? env->ds.obs_dim : 0; }
int32_t marlsc_env_needs_history(const marlsc_env_t* env) { return env ? env->ds.need_hist : 0; }
int32_t marlsc_env_needs_forecast(const marlsc_env_t* env) { return env ? env->ds.need_fcst : 0; }
int32_t marlsc_env_team_size(const marlsc_env_t* env) { return env ? env->team : 0; }

int marlsc_env_set_team_size(marlsc_env_t* env, int32_t tpe) {
  if (!env) return set_error(MARLSC_EINVAL, "null handle");
  if (tpe == 0) {
    env->team = env->team_auto;
    return MARLSC_OK;
  }
  if (tpe < 1 || tpe > 256 || (tpe & (tpe - 1))) return set_error(MARLSC_EINVAL, "team size must be a power of two in [1,256]");
  env->team = tpe;
  return MARLSC_OK;
}

int marlsc_env_reset(marlsc_env_t* env, const marlsc_env_state_t* state, const int32_t* init_inventory, int32_t per_env,
                     float* obs, void* stream) {
  int rc = check_state(env, state);
  if (rc) return rc;
  if (!init_inventory || !obs) return set_error(MARLSC_EINVAL, "init_inventory and obs must not be NULL");
  MARLSC_CUDA(cudaSetDevice(env->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CALL(T) launch_reset<T>(env, *state, init_inventory, per_env, obs, s)
  MARLSC_DISPATCH_TEAM(env->team, CALL)
#undef CALL
}

int marlsc_env_step(marlsc_env_t* env, const marlsc_env_state_t* state, const marlsc_step_io_t* io, int32_t t, void* stream) {
  int rc = check_state(env, state);
  if (rc) return rc;
  if (!io) return set_error(MARLSC_EINVAL, "null io");
  if (!io->actions || !io->order_offsets || !io->rewards || !io->obs)
    return set_error(MARLSC_EINVAL, "io.actions, io.order_offsets, io.rewards and io.obs must not be NULL");
  if (io->order_qty_bytes != 1 && io->order_qty_bytes != 2) return set_error(MARLSC_EINVAL, "order_qty_bytes must be 1 or 2");
  if (env->ds.lead_mode == MARLSC_LEAD_STOCHASTIC && !io->actual_lead)
    return set_error(MARLSC_EINVAL, "io.actual_lead is required with a stochastic lead-time sampler");
  if (io->d_lost_sales && !io->d_unfulfilled) return set_error(MARLSC_EINVAL, "d_lost_sales needs d_unfulfilled");
  if (t < 0) return set_error(MARLSC_EINVAL, "timestep must be >= 0");
  MARLSC_CUDA(cudaSetDevice(env->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CALL(T) launch_step<T>(env, *state, *io, t, s)
  MARLSC_DISPATCH_TEAM(env->team, CALL)
#undef CALL
}

int marlsc_env_step_host(marlsc_env_t* env, const marlsc_env_state_t* state, const marlsc_step_io_t* dev,
                         const marlsc_host_step_t* host, int32_t t, void* stream) {
  int rc = check_state(env, state);
  if (rc) return rc;
  if (!dev || !host) return set_error(MARLSC_EINVAL, "null staging or host descriptor");
  if (!host->actions || !host->order_offsets || !host->rewards) return set_error(MARLSC_EINVAL, "host.actions / order_offsets / rewards are NULL");
  if (host->n_orders > 0 && (!host->order_region || !host->order_qty)) return set_error(MARLSC_EINVAL, "host order arrays are NULL");
  MARLSC_CUDA(cudaSetDevice(env->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t E = state->num_envs, WS = (int64_t)env->ds.W * env->ds.S;
  MARLSC_CUDA(cudaMemcpyAsync(const_cast<float*>(dev->actions), host->actions, sizeof(float) * E * WS, cudaMemcpyHostToDevice, s));
  MARLSC_CUDA(cudaMemcpyAsync(const_cast<int32_t*>(dev->order_offsets), host->order_offsets, sizeof(int32_t) * (E + 1), cudaMemcpyHostToDevice, s));
  if (host->n_orders > 0) {
    MARLSC_CUDA(cudaMemcpyAsync(const_cast<int16_t*>(dev->order_region), host->order_region, sizeof(int16_t) * host->n_orders, cudaMemcpyHostToDevice, s));
    MARLSC_CUDA(cudaMemcpyAsync(const_cast<void*>(dev->order_qty), host->order_qty,
                                (size_t)host->n_orders * env->ds.S * dev->order_qty_bytes, cudaMemcpyHostToDevice, s));
  }
  if (env->ds.lead_mode == MARLSC_LEAD_STOCHASTIC) {
    if (!host->actual_lead) return set_error(MARLSC_EINVAL, "host.actual_lead is required with a stochastic lead-time sampler");
    MARLSC_CUDA(cudaMemcpyAsync(const_cast<uint8_t*>(dev->actual_lead), host->actual_lead, (size_t)E * WS, cudaMemcpyHostToDevice, s));
  }
  rc = marlsc_env_step(env, state, dev, t, stream);
  if (rc) return rc;
  MARLSC_CUDA(cudaMemcpyAsync(host->rewards, dev->rewards, sizeof(float) * E * env->ds.W, cudaMemcpyDeviceToHost, s));
  if (host->obs) MARLSC_CUDA(cudaMemcpyAsync(host->obs, dev->obs, sizeof(float) * E * env->ds.W * env->ds.obs_dim, cudaMemcpyDeviceToHost, s));
  MARLSC_CUDA(cudaStreamSynchronize(s));
  return MARLSC_OK;
}

const char* marlsc_last_error(void) { return g_last_error.c_str(); }
int32_t marlsc_abi_version(void) { return MARLSC_ABI_VERSION; }
int64_t marlsc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
