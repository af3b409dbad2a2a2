// env_core.cuh - per-environment step logic of the fused env-step kernel (K1).
//
// One "team" of TPE threads advances one environment by one timestep. All per-step working data
// of the environment lives in the team's shared-memory scratch; global memory is touched once per
// array element (coalesced over SKUs). The same source compiles
//   * for sm_100a (env_step.cu), TPE in {1,2,4,...,256}, and
//   * as plain C++ with TPE == 1 (tests/emu/emu.cpp, -DMARLSC_HOST_EMU) so the step logic can be
//     checked against the oracle in a container without a GPU. The emulation is test-only.
//
// Reference semantics restated here (paths under the reference repo):
//   step order ................. src/environment/envs/multi_env.py:253-366
//   action rescale ............. multi_env.py:795-848
//   orders / arrivals .......... multi_env.py:850-919
//   greedy allocation .......... src/environment/components/demand_allocator.py:150-208
//   feature buffers ............ multi_env.py:747-793
//   lost sales ................. src/environment/components/lost_sales_handler.py:71-210
//   cost reward ................ src/environment/components/reward_calculator.py:127-188
//   observation ................ multi_env.py:577-745, 941-968
#pragma once
#include <stdint.h>
#include "../../include/marlsc_b200.h"

#ifdef MARLSC_HOST_EMU
#include <cmath>
#include <cstring>
#define MDEV static inline
#define MARLSC_RESTRICT
#else
#define MDEV __device__ __forceinline__
#define MARLSC_RESTRICT __restrict__
#endif

namespace marlsc {

constexpr int kWindow = MARLSC_ROLLING_WINDOW;

// Device-side view of a marlsc_env_spec_t: device table pointers, observation block offsets and
// the per-team scratch layout. Passed by value to the kernels.
struct DevSpec {
  int W, S, R, Rraw, L, D, episode_length;
  int action_type, lead_mode, lost_type, scope, max_splits, norm, id_off, obs_dim;
  uint32_t feat;
  int need_hist, need_fcst, need_ship;
  double scale, alpha;
  const double* action_max;
  const double* out_fixed;
  const double* out_var;
  const double* in_fixed;
  const double* in_var;
  const double* hold_rate;
  const double* pen_rate;
  const double* skw;
  const int32_t* lead_exp;
  const int32_t* home;
  const int32_t* closest;
  const int32_t* region_map;
  const uint8_t* prio;         // [R,W] warehouses in ascending (cost, index) order, valid where prio_static[r]
  const uint8_t* prio_static;  // [R] 1 when the order does not depend on the order's weight
  const float* obs_mean;
  const float* obs_std;
  // observation block offsets inside one warehouse's vector (before the id prefix); -1 = block disabled
  int off_inv, off_pipe, off_dh, off_sh, off_sa, off_so, off_rm, off_fc, off_dos, off_nip, off_dv, off_hist;
  // scratch layout: a double area and a 32-bit word area per team
  int och;                                   // orders staged per chunk
  int d_lostW, d_lostP, d_cout, d_ctot, d_words;            // offsets in doubles, total doubles
  int w_inv, w_q, w_dh, w_sh, w_st, w_rm, w_fc, w_shipq, w_lostN, w_rem, w_prio, w_sreg, w_sqty, w_words;
};

struct Scratch {
  double* d;
  int32_t* w;
};

// ------------------------------------------------------------------------------------------------
// Team: TPE threads working on one environment. Reductions run inside groups of G = min(TPE, 32)
// lanes (a whole warp or an aligned slice of one); teams wider than a warp split per-warehouse work
// over their NG groups and meet at a named barrier.
// ------------------------------------------------------------------------------------------------
template <int TPE>
struct Team {
  static constexpr int G = TPE < 32 ? TPE : 32;
  static constexpr int NG = TPE / G;
  int lane;  // 0..TPE-1 inside the team
  int gl;    // lane inside the group
  int gid;   // group inside the team
#ifndef MARLSC_HOST_EMU
  unsigned gmask;
  int bar_id;
  __device__ __forceinline__ void init(int team_in_block) {
    lane = threadIdx.x % TPE;
    gl = lane % G;
    gid = lane / G;
    const int wl = threadIdx.x & 31;
    gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(G - 1)));
    bar_id = 1 + team_in_block;
  }
  __device__ __forceinline__ void sync() const {
    if (TPE == 1) return;
    if (TPE <= 32) {
      __syncwarp(gmask);
    } else {
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(TPE) : "memory");
    }
  }
  __device__ __forceinline__ void gsync() const {
    if (G > 1) __syncwarp(gmask);
  }
  __device__ __forceinline__ bool g_any(bool p) const {
    if (G == 1) return p;
    return __ballot_sync(gmask, p) != 0u;
  }
  __device__ __forceinline__ int g_sum(int v) const {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
  }
  __device__ __forceinline__ float g_sum(float v) const {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
  }
  __device__ __forceinline__ double g_sum(double v) const {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
  }
#else
  void init(int) { lane = gl = gid = 0; }
  void sync() const {}
  void gsync() const {}
  bool g_any(bool p) const { return p; }
  int g_sum(int v) const { return v; }
  float g_sum(float v) const { return v; }
  double g_sum(double v) const { return v; }
#endif
};

// ---- arithmetic that must round exactly like NumPy (no fused multiply-add) -----------------------
#ifndef MARLSC_HOST_EMU
MDEV float f_add(float a, float b) { return __fadd_rn(a, b); }
MDEV float f_sub(float a, float b) { return __fsub_rn(a, b); }
MDEV float f_mul(float a, float b) { return __fmul_rn(a, b); }
MDEV float f_div(float a, float b) { return __fdiv_rn(a, b); }
MDEV float f_sqrt(float a) { return __fsqrt_rn(a); }
MDEV double d_add(double a, double b) { return __dadd_rn(a, b); }
MDEV double d_mul(double a, double b) { return __dmul_rn(a, b); }
MDEV double d_rint(double a) { return rint(a); }
#else
MDEV float f_add(float a, float b) { volatile float r = a + b; return r; }
MDEV float f_sub(float a, float b) { volatile float r = a - b; return r; }
MDEV float f_mul(float a, float b) { volatile float r = a * b; return r; }
MDEV float f_div(float a, float b) { volatile float r = a / b; return r; }
MDEV float f_sqrt(float a) { return std::sqrt(a); }
MDEV double d_add(double a, double b) { volatile double r = a + b; return r; }
MDEV double d_mul(double a, double b) { volatile double r = a * b; return r; }
MDEV double d_rint(double a) { return std::nearbyint(a); }
#endif
MDEV int imin(int a, int b) { return a < b ? a : b; }
MDEV int imax(int a, int b) { return a > b ? a : b; }
MDEV int pmod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

// Write one observation element (index j inside the un-prefixed local vector) with the fixed
// mean/std normalisation of multi_env.py:700-702 applied when enabled.
MDEV void emit(const DevSpec& sp, float* MARLSC_RESTRICT obs_w, int j, float x) {
  if (sp.norm == MARLSC_NORM_MEANSTD) x = f_div(f_sub(x, sp.obs_mean[j]), sp.obs_std[j]);
  obs_w[sp.id_off + j] = x;
}

// Per-environment global pointers.
struct EnvPtrs {
  int32_t* inv;        // [W,S]
  int32_t* ring_q;     // [D,W,S]
  uint8_t* ring_l;     // [D,W,S] or null
  int32_t* hist;       // [5,W,S] or null
  float* fcst;         // [W,S] or null
};

MDEV EnvPtrs env_ptrs(const DevSpec& sp, const marlsc_env_state_t& st, int64_t e) {
  const int64_t ws = (int64_t)sp.W * sp.S;
  EnvPtrs p;
  p.inv = st.inventory + e * ws;
  p.ring_q = st.ring_qty + e * ws * sp.D;
  p.ring_l = st.ring_lead ? st.ring_lead + e * ws * sp.D : nullptr;
  p.hist = st.demand_hist ? st.demand_hist + e * ws * kWindow : nullptr;
  p.fcst = st.forecast ? st.forecast + e * ws : nullptr;
  return p;
}

// Quantity in expected-arrival slot k (0-based) for cell i = w*S+s at the end of step t
// (multi_env.py:941-968). q_new / lead_new describe the order placed at step t itself.
MDEV int pipeline_value(const DevSpec& sp, const EnvPtrs& p, int t, int i, int k, int q_new, int lead_new) {
  const int WS = sp.W * sp.S;
  const int le = sp.lead_exp[i];
  if (sp.lead_mode == MARLSC_LEAD_FIXED) {
    // an order placed at tau sits in slot tau + le - t; it is in transit iff that slot is >= 1
    if (k + 1 > le) return 0;
    const int tau = t + k + 1 - le;
    if (tau < 0) return 0;
    if (tau == t) return q_new;
    return p.ring_q[(tau % sp.D) * WS + i];
  }
  int v = 0;
  for (int d = 0; d < sp.D; ++d) {
    const int tau = t - pmod(t - d, sp.D);  // youngest placement step <= t that maps to plane d
    if (tau < 0) continue;
    int q, lead;
    if (tau == t) {
      q = q_new;
      lead = lead_new;
    } else {
      q = p.ring_q[d * WS + i];
      lead = p.ring_l[d * WS + i];
    }
    if (q <= 0 || tau + lead <= t) continue;  // nothing placed, or already delivered
    const int slot = tau + le - t;
    if (slot > sp.L) continue;
    const int kk = slot <= 1 ? 0 : slot - 1;  // late orders pile into slot 0
    if (kk == k) v += q;
  }
  return v;
}

// Units on their way to cell i at the start of step t, before this step's arrivals are taken out
// (what the base_stock action space subtracts, multi_env.py:841-845).
MDEV int pending_before(const DevSpec& sp, const EnvPtrs& p, int t, int i) {
  const int WS = sp.W * sp.S;
  int v = 0;
  if (sp.lead_mode == MARLSC_LEAD_FIXED) {
    const int le = sp.lead_exp[i];
    for (int a = 1; a <= le; ++a) {
      const int tau = t - a;
      if (tau < 0) break;
      v += p.ring_q[(tau % sp.D) * WS + i];
    }
    return v;
  }
  for (int d = 0; d < sp.D; ++d) {
    const int tau = (t - 1) - pmod(t - 1 - d, sp.D);
    if (tau < 0) continue;
    const int q = p.ring_q[d * WS + i];
    if (q > 0 && tau + (int)p.ring_l[d * WS + i] >= t) v += q;
  }
  return v;
}

// Units arriving at cell i at step t (actual_arrival == t, multi_env.py:910-919).
MDEV int arrivals_now(const DevSpec& sp, const EnvPtrs& p, int t, int i) {
  const int WS = sp.W * sp.S;
  if (sp.lead_mode == MARLSC_LEAD_FIXED) {
    const int tau = t - sp.lead_exp[i];
    return tau < 0 ? 0 : p.ring_q[(tau % sp.D) * WS + i];
  }
  int v = 0;
  for (int d = 0; d < sp.D; ++d) {
    const int tau = (t - 1) - pmod(t - 1 - d, sp.D);
    if (tau < 0) continue;
    const int q = p.ring_q[d * WS + i];
    if (q > 0 && tau + (int)p.ring_l[d * WS + i] == t) v += q;
  }
  return v;
}

// Action in [-1,1] -> integer order quantity with the reference's mixed fp32/fp64 arithmetic and
// round-half-even (multi_env.py:824-846).
MDEV int rescale_action(const DevSpec& sp, float a, int s, int prev_home_demand, int pending) {
  const double mx = sp.action_max[s];
  if (sp.action_type == MARLSC_ACTION_DIRECT) {
    const float u = f_div(f_add(a, 1.0f), 2.0f);
    double q = d_rint(d_mul((double)u, mx));
    if (q < 0.0) q = 0.0;
    if (q > mx) q = mx;
    return (int)q;
  }
  if (sp.action_type == MARLSC_ACTION_DEMAND_CENTERED) {
    const long long adj = (long long)d_rint(d_mul(mx, (double)a));
    const long long q = adj + (long long)prev_home_demand;
    return q < 0 ? 0 : (int)q;
  }
  const float u = f_div(f_add(a, 1.0f), 2.0f);
  const double target = d_mul((double)u, mx);
  const double x = d_add(d_add(target, -(double)prev_home_demand), -(double)pending);
  const double q = d_rint(x);
  return q < 0.0 ? 0 : (int)q;
}

// ------------------------------------------------------------------------------------------------
// Observation writer for warehouse w by one group (multi_env.py:577-710). Feature values come from
// the team scratch; in-transit orders and older history planes come from global state.
// hist_n = number of valid history entries including the current step (0 right after reset).
// ------------------------------------------------------------------------------------------------
template <int TPE>
MDEV void write_obs_row(const DevSpec& sp, const Team<TPE>& tm, const Scratch& sc, const EnvPtrs& p,
                        float* MARLSC_RESTRICT obs_w, int w, int t, int hist_n,
                        const uint8_t* MARLSC_RESTRICT lead_new_row) {
  constexpr int G = Team<TPE>::G;
  const int S = sp.S, W = sp.W, L = sp.L, WS = sp.W * sp.S;
  const int base = w * S;
  const int32_t* inv = sc.w + sp.w_inv;
  const int32_t* sq = sc.w + sp.w_q;
  const int32_t* dh = sc.w + sp.w_dh;
  const int32_t* sh = sc.w + sp.w_sh;
  const int32_t* st = sc.w + sp.w_st;
  const float* rm = reinterpret_cast<const float*>(sc.w + sp.w_rm);
  const float* fc = reinterpret_cast<const float*>(sc.w + sp.w_fc);
  const bool ratio = sp.norm == MARLSC_NORM_RATIO;
  const uint32_t F = sp.feat;

  if (sp.id_off) {
    for (int j = tm.gl; j < W; j += G) obs_w[j] = (j == w) ? 1.0f : 0.0f;
  }

  // group totals (ratio denominators and aggregates)
  int sumI = 0, sumDh = 0, sumSh = 0, sumSt = 0;
  float sumRm = 0.f, sumFc = 0.f;
  for (int s = tm.gl; s < S; s += G) {
    const int i = base + s;
    sumI += inv[i];
    sumDh += dh[i];
    if (sp.need_ship) {
      sumSh += sh[i];
      sumSt += st[i];
    }
    if (sp.need_hist) sumRm += rm[i];
    if (sp.need_fcst) sumFc += fc[i];
  }
  sumI = tm.g_sum(sumI);
  sumDh = tm.g_sum(sumDh);
  sumSh = tm.g_sum(sumSh);
  sumSt = tm.g_sum(sumSt);
  sumRm = tm.g_sum(sumRm);
  sumFc = tm.g_sum(sumFc);
  const float dhDen = f_add((float)sumDh, 1e-8f);  // float32 total + eps (multi_env.py:614,737)

  // 1. inventory
  if (sp.off_inv >= 0) {
    for (int s = tm.gl; s < S; s += G) {
      const int I = inv[base + s];
      emit(sp, obs_w, sp.off_inv + s, ratio ? (float)((double)I / ((double)sumI + 1e-8)) : (float)I);
    }
    if ((F & MARLSC_F_INVENTORY_AGG) && tm.gl == 0) emit(sp, obs_w, sp.off_inv + S, (float)sumI);
  }

  // 2. pipeline, slot-major (L,S) ravel
  if (sp.off_pipe >= 0) {
    const bool need_total = ratio || (F & MARLSC_F_PIPELINE_AGG);
    float den = 1.0f;
    int total = 0;
    if (need_total) {
      for (int idx = tm.gl; idx < L * S; idx += G) {
        const int k = idx / S, s = idx - k * S;
        total += pipeline_value(sp, p, t, base + s, k, sq[base + s], lead_new_row ? lead_new_row[base + s] : 0);
      }
      total = tm.g_sum(total);
      den = (float)((double)total + 1e-8);
    }
    for (int idx = tm.gl; idx < L * S; idx += G) {
      const int k = idx / S, s = idx - k * S;
      const int v = pipeline_value(sp, p, t, base + s, k, sq[base + s], lead_new_row ? lead_new_row[base + s] : 0);
      emit(sp, obs_w, sp.off_pipe + idx, ratio ? f_div((float)v, den) : (float)v);
    }
    if ((F & MARLSC_F_PIPELINE_AGG) && tm.gl == 0) emit(sp, obs_w, sp.off_pipe + L * S, (float)total);
  }

  // 3. incoming home-region demand
  if (sp.off_dh >= 0) {
    for (int s = tm.gl; s < S; s += G) {
      const float v = (float)dh[base + s];
      emit(sp, obs_w, sp.off_dh + s, ratio ? f_div(v, dhDen) : v);
    }
    if ((F & MARLSC_F_DEMAND_HOME_AGG) && tm.gl == 0) emit(sp, obs_w, sp.off_dh + S, (float)sumDh);
  }
  // 4. units shipped to the home region (float64 in the reference)
  if (sp.off_sh >= 0) {
    for (int s = tm.gl; s < S; s += G) {
      const int v = sh[base + s];
      emit(sp, obs_w, sp.off_sh + s, ratio ? (float)((double)v / (double)dhDen) : (float)v);
    }
  }
  // 5. units shipped to other regions, aggregate = away share of everything shipped
  if (sp.off_sa >= 0) {
    const double den = (double)sumSt + 1e-8;
    for (int s = tm.gl; s < S; s += G) {
      const int v = st[base + s] - sh[base + s];
      emit(sp, obs_w, sp.off_sa + s, ratio ? (float)((double)v / den) : (float)v);
    }
    if ((F & MARLSC_F_SHIPPED_AWAY_AGG) && tm.gl == 0)
      emit(sp, obs_w, sp.off_sa + S, (float)((double)(sumSt - sumSh) / den));
  }
  // 6. stockout = max(home demand - shipped home, 0)
  if (sp.off_so >= 0) {
    for (int s = tm.gl; s < S; s += G) {
      const float v = (float)imax(dh[base + s] - sh[base + s], 0);
      emit(sp, obs_w, sp.off_so + s, ratio ? f_div(v, dhDen) : v);
    }
  }
  // 7. rolling mean of home demand
  if (sp.off_rm >= 0) {
    const float den = f_add(sumRm, 1e-8f);
    for (int s = tm.gl; s < S; s += G) emit(sp, obs_w, sp.off_rm + s, ratio ? f_div(rm[base + s], den) : rm[base + s]);
    if ((F & MARLSC_F_ROLLING_MEAN_AGG) && tm.gl == 0) emit(sp, obs_w, sp.off_rm + S, sumRm);
  }
  // 8. EMA forecast
  if (sp.off_fc >= 0) {
    const float den = f_add(sumFc, 1e-8f);
    for (int s = tm.gl; s < S; s += G) emit(sp, obs_w, sp.off_fc + s, ratio ? f_div(fc[base + s], den) : fc[base + s]);
    if ((F & MARLSC_F_FORECAST_AGG) && tm.gl == 0) emit(sp, obs_w, sp.off_fc + S, sumFc);
  }
  // 9. days of supply
  if (sp.off_dos >= 0) {
    for (int s = tm.gl; s < S; s += G) {
      const float r = rm[base + s];
      emit(sp, obs_w, sp.off_dos + s, (float)((double)inv[base + s] / (double)(r > 1.0f ? r : 1.0f)));
    }
  }
  // 10. net inventory position = on hand + in transit - forecast * expected lead
  if (sp.off_nip >= 0) {
    for (int s = tm.gl; s < S; s += G) {
      const int i = base + s;
      float pt = 0.f;
      for (int k = 0; k < L; ++k)
        pt += (float)pipeline_value(sp, p, t, i, k, sq[i], lead_new_row ? lead_new_row[i] : 0);
      const double v = ((double)inv[i] + (double)pt) - (double)fc[i] * (double)sp.lead_exp[i];
      emit(sp, obs_w, sp.off_nip + s, (float)v);
    }
  }
  // 11. demand variability: population std over the history window, float32 like np.std
  if (sp.off_dv >= 0) {
    for (int s = tm.gl; s < S; s += G) {
      const int i = base + s;
      float v = 0.f;
      if (hist_n > 1) {
        float h[kWindow];
        float sum = 0.f;
        for (int a = 0; a < hist_n; ++a) {  // oldest .. newest, the deque order of the reference
          const int back = hist_n - 1 - a;
          h[a] = back == 0 ? (float)dh[i] : (float)p.hist[pmod(t - back, kWindow) * WS + i];
          sum = f_add(sum, h[a]);
        }
        const float mean = f_div(sum, (float)hist_n);
        float acc = 0.f;
        for (int a = 0; a < hist_n; ++a) {
          const float dlt = f_sub(h[a], mean);
          acc = f_add(acc, f_mul(dlt, dlt));
        }
        v = f_sqrt(f_div(acc, (float)hist_n));
      }
      emit(sp, obs_w, sp.off_dv + s, v);
    }
  }
  // 12. demand history, most recent first, zero padded
  if (sp.off_hist >= 0) {
    for (int idx = tm.gl; idx < kWindow * S; idx += G) {
      const int a = idx / S, s = idx - a * S;
      const int i = base + s;
      float v = 0.f;
      if (a < hist_n) v = a == 0 ? (float)dh[i] : (float)p.hist[pmod(t - a, kWindow) * WS + i];
      emit(sp, obs_w, sp.off_hist + idx, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One environment, one step.
// ------------------------------------------------------------------------------------------------
template <int TPE>
MDEV void step_env(const DevSpec& sp, const Team<TPE>& tm, const Scratch& sc, const marlsc_env_state_t& st,
                   const marlsc_step_io_t& io, int64_t e, int t) {
  constexpr int G = Team<TPE>::G;
  constexpr int NG = Team<TPE>::NG;
  const int W = sp.W, S = sp.S, R = sp.R, WS = W * S;
  const EnvPtrs p = env_ptrs(sp, st, e);

  int32_t* s_inv = sc.w + sp.w_inv;
  int32_t* s_q = sc.w + sp.w_q;
  int32_t* s_dh = sc.w + sp.w_dh;
  int32_t* s_sh = sc.w + sp.w_sh;
  int32_t* s_st = sc.w + sp.w_st;
  float* s_rm = reinterpret_cast<float*>(sc.w + sp.w_rm);
  float* s_fc = reinterpret_cast<float*>(sc.w + sp.w_fc);
  int32_t* s_shipq = sc.w + sp.w_shipq;
  int32_t* s_lostN = sc.w + sp.w_lostN;
  int32_t* s_rem = sc.w + sp.w_rem;
  uint8_t* s_prio = reinterpret_cast<uint8_t*>(sc.w + sp.w_prio);
  int16_t* s_sreg = reinterpret_cast<int16_t*>(sc.w + sp.w_sreg);
  uint8_t* s_sqty = reinterpret_cast<uint8_t*>(sc.w + sp.w_sqty);
  double* s_lostW = sc.d + sp.d_lostW;
  double* s_lostP = sc.d + sp.d_lostP;
  double* s_cout = sc.d + sp.d_cout;
  double* s_ctot = sc.d + sp.d_ctot;

  const float* act = io.actions + e * WS;
  const uint8_t* lead_new = (sp.lead_mode == MARLSC_LEAD_STOCHASTIC) ? io.actual_lead + e * WS : nullptr;
  const int slot_new = t % sp.D;

  // ---- phase 1: orders in, arrivals in (multi_env.py:287-292) ------------------------------------
  for (int i = tm.lane; i < WS; i += TPE) {
    const int s = i % S;
    int prev_dem = 0, pend = 0;
    if (sp.action_type != MARLSC_ACTION_DIRECT) {
      if (t > 0) prev_dem = p.hist[pmod(t - 1, kWindow) * WS + i];
      if (sp.action_type == MARLSC_ACTION_BASE_STOCK) pend = pending_before(sp, p, t, i);
    }
    const int q = rescale_action(sp, act[i], s, prev_dem, pend);
    const int arr = arrivals_now(sp, p, t, i);   // reads the ring before the slot below is reused
    s_inv[i] = p.inv[i] + arr;
    s_q[i] = q;
    p.ring_q[slot_new * WS + i] = q;
    if (lead_new) p.ring_l[slot_new * WS + i] = lead_new[i];
    s_dh[i] = 0;
    if (sp.need_ship) {
      s_sh[i] = 0;
      s_st[i] = 0;
    }
    if (io.d_ordered) io.d_ordered[e * WS + i] = q;
  }
  for (int i = tm.lane; i < W * R; i += TPE) s_shipq[i] = 0;
  for (int i = tm.lane; i < R; i += TPE) {
    s_lostN[i] = 0;
    s_lostW[i] = 0.0;
    s_lostP[i] = 0.0;
  }
  for (int i = tm.lane; i < W; i += TPE) s_cout[i] = 0.0;
  tm.sync();

  // ---- phase 2: sequential greedy allocation of this step's orders (demand_allocator.py:150-208)
  const int o_begin = io.order_offsets[e];
  const int n_orders = io.order_offsets[e + 1] - o_begin;
  const int qb = io.order_qty_bytes;
  const int row_bytes = S * qb;
  const int och = qb == 1 ? sp.och : sp.och / 2;   // the staging area is sized for och one-byte rows
  for (int c0 = 0; c0 < n_orders; c0 += och) {
    const int cn = imin(och, n_orders - c0);
    // stage the chunk: regions (mapped onto included regions) and quantity rows, as aligned words
    for (int j = tm.lane; j < cn; j += TPE) {
      int r = io.order_region[o_begin + c0 + j];
      if (sp.region_map) r = sp.region_map[r];
      s_sreg[j] = (int16_t)r;
    }
    const uint8_t* src = reinterpret_cast<const uint8_t*>(io.order_qty) + (int64_t)(o_begin + c0) * row_bytes;
    const int shift = (int)(reinterpret_cast<uintptr_t>(src) & 3u);
    {
      const uint32_t* src_w = reinterpret_cast<const uint32_t*>(src - shift);
      uint32_t* dst_w = reinterpret_cast<uint32_t*>(s_sqty);
      const int nw = (shift + cn * row_bytes + 3) >> 2;
      for (int k = tm.lane; k < nw; k += TPE) dst_w[k] = src_w[k];
    }
    tm.sync();
    if (tm.gid == 0) {
      for (int j = 0; j < cn; ++j) {
        const int r = s_sreg[j];
        const uint8_t* row = s_sqty + shift + j * row_bytes;
        bool any_d = false;
        double wt = 0.0;
        for (int s = tm.gl; s < S; s += G) {
          const int d = qb == 1 ? (int)row[s] : (int)reinterpret_cast<const uint16_t*>(row)[s];
          s_rem[s] = d;
          if (d > 0) {
            any_d = true;
            wt += (double)d * sp.skw[s];
            for (int w = 0; w < W; ++w)
              if (sp.home[w] == r) s_dh[w * S + s] += d;       // multi_env.py:763-768
          }
        }
        if (!tm.g_any(any_d)) continue;                        // all-zero order: nothing can ship or be lost
        const uint8_t* prio;
        if (sp.prio_static[r]) {
          prio = sp.prio + r * W;
        } else {
          // warehouse order depends on the order's weight: key = fixed + variable * weight in float64,
          // stable ascending (demand_allocator.py:168-173; ties to the lowest index)
          const double wtot = tm.g_sum(wt);
          if (tm.gl == 0) {
            for (int w = 0; w < W; ++w) {
              const double key = d_add(sp.out_fixed[w * R + r], d_mul(sp.out_var[w * R + r], wtot));
              int pos = w;
              while (pos > 0) {
                const int pw = s_prio[pos - 1];
                const double pk = d_add(sp.out_fixed[pw * R + r], d_mul(sp.out_var[pw * R + r], wtot));
                if (pk <= key) break;
                s_prio[pos] = (uint8_t)pw;
                --pos;
              }
              s_prio[pos] = (uint8_t)w;
            }
          }
          tm.gsync();
          prio = s_prio;
        }
        int used = 0;
        for (int jj = 0; jj < W; ++jj) {
          if (used >= sp.max_splits + 1) break;
          const int w = prio[jj];
          const bool is_home = sp.home[w] == r;
          int fsum = 0;
          double wsum = 0.0;
          bool any_f = false, any_rem = false;
          for (int s = tm.gl; s < S; s += G) {
            const int rr = s_rem[s];
            if (rr > 0) {
              const int a = s_inv[w * S + s];
              const int f = imin(rr, a);
              if (f > 0) {
                s_inv[w * S + s] = a - f;
                s_rem[s] = rr - f;
                fsum += f;
                wsum += (double)f * sp.skw[s];
                any_f = true;
                if (sp.need_ship) {
                  s_st[w * S + s] += f;
                  if (is_home) s_sh[w * S + s] += f;
                }
                if (io.d_ship) io.d_ship[((e * W + w) * R + r) * S + s] += f;
              }
              if (rr - f > 0) any_rem = true;
            }
          }
          if (!tm.g_any(any_f)) continue;                      // this warehouse had nothing the order needs
          const int fs = tm.g_sum(fsum);
          const double ws = tm.g_sum(wsum);
          if (tm.gl == 0) {
            s_shipq[w * R + r] += fs;
            s_cout[w] += sp.out_fixed[w * R + r] + sp.out_var[w * R + r] * ws;
            if (io.d_ship_count) io.d_ship_count[(e * W + w) * R + r] += 1;
          }
          ++used;
          if (!tm.g_any(any_rem)) break;
        }
        // whatever is left is lost (demand_allocator.py:205-208)
        bool any_rem = false;
        double lw = 0.0, lp = 0.0;
        for (int s = tm.gl; s < S; s += G) {
          const int rr = s_rem[s];
          if (rr > 0) {
            any_rem = true;
            lw += (double)rr * sp.skw[s];
            lp += (double)rr * sp.pen_rate[s];
            if (io.d_unfulfilled) io.d_unfulfilled[(e * R + r) * S + s] += rr;
          }
        }
        if (tm.g_any(any_rem)) {
          lw = tm.g_sum(lw);
          lp = tm.g_sum(lp);
          if (tm.gl == 0) {
            s_lostN[r] += 1;
            s_lostW[r] += lw;
            s_lostP[r] += lp;
            if (io.d_lost_orders) io.d_lost_orders[e * R + r] += 1;
          }
        }
        tm.gsync();
      }
    }
    tm.sync();
  }

  // ---- phase 3: state write-back, feature buffers, costs (per warehouse, one group each) ----------
  const int hist_n = imin(t + 1, kWindow);
  for (int w = tm.gid; w < W; w += NG) {
    const int base = w * S;
    double hold = 0.0, inb = 0.0;
    for (int s = tm.gl; s < S; s += G) {
      const int i = base + s;
      const int I = s_inv[i];
      p.inv[i] = I;                                            // multi_env.py:307 (never negative)
      hold += (double)I * sp.hold_rate[s];
      const int q = s_q[i];
      if (q > 0) inb += sp.in_fixed[i] + ((double)q * sp.skw[s]) * sp.in_var[i];
      const int dnow = s_dh[i];
      if (sp.need_hist) {
        float sum = 0.f;
        for (int back = hist_n - 1; back >= 1; --back)
          sum = f_add(sum, (float)p.hist[pmod(t - back, kWindow) * WS + i]);
        sum = f_add(sum, (float)dnow);
        s_rm[i] = f_div(sum, (float)hist_n);                   // multi_env.py:785-787
        p.hist[(t % kWindow) * WS + i] = dnow;
      }
      if (sp.need_fcst) {
        const float f = f_add(f_mul(0.3f, (float)dnow), f_mul(0.7f, p.fcst[i]));   // multi_env.py:790-793
        p.fcst[i] = f;
        s_fc[i] = f;
      }
    }
    hold = tm.g_sum(hold);
    inb = tm.g_sum(inb);

    // penalty: lost volume of every region, split over warehouses by the configured handler
    double pen = 0.0;
    for (int r0 = 0; r0 < R; r0 += G) {
      const int r = r0 + tm.gl;
      if (r < R && s_lostN[r] > 0) {
        const double P = s_lostP[r];
        if (sp.lost_type == MARLSC_LOST_CLOSEST) {
          if (sp.closest[r] == w) pen += P;
        } else if (sp.lost_type == MARLSC_LOST_SHIPMENT) {
          int tot = 0;
          for (int ww = 0; ww < W; ++ww) tot += s_shipq[ww * R + r];
          if (tot > 0) pen += ((double)s_shipq[w * R + r] / (double)tot) * P;
          else if (sp.closest[r] == w) pen += P;
        } else {
          const double n = (double)s_lostN[r], lwt = s_lostW[r];
          double zmax = -1e300;
          for (int ww = 0; ww < W; ++ww) {
            const double z = -(sp.out_fixed[ww * R + r] * n + sp.out_var[ww * R + r] * lwt) / sp.alpha;
            if (z > zmax) zmax = z;
          }
          double zsum = 0.0, mine = 0.0;
          for (int ww = 0; ww < W; ++ww) {
            const double z = -(sp.out_fixed[ww * R + r] * n + sp.out_var[ww * R + r] * lwt) / sp.alpha;
            const double ez = exp(z - zmax);
            zsum += ez;
            if (ww == w) mine = ez;
          }
          pen += (mine / zsum) * P;
        }
      }
    }
    pen = tm.g_sum(pen);
    if (tm.gl == 0) {
      const double outc = s_cout[w];
      s_ctot[w] = hold + pen + outc + inb;
      if (io.cost_breakdown) {
        float* cb = io.cost_breakdown + (e * W + w) * 4;
        cb[0] = (float)hold;
        cb[1] = (float)pen;
        cb[2] = (float)outc;
        cb[3] = (float)inb;
      }
    }
  }
  tm.sync();

  // ---- phase 4: rewards, observations (multi_env.py:316-327) --------------------------------------
  for (int w = tm.lane; w < W; w += TPE) {
    double rew;
    if (sp.scope == MARLSC_SCOPE_TEAM) {
      rew = 0.0;
      for (int ww = 0; ww < W; ++ww) rew += -(s_ctot[ww] * sp.scale);
    } else {
      rew = -(s_ctot[w] * sp.scale);
    }
    io.rewards[e * W + w] = (float)rew;
  }
  if (io.truncated && tm.lane == 0) io.truncated[e] = (uint8_t)(t + 1 >= sp.episode_length);

  for (int w = tm.gid; w < W; w += NG)
    write_obs_row<TPE>(sp, tm, sc, p, io.obs + (e * W + w) * (int64_t)sp.obs_dim, w, t, hist_n, lead_new);

  // optional: lost sales per (warehouse, SKU) from the dumped unfulfilled matrix (diagnostics only)
  if (io.d_lost_sales && io.d_unfulfilled) {
    tm.sync();
    for (int i = tm.lane; i < WS; i += TPE) {
      const int w = i / S, s = i - w * S;
      double acc = 0.0;
      for (int r = 0; r < R; ++r) {
        const int u = io.d_unfulfilled[(e * R + r) * S + s];
        if (u == 0) continue;
        double wgt = 0.0;
        if (sp.lost_type == MARLSC_LOST_CLOSEST) {
          wgt = sp.closest[r] == w ? 1.0 : 0.0;
        } else if (sp.lost_type == MARLSC_LOST_SHIPMENT) {
          int tot = 0;
          for (int ww = 0; ww < W; ++ww) tot += s_shipq[ww * R + r];
          wgt = tot > 0 ? (double)s_shipq[w * R + r] / (double)tot : (sp.closest[r] == w ? 1.0 : 0.0);
        } else {
          const double n = (double)s_lostN[r], lwt = s_lostW[r];
          double zmax = -1e300;
          for (int ww = 0; ww < W; ++ww) {
            const double z = -(sp.out_fixed[ww * R + r] * n + sp.out_var[ww * R + r] * lwt) / sp.alpha;
            if (z > zmax) zmax = z;
          }
          double zsum = 0.0, mine = 0.0;
          for (int ww = 0; ww < W; ++ww) {
            const double z = -(sp.out_fixed[ww * R + r] * n + sp.out_var[ww * R + r] * lwt) / sp.alpha;
            const double ez = exp(z - zmax);
            zsum += ez;
            if (ww == w) mine = ez;
          }
          wgt = mine / zsum;
        }
        acc += wgt * (double)u;
      }
      io.d_lost_sales[e * WS + i] = (float)acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Reset of one environment (multi_env.py:233-246): clear pipeline and feature state, load the start
// inventory, emit the first observation (all demand-derived features are zero).
// ------------------------------------------------------------------------------------------------
template <int TPE>
MDEV void reset_env(const DevSpec& sp, const Team<TPE>& tm, const Scratch& sc, const marlsc_env_state_t& st,
                    const int32_t* MARLSC_RESTRICT init_inventory, int per_env, float* MARLSC_RESTRICT obs,
                    int64_t e) {
  constexpr int NG = Team<TPE>::NG;
  const int W = sp.W, S = sp.S, WS = W * S;
  const EnvPtrs p = env_ptrs(sp, st, e);
  const int32_t* init = init_inventory + (per_env ? e * WS : 0);
  int32_t* s_inv = sc.w + sp.w_inv;
  int32_t* s_q = sc.w + sp.w_q;
  int32_t* s_dh = sc.w + sp.w_dh;
  int32_t* s_sh = sc.w + sp.w_sh;
  int32_t* s_st = sc.w + sp.w_st;
  float* s_rm = reinterpret_cast<float*>(sc.w + sp.w_rm);
  float* s_fc = reinterpret_cast<float*>(sc.w + sp.w_fc);
  for (int i = tm.lane; i < WS; i += TPE) {
    const int v = init[i];
    p.inv[i] = v;
    s_inv[i] = v;
    s_q[i] = 0;
    s_dh[i] = 0;
    if (sp.need_ship) {
      s_sh[i] = 0;
      s_st[i] = 0;
    }
    if (sp.need_hist) s_rm[i] = 0.f;
    if (sp.need_fcst) {
      s_fc[i] = 0.f;
      p.fcst[i] = 0.f;
    }
    for (int d = 0; d < sp.D; ++d) {
      p.ring_q[d * WS + i] = 0;
      if (p.ring_l) p.ring_l[d * WS + i] = 0;
    }
    if (p.hist)
      for (int h = 0; h < kWindow; ++h) p.hist[h * WS + i] = 0;
  }
  tm.sync();
  for (int w = tm.gid; w < W; w += NG)
    write_obs_row<TPE>(sp, tm, sc, p, obs + (e * W + w) * (int64_t)sp.obs_dim, w, 0, 0, nullptr);
}

}  // namespace marlsc
