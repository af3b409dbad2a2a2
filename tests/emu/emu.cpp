// emu.cpp - TEST-ONLY host emulation of the K1 step logic.
//
// Compiles marl-sc_b200/csrc/env_core.cuh as plain C++ with one "thread" per environment
// (-DMARLSC_HOST_EMU, Team<1>, up to 128 SKUs per lane) so that the indexing / arithmetic of the CUDA kernel's per-env code can
// be checked against the oracle in a container that has no GPU. It is built by the tests into
// tests/emu/_build/ and is never loaded by the product package, bench.py or smoke().
#define MARLSC_HOST_EMU 1
#include <new>
#include <string>
#include <vector>

#include "../../marl-sc_b200/csrc/spec_build.h"

using namespace marlsc;

constexpr int kEmuSpl = 128;   // SKUs handled by the single emulated lane

struct EmuEnv {
  DevSpec ds;
  HostTables tb;
  Tables tabs;
  std::vector<double> sd;
  std::vector<int32_t> sw;
};

static std::string g_err;

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

int emu_env_create(const marlsc_env_spec_t* spec, void** out) {
  EmuEnv* e = new (std::nothrow) EmuEnv();
  if (!e) return MARLSC_ENOMEM;
  g_err = build_devspec(*spec, e->ds, e->tb);
  if (!g_err.empty()) {
    delete e;
    return MARLSC_EINVAL;
  }
  const HostTables& t = e->tb;
  bind_tables(e->ds, t.action_max.data(), t.out_fixed.data(), t.out_var.data(), t.in_fixed.data(), t.in_var.data(),
              t.hold_rate.data(), t.pen_rate.data(), t.skw.data(), t.lead_exp.data(), t.home.data(),
              t.closest.data(), t.region_map.empty() ? nullptr : t.region_map.data(), t.prio.data(),
              t.prio_static.data(), t.home_mask.empty() ? nullptr : t.home_mask.data(), t.lead_u8.data(),
              t.obs_mean.empty() ? nullptr : t.obs_mean.data(), t.obs_std.empty() ? nullptr : t.obs_std.data());
  if (e->ds.S > kEmuSpl) {
    g_err = "emulation supports at most 128 SKUs";
    delete e;
    return MARLSC_EUNSUPPORTED;
  }
  e->tabs = Tables{t.skw.data(), t.pen_rate.data(), t.hold_rate.data(), t.prio.data(), t.prio_static.data(),
                   t.home_mask.empty() ? nullptr : t.home_mask.data(), t.lead_u8.data()};
  e->sd.assign(e->ds.d_words, 0.0);
  e->sw.assign(e->ds.w_words, 0);
  *out = e;
  return MARLSC_OK;
}

void emu_env_destroy(void* h) { delete static_cast<EmuEnv*>(h); }
int emu_env_obs_dim(void* h) { return static_cast<EmuEnv*>(h)->ds.obs_dim; }
int emu_env_needs_history(void* h) { return static_cast<EmuEnv*>(h)->ds.need_hist; }
int emu_env_needs_forecast(void* h) { return static_cast<EmuEnv*>(h)->ds.need_fcst; }

int emu_env_reset(void* h, const marlsc_env_state_t* st, const int32_t* init, int per_env, float* obs) {
  EmuEnv* e = static_cast<EmuEnv*>(h);
  Team<1> tm;
  tm.init();
  for (int64_t i = 0; i < st->num_envs; ++i) reset_env<1, kEmuSpl, kCapsAll>(e->ds, e->tabs, tm, *st, init, per_env, obs, i);
  return MARLSC_OK;
}

int emu_env_step_lean(void* h, const marlsc_env_state_t* st, const marlsc_step_io_t* io, int t) {
  EmuEnv* e = static_cast<EmuEnv*>(h);
  Team<1> tm;
  tm.init();
  Scratch sc{e->sd.data(), e->sw.data()};
  for (int64_t i = 0; i < st->num_envs; ++i) step_env<1, kEmuSpl, kCapsLean>(e->ds, e->tabs, tm, sc, *st, *io, i, t);
  return MARLSC_OK;
}

int emu_env_step(void* h, const marlsc_env_state_t* st, const marlsc_step_io_t* io, int t) {
  EmuEnv* e = static_cast<EmuEnv*>(h);
  Team<1> tm;
  tm.init();
  Scratch sc{e->sd.data(), e->sw.data()};
  for (int64_t i = 0; i < st->num_envs; ++i) step_env<1, kEmuSpl, kCapsAll>(e->ds, e->tabs, tm, sc, *st, *io, i, t);
  return MARLSC_OK;
}

// The allocation as env_alloc.cuh runs it (one SKU column at a time: availability mask -> priority-order candidates
// through DevSpec::prio_perm -> one shipment per candidate), restated for the host on the library's own tables.
// inv [W,S] in/out, region [n], qty [n,S] one-byte rows, shipq [W,R] and lost_units [R] out (caller-zeroed).
int emu_alloc_avail(void* h, int32_t* inv, int n_orders, const int16_t* region, const uint8_t* qty, int32_t* shipq,
                    int32_t* lost_units) {
  EmuEnv* e = static_cast<EmuEnv*>(h);
  const DevSpec& ds = e->ds;
  const int W = ds.W, S = ds.S, R = ds.R, Wp = (W + 3) & ~3, nc = ds.perm_chunks;
  if (e->tb.prio_perm.empty()) {
    g_err = "no permutation table (W > 16)";
    return MARLSC_EUNSUPPORTED;
  }
  const uint16_t* perm = e->tb.prio_perm.data();
  const uint8_t* prio = e->tb.prio.data();
  for (int s = 0; s < S; ++s) {
    uint32_t am = 0u;
    for (int w = 0; w < W; ++w) am |= (inv[w * S + s] > 0 ? 1u : 0u) << w;
    for (int j = 0; j < n_orders; ++j) {
      int rem = qty[(size_t)j * S + s];
      if (rem == 0) continue;
      const int r = region[j];
      uint32_t cand = 0u;
      for (int c = 0; c < nc; ++c) cand |= perm[((size_t)r * nc + c) * 16 + ((am >> (4 * c)) & 15u)];
      while (rem > 0 && cand != 0u) {
        const int v = __builtin_ctz(cand);
        cand &= cand - 1;
        const int w = prio[r * Wp + v];
        const int a = inv[w * S + s], f = rem < a ? rem : a;
        inv[w * S + s] = a - f;
        shipq[w * R + r] += f;
        rem -= f;
        if (a == f) am &= ~(1u << w);
      }
      if (rem > 0) lost_units[r] += rem;
    }
  }
  return MARLSC_OK;
}
}
