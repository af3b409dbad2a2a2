// env_compact.cu - fused environment step, reset, base-stock policy and order -> line conversion over the compact
// state layout (see env_compact.cuh for the design notes).
//
// Reference semantics restated here (paths under the reference repo):
//   step order ................. src/environment/envs/multi_env.py:253-366
//   action rescale ............. multi_env.py:824-828 (direct)
//   orders / arrivals .......... multi_env.py:850-919
//   greedy allocation .......... src/environment/components/demand_allocator.py:150-208
//   home demand, rolling mean .. multi_env.py:747-793
//   lost sales ................. src/environment/components/lost_sales_handler.py:71-148 (closest, shipment)
//   cost reward ................ src/environment/components/reward_calculator.py:127-188
//   observation ................ multi_env.py:577-710, 941-968
#include "env_compact.cuh"

namespace marlsc {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kSlots = kCompactMaxS / 32;   // SKU slots per lane
constexpr int kPlaneBatch = 5;              // ring planes of a row in flight per lane (times kSlots cells)

__device__ __forceinline__ uint32_t sm_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t ld_s_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t ld_s_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void st_s_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_s_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_g_add(uint32_t* p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_nc_u16(const uint16_t* p) { uint32_t v; asm volatile("ld.global.nc.u16 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ld_cg_u16(const uint16_t* p) { uint32_t v; asm volatile("ld.global.cg.u16 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// observation element j of a warehouse's vector (after the id prefix) with the fixed mean/std normalisation of
// multi_env.py:700-702 when enabled
__device__ __forceinline__ void put(const DevSpec& sp, bool ms, float* __restrict__ out, unsigned j, float x) {
  if (ms) x = f_mul(f_sub(x, sp.obs_mean[j]), sp.obs_std[j]);
  out[j] = x;
}

// ---------------------------------------------------------------------------------------------------------------
// K1 (compact): one warp per environment, lane l owns SKUs l + 32 k.
// ---------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kCompactWarps * 32, 4)
env_step_compact_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ marlsc_env_state_t st,
                        const __grid_constant__ marlsc_step_io_t io, int t) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int W = sp.W, S = sp.S, R = sp.R, L = sp.L, WS = W * S;
  const CompactSmem lay = compact_smem(W, S, R, NCH, sp.pen_uniform);
  {  // per-CTA tables: availability -> priority-order permutation, priority rows, home warehouse of a region
    const int n_perm = (R * NCH * 32) >> 1, n_prio = (R * 16) >> 2, n_home = (R + 3) >> 2;
    for (int i = threadIdx.x; i < n_perm + n_prio + n_home; i += blockDim.x) {
      if (i < n_perm) reinterpret_cast<uint32_t*>(smem + lay.t_perm)[i] = reinterpret_cast<const uint32_t*>(sp.perm5)[i];
      else if (i < n_perm + n_prio) reinterpret_cast<uint32_t*>(smem + lay.t_prio)[i - n_perm] = reinterpret_cast<const uint32_t*>(sp.prio16)[i - n_perm];
      else reinterpret_cast<uint32_t*>(smem + lay.t_home)[i - n_perm - n_prio] = reinterpret_cast<const uint32_t*>(sp.home_wh)[i - n_perm - n_prio];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * kCompactWarps + wid;
  if (e >= st.num_envs) return;                       // whole warps leave together; everything below is warp-local

  // this environment's lines: read first, everything the allocation loads hangs on these two values
  const int64_t round0 = io.line_counts ? e * (int64_t)io.line_stride : (int64_t)io.line_offsets[e];
  const int n_rounds = io.line_counts ? io.line_counts[e] : io.line_offsets[e + 1] - (int)round0;
  const uint16_t* lp = io.lines + round0 * 32 + lane;
  if (lane == 0 && n_rounds > 0) prefetch_l2_bulk(io.lines + round0 * 32, (uint32_t)n_rounds * 64u);   // in flight during phase 1

  unsigned char* const wbase = smem + lay.t_bytes + (size_t)wid * lay.warp_bytes;
  uint16_t* const s_inv = reinterpret_cast<uint16_t*>(wbase + lay.inv);
  uint32_t* const s_shipq = reinterpret_cast<uint32_t*>(wbase + lay.shipq);
  uint32_t* const s_lostU = reinterpret_cast<uint32_t*>(wbase + lay.lostU);
  double* const s_lostP = reinterpret_cast<double*>(wbase + lay.lostP);
  const bool pen_uniform = sp.pen_uniform != 0;
  for (int i = lane; i < W * R; i += 32) s_shipq[i] = 0u;
  for (int i = lane; i < R; i += 32) {
    s_lostU[i] = 0u;
    if (!pen_uniform) s_lostP[i] = 0.0;
  }

  uint16_t* const g_inv = pinned(static_cast<uint16_t*>(st.inventory) + e * WS);
  uint8_t* const g_ring = pinned(static_cast<uint8_t*>(st.ring_qty) + e * (int64_t)WS * L);
  const bool need_hist = sp.need_hist != 0;
  uint16_t* const g_hist = need_hist ? pinned(static_cast<uint16_t*>(st.demand_hist) + e * (int64_t)kWindow * WS) : nullptr;
  float* const g_obs = pinned(io.obs + e * (int64_t)W * sp.obs_dim);
  const bool ms = sp.norm == MARLSC_NORM_MEANSTD;
  const int pa = t % L;                               // plane of the orders arriving now
  uint16_t* const hist_now = need_hist ? g_hist + (t % kWindow) * WS : nullptr;
  const bool by_row = sp.row_rates_uniform != 0;
  bool own[kSlots];
#pragma unroll
  for (int k = 0; k < kSlots; ++k) own[k] = lane + 32 * k < S;

  // ---- phase 1: per warehouse row - orders in, arrivals in, pipeline block of the observation -------------------
  uint32_t avlo = 0u, avhi = 0u;                      // which warehouses hold SKU slot k: bits 16 (k & 1) .. of (k & 2 ? avhi : avlo)
  int rowQ = 0, rowPos = 0;                           // lane w keeps row w's ordered units / ordered cells
  double rowInb = 0.0;                                // ... or its inbound cost when the rates vary over the row
  const float* const act = io.action_qty ? nullptr : pinned(io.actions + e * WS);
  const uint8_t* const aq = io.action_qty ? io.action_qty + e * WS : nullptr;
#pragma unroll 1
  for (int w = 0; w < W; ++w) {
    const int base = w * S;
    uint8_t* const ring_w = g_ring + (size_t)w * L * S + lane;     // this lane's first cell of every plane
    float* const out = g_obs + (size_t)w * sp.obs_dim + sp.id_off;
    float a_in[kSlots];
    int inv_in[kSlots], arr_in[kSlots], le[kSlots], q[kSlots];
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {                // every load of the row's cells first
      const int i = base + lane + 32 * k;
      a_in[k] = 0.f;
      inv_in[k] = arr_in[k] = 0;
      le[k] = 1;
      if (own[k]) {
        a_in[k] = aq ? (float)aq[i] : act[i];
        inv_in[k] = g_inv[i];
        arr_in[k] = ring_w[pa * S + 32 * k];
        le[k] = sp.lead_u8[i];
      }
    }
    // pipeline slots 0 .. L-2 are the planes after the arrival plane, unchanged by this step except for the order a
    // cell places now, which lands in slot lead-1 (the byte there is 0: the plane was cleared when it last arrived and
    // only this cell's order of exactly that lead writes it); slot L-1 is the arrival plane itself, which after this
    // step only holds the new orders of lead L.
    int plane = pa + 1;
    plane -= plane >= L ? L : 0;
    int nQ = 0, nPos = 0;
    double inb = 0.0;
#pragma unroll 1
    for (int k0 = 0; k0 < L; k0 += kPlaneBatch) {
      uint32_t v[kPlaneBatch][kSlots];
      int pl = plane;
#pragma unroll
      for (int kk = 0; kk < kPlaneBatch; ++kk) {
#pragma unroll
        for (int k = 0; k < kSlots; ++k) v[kk][k] = (own[k] && k0 + kk < L - 1) ? (uint32_t)ring_w[pl * S + 32 * k] : 0u;
        ++pl;
        pl -= pl >= L ? L : 0;
      }
      plane = pl;
      if (k0 == 0) {                                  // the row's own cells: order quantity, stock, ring, history plane
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
          q[k] = 0;
          if (own[k]) {
            const int i = base + lane + 32 * k;
            const double mx = sp.action_max[lane + 32 * k];
            q[k] = aq ? imin((int)a_in[k], (int)mx) : rescale_action<kCapsLean>(sp, a_in[k], mx, 0, 0);
            const int ni = inv_in[k] + arr_in[k];
            s_inv[i] = (uint16_t)ni;
            const uint32_t bit = (ni > 0 ? 1u : 0u) << (w + 16 * (k & 1));
            if (k & 2) avhi |= bit; else avlo |= bit;
            ring_w[pa * S + 32 * k] = (uint8_t)(le[k] == L ? q[k] : 0);          // arrivals consumed, lead-L orders in
            if (le[k] < L && q[k] > 0) {
              int p2 = pa + le[k];
              p2 -= p2 >= L ? L : 0;
              ring_w[p2 * S + 32 * k] = (uint8_t)q[k];
            }
            if (need_hist) hist_now[i] = 0;
            nQ += q[k];
            nPos += q[k] > 0 ? 1 : 0;
            if (!by_row && q[k] > 0) inb += sp.in_fixed[i] + ((double)q[k] * sp.skw[lane + 32 * k]) * sp.in_var[i];
          }
        }
      }
#pragma unroll
      for (int kk = 0; kk < kPlaneBatch; ++kk) {
        const int slot = k0 + kk;
        if (slot < L) {
#pragma unroll
          for (int k = 0; k < kSlots; ++k)
            if (own[k]) {
              const unsigned idx = (unsigned)(sp.off_pipe + slot * S + lane + 32 * k);
              put(sp, ms, out, idx, (float)(slot == le[k] - 1 ? (uint32_t)q[k] : v[kk][k]));
            }
        }
      }
    }
    if (by_row) {
      nQ = __reduce_add_sync(FULL, nQ);
      nPos = __reduce_add_sync(FULL, nPos);
      if (lane == w) {
        rowQ = nQ;
        rowPos = nPos;
      }
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) inb += __shfl_xor_sync(FULL, inb, o);
      if (lane == w) rowInb = inb;
    }
  }
  __syncwarp();                                       // stock staged, history plane cleared

  // ---- phase 2: greedy allocation of this step's lines (demand_allocator.py:150-208) -----------------------------
  // A trip of the loop: if the lane's current line is done, take the next entry of its stream (two entries are kept
  // requested ahead); then one shipment from the cheapest warehouse that holds the SKU, or the lost-sales bookkeeping
  // when none does. Lanes only meet in the exit vote.
  {
    const uint32_t a_perm = sm_addr(smem + lay.t_perm), a_prio = sm_addr(smem + lay.t_prio), a_home = sm_addr(smem + lay.t_home),
                   a_inv = sm_addr(s_inv) + 2u * lane, a_shipq = sm_addr(s_shipq), a_lostU = sm_addr(s_lostU);
    const uint16_t* const lend = lp + (int64_t)n_rounds * 32;
    uint32_t n0 = 0u, n1 = 0u;                        // the next two entries of this lane's stream (0 = none)
    if (lp < lend) n0 = ld_nc_u16(lp);
    if (lp + 32 < lend) n1 = ld_nc_u16(lp + 32);
    lp += 64;
    uint32_t rem = 0u, r = 0u, sl = 0u, cand = 0u;    // units left of the current line, its region, SKU slot, candidate bits
    const uint32_t S2 = 2u * S, R4 = 4u * R;
    while (true) {
      if (rem == 0u && n0 != 0u) {
        rem = n0 & 0xffu;
        r = (n0 >> 8) & 0x3fu;
        sl = n0 >> 14;
        n0 = n1;
        n1 = lp < lend ? ld_nc_u16(lp) : 0u;
        lp += 32;
        const uint32_t am = ((sl & 2u ? avhi : avlo) >> (16u * (sl & 1u))) & 0xffffu;
        const uint32_t pm = a_perm + r * (NCH * 64u);
        cand = ld_s_u16(pm + 2u * (am & 31u));
        if (NCH > 1) cand |= ld_s_u16(pm + 64u + 2u * ((am >> 5) & 31u));
        if (NCH > 2) cand |= ld_s_u16(pm + 128u + 2u * ((am >> 10) & 31u));
        if (NCH > 3) cand |= ld_s_u16(pm + 192u + 2u * ((am >> 15) & 31u));
        if (need_hist) {                              // home-region demand of this step (multi_env.py:763-768)
          const uint32_t hw = ld_s_u8(a_home + r);
          if (hw != 255u) {
            const uint32_t s = lane + 32u * sl;
            if (hw != 254u) {                         // two uint16 cells share a word: add into the right half, nobody waits
              const uint32_t c = hw * S + s;
              red_g_add(reinterpret_cast<uint32_t*>(hist_now) + (c >> 1), rem << (16u * (c & 1u)));
            } else {
              uint32_t hm = sp.home_mask[r];
              while (hm) {
                const uint32_t c = (uint32_t)lowest_bit(hm) * S + s;
                hm &= hm - 1;
                red_g_add(reinterpret_cast<uint32_t*>(hist_now) + (c >> 1), rem << (16u * (c & 1u)));
              }
            }
          }
        }
      }
      if (rem != 0u) {
        if (cand != 0u) {                             // ship from the cheapest warehouse that has the SKU
          const uint32_t v = (uint32_t)lowest_bit(cand);
          cand &= cand - 1;
          const uint32_t w = ld_s_u8(a_prio + r * 16u + v);
          const uint32_t cell = a_inv + w * S2 + 64u * sl;
          const uint32_t a = ld_s_u16(cell);          // the cells of a SKU belong to this lane
          const uint32_t f = rem < a ? rem : a;
          st_s_u16(cell, a - f);
          red_s_add(a_shipq + w * R4 + 4u * r, f);
          rem -= f;
          if (a == f) {                               // emptied
            const uint32_t clr = ~(1u << (w + 16u * (sl & 1u)));
            if (sl & 2u) avhi &= clr; else avlo &= clr;
          }
        }
        if (rem != 0u && cand == 0u) {
          // no warehouse can supply the rest: lost (demand_allocator.py:205-208); units are enough when every SKU
          // carries the same penalty rate
          red_s_add(a_lostU + 4u * r, rem);
          if (!pen_uniform) atomicAdd(&s_lostP[r], (double)rem * sp.pen_rate[lane + 32u * sl]);
          rem = 0u;
        }
      }
      if (!__any_sync(FULL, (rem | n0) != 0u)) break;
    }
  }
  __syncwarp();

  // Outbound cost and lost-sales penalty of every warehouse (reward_calculator.py:134-142). Lanes take regions (two per
  // lane cover R <= 64). Per region the lost volume's cost is either spread over the warehouses in proportion to what
  // they shipped there (shipment handler, lost_sales_handler.py:113-148: a rate per shipped unit, one division per
  // region) or goes to the closest warehouse (closest handler, and the shipment handler's fallback when nothing was
  // shipped). Then warehouse by warehouse the lanes' partial sums meet in a shuffle reduction; lane w keeps row w's.
  double costA = 0.0;
  {
    const double pen0 = sp.pen_rate[0];
    double rate[2], lump[2];
    int close_w[2];
#pragma unroll
    for (int qq = 0; qq < 2; ++qq) {
      const int r = lane + 32 * qq;
      rate[qq] = lump[qq] = 0.0;
      close_w[qq] = -1;
      if (r < R) {
        const uint32_t lu = s_lostU[r];
        if (lu > 0u) {
          const double lpn = pen_uniform ? (double)lu * pen0 : s_lostP[r];
          uint32_t shipped_r = 0u;
          if (sp.lost_type == MARLSC_LOST_SHIPMENT)
            for (int w = 0; w < W; ++w) shipped_r += s_shipq[w * R + r];
          if (shipped_r > 0u) {
            rate[qq] = lpn / (double)shipped_r;
          } else {
            lump[qq] = lpn;
            close_w[qq] = sp.closest[r];
          }
        }
      }
    }
    for (int w = 0; w < W; ++w) {
      double c = 0.0;
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int r = lane + 32 * qq;
        if (r < R) {
          const uint32_t sq = s_shipq[w * R + r];
          if (sq > 0u) c += (double)sq * (sp.out_var[w * R + r] + rate[qq]);
          if (close_w[qq] == w) c += lump[qq];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
      if (lane == w) costA = c;
    }
  }

  // ---- phase 3: per warehouse row - stock out, rolling mean, the remaining observation blocks, holding cost ------
  const int hist_n = imin(t + 1, kWindow);
  double costR = 0.0;
#pragma unroll 1
  for (int w = 0; w < W; ++w) {
    const int base = w * S;
    float* const obs_w = g_obs + (size_t)w * sp.obs_dim;
    float* const out = obs_w + sp.id_off;
    int vI[kSlots], vdh[kSlots], hv[kSlots][kWindow - 1];
#pragma unroll
    for (int k = 0; k < kSlots; ++k) {                // loads first
      const int i = base + lane + 32 * k;
      vI[k] = own[k] ? (int)s_inv[i] : 0;
      vdh[k] = own[k] && need_hist ? (int)ld_cg_u16(hist_now + i) : 0;   // accumulated with fire-and-forget adds above
#pragma unroll
      for (int back = 1; back < kWindow; ++back)
        hv[k][back - 1] = own[k] && need_hist && back < hist_n ? (int)g_hist[pmod(t - back, kWindow) * WS + i] : 0;
    }
    int nI = 0;
    double hold = 0.0;
#pragma unroll
    for (int k = 0; k < kSlots; ++k)
      if (own[k]) {
        const int s = lane + 32 * k, i = base + s;
        g_inv[i] = (uint16_t)vI[k];                   // multi_env.py:307 (never negative)
        nI += vI[k];
        if (!by_row) hold += (double)vI[k] * sp.hold_rate[s];
        put(sp, ms, out, (unsigned)(sp.off_inv + s), (float)vI[k]);
        if (sp.off_dh >= 0) put(sp, ms, out, (unsigned)(sp.off_dh + s), (float)vdh[k]);
        if (sp.off_rm >= 0) {
          // integer-valued float32 sum over the window is exact in any order (multi_env.py:785-787)
          const int hsum = (hv[k][0] + hv[k][1]) + (hv[k][2] + hv[k][3]) + vdh[k];
          put(sp, ms, out, (unsigned)(sp.off_rm + s), f_div((float)hsum, (float)hist_n));
        }
      }
    nI = __reduce_add_sync(FULL, nI);
    if (lane == 0 && (sp.feat & MARLSC_F_INVENTORY_AGG)) put(sp, ms, out, (unsigned)(sp.off_inv + S), (float)nI);
    if (sp.id_off && lane < W) obs_w[lane] = lane == w ? 1.0f : 0.0f;
    if (by_row) {
      if (lane == w)
        costR = (double)nI * sp.hold_rate[0] + ((double)rowPos * sp.in_fixed[base] + ((double)rowQ * sp.skw[0]) * sp.in_var[base]);
    } else {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hold += __shfl_xor_sync(FULL, hold, o);
      if (lane == w) costR = hold + rowInb;
    }
  }

  // ---- rewards (multi_env.py:316-327) ----------------------------------------------------------------------------
  double rew = lane < W ? -((costR + costA) * sp.scale) : 0.0;
  if (sp.scope == MARLSC_SCOPE_TEAM) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rew += __shfl_xor_sync(FULL, rew, o);
  }
  if (lane < W) io.rewards[e * W + lane] = (float)rew;
  if (io.truncated && lane == 0) io.truncated[e] = (uint8_t)(t + 1 >= sp.episode_length);
}

// ---------------------------------------------------------------------------------------------------------------
// Reset (multi_env.py:233-246): clear ring and history, load the start inventory, first observation. One thread per
// (environment, warehouse, SKU) cell; not a hot kernel.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
env_reset_compact_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ marlsc_env_state_t st,
                         const int32_t* __restrict__ init, int per_env, float* __restrict__ obs) {
  const int W = sp.W, S = sp.S, L = sp.L, WS = W * S;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= st.num_envs * (long long)WS) return;
  const long long e = idx / WS;
  const int i = (int)(idx - e * WS), w = i / S, s = i - w * S;
  const int v = init[per_env ? idx : i];
  static_cast<uint16_t*>(st.inventory)[idx] = (uint16_t)(v < 0 ? 0 : (v > 65535 ? 65535 : v));
  uint8_t* ring = static_cast<uint8_t*>(st.ring_qty) + (e * WS + (long long)w * S) * L + s;
  for (int d = 0; d < L; ++d) ring[d * S] = 0;
  if (st.demand_hist) {
    uint16_t* h = static_cast<uint16_t*>(st.demand_hist) + e * (long long)kWindow * WS + i;
    for (int b = 0; b < kWindow; ++b) h[b * WS] = 0;
  }
  const bool ms = sp.norm == MARLSC_NORM_MEANSTD;
  float* obs_w = obs + (e * W + w) * (long long)sp.obs_dim;
  float* out = obs_w + sp.id_off;
  put(sp, ms, out, (unsigned)(sp.off_inv + s), (float)v);
  for (int k = 0; k < L; ++k) put(sp, ms, out, (unsigned)(sp.off_pipe + k * S + s), 0.f);
  if (sp.off_dh >= 0) put(sp, ms, out, (unsigned)(sp.off_dh + s), 0.f);
  if (sp.off_rm >= 0) put(sp, ms, out, (unsigned)(sp.off_rm + s), 0.f);
  if (s == 0) {
    if (sp.feat & MARLSC_F_INVENTORY_AGG) {
      int tot = 0;
      for (int ss = 0; ss < S; ++ss) tot += init[(per_env ? e * WS : 0) + w * S + ss];
      put(sp, ms, out, (unsigned)(sp.off_inv + S), (float)tot);
    }
    for (int j = 0; j < sp.id_off; ++j) obs_w[j] = j == w ? 1.0f : 0.0f;
  }
}

// K5 (compact): units in transit to a cell are the sum of its ring column (every plane is an arrival still to come).
template <typename IDX>
__global__ void __launch_bounds__(256)
base_stock_compact_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ marlsc_env_state_t st,
                          const float* __restrict__ level, int level_per_env, float* __restrict__ actions) {
  const IDX WS = (IDX)(sp.W * sp.S);
  const IDX idx = (IDX)blockIdx.x * (IDX)blockDim.x + (IDX)threadIdx.x;
  if (idx >= (IDX)st.num_envs * WS) return;
  const IDX e = idx / WS;
  const int i = (int)(idx - e * WS), w = i / sp.S, s = i - w * sp.S;
  const int L = sp.L;
  const uint8_t* ring = static_cast<const uint8_t*>(st.ring_qty) + ((long long)e * (long long)WS + (long long)w * sp.S) * L + s;
  int pending = 0;
  for (int d0 = 0; d0 < L; d0 += 4) {
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = d0 + k < L ? (int)ring[(d0 + k) * sp.S] : 0;
    pending += (v[0] + v[1]) + (v[2] + v[3]);
  }
  const double mx = sp.action_max[s];
  double q = (double)(level_per_env ? level[idx] : level[i]) - (double)static_cast<const uint16_t*>(st.inventory)[idx] - (double)pending;
  q = q < 0.0 ? 0.0 : (q > mx ? mx : q);
  actions[idx] = (float)(2.0 * q / mx - 1.0);
}

// Dense order rows -> lines (padded layout): one warp per environment, lane l appends the non-zero cells of its SKUs
// order by order, so every stream keeps the order sequence the allocation needs.
__global__ void __launch_bounds__(128)
lines_from_orders_kernel(const __grid_constant__ DevSpec sp, long long E, const __grid_constant__ marlsc_step_io_t io,
                         int stride, uint16_t* __restrict__ lines, int32_t* __restrict__ counts, int32_t* __restrict__ overflow) {
  const long long e = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (e >= E) return;
  const int lane = threadIdx.x & 31, S = sp.S;
  const long long o_begin = io.order_counts ? e * (long long)io.order_stride : (long long)io.order_offsets[e];
  const int n_orders = io.order_counts ? io.order_counts[e] : io.order_offsets[e + 1] - (int)o_begin;
  const uint8_t* qty = static_cast<const uint8_t*>(io.order_qty) + o_begin * S + lane;
  uint16_t* out = lines + e * (long long)stride * 32 + lane;
  int cnt = 0;
  bool over = false;
  for (int j0 = 0; j0 < n_orders; j0 += 4) {          // four orders' cells in flight per lane
    uint32_t v[4][kSlots];
    int rg[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const bool live = j0 + jj < n_orders;
      rg[jj] = live ? (int)io.order_region[o_begin + j0 + jj] : 0;
#pragma unroll
      for (int k = 0; k < kSlots; ++k) v[jj][k] = (live && lane + 32 * k < S) ? (uint32_t)qty[(long long)(j0 + jj) * S + 32 * k] : 0u;
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int r = sp.region_map ? sp.region_map[rg[jj]] : rg[jj];
#pragma unroll
      for (int k = 0; k < kSlots; ++k)
        if (v[jj][k] != 0u) {
          if (cnt < stride) out[(long long)cnt * 32] = line_entry((int)v[jj][k], r, k);
          else over = true;
          ++cnt;
        }
    }
  }
  cnt = imin(cnt, stride);
  const int rounds = __reduce_max_sync(FULL, cnt);
  for (int c = cnt; c < rounds; ++c) out[(long long)c * 32] = 0;     // pad this stream to the environment's round count
  if (lane == 0) counts[e] = rounds;
  if (over) atomicExch(overflow, 1);
}

template <int NCH>
int launch_step_nch(const LaunchArgs& a, const marlsc_step_io_t& io, int t, cudaStream_t s) {
  const CompactSmem lay = compact_smem(a.ds.W, a.ds.S, a.ds.R, NCH, a.ds.pen_uniform);
  const size_t smem = (size_t)lay.t_bytes + (size_t)kCompactWarps * lay.warp_bytes;
  if ((int)smem > a.max_smem_optin)
    return set_error(MARLSC_EUNSUPPORTED, "compact step: shared-memory scratch of " + std::to_string(smem) + " bytes per CTA does not fit");
  static int configured_for[64] = {0};                // per device: the attributes belong to the device's context
  int dev = 0;
  MARLSC_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && configured_for[dev] < (int)smem) {
    MARLSC_CUDA(cudaFuncSetAttribute((const void*)env_step_compact_kernel<NCH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     (int)cudaSharedmemCarveoutMaxShared));
    MARLSC_CUDA(cudaFuncSetAttribute((const void*)env_step_compact_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured_for[dev] = (int)smem;
  }
  const unsigned grid = (unsigned)((a.st.num_envs + kCompactWarps - 1) / kCompactWarps);
  env_step_compact_kernel<NCH><<<grid, kCompactWarps * 32, smem, s>>>(a.ds, a.st, io, t);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

}  // namespace

int launch_step_compact(const LaunchArgs& a, const marlsc_step_io_t& io, int t, cudaStream_t s) {
  switch (a.ds.perm5_chunks) {
    case 1: return launch_step_nch<1>(a, io, t, s);
    case 2: return launch_step_nch<2>(a, io, t, s);
    case 3: return launch_step_nch<3>(a, io, t, s);
    case 4: return launch_step_nch<4>(a, io, t, s);
    default: return set_error(MARLSC_EUNSUPPORTED, "compact step needs W <= 16");
  }
}

int launch_reset_compact(const LaunchArgs& a, const int32_t* init, int per_env, float* obs, cudaStream_t s) {
  const long long n = a.st.num_envs * (long long)a.ds.W * a.ds.S;
  env_reset_compact_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a.ds, a.st, init, per_env, obs);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

int launch_base_stock_compact(const DevSpec& ds, const marlsc_env_state_t& st, const float* level, int level_per_env, int t,
                              float* actions, cudaStream_t s) {
  (void)t;
  const long long n = st.num_envs * (long long)ds.W * ds.S;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (n < (1LL << 32)) base_stock_compact_kernel<unsigned><<<grid, 256, 0, s>>>(ds, st, level, level_per_env, actions);
  else base_stock_compact_kernel<long long><<<grid, 256, 0, s>>>(ds, st, level, level_per_env, actions);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

int launch_lines_from_orders(const DevSpec& ds, int64_t num_envs, const marlsc_step_io_t& io, int32_t line_stride, uint16_t* lines,
                             int32_t* line_counts, int32_t* overflow, cudaStream_t s) {
  lines_from_orders_kernel<<<(unsigned)((num_envs + 3) / 4), 128, 0, s>>>(ds, num_envs, io, line_stride, lines, line_counts, overflow);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

}  // namespace marlsc
