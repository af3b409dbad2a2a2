// env_alloc.cuh - K1b of the split step for one-warp teams: the greedy allocation of one environment's orders
// (demand_allocator.py:118-217) as independent per-lane SKU chains over the WHOLE step's orders.
//
// The allocation of one SKU never looks at another SKU when no split limit binds, no per-shipment fixed cost
// applies and the warehouse priority of a region is static (the lean capability set), so the sequential semantics
// of demand_allocator.py:150-208 only order the lines of the same SKU. Lane l of the warp owns SKUs l + 32k:
//
//   mask pass  - the lane reads its cells of every order row straight from global memory (each row sector is
//                touched once) and notes the non-zero ones as bits, order-major, in a per-lane word list in
//                shared memory; orders of a home region add their quantities to this step's home-demand plane
//                (multi_env.py:763-768) with fire-and-forget reductions - a warp-uniform branch per order.
//   chain pass - "take my next line, ship it from the warehouses that HAVE the SKU, cheapest first, until it is
//                filled or lost". Which warehouses hold a SKU is kept as a bit mask per SKU (bit w); a few table
//                lookups turn it into the same bits in the region's priority order (DevSpec::prio_perm), so a
//                trip of the loop is one shipment - no visits to empty warehouses, which is most of them once
//                stock is scarce. The quantity of the line after the current one is requested one line ahead.
//                A pass covers 64 orders, so the lanes meet once per step and the busiest lane of a warp sets
//                its duration over ~40 lines instead of ~10 (the per-chunk version lost 45 % of its lanes).
//
// Results are identical to allocate_orders (env_core.cuh), which remains the path of every other team width and
// of the fused kernel.
#pragma once
#include <type_traits>

#include "env_kernels.cuh"

namespace marlsc {

template <int SPL>
struct AllocCfg {
  static constexpr int NA = SPL <= 1 ? 1 : SPL <= 2 ? 2 : SPL <= 4 ? 4 : SPL <= 8 ? 8 : 16;   // mask bits per order
  static constexpr int OPW = 32 / NA;      // orders per mask word
  static constexpr int kPass = 64;         // orders per pass
  static constexpr int MW = kPass / OPW;   // mask words per lane and pass
};

// Shared memory of the kernel, in bytes: [priority table | permutation table | home masks | team 0 | team 1 | ...]
struct AllocLayout {
  int t_prio, t_perm, t_hmask, t_bytes;                        // per CTA
  int lostP, inv, shipq, lostU, mask, reg, avail, team_bytes;  // per team, from the team's base
  int SP;                                                      // stock row stride
};
__host__ __device__ inline AllocLayout alloc_layout(int W, int S, int R, int nch, int MW, int kPass) {
  AllocLayout l;
  int o = 0;
  l.t_prio = o; o += (R * ((W + 3) & ~3) + 15) & ~15;
  l.t_perm = o; o += (R * nch * 16 * 2 + 15) & ~15;
  l.t_hmask = o; o += (R * 4 + 15) & ~15;
  l.t_bytes = o;
  l.SP = (S + 3) & ~3;
  o = 0;
  l.lostP = o; o += R * 8;
  l.inv = o; o += W * l.SP * 4;
  l.shipq = o; o += W * R * 4;
  l.lostU = o; o += R * 4;
  l.mask = o; o += MW * 32 * 4;
  l.reg = o; o += kPass * 2;
  l.avail = o; o += ((S + 31) & ~31) * 2;
  l.team_bytes = (o + 15) & ~15;
  return l;
}

// 1 where the byte of x is non-zero, gathered into bits 0..3
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t x) {
  const uint32_t t = (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;   // bit 7 of every non-zero byte
  return (t * 0x00204081u) >> 28;                                           // bits 7, 15, 23, 31 -> 28..31 (no carries)
}

template <int SPL, int NCH>
__global__ void __launch_bounds__(128, 6)
env_alloc_warp_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ marlsc_env_state_t st,
                      const __grid_constant__ marlsc_step_io_t io, double* __restrict__ cost_alloc, int t) {
  constexpr uint32_t CAPS = kCapsLean;
  using Cfg = AllocCfg<SPL>;
  constexpr int NA = Cfg::NA, OPW = Cfg::OPW, MW = Cfg::MW, kPass = Cfg::kPass;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem[];
  const int W = sp.W, S = sp.S, R = sp.R, WS = W * S, Wp = (W + 3) & ~3;
  const AllocLayout lay = alloc_layout(W, S, R, NCH, MW, kPass);
  {  // per-CTA tables: warehouse priority per region (rows padded to whole words), availability -> priority-order
     // permutation, home-warehouse masks
    const int n_prio = (R * Wp) >> 2, n_perm = (R * NCH * 16) >> 1;
    for (int i = threadIdx.x; i < n_prio + n_perm + R; i += blockDim.x) {
      if (i < n_prio) reinterpret_cast<uint32_t*>(smem + lay.t_prio)[i] = reinterpret_cast<const uint32_t*>(sp.prio)[i];
      else if (i < n_prio + n_perm)
        reinterpret_cast<uint32_t*>(smem + lay.t_perm)[i - n_prio] = reinterpret_cast<const uint32_t*>(sp.prio_perm)[i - n_prio];
      else reinterpret_cast<uint32_t*>(smem + lay.t_hmask)[i - n_prio - n_perm] = sp.home_mask ? sp.home_mask[i - n_prio - n_perm] : 0u;
    }
  }
  __syncthreads();
  const uint8_t* const t_prio = smem + lay.t_prio;
  const uint16_t* const t_perm = reinterpret_cast<const uint16_t*>(smem + lay.t_perm);
  const uint32_t* const t_hmask = reinterpret_cast<const uint32_t*>(smem + lay.t_hmask);
  const int lane = threadIdx.x & 31, team = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * 4 + team;
  if (e >= st.num_envs) return;                      // whole warps leave together
  unsigned char* const base = smem + lay.t_bytes + (size_t)team * lay.team_bytes;
  double* const s_lostP = reinterpret_cast<double*>(base + lay.lostP);
  int32_t* const s_inv = reinterpret_cast<int32_t*>(base + lay.inv);
  int32_t* const s_shipq = reinterpret_cast<int32_t*>(base + lay.shipq);
  int32_t* const s_lostU = reinterpret_cast<int32_t*>(base + lay.lostU);
  uint32_t* const s_mask = reinterpret_cast<uint32_t*>(base + lay.mask);
  int16_t* const s_reg = reinterpret_cast<int16_t*>(base + lay.reg);
  uint16_t* const s_avail = reinterpret_cast<uint16_t*>(base + lay.avail);
  const int SP = lay.SP;
  const EnvPtrs p = env_ptrs(sp, st, e);
  const bool pen_uniform = sp.pen_uniform != 0;

  // stock in: the environment's [W,S] block, row stride SP; bit w of a SKU's availability mask says warehouse w has it
  {
    uint32_t av[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) av[k] = 0u;
    for (int w0 = 0; w0 < W; w0 += 2) {
      int v[2][SPL];
#pragma unroll
      for (int d = 0; d < 2; ++d)
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
          const int s = lane + 32 * k;
          v[d][k] = (w0 + d < W && s < S) ? p.inv[(w0 + d) * S + s] : 0;
        }
#pragma unroll
      for (int d = 0; d < 2; ++d)
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
          const int s = lane + 32 * k;
          if (w0 + d < W && s < S) s_inv[(w0 + d) * SP + s] = v[d][k];
          av[k] |= (v[d][k] > 0 ? 1u : 0u) << (w0 + d);
        }
    }
#pragma unroll
    for (int k = 0; k < SPL; ++k) s_avail[lane + 32 * k] = (uint16_t)av[k];
  }
  for (int i = lane; i < W * R; i += 32) s_shipq[i] = 0;
  for (int i = lane; i < R; i += 32) {
    s_lostU[i] = 0;
    s_lostP[i] = 0.0;
  }

  // CSR (offsets) or padded layout (row e * stride, count[e]); the latter is what the device sampler writes
  const long long o_begin = io.order_counts ? e * (long long)io.order_stride : (long long)io.order_offsets[e];
  const int n_orders = io.order_counts ? io.order_counts[e] : io.order_offsets[e + 1] - (int)o_begin;
  const uint8_t* const qty = reinterpret_cast<const uint8_t*>(io.order_qty) + o_begin * S;
  const int16_t* const reg = io.order_region + o_begin;
  int32_t* const dh_acc = sp.dh_mode == 1 ? p.hist + (t % kWindow) * WS : nullptr;
  bool own[SPL];
#pragma unroll
  for (int k = 0; k < SPL; ++k) own[k] = lane + 32 * k < S;
  __syncwarp();

  for (int p0 = 0; p0 < n_orders; p0 += kPass) {
    const int pn = imin(kPass, n_orders - p0);
    const int nw = (pn + OPW - 1) / OPW;
    // ---- mask pass ---------------------------------------------------------------------------------
#pragma unroll
    for (int h = 0; h < kPass / 32; ++h)
      if (lane + 32 * h < pn) s_reg[lane + 32 * h] = reg[p0 + lane + 32 * h];
    __syncwarp();
    const uint8_t* const rows = qty + (long long)p0 * S + lane;     // this lane's column of the pass
    auto mask_word = [&](int wi, auto tail) {
      constexpr bool kTail = decltype(tail)::value;
      const uint8_t* rp = rows + (unsigned)(wi * OPW) * (unsigned)S;
      uint32_t b[OPW][SPL];
#pragma unroll
      for (int oo = 0; oo < OPW; ++oo) {              // every load of the word's orders first
#pragma unroll
        for (int k = 0; k < SPL; ++k) b[oo][k] = (own[k] && (!kTail || wi * OPW + oo < pn)) ? (uint32_t)rp[32 * k] : 0u;
        rp += S;
      }
      uint32_t word = 0u;
#pragma unroll
      for (int oo = 0; oo < OPW; ++oo) {
        const int j = wi * OPW + oo;
        if constexpr (SPL == 4) {
          word |= nonzero_nibble(b[oo][0] | (b[oo][1] << 8) | (b[oo][2] << 16) | (b[oo][3] << 24)) << (oo * NA);
        } else {
#pragma unroll
          for (int k = 0; k < SPL; ++k) word |= (b[oo][k] != 0u ? 1u : 0u) << (oo * NA + k);
        }
        if (dh_acc && (!kTail || j < pn)) {           // uniform: the region belongs to the order
          uint32_t hm = t_hmask[s_reg[j]];
          while (hm) {
            const int w = lowest_bit(hm);
            hm &= hm - 1;
#pragma unroll
            for (int k = 0; k < SPL; ++k)
              if (b[oo][k] != 0u) global_add(&dh_acc[w * S + lane + 32 * k], (int)b[oo][k]);   // nobody waits for the sum
          }
        }
      }
      s_mask[wi * 32 + lane] = word;                  // read back by this lane only
    };
    const int nfull = pn / OPW;
#pragma unroll 1
    for (int wi = 0; wi < nfull; ++wi) mask_word(wi, std::false_type());
    if (nfull < nw) mask_word(nfull, std::true_type());
    // ---- chain pass --------------------------------------------------------------------------------
    int wi = 0;
    uint32_t cur = s_mask[lane];
    int rem = 0, r = 0, s = 0;
    uint32_t am = 0u, cand = 0u;                      // warehouses holding s (bit w); the same in r's priority order
    int n_q = 0;
    unsigned n_k = 0, n_oj = 0;                       // the line after the current one: SKU slot, order
    bool n_ok = false;
    while (true) {
      if (rem == 0 && n_ok) {                         // take the requested line
        rem = n_q;
        s = lane + 32 * (int)n_k;
        r = s_reg[n_oj];
        n_ok = false;
        am = s_avail[s];
        const uint16_t* pm = t_perm + r * (NCH * 16);
        cand = pm[am & 15u];
        if (NCH > 1) cand |= pm[16 + ((am >> 4) & 15u)];
        if (NCH > 2) cand |= pm[32 + ((am >> 8) & 15u)];
        if (NCH > 3) cand |= pm[48 + ((am >> 12) & 15u)];
      }
      if (!n_ok) {                                    // request the one after it
        if (cur == 0u && wi + 1 < nw) cur = s_mask[(++wi) * 32 + lane];
        if (cur != 0u) {
          const unsigned bit = (unsigned)lowest_bit(cur);
          cur &= cur - 1;
          n_oj = (unsigned)wi * OPW + bit / NA;
          n_k = bit % NA;
          n_q = rows[n_oj * (unsigned)S + 32u * n_k];
          n_ok = true;
        }
      }
      if (rem > 0) {
        if (cand != 0u) {                             // ship from the cheapest warehouse that has the SKU
          const int v = lowest_bit(cand);
          cand &= cand - 1;
          const int w = t_prio[r * Wp + v];
          const int cell = w * SP + s;
          const int a = s_inv[cell];                  // the cells of a line's SKU belong to this lane
          const int f = imin(rem, a);
          s_inv[cell] = a - f;
          atomicAdd(&s_shipq[w * R + r], f);
          rem -= f;
          if (a == f) {                               // emptied
            am &= ~(1u << w);
            s_avail[s] = (uint16_t)am;
          }
        }
        if (rem > 0 && cand == 0u) {
          // no warehouse can supply the rest: lost (demand_allocator.py:205-208); units are enough when every
          // SKU carries the same penalty rate
          atomicAdd(&s_lostU[r], rem);
          if (!pen_uniform) atomicAdd(&s_lostP[r], (double)rem * sp.pen_rate[s]);
          rem = 0;
        }
      }
      if (!__any_sync(FULL, rem > 0 || n_ok || cur != 0u || wi + 1 < nw)) break;
    }
    __syncwarp();
  }

  // stock out (multi_env.py:307; never negative)
  for (int w = 0; w < W; ++w)
#pragma unroll
    for (int k = 0; k < SPL; ++k) {
      const int s = lane + 32 * k;
      if (s < S) p.inv[w * S + s] = s_inv[w * SP + s];
    }
  __syncwarp();
  // Outbound cost and lost-sales penalty of every warehouse (reward_calculator.py:150-175). Lanes take regions;
  // a lane's partial sums per warehouse go to a [W][33] tile laid over the stock scratch (written back above),
  // then lane w adds up row w.
  const double pen0 = sp.pen_rate[0];
  const bool tiled = W * 33 * 8 <= W * SP * 4 && W <= 32;
  if (tiled) {
    double* const tile = reinterpret_cast<double*>(s_inv);
    for (int w = 0; w < W; ++w) tile[w * 33 + lane] = 0.0;
    for (int r = lane; r < R; r += 32) {
      const int lu = s_lostU[r];
      const bool lost = lu > 0;
      const double lp = !lost ? 0.0 : (pen_uniform ? (double)lu * pen0 : s_lostP[r]);
      int shipped_r = 0;
      if (lost)
        for (int w = 0; w < W; ++w) shipped_r += s_shipq[w * R + r];
      for (int w = 0; w < W; ++w) {
        const int sq = s_shipq[w * R + r];
        double c = 0.0;
        if (sq > 0) c = (double)sq * sp.out_var[w * R + r];
        if (lost) c += lost_weight<CAPS>(sp, s_shipq, s_lostU, nullptr, w, r, shipped_r) * lp;
        if (sq > 0 || lost) tile[w * 33 + lane] += c;
      }
    }
    __syncwarp();
    if (lane < W) {
      double c = 0.0;
      for (int l = 0; l < 32; ++l) c += tile[lane * 33 + l];
      cost_alloc[e * W + lane] = c;
    }
  } else {
    for (int w = 0; w < W; ++w) {
      double c = 0.0;
      for (int r = lane; r < R; r += 32) {
        const int sq = s_shipq[w * R + r];
        if (sq > 0) c += (double)sq * sp.out_var[w * R + r];
        const int lu = s_lostU[r];
        if (lu > 0) c += lost_weight<CAPS>(sp, s_shipq, s_lostU, nullptr, w, r) * (pen_uniform ? (double)lu * pen0 : s_lostP[r]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
      if (lane == 0) cost_alloc[e * W + w] = c;
    }
  }
}

}  // namespace marlsc
