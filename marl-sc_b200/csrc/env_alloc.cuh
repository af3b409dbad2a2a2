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
__host__ __device__ inline AllocLayout alloc_layout(int W, int S, int R, int nch, int MW, int kPass, int spl) {
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
  l.avail = o; o += 32 * spl * 2;                        // one entry per (lane, SKU slot), also the slots beyond S
  l.team_bytes = (o + 15) & ~15;
  return l;
}

// 1 where the byte of x is non-zero, gathered into bits 0..3
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t x) {
  const uint32_t t = (x | ((x & 0x7f7f7f7fu) + 0x7f7f7f7fu)) & 0x80808080u;   // bit 7 of every non-zero byte
  return (t * 0x00204081u) >> 28;                                           // bits 7, 15, 23, 31 -> 28..31 (no carries)
}

// Shared-memory accesses of the chain loop by 32-bit shared-window address: the bases stay in registers and an
// access is one instruction plus its offset arithmetic (generic pointers were re-derived on every trip).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_s16(uint32_t a) { int v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void red_shared_add(uint32_t a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ldg_nc_u8(const uint8_t* p) { uint32_t v; asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

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
  const AllocLayout lay = alloc_layout(W, S, R, NCH, MW, kPass, SPL);
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
  // CSR (offsets) or padded layout (row e * stride, count[e]); the latter is what the device sampler writes. Read
  // first: everything the mask pass loads hangs on these two values.
  const long long o_begin = io.order_counts ? e * (long long)io.order_stride : (long long)io.order_offsets[e];
  const int n_orders = io.order_counts ? io.order_counts[e] : io.order_offsets[e + 1] - (int)o_begin;

  // stock in: the environment's [W,S] block, row stride SP; bit w of a SKU's availability mask says warehouse w has it
  {
    constexpr int kRows = SPL <= 4 ? 5 : 2;
    uint32_t av[SPL];
#pragma unroll
    for (int k = 0; k < SPL; ++k) av[k] = 0u;
    for (int w0 = 0; w0 < W; w0 += kRows) {              // kRows x SPL loads in flight per lane
      int v[kRows][SPL];
#pragma unroll
      for (int d = 0; d < kRows; ++d)
#pragma unroll
        for (int k = 0; k < SPL; ++k) {
          const int s = lane + 32 * k;
          v[d][k] = (w0 + d < W && s < S) ? p.inv[(w0 + d) * S + s] : 0;
        }
#pragma unroll
      for (int d = 0; d < kRows; ++d)
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
    static_assert(kPass == 64, "home_orders is one 64-bit mask");
    unsigned long long home_orders = 0ull;            // bit j: order j goes to some warehouse's home region
#pragma unroll
    for (int h = 0; h < kPass / 32; ++h) {
      int rg = -1;
      if (lane + 32 * h < pn) {
        rg = reg[p0 + lane + 32 * h];
        s_reg[lane + 32 * h] = (int16_t)rg;
      }
      home_orders |= (unsigned long long)__ballot_sync(FULL, dh_acc != nullptr && rg >= 0 && t_hmask[rg] != 0u) << (32 * h);
    }
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
        if ((home_orders >> j) & 1ull) {              // uniform: the region belongs to the order
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
    // Four SKUs per lane and rows that are whole aligned words: lane q reads WORD q of a row (SKUs 4q..4q+3, one
    // coalesced load per order instead of four byte loads), packs the non-zero nibbles of eight orders into one
    // register, and every lane fetches the four registers that hold its own SKUs (l + 32k lives in word l/4 + 8k,
    // byte l%4) with shuffles. The home-demand adds are done by the lane that holds the bytes.
    auto mask_word_w = [&](int wi, auto tail) {
      constexpr bool kTail = decltype(tail)::value;
      const uint32_t* rp = reinterpret_cast<const uint32_t*>(qty + (long long)p0 * S + (unsigned)(wi * OPW) * (unsigned)S) + lane;
      const bool has_word = 4 * lane < S;
      uint32_t x[OPW];
#pragma unroll
      for (int oo = 0; oo < OPW; ++oo) {
        x[oo] = (has_word && (!kTail || wi * OPW + oo < pn)) ? *rp : 0u;
        rp += S >> 2;
      }
      uint32_t packed = 0u;
      const uint32_t home_w = (uint32_t)(home_orders >> (wi * OPW)) & (OPW >= 32 ? 0xffffffffu : ((1u << (OPW & 31)) - 1u));   // this word's orders
#pragma unroll
      for (int oo = 0; oo < OPW; ++oo) {
        const int j = wi * OPW + oo;
        packed |= nonzero_nibble(x[oo]) << (oo * 4);
        if (home_w & (1u << oo)) {                    // uniform: the region belongs to the order
          uint32_t hm = t_hmask[s_reg[j]];
          while (hm) {
            const int w = lowest_bit(hm);
            hm &= hm - 1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int v = (int)((x[oo] >> (8 * q)) & 0xffu);
              if (v != 0) global_add(&dh_acc[w * S + 4 * lane + q], v);   // nobody waits for the sum
            }
          }
        }
      }
      uint32_t word = 0u;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t pk = __shfl_sync(FULL, packed, (lane >> 2) + 8 * k);
        word |= ((pk >> (lane & 3)) & 0x11111111u) << k;
      }
      s_mask[wi * 32 + lane] = word;
    };
    const int nfull = pn / OPW;
    bool by_words = false;
    if constexpr (SPL == 4 && OPW == 8) by_words = (S & 3) == 0 && (reinterpret_cast<uintptr_t>(io.order_qty) & 3u) == 0;
    if (by_words) {
      if constexpr (SPL == 4 && OPW == 8) {
#pragma unroll 1
        for (int wi = 0; wi < nfull; ++wi) mask_word_w(wi, std::false_type());
        if (nfull < nw) mask_word_w(nfull, std::true_type());
      }
    } else {
#pragma unroll 1
      for (int wi = 0; wi < nfull; ++wi) mask_word(wi, std::false_type());
      if (nfull < nw) mask_word(nfull, std::true_type());
    }
    // ---- chain pass --------------------------------------------------------------------------------
    const uint8_t* rows_l = rows;
    uint32_t a_mask = smem_addr(s_mask) + 4u * lane, a_reg = smem_addr(s_reg), a_avail = smem_addr(s_avail) + 2u * lane,
             a_perm = smem_addr(t_perm), a_prio = smem_addr(t_prio), a_inv = smem_addr(s_inv) + 4u * lane,
             a_shipq = smem_addr(s_shipq), a_lostU = smem_addr(s_lostU);
    uint32_t Wp_l = Wp, SP4 = 4u * SP, S_l = S, R4 = 4u * R;
    asm volatile("" : "+l"(rows_l), "+r"(a_mask), "+r"(a_reg), "+r"(a_avail), "+r"(a_perm), "+r"(a_prio), "+r"(a_inv),
                 "+r"(a_shipq), "+r"(a_lostU), "+r"(Wp_l), "+r"(SP4), "+r"(S_l), "+r"(R4));   // keep them in registers
    const uint32_t a_mask_end = a_mask + 128u * (nw - 1);
    uint32_t cur = lds_u32(a_mask);
    uint32_t oj_base = 0;                             // first order of the lane's current mask word
    int rem = 0;
    uint32_t r = 0, k32 = 0;                          // region; 32 * SKU slot of the current line
    uint32_t am = 0u, cand = 0u;                      // warehouses holding the SKU (bit w); the same in r's priority order
    uint32_t n_q = 0, n_k32 = 0, n_oj = 0;            // the line after the current one: quantity, 32 * SKU slot, order
    bool n_ok = false;
    while (true) {
      if (rem == 0) {
        if (n_ok) {                                   // take the requested line
          rem = (int)n_q;
          k32 = n_k32;
          r = (uint32_t)lds_s16(a_reg + 2u * n_oj);
          am = lds_u16(a_avail + 2u * k32);
          const uint32_t pm = a_perm + r * (NCH * 32u);
          cand = lds_u16(pm + 2u * (am & 15u));
          if (NCH > 1) cand |= lds_u16(pm + 32u + 2u * ((am >> 4) & 15u));
          if (NCH > 2) cand |= lds_u16(pm + 64u + 2u * ((am >> 8) & 15u));
          if (NCH > 3) cand |= lds_u16(pm + 96u + 2u * ((am >> 12) & 15u));
        }
        // find the lane's next line and ask for its quantity (an all-zero word costs the lane one idle trip)
        if (cur == 0u && a_mask != a_mask_end) {
          cur = lds_u32(a_mask += 128u);
          oj_base += OPW;
        }
        n_ok = cur != 0u;
        if (n_ok) {
          const uint32_t bit = (uint32_t)lowest_bit(cur);
          cur &= cur - 1;
          n_oj = oj_base + bit / NA;
          n_k32 = 32u * (bit % NA);
          n_q = ldg_nc_u8(rows_l + (n_oj * S_l + n_k32));
        }
      }
      if (rem > 0) {
        if (cand != 0u) {                             // ship from the cheapest warehouse that has the SKU
          const uint32_t v = (uint32_t)lowest_bit(cand);
          cand &= cand - 1;
          const uint32_t w = lds_u8(a_prio + r * Wp_l + v);
          const uint32_t cell = a_inv + w * SP4 + 4u * k32;
          const int a = (int)lds_u32(cell);           // the cells of a line's SKU belong to this lane
          const int f = imin(rem, a);
          sts_u32(cell, (uint32_t)(a - f));
          red_shared_add(a_shipq + w * R4 + 4u * r, f);
          rem -= f;
          if (a == f) {                               // emptied
            am &= ~(1u << w);
            sts_u16(a_avail + 2u * k32, am);
          }
        }
        if (rem > 0 && cand == 0u) {
          // no warehouse can supply the rest: lost (demand_allocator.py:205-208); units are enough when every
          // SKU carries the same penalty rate
          red_shared_add(a_lostU + 4u * r, rem);
          if (!pen_uniform) atomicAdd(&s_lostP[r], (double)rem * sp.pen_rate[lane + k32]);
          rem = 0;
        }
      }
      if (!__any_sync(FULL, ((uint32_t)rem | (uint32_t)n_ok | (a_mask ^ a_mask_end)) != 0u)) break;
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
  // Outbound cost and lost-sales penalty of every warehouse (reward_calculator.py:150-175). Lanes take regions
  // (two per lane cover R <= 64). Per region the lost volume's cost lp is either spread over the warehouses in
  // proportion to what they shipped there (shipment handler, lost_sales_handler.py:113-148: a rate per shipped unit,
  // one division per region) or goes to the closest warehouse (closest handler, and the shipment handler's fallback
  // when nothing was shipped). Then warehouse by warehouse: the lanes' partial sums meet in a shuffle reduction.
  const double pen0 = sp.pen_rate[0];
  constexpr int kRL = 2;                                // regions per lane in the fast path
  if (R <= 32 * kRL) {
    double rate[kRL], lump[kRL];                        // per shipped unit; to the closest warehouse
    int close_w[kRL];
#pragma unroll
    for (int q = 0; q < kRL; ++q) {
      const int r = lane + 32 * q;
      rate[q] = lump[q] = 0.0;
      close_w[q] = -1;
      if (r < R) {
        const int lu = s_lostU[r];
        if (lu > 0) {
          const double lp = pen_uniform ? (double)lu * pen0 : s_lostP[r];
          int shipped_r = 0;
          if (sp.lost_type == MARLSC_LOST_SHIPMENT)
            for (int w = 0; w < W; ++w) shipped_r += s_shipq[w * R + r];
          if (shipped_r > 0) {
            rate[q] = lp / (double)shipped_r;
          } else {
            lump[q] = lp;
            close_w[q] = sp.closest[r];
          }
        }
      }
    }
    for (int w = 0; w < W; ++w) {
      double c = 0.0;
#pragma unroll
      for (int q = 0; q < kRL; ++q) {
        const int r = lane + 32 * q;
        if (r < R) {
          const int sq = s_shipq[w * R + r];
          if (sq > 0) c += (double)sq * (sp.out_var[w * R + r] + rate[q]);
          if (close_w[q] == w) c += lump[q];
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
      if (lane == 0) cost_alloc[e * W + w] = c;
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
