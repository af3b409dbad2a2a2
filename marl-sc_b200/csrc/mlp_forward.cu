// mlp_forward.cu - K7: forward pass of a one-hidden-layer MLP head for every agent-sample of a rollout step, fp32.
//
// What the rollout runs per env step for every agent (reference: ActorCriticRLModule._forward_actor / _forward_critic,
// src/algorithms/models/rlmodules/base.py:412-457, over an "mlp" network with one hidden layer - the IPPO configuration
// of BASELINE configs[3]: hidden_sizes [256], local observations of 14 floats, 2 action means / 1 value):
//   out = W2 act(W1 x + b1) + b2
// As two library GEMMs the hidden activations ([786,432 x 256] floats = 805 MB per network and step at 262,144 small
// environments) are written to HBM and read back, and the step of the whole rollout is bound by that traffic (1.5 of
// 1.8 ms). Here a thread owns one or two rows of x in registers and walks the hidden units: the unit's weights
// [W1[j,:], b1[j], W2[:,j]] are one shared-memory record, read as broadcast 16-byte words, the unit's activation never
// leaves the register file, and the only HBM traffic is x in and out back. Plain fp32 FFMA chains (inputs in ascending
// order per unit, units in ascending order per output), no tensor cores, no reduced precision: results agree with the
// library GEMMs to float32 rounding of a differently ordered sum (tests: rtol 1e-5).
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "lib_common.h"

namespace marlsc {
namespace {

// DP: input width padded to whole 16-byte words (x and W1 rows zero-padded); ROWS rows per thread (four for narrow inputs:
// a record is read once per ROWS rows); ACT 0 ReLU, 1 tanh; NO outputs.
// Shared-memory record of hidden unit j: DP floats of W1[j,:], then b1[j], W2[0,j], W2[1,j], W2[2,j] (O <= 3).
template <int DP, int ROWS, int ACT, int NO>
__global__ void __launch_bounds__(128)
mlp1_forward_kernel(const float* __restrict__ x, long long N, int D, const float* __restrict__ w1, const float* __restrict__ b1, int H,
                    const float* __restrict__ w2, const float* __restrict__ b2, int O, float* __restrict__ out) {
  extern __shared__ float4 s_rec[];
  constexpr int P = DP + 4;
  float* const s = reinterpret_cast<float*>(s_rec);
  for (int i = threadIdx.x; i < H * P; i += blockDim.x) {
    const int j = i / P, c = i - j * P;
    float v = 0.0f;
    if (c < D) v = w1[(long long)j * D + c];
    else if (c == DP) v = b1[j];
    else if (c > DP && c - DP - 1 < O) v = w2[(long long)(c - DP - 1) * H + j];
    s[i] = v;
  }
  __syncthreads();
  const float o0 = b2[0], o1 = O > 1 ? b2[1] : 0.0f, o2 = O > 2 ? b2[2] : 0.0f;
  const long long stride = (long long)gridDim.x * blockDim.x * ROWS;
  for (long long row0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * ROWS; row0 < N; row0 += stride) {
    float xr[ROWS][DP];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int d = 0; d < DP; ++d) xr[r][d] = (d < D && row0 + r < N) ? x[(row0 + r) * D + d] : 0.0f;
    float acc[ROWS][3];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      acc[r][0] = o0;
      acc[r][1] = o1;
      acc[r][2] = o2;
    }
#pragma unroll 2
    for (int j = 0; j < H; ++j) {
      const float4* const rec = s_rec + j * (P / 4);
      float w[DP];
#pragma unroll
      for (int q = 0; q < DP / 4; ++q) {
        const float4 t = rec[q];
        w[4 * q] = t.x;
        w[4 * q + 1] = t.y;
        w[4 * q + 2] = t.z;
        w[4 * q + 3] = t.w;
      }
      const float4 tail = rec[DP / 4];                // b1[j], W2[0..2, j]
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        float h = tail.x;
#pragma unroll
        for (int d = 0; d < DP; ++d) h = fmaf(xr[r][d], w[d], h);
        h = ACT == 0 ? fmaxf(h, 0.0f) : tanhf(h);
        acc[r][0] = fmaf(h, tail.y, acc[r][0]);
        if (NO > 1) acc[r][1] = fmaf(h, tail.z, acc[r][1]);
        if (NO > 2) acc[r][2] = fmaf(h, tail.w, acc[r][2]);
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      if (row0 + r < N) {
#pragma unroll
        for (int k = 0; k < NO; ++k) out[(row0 + r) * NO + k] = acc[r][k];
      }
  }
}

template <int DP, int ROWS>
int launch_mlp1(const float* x, long long N, int D, const float* w1, const float* b1, int H, const float* w2, const float* b2, int O,
                int act, float* out, cudaStream_t s) {
  const size_t smem = (size_t)H * (DP + 4) * sizeof(float);
  int dev = 0, sms = 0, optin = 0;
  MARLSC_CUDA(cudaGetDevice(&dev));
  MARLSC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  MARLSC_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if ((int)smem > optin) return set_error(MARLSC_EUNSUPPORTED, "mlp1_forward: the hidden layer's weights do not fit shared memory");
  const long long want = (N + 128LL * ROWS - 1) / (128LL * ROWS);
  const unsigned grid = (unsigned)(want < (long long)sms * 8 ? (want > 0 ? want : 1) : (long long)sms * 8);
#define MARLSC_MLP1(A, K)                                                                                                  \
  {                                                                                                                        \
    if (smem > 48 * 1024)                                                                                                  \
      MARLSC_CUDA(cudaFuncSetAttribute((const void*)mlp1_forward_kernel<DP, ROWS, A, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    mlp1_forward_kernel<DP, ROWS, A, K><<<grid, 128, smem, s>>>(x, N, D, w1, b1, H, w2, b2, O, out);                        \
  }
  if (act == 0) {
    if (O == 1) MARLSC_MLP1(0, 1) else if (O == 2) MARLSC_MLP1(0, 2) else MARLSC_MLP1(0, 3)
  } else {
    if (O == 1) MARLSC_MLP1(1, 1) else if (O == 2) MARLSC_MLP1(1, 2) else MARLSC_MLP1(1, 3)
  }
#undef MARLSC_MLP1
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MARLSC_CUDA(cudaGetLastError());
  return MARLSC_OK;
}

}  // namespace
}  // namespace marlsc

using namespace marlsc;

extern "C" int marlsc_mlp1_forward(const float* x, int64_t n_rows, int32_t in_dim, const float* w1, const float* b1, int32_t hidden,
                                   const float* w2, const float* b2, int32_t out_dim, int32_t activation, float* out, void* stream) {
  if (!x || !w1 || !b1 || !w2 || !b2 || !out) return set_error(MARLSC_EINVAL, "mlp1_forward: null pointer");
  if (n_rows < 0 || in_dim < 1 || in_dim > 64 || hidden < 1 || out_dim < 1 || out_dim > 3 || (activation != 0 && activation != 1))
    return set_error(MARLSC_EUNSUPPORTED, "mlp1_forward: in_dim <= 64, out_dim <= 3, activation 0 (ReLU) or 1 (tanh)");
  if (n_rows == 0) return MARLSC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (in_dim <= 16) return launch_mlp1<16, 4>(x, n_rows, in_dim, w1, b1, hidden, w2, b2, out_dim, activation, out, s);
  if (in_dim <= 32) return launch_mlp1<32, 2>(x, n_rows, in_dim, w1, b1, hidden, w2, b2, out_dim, activation, out, s);
  return launch_mlp1<64, 1>(x, n_rows, in_dim, w1, b1, hidden, w2, b2, out_dim, activation, out, s);
}
