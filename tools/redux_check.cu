// Scratch check of warp/partial-warp integer reductions (REDUX) against shuffle reductions on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int G>
__global__ void k(const int* in, int* out_redux, int* out_shfl) {
  const int wl = threadIdx.x & 31;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (wl & ~(G - 1)));
  int v = in[blockIdx.x * blockDim.x + threadIdx.x];
  int a = __reduce_add_sync(gmask, v);
  int b = v;
  for (int o = G / 2; o > 0; o >>= 1) b += __shfl_xor_sync(gmask, b, o);
  // divergent use: only lanes with v > 0 contribute through a predicate-selected value
  int c = __reduce_add_sync(gmask, v > 3 ? v : 0);
  int d = v > 3 ? v : 0;
  for (int o = G / 2; o > 0; o >>= 1) d += __shfl_xor_sync(gmask, d, o);
  out_redux[blockIdx.x * blockDim.x + threadIdx.x] = a * 1000003 + c;
  out_shfl[blockIdx.x * blockDim.x + threadIdx.x] = b * 1000003 + d;
}
template <int G>
int run() {
  const int n = 148 * 8 * 128;
  int *in, *a, *b;
  cudaMallocManaged(&in, n * 4); cudaMallocManaged(&a, n * 4); cudaMallocManaged(&b, n * 4);
  for (int i = 0; i < n; ++i) in[i] = (i * 2654435761u >> 20) % 11 - 2;
  k<G><<<n / 128, 128>>>(in, a, b);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("G=%d cuda error\n", G); return 1; }
  int bad = 0;
  for (int i = 0; i < n; ++i) bad += a[i] != b[i];
  printf("G=%2d mismatches=%d of %d\n", G, bad, n);
  return bad != 0;
}
int main() { return run<32>() | run<16>() | run<8>() | run<4>() | run<2>(); }
