// MUFU.EX2 issue ceiling for the attention softmax (B200, sm_100a).  Development aid, not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_mufu.bin scripts/ubench_mufu.cu
// Variants (per thread, per round, 64 values):
//   0: 64 x (FFMA -> EX2), then 64 FADD + 32 PRMT            (the kernel's two-phase loop)
//   1: 64 x EX2 only (arguments precomputed), sums afterwards
//   2: interleaved EX2 / FADD as ptxas schedules a single loop
//   3: 48 EX2 + 16 polynomial exp2 on the FMA pipe
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

template <int V>
__global__ void __launch_bounds__(512, 1) bench(int reps, long long* out, float* sink, float scale, float negm) {
  float s[64];
  for (int i = 0; i < 64; ++i) s[i] = -0.01f * (threadIdx.x + i);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    float e[64];
    if (V == 0) {
#pragma unroll
      for (int c = 0; c < 64; ++c) e[c] = ex2(fmaf(s[c], scale, negm));
    } else if (V == 1) {
#pragma unroll
      for (int c = 0; c < 64; ++c) e[c] = ex2(s[c]);
    } else if (V == 2) {
#pragma unroll
      for (int c = 0; c < 64; ++c) { e[c] = ex2(fmaf(s[c], scale, negm)); acc += e[c]; }
    } else {
#pragma unroll
      for (int c = 0; c < 64; ++c) e[c] = (c & 3) == 3 ? poly_exp2(fmaf(s[c], scale, negm)) : ex2(fmaf(s[c], scale, negm));
    }
    float sum8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 64; ++c) sum8[c & 7] += e[c];
    unsigned pk = 0;
#pragma unroll
    for (int c = 0; c < 32; ++c) pk ^= __byte_perm(__float_as_uint(e[2 * c]), __float_as_uint(e[2 * c + 1]), 0x7632);
    acc += sum8[0] + sum8[1] + sum8[2] + sum8[3] + sum8[4] + sum8[5] + sum8[6] + sum8[7] + __uint_as_float(pk & 0x3fffffff);
#pragma unroll
    for (int c = 0; c < 64; ++c) s[c] = s[c] * 0.999f - 1e-6f * e[c];
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  long long* d_out; float* sink;
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  cudaMalloc(&sink, 148 * 1024 * sizeof(float));
  std::vector<long long> h(148);
  const int reps = 200;
  for (int v = 0; v < 4; ++v)
    for (int warps : {4, 8, 16}) {
      switch (v) {
        case 0: bench<0><<<148, warps * 32>>>(reps, d_out, sink, 0.17f, -0.3f); break;
        case 1: bench<1><<<148, warps * 32>>>(reps, d_out, sink, 0.17f, -0.3f); break;
        case 2: bench<2><<<148, warps * 32>>>(reps, d_out, sink, 0.17f, -0.3f); break;
        default: bench<3><<<148, warps * 32>>>(reps, d_out, sink, 0.17f, -0.3f); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), d_out, sizeof(long long), cudaMemcpyDeviceToHost);
      const double cyc = (double)h[0] / reps;
      printf("variant %d warps/SM=%2d: %8.1f cycles per round of 64 values per thread  -> %5.2f exp/clk/SM (%.1f cycles per warp-EX2 slot per scheduler)\n",
             v, warps, cyc, warps * 32 * 64 / cyc, cyc / (64.0 * warps / 4));
    }
  return 0;
}
