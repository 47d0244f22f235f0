// Inner loop of the attention softmax on B200 (sm_100a): cycles per 64-key row step for 1 / 2 / 4 warps per scheduler.
// Development aid, not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_softmax.bin scripts/ubench_softmax.cu
// One round per thread = what a softmax warp of attention_dq.cu does per 64-key S tile: fetch 64 fp32 scores (here: shared
// memory instead of TMEM), row maximum, p = exp2(s·scale − m), row sum, pack to bf16 pairs, write 32 words back.
//   MAXV 0: eight scalar FMNMX chains            1: three-input max (FMNMX3)
//   EXPV 0: scalar FFMA -> MUFU.EX2 -> FADD (round-2 kernel)
//        1: packed FFMA2 / FADD2 around two MUFU.EX2 per pair
//        R >= 2: as 1, and every R-th pair takes a degree-3 polynomial exp2 on the FMA pipe (Cody-Waite, packed)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float max3(float a, float b, float c) { float d; asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// 2^x for a pair, x <= ~8: round to nearest integer with the 1.5·2^23 trick, degree-3 minimax on [-0.5, 0.5] (7.5e-5 rel),
// exponent added as an integer
__device__ __forceinline__ float2 poly_exp2x2(float2 x) {
  const float kMagic = 12582912.0f;
  x.x = fmaxf(x.x, -120.0f);
  x.y = fmaxf(x.y, -120.0f);
  const float2 t = fadd2(x, make_float2(kMagic, kMagic));
  const float2 n = fadd2(t, make_float2(-kMagic, -kMagic));
  const float2 f = fadd2(x, make_float2(-n.x, -n.y));
  float2 p = ffma2(f, make_float2(0.05517153f, 0.05517153f), make_float2(0.24261111f, 0.24261111f));
  p = ffma2(p, f, make_float2(0.69326103f, 0.69326103f));
  p = ffma2(p, f, make_float2(0.99992806f, 0.99992806f));
  float2 r;
  r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(t.x) << 23));
  r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(t.y) << 23));
  return r;
}

__device__ __forceinline__ uint32_t cvt_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t r;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t h) {
  float2 r;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}\n" : "=f"(r.x), "=f"(r.y) : "r"(h));
  return r;
}

template <int MAXV, int EXPV>
__global__ void __launch_bounds__(512, 1) bench(int reps, long long* out, float* sink, float scale) {
  extern __shared__ float4 sm4[];
  float4* mine = sm4 + threadIdx.x;        // [16][blockDim] float4: conflict free
  const int nt = blockDim.x;
  for (int i = 0; i < 16; ++i)
    mine[i * nt] = make_float4(-0.37f * ((threadIdx.x * 7 + i * 4) % 61), -0.11f * ((threadIdx.x + i) % 53),
                               -0.23f * ((threadIdx.x * 3 + i) % 47), -0.05f * (i + 1));
  float acc = 0.f, m_run = -1e30f;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    float s[64];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 v = mine[i * nt];
      s[4 * i] = v.x; s[4 * i + 1] = v.y; s[4 * i + 2] = v.z; s[4 * i + 3] = v.w;
    }
    float mx;
    if (MAXV == 0) {
      float m8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) m8[c] = s[c];
#pragma unroll
      for (int c = 8; c < 64; ++c) m8[c & 7] = fmaxf(m8[c & 7], s[c]);
      mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
    } else {
      float m4[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) m4[c] = max3(s[c], s[4 + c], s[8 + c]);
#pragma unroll
      for (int c = 12; c < 60; c += 8)
#pragma unroll
        for (int k = 0; k < 4; ++k) m4[k] = max3(m4[k], s[c + k], s[c + 4 + k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) m4[k] = fmaxf(m4[k], s[60 + k]);
      mx = fmaxf(max3(m4[0], m4[1], m4[2]), m4[3]);
    }
    mx *= scale;
    m_run = (mx > m_run + 8.0f) ? mx : m_run;
    const float neg_m = -m_run;
    uint32_t pk[32];
    float lsum;
    if (EXPV == 0) {
      float sum8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p0 = ex2(fmaf(s[2 * c], scale, neg_m)), p1 = ex2(fmaf(s[2 * c + 1], scale, neg_m));
        sum8[(2 * c) & 7] += p0;
        sum8[(2 * c + 1) & 7] += p1;
        pk[c] = pack2(p0, p1);
      }
      lsum = ((sum8[0] + sum8[1]) + (sum8[2] + sum8[3])) + ((sum8[4] + sum8[5]) + (sum8[6] + sum8[7]));
    } else if (EXPV == 10 || EXPV == 11) {
      // half-precision exponentials: one MUFU.EX2 on an f16 pair; P would be an f16 A operand
      const float2 sc2 = make_float2(scale, scale), nm2 = make_float2(neg_m, neg_m);
      float2 sum4[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
      if (EXPV == 10) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float2 x = ffma2(make_float2(s[2 * c], s[2 * c + 1]), sc2, nm2);
          pk[c] = ex2_f16x2(cvt_f16x2(x.x, x.y));
          sum4[c & 3] = fadd2(sum4[c & 3], h2_to_f2(pk[c]));
        }
      } else {
#pragma unroll
        for (int c8 = 0; c8 < 32; c8 += 8) {
#pragma unroll
          for (int c = c8; c < c8 + 8; ++c) {
            const float2 x = ffma2(make_float2(s[2 * c], s[2 * c + 1]), sc2, nm2);
            pk[c] = ex2_f16x2(cvt_f16x2(x.x, x.y));
          }
          const uint32_t h = hadd2(hadd2(hadd2(pk[c8], pk[c8 + 1]), hadd2(pk[c8 + 2], pk[c8 + 3])),
                                   hadd2(hadd2(pk[c8 + 4], pk[c8 + 5]), hadd2(pk[c8 + 6], pk[c8 + 7])));
          sum4[(c8 >> 3) & 3] = fadd2(sum4[(c8 >> 3) & 3], h2_to_f2(h));
        }
      }
      const float2 a = fadd2(fadd2(sum4[0], sum4[1]), fadd2(sum4[2], sum4[3]));
      lsum = a.x + a.y;
    } else {
      float2 sum4[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
      const float2 sc2 = make_float2(scale, scale), nm2 = make_float2(neg_m, neg_m);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float2 x = ffma2(make_float2(s[2 * c], s[2 * c + 1]), sc2, nm2);
        float2 p;
        if (EXPV >= 2 && (c % EXPV) == EXPV - 1) p = poly_exp2x2(x);
        else p = make_float2(ex2(x.x), ex2(x.y));
        sum4[c & 3] = fadd2(sum4[c & 3], p);
        pk[c] = pack2(p.x, p.y);
      }
      const float2 a = fadd2(fadd2(sum4[0], sum4[1]), fadd2(sum4[2], sum4[3]));
      lsum = a.x + a.y;
    }
    acc = acc * 0.5f + lsum;
    // P goes back (the kernel: tcgen05.st); the next round reads perturbed scores
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = mine[i * nt];
      v.x += __uint_as_float((pk[4 * i] & 0x7fffu) | 0x30000000u);
      v.y -= __uint_as_float((pk[4 * i + 1] & 0x7fffu) | 0x30000000u);
      v.z += __uint_as_float((pk[4 * i + 2] & 0x7fffu) | 0x30000000u);
      v.w -= __uint_as_float((pk[4 * i + 3] & 0x7fffu) | 0x30000000u);
      mine[i * nt] = v;
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc + m_run;
}

template <int MAXV, int EXPV>
void run(const char* name, long long* d_out, float* sink) {
  const int reps = 400;
  for (int warps : {4, 8, 16}) {
    const size_t smem = (size_t)warps * 32 * 16 * sizeof(float4);
    cudaFuncSetAttribute(bench<MAXV, EXPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bench<MAXV, EXPV><<<148, warps * 32, smem>>>(reps, d_out, sink, 0.17f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long h = 0;
    cudaMemcpy(&h, d_out, sizeof(long long), cudaMemcpyDeviceToHost);
    const double cyc = (double)h / reps;
    printf("%-34s warps/scheduler=%d: %7.1f cycles per 64-key row step  (%.2f cycles per exponential per scheduler)\n", name,
           warps / 4, cyc, cyc / (64.0 * warps / 4));
  }
}

int main() {
  long long* d_out; float* sink;
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  cudaMalloc(&sink, 148 * 1024 * sizeof(float));
  // accuracy of the polynomial
  run<0, 0>("max8 + scalar EX2 (round-2 kernel)", d_out, sink);
  run<1, 0>("max3 + scalar EX2", d_out, sink);
  run<1, 1>("max3 + packed, all MUFU", d_out, sink);
  run<1, 4>("max3 + packed, 1/4 polynomial", d_out, sink);
  run<1, 3>("max3 + packed, 1/3 polynomial", d_out, sink);
  run<1, 2>("max3 + packed, 1/2 polynomial", d_out, sink);
  run<1, 10>("max3 + f16x2 EX2, fp32 sums", d_out, sink);
  run<1, 11>("max3 + f16x2 EX2, f16 sums of 16", d_out, sink);
  return 0;
}
