// Micro-benchmarks behind the attention kernel's design choices (B200, sm_100a): cycles per tcgen05.mma for the
// operand forms the kernel uses, TMEM load throughput, MUFU ex2 throughput.  Development aid, not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_tc.bin scripts/ubench_tc.cu
#include <cstdio>
#include <vector>

#include "../deepfake-detection-using-clip-based-siglip-2-vision-transformers_b200/csrc/dfd_common.cuh"

using namespace dfd;

struct MmaCase {
  const char* name;
  int a_tmem;     // A operand from TMEM
  int n;          // N
  int b_mn;       // B MN-major
  int layout;     // 2 = SW128, 6 = SW32
  int lbo, sbo;
  int a_layout, a_sbo;
};

template <int mode>
__global__ void __launch_bounds__(128, 2) mma_bench(MmaCase c, int reps, int nacc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc<256>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  // mode 0: `if (threadIdx.x == 0)` (divergent branch);  mode 1: whole warp enters, one lane chosen by elect.sync
  bool issuer;
  if (mode == 0) issuer = threadIdx.x == 0;
  else issuer = (threadIdx.x < 32) ? elect_one() : false;
  if (issuer) {
    const uint64_t dA = umma_desc(smem_u32(smem), 16, c.a_sbo, c.a_layout);
    const uint64_t dB = umma_desc(smem_u32(smem + 32768), c.lbo, c.sbo, c.layout);
    const uint32_t idesc = umma_idesc_bf16_major(128, c.n, 0, c.b_mn);
    // warm
    for (int r = 0; r < 8; ++r) {
      if (c.a_tmem) umma_bf16_ts(tm + 32, tm, dB, idesc, 1);
      else umma_bf16_ss(tm + 32, dA, dB, idesc, 1);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t0 = clock64();
    int a = 0;
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tm + 32 + a * c.n;
      if (c.a_tmem) umma_bf16_ts(d, tm, dB, idesc, 1);
      else umma_bf16_ss(d, dA, dB, idesc, 1);
      a = (a + 1 == nacc) ? 0 : a + 1;
    }
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<256>(tm);
  }
}

// The attention kernel's per-tile MMA mix, issued back to back with no softmax in between:
//   variant 0: 5 x (SS N=64 -> S)            then 4 x (TS N=64 -> O, TS N=16 -> O tail)      (64-key tile, as shipped)
//   variant 1: 5 x (SS N=64 -> S)            then 4 x (TS N=80 -> O)                         (single P.V MMA per k-step)
//   variant 2: 5 x (SS N=128 -> S)           then 8 x (TS N=64, TS N=16)                     (128-key tile)
//   variant 3: 5 x (TS N=64 -> S, Q in TMEM) then 4 x (TS N=80 -> O)
template <int mode>
__global__ void __launch_bounds__(128, 2) seq_bench(int variant, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 80 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (threadIdx.x < 32) tmem_alloc<256>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  bool issuer = (threadIdx.x < 32) ? elect_one() : false;
  if (issuer) {
    const uint32_t q = smem_u32(smem), k = smem_u32(smem + 20480), v = smem_u32(smem + 40960);
    const uint64_t dQ = umma_desc(q, 16, 1024, 2), dQt = umma_desc(q + 16384, 16, 256, 6);
    const uint64_t dK = umma_desc(k, 16, 1024, 2), dKt = umma_desc(k + 16384, 16, 256, 6);
    const uint64_t dV = umma_desc(v, 16, 1024, 2), dVt = umma_desc(v + 16384, 16, 256, 6);
    const uint64_t dV80 = umma_desc(v, 4096, 256, 6);
    const uint32_t tS = tm, tO = tm + 128, tQ = tm + 216;
    const int nqk = variant == 2 ? 128 : 64;
    const uint32_t id_qk = umma_idesc_bf16_major(128, nqk, 0, 0);
    const uint32_t id_pv = umma_idesc_bf16_major(128, 64, 0, 1), id_pvt = umma_idesc_bf16_major(128, 16, 0, 1);
    const uint32_t id_pv80 = umma_idesc_bf16_major(128, 80, 0, 1);
    const int ksteps = variant == 2 ? 8 : 4;
    auto tile = [&]() {
      for (int kk = 0; kk < 4; ++kk) {
        if (variant == 3) umma_bf16_ts(tS, tQ + 8 * kk, dK + 2 * kk, id_qk, kk != 0);
        else umma_bf16_ss(tS, dQ + 2 * kk, dK + 2 * kk, id_qk, kk != 0);
      }
      if (variant == 3) umma_bf16_ts(tS, tQ + 32, dKt, id_qk, 1);
      else umma_bf16_ss(tS, dQt, dKt, id_qk, 1);
      for (int kk = 0; kk < ksteps; ++kk) {
        if (variant == 1 || variant == 3) {
          umma_bf16_ts(tO, tS + 8 * kk, dV80 + kk * 32, id_pv80, 1);
        } else {
          umma_bf16_ts(tO, tS + 8 * kk, dV + kk * 128, id_pv, 1);
          umma_bf16_ts(tO + 64, tS + 8 * kk, dVt + kk * 32, id_pvt, 1);
        }
      }
    };
    tile();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) tile();
    umma_commit(&bar);
    mbar_wait(&bar, 1);
    out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<256>(tm);
  }
}

// every warp loads `cols` fp32 columns of its lane quadrant `reps` times
template <int X>
__global__ void __launch_bounds__(512, 1) ldtm_bench(int reps, long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) tmem_alloc<256>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot + (((threadIdx.x >> 5) & 3) * 32u << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (X == 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tm + (r & 1) * 32, v);
      tmem_ld_wait();
      acc += v[0] ^ v[31];
    } else {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tm + (r & 3) * 16, v);
      tmem_ld_wait();
      acc += v[0] ^ v[15];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<256>(slot);
  }
}

__global__ void __launch_bounds__(1024, 1) mufu_bench(int reps, long long* out, float* sink, int poly) {
  float x[8];
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = fast_exp2(x[i]) - 1.0f;
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s + poly;
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  uint32_t* sink;
  cudaMalloc(&sink, 1 << 22);
  std::vector<long long> h(1024);
  const int reps = 2000;
  MmaCase cases[] = {
      {"ss  A=K-major SW128, B=K-major SW128, N=64 ", 0, 64, 0, 2, 16, 1024, 2, 1024},
      {"ss  A=K-major SW128, B=K-major SW128, N=128", 0, 128, 0, 2, 16, 1024, 2, 1024},
      {"ss  A=K-major SW32,  B=K-major SW32,  N=64 ", 0, 64, 0, 6, 16, 256, 6, 256},
      {"ts  A=TMEM,          B=K-major SW128, N=64 ", 1, 64, 0, 2, 16, 1024, 2, 1024},
      {"ts  A=TMEM,          B=MN-major SW128, N=64", 1, 64, 1, 2, 16, 1024, 2, 1024},
      {"ts  A=TMEM,          B=MN-major SW32,  N=16", 1, 16, 1, 6, 16, 256, 2, 1024},
      {"ts  A=TMEM,          B=MN-major SW32,  N=80 (LBO 2048)", 1, 80, 1, 6, 2048, 256, 2, 1024},
      {"ts  A=TMEM,          B=MN-major SW128, N=128 (LBO 8192)", 1, 128, 1, 2, 8192, 1024, 2, 1024},
      {"ss  A=K-major SW128, B=K-major SW128, N=16 ", 0, 16, 0, 2, 16, 1024, 2, 1024},
      {"ss  A=K-major SW128, B=K-major SW128, N=32 ", 0, 32, 0, 2, 16, 1024, 2, 1024},
  };
  cudaFuncSetAttribute(mma_bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(mma_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (auto& c : cases) {
    for (int cfg = 0; cfg < 6; ++cfg) {
      const int ctas = cfg < 3 ? 148 : 296;
      const int nacc = cfg % 3 + 1;
      if (nacc == 3) continue;
      if (nacc * c.n > 224) continue;
      for (int mode = 0; mode < 2; ++mode) {
      if (mode) mma_bench<1><<<ctas, 128, 100 * 1024>>>(c, reps, nacc, d_out);
      else mma_bench<0><<<ctas, 128, 100 * 1024>>>(c, reps, nacc, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("%s: %s\n", c.name, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h.data(), d_out, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
      printf("mma %-58s ctas/SM=%d accumulators=%d %s: %7.1f cycles/MMA (per CTA)\n", c.name, ctas / 148, nacc, mode ? "elect.sync" : "tid==0    ", (double)h[0] / reps);
      }
    }
  }
  cudaFuncSetAttribute(seq_bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int variant = 0; variant < 4; ++variant)
    for (int ctas : {148, 296}) {
      seq_bench<1><<<ctas, 128, 100 * 1024>>>(variant, 500, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("seq variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h.data(), d_out, sizeof(long long), cudaMemcpyDeviceToHost);
      const double keys = variant == 2 ? 128 : 64;
      printf("attention MMA mix variant %d ctas/SM=%d: %7.1f cycles per tile per CTA = %6.1f cycles per 64 keys per SM\n", variant,
             ctas / 148, (double)h[0] / 500, (double)h[0] / 500 / (keys / 64) / (296 / ctas == 1 ? 2.0 : 1.0) * (ctas == 296 ? 2.0 : 1.0) / 2.0 * (ctas == 296 ? 1.0 : 2.0));
    }
  for (int warps : {4, 8, 16}) {
    ldtm_bench<32><<<148, warps * 32>>>(reps, d_out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d_out, sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = (double)h[0] / reps;
    printf("ldtm x32 warps=%2d: %7.1f cycles per round  -> %6.1f B/clk/SM\n", warps, cyc, warps * 32 * 32 * 4 / cyc);
    ldtm_bench<16><<<148, warps * 32>>>(reps, d_out, sink);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d_out, sizeof(long long), cudaMemcpyDeviceToHost);
    cyc = (double)h[0] / reps;
    printf("ldtm x16 warps=%2d: %7.1f cycles per round  -> %6.1f B/clk/SM\n", warps, cyc, warps * 32 * 16 * 4 / cyc);
  }
  for (int warps : {4, 8, 16, 32}) {
    mufu_bench<<<148, warps * 32>>>(reps, d_out, reinterpret_cast<float*>(sink), 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d_out, sizeof(long long), cudaMemcpyDeviceToHost);
    double cyc = (double)h[0] / reps;
    printf("mufu ex2 warps=%2d: %6.2f ex2/clk/SM\n", warps, warps * 32 * 8 / cyc);
  }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
