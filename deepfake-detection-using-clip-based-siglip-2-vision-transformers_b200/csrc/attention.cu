// Non-causal multi-head attention for the 196/576/729/1024-token SigLIP sequences, and the
// single-query MAP pooling attention.
//
//   attention_fwd_kernel<HD>  flash-style: one CTA = 128 query rows of one (image, head); K/V streamed in
//                             64-key tiles through a cp.async double buffer; S = QKᵀ and O += PV on the
//                             warp-level tensor-core path (mma.sync m16n8k16 / m16n8k8 bf16, fp32 accumulate),
//                             online softmax in fp32 registers.  hd=72 needs no padding: 72 = 4·16 + 8 along
//                             K for QKᵀ and 9 n8-tiles for PV.  Shared rows are 144 B apart for both head
//                             sizes, which makes every ldmatrix phase conflict-free.
//                             (HF:modeling_siglip.py:229-249,293-306 — SDPA, is_causal=False, fp32 softmax)
//   map_attention_kernel      softmax_n(q_h·k_n)·v_n for the one learned probe query (HF:modeling_siglip.py:639-646)
#include "dfd_common.cuh"

#include <algorithm>
#include <atomic>

namespace dfd {

extern std::atomic<int64_t> g_launches;

namespace {

constexpr int kBQ = 128;        // query rows per CTA (8 warps x 16)
constexpr int kBKV = 64;        // keys per tile
constexpr int kLds = 72;        // shared row stride in elements (144 B)
constexpr int kAttnThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = smem_u32(smem);
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_k16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                        uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}

template <int HD>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int rows,
                                          int row0, int N) {
  constexpr int kChunks = HD / 8;
  for (int i = threadIdx.x; i < rows * kChunks; i += kAttnThreads) {
    const int r = i / kChunks, c = i - r * kChunks;
    const bool ok = (row0 + r) < N;
    const __nv_bfloat16* g = src + (int64_t)(ok ? (row0 + r) : 0) * ld + c * 8;
    cp_async16(dst + r * kLds + c * 8, g, ok);
  }
}

template <int HD>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, int64_t ldqkv, __nv_bfloat16* __restrict__ out,
                     int64_t ldo, int N, int H, float scale_log2) {
  static_assert(HD == 64 || HD == 72, "head dim");
  constexpr int kK16 = HD / 16;       // 4
  constexpr bool kTail8 = (HD % 16) != 0;
  constexpr int kNT = HD / 8;         // PV n-tiles: 8 or 9
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + kBQ * kLds;
  __nv_bfloat16* sV = sK + 2 * kBKV * kLds;

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* base = qkv + (int64_t)b * N * ldqkv;
  const __nv_bfloat16* gQ = base + h * HD;
  const __nv_bfloat16* gK = base + D + h * HD;
  const __nv_bfloat16* gV = base + 2 * D + h * HD;
  const int q0 = qb * kBQ;
  const int num_kv = (N + kBKV - 1) / kBKV;

  load_tile<HD>(sQ, gQ, ldqkv, kBQ, q0, N);
  load_tile<HD>(sK, gK, ldqkv, kBKV, 0, N);
  load_tile<HD>(sV, gV, ldqkv, kBKV, 0, N);
  cp_async_commit();

  uint32_t qf[kK16][4];
  uint32_t qt[2] = {0u, 0u};
  float o[kNT][4];
#pragma unroll
  for (int j = 0; j < kNT; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const int lrow = lane & 7, lmat = lane >> 3;

  for (int t = 0; t < num_kv; ++t) {
    const int buf = t & 1;
    if (t + 1 < num_kv) {
      load_tile<HD>(sK + (buf ^ 1) * kBKV * kLds, gK, ldqkv, kBKV, (t + 1) * kBKV, N);
      load_tile<HD>(sV + (buf ^ 1) * kBKV * kLds, gV, ldqkv, kBKV, (t + 1) * kBKV, N);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (t == 0) {
      // Q fragments stay in registers for the whole CTA lifetime
      const uint32_t qbase = smem_u32(sQ + (warp * 16 + (lmat & 1) * 8 + lrow) * kLds + (lmat >> 1) * 8);
#pragma unroll
      for (int ks = 0; ks < kK16; ++ks) ldsm_x4(qbase + ks * 32, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      if (kTail8) {
        const uint32_t qa = smem_u32(sQ + (warp * 16 + (lmat & 1) * 8 + lrow) * kLds + kK16 * 16);
        ldsm_x2(qa, qt[0], qt[1]);
      }
    }

    const __nv_bfloat16* tK = sK + buf * kBKV * kLds;
    const __nv_bfloat16* tV = sV + buf * kBKV * kLds;

    // ---- S = Q K^T -------------------------------------------------------------------------
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {  // pairs of key n-tiles
      const uint32_t kb = smem_u32(tK + ((2 * jp + (lmat >> 1)) * 8 + lrow) * kLds + (lmat & 1) * 8);
#pragma unroll
      for (int ks = 0; ks < kK16; ++ks) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb + ks * 32, b0, b1, b2, b3);
        mma_k16(s[2 * jp], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
        mma_k16(s[2 * jp + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
      }
    }
    if (kTail8) {
#pragma unroll
      for (int jq = 0; jq < 2; ++jq) {  // 4 key n-tiles per ldmatrix.x4, d = 64..71
        const uint32_t kb = smem_u32(tK + ((4 * jq + lmat) * 8 + lrow) * kLds + kK16 * 16);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(kb, b0, b1, b2, b3);
        mma_k8(s[4 * jq + 0], qt[0], qt[1], b0);
        mma_k8(s[4 * jq + 1], qt[0], qt[1], b1);
        mma_k8(s[4 * jq + 2], qt[0], qt[1], b2);
        mma_k8(s[4 * jq + 3], qt[0], qt[1], b3);
      }
    }

    // ---- online softmax (rows g = lane/4 and g+8) ------------------------------------------------
    const int key0 = t * kBKV + (lane & 3) * 2;
    const bool tail = (t + 1) * kBKV > N;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = s[j][e] * scale_log2;
        if (tail && (key0 + j * 8 + (e & 1)) >= N) v = -INFINITY;
        s[j][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = exp2f(m_run[r] - m_new);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p = exp2f(s[j][e] - m_run[e >> 1]);
        s[j][e] = p;
        rs[e >> 1] += p;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
    for (int j = 0; j < kNT; ++j) {
      o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
      o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
    }

    // ---- O += P V ------------------------------------------------------------------------------
#pragma unroll
    for (int kk = 0; kk < kBKV / 16; ++kk) {
      const uint32_t a0 = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      const uint32_t a1 = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      const uint32_t a2 = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      const uint32_t a3 = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const uint32_t vb = smem_u32(tV + (kk * 16 + (lmat & 1) * 8 + lrow) * kLds + (lmat >> 1) * 8);
#pragma unroll
      for (int jd = 0; jd < kNT / 2; ++jd) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vb + jd * 32, b0, b1, b2, b3);
        mma_k16(o[2 * jd], a0, a1, a2, a3, b0, b1);
        mma_k16(o[2 * jd + 1], a0, a1, a2, a3, b2, b3);
      }
      if (kNT & 1) {
        const uint32_t va = smem_u32(tV + (kk * 16 + (lmat & 1) * 8 + lrow) * kLds + (kNT - 1) * 8);
        uint32_t b0, b1;
        ldsm_x2_t(va, b0, b1);
        mma_k16(o[kNT - 1], a0, a1, a2, a3, b0, b1);
      }
    }
    __syncthreads();  // everyone is done with buffer `buf` before it is refilled
  }

  // ---- finalise: O / l, bf16 store ---------------------------------------------------------------
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int row0 = q0 + warp * 16 + (lane >> 2);
  __nv_bfloat16* ob = out + (int64_t)b * N * ldo + h * HD + (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < kNT; ++j) {
    if (row0 < N)
      *reinterpret_cast<uint32_t*>(ob + (int64_t)row0 * ldo + j * 8) = pack_bf16x2(o[j][0] * inv0, o[j][1] * inv0);
    if (row0 + 8 < N)
      *reinterpret_cast<uint32_t*>(ob + (int64_t)(row0 + 8) * ldo + j * 8) =
          pack_bf16x2(o[j][2] * inv1, o[j][3] * inv1);
  }
}

// ---- MAP pooling attention: one CTA per (image, head), one query --------------------------------
// HBM-bound (K and V of one head are read exactly once, N x 2 x HD bf16).  Both passes walk the rows in 16-byte chunks
// with consecutive threads on consecutive chunks, so every row's HD*2 bytes are fetched as whole sectors and a CTA keeps
// 256 independent 16-byte loads in flight: the scores pass leaves one partial dot product per chunk in shared memory, the
// P·V pass gives each thread one chunk column and every (256 / chunks)-th key.
constexpr int kMapThreads = 256;

template <int HD>
__global__ void __launch_bounds__(kMapThreads)
map_attention_kernel(const __nv_bfloat16* __restrict__ kv, int64_t ldkv, const float* __restrict__ q,
                     __nv_bfloat16* __restrict__ out, int64_t ldo, int N, int H, float scale) {
  constexpr int kC = HD / 8;             // 16-byte chunks per row
  constexpr int kG = kMapThreads / kC;   // key groups of the P·V pass
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* sc = reinterpret_cast<float*>(smem_raw);   // [N] scores -> probabilities
  float* ps = sc + ((N + 3) & ~3);                  // [max(N*kC, kG*HD)] chunk partials, later the P·V partials
  __shared__ __align__(16) float sq[HD];
  __shared__ float red[kMapThreads / 32];
  const int h = blockIdx.x, b = blockIdx.y;
  const int D = H * HD;
  const __nv_bfloat16* gK = kv + (int64_t)b * N * ldkv + h * HD;
  const __nv_bfloat16* gV = gK + D;

  if (threadIdx.x < HD) sq[threadIdx.x] = __ldg(q + h * HD + threadIdx.x) * scale;
  __syncthreads();
  for (int i = threadIdx.x; i < N * kC; i += kMapThreads) {
    const int n = i / kC, c = i - n * kC;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(gK + (int64_t)n * ldkv) + c);
    const float4 q0 = *reinterpret_cast<const float4*>(sq + c * 8), q1 = *reinterpret_cast<const float4*>(sq + c * 8 + 4);
    const float2 a = unpack_bf16x2(v.x), b2 = unpack_bf16x2(v.y), c2 = unpack_bf16x2(v.z), d2 = unpack_bf16x2(v.w);
    ps[i] = ((a.x * q0.x + a.y * q0.y) + (b2.x * q0.z + b2.y * q0.w)) + ((c2.x * q1.x + c2.y * q1.y) + (d2.x * q1.z + d2.y * q1.w));
  }
  __syncthreads();
  float lmax = -INFINITY;
  for (int n = threadIdx.x; n < N; n += kMapThreads) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < kC; ++c) acc += ps[n * kC + c];
    sc[n] = acc;
    lmax = fmaxf(lmax, acc);
  }
  lmax = warp_max(lmax);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lmax;
  __syncthreads();
  float gmax = red[0];
#pragma unroll
  for (int i = 1; i < kMapThreads / 32; ++i) gmax = fmaxf(gmax, red[i]);
  __syncthreads();
  float lsum = 0.f;
  for (int n = threadIdx.x; n < N; n += kMapThreads) {
    const float p = __expf(sc[n] - gmax);
    sc[n] = p;
    lsum += p;
  }
  lsum = warp_sum(lsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
  __syncthreads();
  float gsum = 0.f;
#pragma unroll
  for (int i = 0; i < kMapThreads / 32; ++i) gsum += red[i];

  // out[d] = sum_n p[n] v[n,d] / gsum
  const int g = threadIdx.x / kC, c = threadIdx.x - g * kC;
  if (g < kG) {
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int n = g; n < N; n += kG) {
      const float p = sc[n];
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(gV + (int64_t)n * ldkv) + c);
      const float2 v0 = unpack_bf16x2(v.x), v1 = unpack_bf16x2(v.y), v2 = unpack_bf16x2(v.z), v3 = unpack_bf16x2(v.w);
      a[0] += p * v0.x; a[1] += p * v0.y; a[2] += p * v1.x; a[3] += p * v1.y;
      a[4] += p * v2.x; a[5] += p * v2.y; a[6] += p * v3.x; a[7] += p * v3.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ps[g * HD + c * 8 + j] = a[j];
  }
  __syncthreads();
  if (threadIdx.x < HD) {
    const int d = threadIdx.x;
    float v = 0.f;
#pragma unroll 4
    for (int k = 0; k < kG; ++k) v += ps[k * HD + d];
    out[(int64_t)b * ldo + h * HD + d] = __float2bfloat16(v / gsum);
  }
}

}  // namespace

int attention_bf16(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int B, int N, int H, int hd,
                   float scale, cudaStream_t st) {
  DFD_REQUIRE(qkv && out, DFD_ERR_BAD_ARG, "attention: null pointer");
  DFD_REQUIRE(B > 0 && N > 0 && H > 0, DFD_ERR_SHAPE, "attention: B, N, H must be positive");
  DFD_REQUIRE(hd == 64 || hd == 72, DFD_ERR_UNSUPPORTED, "attention: head dim %d not supported (64, 72)", hd);
  DFD_REQUIRE(ldqkv % 8 == 0 && ldqkv >= 3 * H * hd && ldo % 8 == 0 && ldo >= H * hd, DFD_ERR_SHAPE,
              "attention: bad leading dimensions");
  DFD_REQUIRE(B <= 65535 && H <= 65535, DFD_ERR_SHAPE, "attention: B and H must be <= 65535");
  DFD_REQUIRE(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 4 == 0), DFD_ERR_BAD_ARG,
              "attention: pointers must be 16-byte aligned");
  const int smem = (kBQ + 4 * kBKV) * kLds * 2;
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((N + kBQ - 1) / kBQ, H, B);
  static SmemOptIn smem_once[2];
  if (hd == 64) {
    if (int rc = ensure_dynamic_smem(smem_once[0], attention_fwd_kernel<64>, smem)) return rc;
    attention_fwd_kernel<64><<<grid, kAttnThreads, smem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), ldqkv, reinterpret_cast<__nv_bfloat16*>(out), ldo, N, H, scale_log2);
  } else {
    if (int rc = ensure_dynamic_smem(smem_once[1], attention_fwd_kernel<72>, smem)) return rc;
    attention_fwd_kernel<72><<<grid, kAttnThreads, smem, st>>>(
        reinterpret_cast<const __nv_bfloat16*>(qkv), ldqkv, reinterpret_cast<__nv_bfloat16*>(out), ldo, N, H, scale_log2);
  }
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

int map_attention_bf16(const void* kv, int64_t ldkv, const float* q, void* out, int64_t ldo, int B, int N,
                       int H, int hd, float scale, cudaStream_t st) {
  DFD_REQUIRE(kv && q && out, DFD_ERR_BAD_ARG, "map_attention: null pointer");
  DFD_REQUIRE(B > 0 && N > 0 && H > 0, DFD_ERR_SHAPE, "map_attention: B, N, H must be positive");
  DFD_REQUIRE(hd == 64 || hd == 72, DFD_ERR_UNSUPPORTED, "map_attention: head dim %d not supported (64, 72)", hd);
  DFD_REQUIRE(ldkv % 8 == 0 && ldkv >= 2 * H * hd && ldo >= H * hd, DFD_ERR_SHAPE,
              "map_attention: bad leading dimensions");
  DFD_REQUIRE(B <= 65535, DFD_ERR_SHAPE, "map_attention: B must be <= 65535");
  DFD_REQUIRE(((uintptr_t)kv % 16 == 0), DFD_ERR_BAD_ARG, "map_attention: kv must be 16-byte aligned");
  const int chunks = hd / 8, groups = kMapThreads / chunks;
  const int64_t part_words = std::max<int64_t>((int64_t)N * chunks, (int64_t)groups * hd);
  const int64_t smem64 = (((int64_t)N + 3) & ~3ll) * 4 + part_words * 4;
  DFD_REQUIRE(smem64 <= 200 * 1024, DFD_ERR_UNSUPPORTED, "map_attention: N=%d too long", N);
  const int smem = (int)smem64;
  dim3 grid(H, B);
  static SmemOptIn smem_once[2];
  if (hd == 64) {
    if (int rc = ensure_dynamic_smem(smem_once[0], map_attention_kernel<64>, 200 * 1024)) return rc;
    map_attention_kernel<64><<<grid, kMapThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(kv), ldkv, q,
                                                             reinterpret_cast<__nv_bfloat16*>(out), ldo, N, H, scale);
  } else {
    if (int rc = ensure_dynamic_smem(smem_once[1], map_attention_kernel<72>, 200 * 1024)) return rc;
    map_attention_kernel<72><<<grid, kMapThreads, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(kv), ldkv, q,
                                                             reinterpret_cast<__nv_bfloat16*>(out), ldo, N, H, scale);
  }
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

}  // namespace dfd

extern "C" DFD_API int dfd_attention_bf16(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int B,
                                          int N, int H, int hd, float scale, void* stream) {
  return dfd::attention_auto_bf16(qkv, ldqkv, out, ldo, B, N, H, hd, scale, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" DFD_API int dfd_map_attention_bf16(const void* kv, int64_t ldkv, const float* q, void* out,
                                              int64_t ldo, int B, int N, int H, int hd, float scale,
                                              void* stream) {
  return dfd::map_attention_bf16(kv, ldkv, q, out, ldo, B, N, H, hd, scale,
                                 reinterpret_cast<cudaStream_t>(stream));
}
