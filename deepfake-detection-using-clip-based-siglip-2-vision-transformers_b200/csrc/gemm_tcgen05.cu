// Persistent, warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epi(A[M,K] · W[N,K]ᵀ)
//
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, kStages-deep mbarrier ring)
//   warp 1      MMA issuer     (one thread issues tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16;
//                               accumulators live in TMEM, two accumulator stages)
//   warps 2..5  epilogue       (tcgen05.ld 32x32b → registers → fused epilogue → bf16 global stores)
//
// Replaces the cuBLAS/cuDNN calls behind nn.Linear / nn.Conv2d in the reference's backbone
// (HF:modeling_siglip.py:178,285-287,308,324-326; SURVEY.md §2.2 K2,K4,K6,K7,K8,K9).
#include "dfd_common.cuh"

#include <atomic>
#include <mutex>

namespace dfd {

extern std::atomic<int64_t> g_launches;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kGemmThreads = 192;

struct EpiArgs {
  const float* bias;
  int act;
  const float* pos;
  int pos_rows;
  const __nv_bfloat16* residual;
  int64_t ldr;
  const float* ln_rowstats;
  const float* ln_colsum;
  float ln_inv_dim;
  float ln_eps;
  float* stats_out;
};

template <int BN>
struct GemmSmem {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 192 ? 5 : 6);
  static constexpr int kTileBytes = kStages * kStageBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kTileBytes + kBarBytes + 1024;  // +1024 alignment slack
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, __nv_bfloat16* C,
                         int64_t ldc, int M, int N, int K, EpiArgs epi) {
  using S = GemmSmem<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kTileBytes);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tfull_bar = empty_bar + S::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
#pragma unroll
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<S::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / num_n) * BM;
        const int n0 = (tile % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * S::kStageBytes;
          uint8_t* sb = sa + S::kABytes;
          mbar_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_2d(&tmA, &full_bar[stage], sa, kb * BK, m0);
          tma_load_2d(&tmB, &full_bar[stage], sb, kb * BK, n0);
          if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::kStageBytes);
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sa + S::kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-byte swizzle row: +2 in 16-byte units
            umma_bf16_ss(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                         idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage when the MMAs retire
          if (++stage == S::kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / num_n) * BM;
      const int n0 = (tile % num_n) * BN;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < M;

      float ln_mean = 0.f, ln_rstd = 1.f;
      if (epi.ln_colsum != nullptr && row_ok) {
        const float2 st = *reinterpret_cast<const float2*>(epi.ln_rowstats + 2 * (int64_t)row);
        ln_mean = st.x * epi.ln_inv_dim;
        const float var = fmaxf(st.y * epi.ln_inv_dim - ln_mean * ln_mean, 0.f);
        ln_rstd = rsqrtf(var + epi.ln_eps);
      }
      const float* pos_row =
          epi.pos != nullptr ? epi.pos + (int64_t)(row % epi.pos_rows) * N : nullptr;
      const __nv_bfloat16* res_row =
          epi.residual != nullptr ? epi.residual + (int64_t)row * epi.ldr : nullptr;
      __nv_bfloat16* c_row = C + (int64_t)row * ldc;
      float st_sum = 0.f, st_sq = 0.f;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                             static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        const int c0 = n0 + ch * 32;
        if (c0 >= N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + static_cast<uint32_t>(ch * 32), r);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int c = c0 + g * 8;
            if (c < N) {
              float v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
              if (epi.ln_colsum != nullptr) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(epi.ln_colsum + c));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(epi.ln_colsum + c + 4));
                const float cs[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = ln_rstd * (v[j] - ln_mean * cs[j]);
              }
              if (epi.bias != nullptr) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(epi.bias + c));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(epi.bias + c + 4));
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (epi.act == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = gelu_tanh(v[j]);
              }
              if (pos_row != nullptr) {
                const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos_row + c));
                const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos_row + c + 4));
                v[0] += p0.x; v[1] += p0.y; v[2] += p0.z; v[3] += p0.w;
                v[4] += p1.x; v[5] += p1.y; v[6] += p1.z; v[7] += p1.w;
              }
              if (res_row != nullptr) {
                const uint4 rr = *reinterpret_cast<const uint4*>(res_row + c);
                const float2 a0 = unpack_bf16x2(rr.x), a1 = unpack_bf16x2(rr.y),
                             a2 = unpack_bf16x2(rr.z), a3 = unpack_bf16x2(rr.w);
                v[0] += a0.x; v[1] += a0.y; v[2] += a1.x; v[3] += a1.y;
                v[4] += a2.x; v[5] += a2.y; v[6] += a3.x; v[7] += a3.y;
              }
              uint4 o;
              o.x = pack_bf16x2(v[0], v[1]);
              o.y = pack_bf16x2(v[2], v[3]);
              o.z = pack_bf16x2(v[4], v[5]);
              o.w = pack_bf16x2(v[6], v[7]);
              if (epi.stats_out != nullptr) {
                const float2 q0 = unpack_bf16x2(o.x), q1 = unpack_bf16x2(o.y),
                             q2 = unpack_bf16x2(o.z), q3 = unpack_bf16x2(o.w);
                st_sum += ((q0.x + q0.y) + (q1.x + q1.y)) + ((q2.x + q2.y) + (q3.x + q3.y));
                st_sq += ((q0.x * q0.x + q0.y * q0.y) + (q1.x * q1.x + q1.y * q1.y)) +
                         ((q2.x * q2.x + q2.y * q2.y) + (q3.x * q3.x + q3.y * q3.y));
              }
              *reinterpret_cast<uint4*>(c_row + c) = o;
            }
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (wait::ld above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (epi.stats_out != nullptr && row_ok) {
        atomicAdd(epi.stats_out + 2 * (int64_t)row, st_sum);
        atomicAdd(epi.stats_out + 2 * (int64_t)row + 1, st_sq);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<S::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  });
  return fn;
}

}  // namespace

// bf16 row-major [rows, cols] with leading dimension ld (elements) -> 2-D tiled map, box = {64, box_rows}
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld,
                      int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  DFD_REQUIRE(fn != nullptr, DFD_ERR_NO_DEVICE,
              "cuTensorMapEncodeTiled unavailable (no CUDA driver on this host)");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstr[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFD_REQUIRE(r == CUDA_SUCCESS, DFD_ERR_CUDA,
              "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r,
              (long long)rows, (long long)cols, (long long)ld, box_rows);
  return DFD_OK;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, __nv_bfloat16* C, int64_t ldc,
                       int M, int N, int K, const EpiArgs& ea, cudaStream_t st) {
  using S = GemmSmem<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    DFD_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  const int num_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = num_tiles < kNumSMs ? num_tiles : kNumSMs;
  gemm_bf16_tcgen05_kernel<BN><<<grid, kGemmThreads, S::kTotal, st>>>(tmA, tmB, C, ldc, M, N, K, ea);
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

// Tile-N choice: the widest tile that does not waste more than ~6 % of the MMA work on padding,
// preferring tiles that give every SM work.
static int pick_bn(int M, int N) {
  const int cands[3] = {256, 192, 128};
  int best = 128;
  double best_cost = 1e30;
  const int num_m = (M + BM - 1) / BM;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const int num_n = (N + bn - 1) / bn;
    const long tiles = (long)num_m * num_n;
    const long waves = (tiles + kNumSMs - 1) / kNumSMs;
    // time ~ waves * (bn columns per tile) with a mild bonus for wider tiles (better operand reuse)
    const double cost = (double)waves * bn * (bn == 256 ? 1.0 : (bn == 192 ? 1.03 : 1.10));
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

int gemm_bf16_dispatch(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc,
                       int M, int N, int K, const dfd_gemm_epilogue* epi, int force_bn,
                       cudaStream_t st) {
  DFD_REQUIRE(A && W && C, DFD_ERR_BAD_ARG, "gemm: null pointer");
  DFD_REQUIRE(M > 0 && N > 0 && K > 0, DFD_ERR_SHAPE, "gemm: M,N,K must be positive (%d,%d,%d)", M, N, K);
  DFD_REQUIRE(K % 8 == 0 && N % 8 == 0, DFD_ERR_SHAPE, "gemm: K and N must be multiples of 8 (K=%d N=%d)", K, N);
  DFD_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && ldc % 8 == 0, DFD_ERR_SHAPE,
              "gemm: leading dimensions must be multiples of 8 elements");
  DFD_REQUIRE(lda >= K && ldw >= K && ldc >= N, DFD_ERR_SHAPE, "gemm: leading dimension too small");
  DFD_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)W % 16 == 0) && ((uintptr_t)C % 16 == 0),
              DFD_ERR_BAD_ARG, "gemm: pointers must be 16-byte aligned");
  EpiArgs ea{};
  if (epi != nullptr) {
    ea.bias = epi->bias;
    ea.act = epi->act;
    ea.pos = epi->pos;
    ea.pos_rows = epi->pos_rows > 0 ? epi->pos_rows : 1;
    ea.residual = reinterpret_cast<const __nv_bfloat16*>(epi->residual);
    ea.ldr = epi->ldr;
    ea.ln_rowstats = epi->ln_rowstats;
    ea.ln_colsum = epi->ln_colsum;
    ea.ln_inv_dim = epi->ln_dim > 0 ? 1.0f / (float)epi->ln_dim : 0.f;
    ea.ln_eps = epi->ln_eps;
    ea.stats_out = epi->stats_out;
    DFD_REQUIRE(ea.act == 0 || ea.act == 1, DFD_ERR_BAD_ARG, "gemm: act must be 0 or 1");
    DFD_REQUIRE((ea.ln_colsum == nullptr) == (ea.ln_rowstats == nullptr), DFD_ERR_BAD_ARG,
                "gemm: ln_rowstats and ln_colsum must be given together");
    DFD_REQUIRE(ea.ln_colsum == nullptr || epi->ln_dim > 0, DFD_ERR_BAD_ARG, "gemm: ln_dim must be > 0");
    DFD_REQUIRE(ea.residual == nullptr || (ea.ldr % 8 == 0 && ea.ldr >= N), DFD_ERR_SHAPE,
                "gemm: residual leading dimension invalid");
  }
  const int bn = force_bn > 0 ? force_bn : pick_bn(M, N);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, A, M, K, lda, BM);
  if (rc != DFD_OK) return rc;
  rc = make_tmap_bf16_2d(&tmB, W, N, K, ldw, bn);
  if (rc != DFD_OK) return rc;
  __nv_bfloat16* Cb = reinterpret_cast<__nv_bfloat16*>(C);
  switch (bn) {
    case 256: return launch_gemm<256>(tmA, tmB, Cb, ldc, M, N, K, ea, st);
    case 192: return launch_gemm<192>(tmA, tmB, Cb, ldc, M, N, K, ea, st);
    case 128: return launch_gemm<128>(tmA, tmB, Cb, ldc, M, N, K, ea, st);
    default: break;
  }
  set_last_error("gemm: unsupported tile N %d", bn);
  return DFD_ERR_UNSUPPORTED;
}

}  // namespace dfd

extern "C" DFD_API int dfd_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C,
                                     int64_t ldc, int M, int N, int K, const dfd_gemm_epilogue* epi,
                                     void* stream) {
  return dfd::gemm_bf16_dispatch(A, lda, W, ldw, C, ldc, M, N, K, epi, 0,
                                 reinterpret_cast<cudaStream_t>(stream));
}

// Test hook: same as dfd_gemm_bf16 with the tile width forced (128/192/256).
extern "C" DFD_API int dfd_gemm_bf16_tile(const void* A, int64_t lda, const void* W, int64_t ldw,
                                          void* C, int64_t ldc, int M, int N, int K,
                                          const dfd_gemm_epilogue* epi, int tile_n, void* stream) {
  return dfd::gemm_bf16_dispatch(A, lda, W, ldw, C, ldc, M, N, K, epi, tile_n,
                                 reinterpret_cast<cudaStream_t>(stream));
}
