// Non-causal attention on the 5th-gen tensor cores (tcgen05 + TMEM), for the 196..1024-token SigLIP sequences.
//
// One CTA = 128 query rows of one (image, head); two CTAs share an SM (256 TMEM columns and ~101 KB smem each).
// Keys are processed in tiles of 64; S is double buffered in TMEM so the MMA warp computes S_{j+1} (and S_{j+2})
// while the softmax warps work on S_j.
//
//   warp 0      TMA loader: Q once, then K/V tiles of 64 keys through a 4-stage mbarrier ring.  The operand is
//               addressed through a 4-D tensor map (head-dim, head, token, image) so that rows past the image's
//               last token are zero-filled by the TMA unit.  hd = 72 is covered by a 64-column box plus a 16-column
//               "tail" box over columns 56..71 — IN bounds: a tail box over columns 64..79 (zero-filled past 72) is
//               the obvious choice but boxes that cross the tensor edge are served far more slowly (handshake-only
//               pipeline 0.257 ms with them, 0.225 ms in bounds, 0.181 ms without any tail box).  The 8 columns the
//               two boxes share are cancelled on the Q side (the softmax warps zero columns 56..63 of the Q tail once
//               per CTA) and, for P·V, land in accumulator columns nobody reads.
//   warp 1      MMA issuer: S = Q·Kᵀ   (M=128, N<=64, K=64 via SWIZZLE_128B tiles + K=16 tail via SWIZZLE_32B tiles)
//                           O += P·V   (A = P read from TMEM, B = V tile as an MN-major operand; N=64 + N=16 tail)
//   warps 2..5  softmax: one thread per query row reads its S row from TMEM (no shuffles), keeps the running
//               max/sum in fp32, writes P (bf16) back into the S columns, rescales O in TMEM only when the running
//               max moved by more than 2^8 (exact: O and l always share the same reference max), and finally
//               normalises and stores O.
//
// Reference semantics: HF:modeling_siglip.py:229-249,293-306 (softmax(q·kᵀ/sqrt(hd)) v, fp32 softmax, no mask).
#include "dfd_common.cuh"

#include <atomic>
#include <mutex>

namespace dfd {

extern std::atomic<int64_t> g_launches;

namespace {

constexpr int kQ = 128;     // query rows per CTA
constexpr int kKV = 64;     // keys per tile
constexpr int kAttnThreads = 192;
constexpr int kStagesKV = 4;
constexpr int kTmemCols = 256;
constexpr int kColS = 0;    // S: two fp32 [128 x 64] buffers at columns 0 and 64; P_j (bf16 pairs) aliases the
                            // first 32 columns of S_j
constexpr int kColO = 128;  // O: fp32 [128 x 80]

template <int HD>
struct AttnSmem {
  static constexpr bool kTail = (HD % 64) != 0;
  static constexpr int kMainBytes = kKV * 64 * 2;              // 64 rows x 128 B, SWIZZLE_128B
  static constexpr int kTailBytes = kTail ? kKV * 16 * 2 : 0;  // 64 rows x 32 B, SWIZZLE_32B
  static constexpr int kQMain = 2 * kMainBytes, kQTail = 2 * kTailBytes;  // Q = two 64-row boxes per part
  static constexpr int kQBytes = kQMain + kQTail;
  static constexpr int kTileBytes = kMainBytes + kTailBytes;   // one K or V tile
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kQBytes + 2 * kStagesKV * kTileBytes + kBarBytes + 1024;
};

template <int HD>
__global__ void __launch_bounds__(kAttnThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmMain, const __grid_constant__ CUtensorMap tmTail,
                    __nv_bfloat16* __restrict__ out, int64_t ldo, int N, int H, float scale_log2) {
  using S = AttnSmem<HD>;
  constexpr bool kTail = S::kTail;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + S::kQBytes;                                 // [stage]
  uint8_t* sV = sK + kStagesKV * S::kTileBytes;                    // [stage]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStagesKV * S::kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = kv_full + kStagesKV;
  uint64_t* s_full = kv_empty + kStagesKV;   // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* o_done = p_full + 2;             // phase k completes when P_k·V_k has retired
  uint64_t* o_full = o_done + 1;             // completes once, when the last P·V has retired
  uint64_t* q_fixed = o_full + 1;            // the Q tail's shared columns have been zeroed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_fixed + 1);
  constexpr int kTailCol = HD - 16;          // first column of the tail boxes

  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = (N + kKV - 1) / kKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmMain);
    if (kTail) tma_prefetch_desc(&tmTail);
    mbar_init(q_full, 1);
#pragma unroll
    for (int s = 0; s < kStagesKV; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4);
    }
    mbar_init(o_done, 1);
    mbar_init(o_full, 1);
    mbar_init(q_fixed, 4);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------ TMA loader ------------------------------------
    if (elect_one()) {
      mbar_expect_tx(q_full, S::kQBytes);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        tma_load_4d(&tmMain, q_full, sQ + i * S::kMainBytes, 0, h, qt * kQ + i * kKV, b);
        if (kTail) tma_load_4d(&tmTail, q_full, sQ + S::kQMain + i * S::kTailBytes, kTailCol, h, qt * kQ + i * kKV, b);
      }
      for (int j = 0; j < T; ++j) {
        const int st = j % kStagesKV;
        mbar_wait(&kv_empty[st], ((j / kStagesKV) & 1u) ^ 1u);
        mbar_expect_tx(&kv_full[st], 2 * S::kTileBytes);
        uint8_t* k = sK + st * S::kTileBytes;
        uint8_t* v = sV + st * S::kTileBytes;
        tma_load_4d(&tmMain, &kv_full[st], k, 0, H + h, j * kKV, b);
        tma_load_4d(&tmMain, &kv_full[st], v, 0, 2 * H + h, j * kKV, b);
        if (kTail) {
          tma_load_4d(&tmTail, &kv_full[st], k + S::kMainBytes, kTailCol, H + h, j * kKV, b);
          tma_load_4d(&tmTail, &kv_full[st], v + S::kMainBytes, kTailCol, 2 * H + h, j * kKV, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    // One thread issues ~13 MMAs and 3-4 commits per key tile; the first version of this loop spent ~110 scalar /
    // uniform-datapath instructions per tile on descriptor arithmetic, dynamic stage indexing and a runtime k-step loop
    // and was ISSUE bound (ncu: the thread was busy ~1450 cycles per tile while the tensor pipe needs ~960).  Now the
    // key-tile loop is unrolled by the ring depth, so stage / S-buffer indices, descriptors and the full-tile shapes
    // are compile-time constants.
    if (elect_one()) {
      const uint32_t tO = tmem_base + kColO;
      const uint64_t dQ = umma_desc(smem_u32(sQ), 16, 1024, 2);
      const uint64_t dQt = umma_desc(smem_u32(sQ + S::kQMain), 16, 256, 6);
      const uint64_t dK0 = umma_desc(smem_u32(sK), 16, 1024, 2);
      const uint64_t dKt0 = umma_desc(smem_u32(sK + S::kMainBytes), 16, 256, 6);
      const uint64_t dV0 = umma_desc(smem_u32(sV), 16, 1024, 2);  // MN-major, 8-key groups 1024 B apart
      const uint64_t dVt0 = umma_desc(smem_u32(sV + S::kMainBytes), 16, 256, 6);
      constexpr uint64_t kStageStep = S::kTileBytes >> 4;        // descriptor start addresses count 16-byte units
      constexpr uint32_t idesc_qk_full = umma_idesc_bf16_major(kQ, kKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_major(kQ, 64, 0, 1);
      constexpr uint32_t idesc_pvt = umma_idesc_bf16_major(kQ, 16, 0, 1);
      // S_j = Q · K_j^T into S buffer sb (runs up to two tiles ahead of the softmax); st, sb are constants after unrolling
      auto issue_qk = [&](int j, int st, int sb) {
        mbar_wait(&kv_full[st], (j / kStagesKV) & 1u);
        tc_fence_after();
        const uint32_t tS = tmem_base + kColS + static_cast<uint32_t>(sb * kKV);
        const uint64_t dK = dK0 + st * kStageStep;
        const int valid = N - j * kKV;
        const uint32_t idesc_qk = valid >= kKV ? idesc_qk_full : umma_idesc_bf16_major(kQ, (valid + 15) & ~15, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, dQ + static_cast<uint64_t>(2 * k), dK + static_cast<uint64_t>(2 * k), idesc_qk, k != 0);
        if (kTail) {
          if (j == 0) {  // the four main k-steps above do not touch the Q tail; this one needs its zeroed columns
            mbar_wait(q_fixed, 0);
            tc_fence_after();
          }
          umma_bf16_ss(tS, dQt, dKt0 + st * kStageStep, idesc_qk, 1u);
        }
        umma_commit(&s_full[sb]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0, 0, 0);
      if (T > 1) issue_qk(1, 1, 1);
      for (int j4 = 0; j4 < T; j4 += kStagesKV) {
#pragma unroll
        for (int u = 0; u < kStagesKV; ++u) {
          const int j = j4 + u;
          if (j < T) {
            // ---- O += P_j · V_j ----  (P_j written by the softmax warps into the S_j columns)
            mbar_wait(&p_full[u & 1], (j >> 1) & 1u);
            tc_fence_after();
            const uint32_t tP = tmem_base + kColS + static_cast<uint32_t>((u & 1) * kKV);
            const uint64_t dV = dV0 + u * kStageStep, dVt = dVt0 + u * kStageStep;
            const int valid = N - j * kKV;
            if (valid >= kKV) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                // 16 keys per step: 16 rows x 128 B (main) / 16 rows x 32 B (tail); P: 8 packed columns per step
                const uint32_t acc = (kk != 0) ? 1u : (j != 0 ? 1u : 0u);
                umma_bf16_ts(tO, tP + static_cast<uint32_t>(8 * kk), dV + static_cast<uint64_t>(kk * 128), idesc_pv, acc);
                if (kTail)
                  umma_bf16_ts(tO + 64, tP + static_cast<uint32_t>(8 * kk), dVt + static_cast<uint64_t>(kk * 32),
                               idesc_pvt, acc);
              }
            } else {
              const int ksteps = (valid + 15) >> 4;
              for (int kk = 0; kk < ksteps; ++kk) {
                const uint32_t acc = (j | kk) != 0 ? 1u : 0u;
                umma_bf16_ts(tO, tP + static_cast<uint32_t>(8 * kk), dV + static_cast<uint64_t>(kk * 128), idesc_pv, acc);
                if (kTail)
                  umma_bf16_ts(tO + 64, tP + static_cast<uint32_t>(8 * kk), dVt + static_cast<uint64_t>(kk * 32),
                               idesc_pvt, acc);
              }
            }
            umma_commit(&kv_empty[u]);  // K/V stage back to the loader once these MMAs retire
            umma_commit(o_done);
            // the tensor pipe executes in issue order, so S_{j&1} / P_j are free for tile j+2 right after P_j·V_j
            if (j + 2 < T) issue_qk(j + 2, (u + 2) % kStagesKV, u & 1);
          }
        }
      }
      umma_commit(o_full);
    }
  } else {
    // ------------------------------------ softmax / output ------------------------------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int grow = qt * kQ + row;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + kColO;
    constexpr int kOChunks = (HD + 15) / 16;  // 16-column chunks of O
    float m = -INFINITY, l = 0.f;             // m is kept in log2 units (already multiplied by scale_log2)
    if (kTail) {
      // Q tail = columns 56..71; 56..63 are already covered by the main box: zero them (SWIZZLE_32B: the 16-byte chunk
      // index is XORed with bit 2 of the row), then hand the tile to the async proxy
      mbar_wait(q_full, 0);
      sts_zero16(smem_u32(sQ + S::kQMain) + static_cast<uint32_t>(row * 32 + (((row >> 2) & 1) << 4)));
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_fixed);
    }
    for (int j = 0; j < T; ++j) {
      const int valid = min(kKV, N - j * kKV);
      const uint32_t tS = tmem_base + lane_off + kColS + static_cast<uint32_t>((j & 1) * kKV);
      mbar_wait(&s_full[j & 1], (j >> 1) & 1u);
      tc_fence_after();
      uint32_t s[64];
      tmem_ld_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_32x32b_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      if (valid < kKV) {  // last tile: keys past the sequence end (zero-filled K rows) never win
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= valid) s[c] = __float_as_uint(-INFINITY);
      }
      // 8 independent chains (a single running max would be a 64-deep dependent chain)
      float mx8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mx8[c] = __uint_as_float(s[c]);
#pragma unroll
      for (int c = 8; c < 64; ++c) mx8[c & 7] = fmaxf(mx8[c & 7], __uint_as_float(s[c]));
      float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                       fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      mx *= scale_log2;  // scale > 0
      // lazy rescale: keep the old reference max unless the new one is more than 2^8 larger
      const float m_new = (mx > m + 8.0f) ? mx : m;
      const bool moved = m_new != m;
      const float alpha = (j == 0) ? 0.f : fast_exp2(m - m_new);
      if (j > 0 && __any_sync(0xffffffffu, moved)) {
        mbar_wait(o_done, (j - 1) & 1u);  // P_{j-1}·V_{j-1} has retired: O is stable
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < kOChunks; ++c) {
          uint32_t o[16];
          tmem_ld_32x32b_x16(tO + static_cast<uint32_t>(16 * c), o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32b_x16(tO + static_cast<uint32_t>(16 * c), o);
        }
      }
      m = m_new;
      float sum8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const float neg_m = -m;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(s[2 * c]), scale_log2, neg_m));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s[2 * c + 1]), scale_log2, neg_m));
        sum8[(2 * c) & 7] += p0;
        sum8[(2 * c + 1) & 7] += p1;
        s[c] = pack_bf16x2(p0, p1);  // in place: s[2c], s[2c+1] (indices >= c) are consumed first
      }
      l = l * alpha + (((sum8[0] + sum8[1]) + (sum8[2] + sum8[3])) + ((sum8[4] + sum8[5]) + (sum8[6] + sum8[7])));
      tmem_st_32x32b_x32(tS, *reinterpret_cast<const uint32_t(*)[32]>(&s[0]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[j & 1]);
    }
    // ---- O / l -> bf16 -> global ----
    // (o_done cannot be used here: with S running a tile ahead, its previous same-parity phase may already
    //  satisfy the wait before the last two P·V products have retired)
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.0f / l;
    __nv_bfloat16* orow = out + ((int64_t)b * N + grow) * ldo + h * HD;
#pragma unroll
    for (int c = 0; c < kOChunks; ++c) {
      uint32_t o[16];
      tmem_ld_32x32b_x16(tO + static_cast<uint32_t>(16 * c), o);
      tmem_ld_wait();
      if (kTail && c == kOChunks - 1) {  // tail accumulator = columns 56..71: its upper half is columns 64..71
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = o[8 + i];
      }
      if (grow < N) {
        uint4 lo, hi;
        lo.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
        lo.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
        lo.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
        lo.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
        hi.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
        hi.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
        hi.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
        hi.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
        *reinterpret_cast<uint4*>(orow + 16 * c) = lo;
        if (16 * c + 8 < HD) *reinterpret_cast<uint4*>(orow + 16 * c + 8) = hi;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// qkv [B*N, ld] viewed as (hd, 3H, N, B); box = box_cols x 1 x 128 x 1
int make_tmap_qkv(CUtensorMap* out, const void* base, int hd, int heads3, int N, int B, int64_t ld, int box_cols,
                  CUtensorMapSwizzle sw, int box_rows = kKV) {
  PFN_encodeTiled fn = encode_fn();
  DFD_REQUIRE(fn != nullptr, DFD_ERR_NO_DEVICE, "cuTensorMapEncodeTiled unavailable (no CUDA driver on this host)");
  cuuint64_t gdim[4] = {(cuuint64_t)hd, (cuuint64_t)heads3, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)hd * 2, (cuuint64_t)ld * 2, (cuuint64_t)N * (cuuint64_t)ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_cols, 1, (cuuint32_t)box_rows, 1};  // box_rows token rows per box
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DFD_REQUIRE(r == CUDA_SUCCESS, DFD_ERR_CUDA, "cuTensorMapEncodeTiled(qkv 4-D) failed (%d)", (int)r);
  return DFD_OK;
}

}  // namespace

// shared with attention_ws.cu: qkv [B*N, ld] viewed as (hd, 3H, N, B), box = box_cols x 1 x box_rows x 1
int make_tmap_qkv_4d(CUtensorMap* out, const void* base, int hd, int heads3, int N, int B, int64_t ld, int box_cols,
                     int box_rows, int swizzle32) {
  return make_tmap_qkv(out, base, hd, heads3, N, B, ld, box_cols,
                       swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, box_rows);
}

int attention_tc_bf16(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int B, int N, int H, int hd,
                      float scale, cudaStream_t st) {
  DFD_REQUIRE(qkv && out, DFD_ERR_BAD_ARG, "attention: null pointer");
  DFD_REQUIRE(B > 0 && N > 0 && H > 0, DFD_ERR_SHAPE, "attention: B, N, H must be positive");
  DFD_REQUIRE(hd == 64 || hd == 72, DFD_ERR_UNSUPPORTED, "attention: head dim %d not supported (64, 72)", hd);
  DFD_REQUIRE(ldqkv % 8 == 0 && ldqkv >= 3 * H * hd && ldo % 8 == 0 && ldo >= H * hd, DFD_ERR_SHAPE,
              "attention: bad leading dimensions");
  DFD_REQUIRE(B <= 65535 && H <= 65535, DFD_ERR_SHAPE, "attention: B and H must be <= 65535");
  DFD_REQUIRE(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), DFD_ERR_BAD_ARG,
              "attention: pointers must be 16-byte aligned");
  CUtensorMap tmMain, tmTail;
  int rc = make_tmap_qkv(&tmMain, qkv, hd, 3 * H, N, B, ldqkv, 64, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != DFD_OK) return rc;
  tmTail = tmMain;
  if (hd == 72) {
    rc = make_tmap_qkv(&tmTail, qkv, hd, 3 * H, N, B, ldqkv, 16, CU_TENSOR_MAP_SWIZZLE_32B);
    if (rc != DFD_OK) return rc;
  }
  const float scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((N + kQ - 1) / kQ, H, B);
  static SmemOptIn smem_once[2];
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (hd == 64) {
    if (int rc2 = ensure_dynamic_smem(smem_once[0], attention_tc_kernel<64>, AttnSmem<64>::kTotal)) return rc2;
    attention_tc_kernel<64><<<grid, kAttnThreads, AttnSmem<64>::kTotal, st>>>(tmMain, tmTail, o, ldo, N, H, scale_log2);
  } else {
    if (int rc2 = ensure_dynamic_smem(smem_once[1], attention_tc_kernel<72>, AttnSmem<72>::kTotal)) return rc2;
    attention_tc_kernel<72><<<grid, kAttnThreads, AttnSmem<72>::kTotal, st>>>(tmMain, tmTail, o, ldo, N, H, scale_log2);
  }
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

// Product dispatch (measured on B200, scripts/kbench.py attn): short sequences (base-224: 196 tokens, 2 query tiles
// and 4 key tiles per head) gain 4-5 % from the persistent kernel, which hides the per-CTA prologue; at 729 tokens the
// per-tile kernel is 3 % ahead.
int attention_auto_bf16(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int B, int N, int H, int hd,
                        float scale, cudaStream_t st) {
  if (N <= 256) return attention_ws_bf16(qkv, ldqkv, out, ldo, B, N, H, hd, scale, 64, st);
  // attention_pq.cu (persistent, Q in TMEM) is 3-4 % faster timed alone at N = 729 but not inside the power-capped
  // detect step (same-box A/B: 1465 / 1449 img/s with this kernel, 1432 / 1443 with pq) - see profiles/r01_attention_full.md
  return attention_tc_bf16(qkv, ldqkv, out, ldo, B, N, H, hd, scale, st);
}

}  // namespace dfd

// Test / A-B hook: impl 0 = warp-level mma.sync kernel (attention.cu), 1 = tcgen05 kernel (this file),
// 2 / 3 = persistent tcgen05 kernel (attention_ws.cu) with 64- / 128-key tiles, 4 = persistent with Q in TMEM
// (attention_pq.cu).
extern "C" DFD_API int dfd_attention_bf16_impl(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int B, int N,
                                               int H, int hd, float scale, int impl, void* stream) {
  if (impl == 5)
    return dfd::attention_dq_bf16(qkv, ldqkv, out, ldo, B, N, H, hd, scale, reinterpret_cast<cudaStream_t>(stream));
  if (impl == 4)
    return dfd::attention_pq_bf16(qkv, ldqkv, out, ldo, B, N, H, hd, scale, reinterpret_cast<cudaStream_t>(stream));
  if (impl == 2 || impl == 3)
    return dfd::attention_ws_bf16(qkv, ldqkv, out, ldo, B, N, H, hd, scale, impl == 2 ? 64 : 128,
                                  reinterpret_cast<cudaStream_t>(stream));
  if (impl == 1)
    return dfd::attention_tc_bf16(qkv, ldqkv, out, ldo, B, N, H, hd, scale, reinterpret_cast<cudaStream_t>(stream));
  return dfd::attention_bf16(qkv, ldqkv, out, ldo, B, N, H, hd, scale, reinterpret_cast<cudaStream_t>(stream));
}
