// Non-causal attention on tcgen05 + TMEM: persistent CTAs with Q held in TMEM (fourth schedule of the round).
//
// What the earlier kernels taught (profiles/r01_attention_full.md, scripts/ubench_tc.cu):
//   * S = Q·Kᵀ in SS mode re-reads the 128-row Q operand from shared memory for every key tile: an M128 N64 K16 MMA
//     then takes 62-75 cycles against a 32-cycle floor.  With A = Q in TMEM it takes 38-42.
//   * writing Q to TMEM in the prologue of a per-tile CTA exposes the load (-6 %); a per-tile CTA also pays TMEM
//     allocation, barrier initialisation and the first K/V round trip once per 12 key tiles (~10 % of its softmax
//     warps' time) and cannot overlap its output phase with anything.
// Here: 2 persistent CTAs per SM loop over (image, head, 128-query tile) items.  The loader prefetches the next
// item's Q into shared memory while the current item runs; the four softmax warps copy it smem -> registers -> TMEM
// right after their last P of the current item, BEFORE they write the current item's output, so the MMA warp can
// already compute S'0, S'1 of the next item during that output phase.  One thread per query row (no cross-warp
// exchange), 64-key tiles, S double buffered, P aliases S, O accumulates in TMEM (as attention_tc.cu).
// The key-tile loop of the MMA thread is unrolled by the ring depth; ring stages and S buffers restart at 0 with every
// item and carry per-stage use counters, so the barrier parities stay right for any number of tiles per item.
//
// Reference semantics: HF:modeling_siglip.py:229-249,293-306 (softmax(q·kᵀ/sqrt(hd)) v, fp32 softmax, no mask).
#include "dfd_common.cuh"

#include <atomic>

namespace dfd {

extern std::atomic<int64_t> g_launches;
int make_tmap_qkv_4d(CUtensorMap* out, const void* base, int hd, int heads3, int N, int B, int64_t ld, int box_cols,
                     int box_rows, int swizzle32);

namespace {

constexpr int kQ = 128;            // query rows per item
constexpr int kKV = 64;            // keys per tile
constexpr int kThreads = 192;      // loader, MMA, 4 softmax warps
constexpr int kStagesKV = 4;
constexpr int kTmemCols = 256;
constexpr int kColS = 0;           // S: two fp32 [128 x 64] buffers (columns 0 and 64); P (bf16 pairs) aliases the first 32
constexpr int kColO = 128;         // O: fp32 [128 x 80]
constexpr int kColQ = 208;         // Q: bf16 pairs [128 x 40] (hd padded to 80 with zeros)

template <int HD>
struct PqSmem {
  static constexpr bool kTail = (HD % 64) != 0;
  static constexpr int kMainBytes = kKV * 64 * 2;              // 64 rows x 128 B, SWIZZLE_128B
  static constexpr int kTailBytes = kTail ? kKV * 16 * 2 : 0;  // 64 rows x 32 B, SWIZZLE_32B
  static constexpr int kQMain = 2 * kMainBytes, kQTail = 2 * kTailBytes;
  static constexpr int kQBytes = kQMain + kQTail;
  static constexpr int kTileBytes = kMainBytes + kTailBytes;   // one K or V tile
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kQBytes + 2 * kStagesKV * kTileBytes + kBarBytes + 1024;
};

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

template <int HD>
__global__ void __launch_bounds__(kThreads, 2)
attention_pq_kernel(const __grid_constant__ CUtensorMap tmMain, const __grid_constant__ CUtensorMap tmTail,
                    __nv_bfloat16* __restrict__ out, int64_t ldo, int N, int H, int n_items, float scale_log2) {
  using S = PqSmem<HD>;
  constexpr bool kTail = S::kTail;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + S::kQBytes;                                 // [stage]
  uint8_t* sV = sK + kStagesKV * S::kTileBytes;                    // [stage]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kStagesKV * S::kTileBytes);
  uint64_t* q_full = bars;                   // TMA: the item's Q tile is in shared memory
  uint64_t* q_copied = bars + 1;             // 4 softmax warps: Q is in TMEM (and the shared-memory tile is free again)
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + kStagesKV;
  uint64_t* s_full = kv_empty + kStagesKV;   // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* o_done = p_full + 2;             // phase g completes when the g-th P·V of this CTA has retired
  uint64_t* o_full = o_done + 1;             // phase i completes when the last P·V of this CTA's i-th item has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = (N + kKV - 1) / kKV;
  const int QT = (N + kQ - 1) / kQ;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmMain);
    if (kTail) tma_prefetch_desc(&tmTail);
    mbar_init(q_full, 1);
    mbar_init(q_copied, 4);
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
      uint32_t ku[kStagesKV] = {0, 0, 0, 0};   // uses of each ring stage so far
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qt = item % QT, h = (item / QT) % H, b = item / (QT * H);
        mbar_wait(q_copied, (it & 1u) ^ 1u);  // the previous item's Q has been copied out of shared memory
        mbar_expect_tx(q_full, S::kQBytes);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          tma_load_4d(&tmMain, q_full, sQ + i * S::kMainBytes, 0, h, qt * kQ + i * kKV, b);
          if (kTail) tma_load_4d(&tmTail, q_full, sQ + S::kQMain + i * S::kTailBytes, HD - 16, h, qt * kQ + i * kKV, b);
        }
        for (int j = 0; j < T; ++j) {
          const int st = j % kStagesKV;
          const uint32_t n = ku[st]++;
          mbar_wait(&kv_empty[st], (n & 1u) ^ 1u);
          mbar_expect_tx(&kv_full[st], 2 * S::kTileBytes);
          uint8_t* k = sK + st * S::kTileBytes;
          uint8_t* v = sV + st * S::kTileBytes;
          tma_load_4d(&tmMain, &kv_full[st], k, 0, H + h, j * kKV, b);
          tma_load_4d(&tmMain, &kv_full[st], v, 0, 2 * H + h, j * kKV, b);
          if (kTail) {
            // tail boxes = columns 56..71, in bounds (boxes crossing the tensor edge are served slowly, see attention_tc.cu)
            tma_load_4d(&tmTail, &kv_full[st], k + S::kMainBytes, HD - 16, H + h, j * kKV, b);
            tma_load_4d(&tmTail, &kv_full[st], v + S::kMainBytes, HD - 16, 2 * H + h, j * kKV, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (elect_one()) {
      const uint32_t tO = tmem_base + kColO, tQ = tmem_base + kColQ;
      const uint64_t dK0 = umma_desc(smem_u32(sK), 16, 1024, 2);
      const uint64_t dKt0 = umma_desc(smem_u32(sK + S::kMainBytes), 16, 256, 6);
      const uint64_t dV0 = umma_desc(smem_u32(sV), 16, 1024, 2);  // MN-major, 8-key groups 1024 B apart
      const uint64_t dVt0 = umma_desc(smem_u32(sV + S::kMainBytes), 16, 256, 6);
      constexpr uint64_t kStageStep = S::kTileBytes >> 4;
      constexpr uint32_t idesc_qk_full = umma_idesc_bf16_major(kQ, kKV, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16_major(kQ, 64, 0, 1);
      constexpr uint32_t idesc_pvt = umma_idesc_bf16_major(kQ, 16, 0, 1);
      uint32_t ku[kStagesKV] = {0, 0, 0, 0};   // kv_full uses per stage
      uint32_t pu[2] = {0, 0};                 // p_full uses per S buffer
      // S = Q · K_j^T into S buffer sb; st, sb are constants after unrolling.  A = Q from TMEM.
      auto issue_qk = [&](int j, int st, int sb) {
        mbar_wait(&kv_full[st], ku[st]++ & 1u);
        tc_fence_after();
        const uint32_t tS = tmem_base + kColS + static_cast<uint32_t>(sb * kKV);
        const uint64_t dK = dK0 + st * kStageStep;
        const int valid = N - j * kKV;
        const uint32_t idesc_qk = valid >= kKV ? idesc_qk_full : umma_idesc_bf16_major(kQ, (valid + 15) & ~15, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ts(tS, tQ + static_cast<uint32_t>(8 * k), dK + static_cast<uint64_t>(2 * k), idesc_qk, k != 0);
        if (kTail) umma_bf16_ts(tS, tQ + 32, dKt0 + st * kStageStep, idesc_qk, 1u);
        umma_commit(&s_full[sb]);
      };
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        mbar_wait(q_copied, it & 1u);
        tc_fence_after();
        issue_qk(0, 0, 0);
        if (T > 1) issue_qk(1, 1, 1);
        for (int j4 = 0; j4 < T; j4 += kStagesKV) {
#pragma unroll
          for (int u = 0; u < kStagesKV; ++u) {
            const int j = j4 + u;
            if (j < T) {
              // ---- O (+)= P_j · V_j ----
              mbar_wait(&p_full[u & 1], pu[u & 1]++ & 1u);
              tc_fence_after();
              const uint32_t tP = tmem_base + kColS + static_cast<uint32_t>((u & 1) * kKV);
              const uint64_t dV = dV0 + u * kStageStep, dVt = dVt0 + u * kStageStep;
              const int valid = N - j * kKV;
              if (valid >= kKV) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
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
              if (j == T - 1) umma_commit(o_full);
              // the tensor pipe executes in issue order, so the S buffer / P_j are free for tile j+2 right after P_j·V_j
              if (j + 2 < T) issue_qk(j + 2, (u + 2) % kStagesKV, u & 1);
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------ softmax / output ------------------------------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + kColO;
    const uint32_t tQ = tmem_base + lane_off + kColQ;
    constexpr int kOChunks = (HD + 15) / 16;  // 16-column chunks of O
    const uint32_t q_row = smem_u32(sQ) + static_cast<uint32_t>(row * 128);
    const uint32_t q_tail = smem_u32(sQ + S::kQMain) + static_cast<uint32_t>(row * 32);
    // Q row of item `it_q`: shared memory (TMA, swizzled) -> registers -> TMEM as bf16 pairs
    auto copy_q = [&](uint32_t it_q) {
      mbar_wait(q_full, it_q & 1u);
      uint32_t qw[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = lds_u4(q_row + static_cast<uint32_t>((c ^ (row & 7)) << 4));
        qw[4 * c] = v.x; qw[4 * c + 1] = v.y; qw[4 * c + 2] = v.z; qw[4 * c + 3] = v.w;
      }
      tmem_st_32x32b_x32(tQ, qw);
      if (kTail) {
        // tail k-step = columns 56..71; 56..63 already went through the main box: Q contributes zeros there
        uint32_t t8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        const uint4 v = lds_u4(q_tail + static_cast<uint32_t>((1 ^ ((row >> 2) & 1)) << 4));
        t8[4] = v.x; t8[5] = v.y; t8[6] = v.z; t8[7] = v.w;
        tmem_st_32x32b_x8(tQ + 32, t8);
      }
      tmem_st_wait();
      // generic-proxy reads of the Q tile precede its TMA refill: proxy fence before the release (see gemm_tcgen05.cu)
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_copied);
    };
    uint32_t su[2] = {0, 0};   // s_full uses per S buffer
    uint32_t g = 0, it = 0;    // running tile count (o_done phases), item count
    if (blockIdx.x < n_items) copy_q(0);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % QT, h = (item / QT) % H, b = item / (QT * H);
      const int grow = qt * kQ + row;
      float m = -INFINITY, l = 0.f;             // m is kept in log2 units (already multiplied by scale_log2)
      for (int j = 0; j < T; ++j, ++g) {
        const int sb = j & 1;
        const int valid = min(kKV, N - j * kKV);
        const uint32_t tS = tmem_base + lane_off + kColS + static_cast<uint32_t>(sb * kKV);
        mbar_wait(&s_full[sb], su[sb]++ & 1u);
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
          mbar_wait(o_done, (g - 1) & 1u);  // the previous P·V has retired: O is stable
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
        if (lane == 0) mbar_arrive(&p_full[sb]);
      }
      // every Q·Kᵀ of this item has retired (its last S tile has been consumed): the next item's Q may replace it in
      // TMEM now, so that the MMA warp can start S'0 / S'1 while this item's output is written
      if (item + (int)gridDim.x < n_items) copy_q(it + 1);
      // ---- O / l -> bf16 -> global ----
      mbar_wait(o_full, it & 1u);
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
      // order this item's TMEM reads before the p_full arrive that lets the next item's first P·V overwrite O
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace

int attention_pq_bf16(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int B, int N, int H, int hd,
                      float scale, cudaStream_t st) {
  DFD_REQUIRE(qkv && out, DFD_ERR_BAD_ARG, "attention: null pointer");
  DFD_REQUIRE(B > 0 && N > 0 && H > 0, DFD_ERR_SHAPE, "attention: B, N, H must be positive");
  DFD_REQUIRE(hd == 64 || hd == 72, DFD_ERR_UNSUPPORTED, "attention: head dim %d not supported (64, 72)", hd);
  DFD_REQUIRE(ldqkv % 8 == 0 && ldqkv >= 3 * H * hd && ldo % 8 == 0 && ldo >= H * hd, DFD_ERR_SHAPE,
              "attention: bad leading dimensions");
  DFD_REQUIRE(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), DFD_ERR_BAD_ARG,
              "attention: pointers must be 16-byte aligned");
  const int64_t items64 = (int64_t)((N + kQ - 1) / kQ) * H * B;
  DFD_REQUIRE(items64 < (1ll << 31), DFD_ERR_SHAPE, "attention: too many work items");
  CUtensorMap tmMain, tmTail;
  int rc = make_tmap_qkv_4d(&tmMain, qkv, hd, 3 * H, N, B, ldqkv, 64, kKV, 0);
  if (rc != DFD_OK) return rc;
  tmTail = tmMain;
  if (hd == 72) {
    rc = make_tmap_qkv_4d(&tmTail, qkv, hd, 3 * H, N, B, ldqkv, 16, kKV, 1);
    if (rc != DFD_OK) return rc;
  }
  const float scale_log2 = scale * 1.4426950408889634f;
  const int n_items = (int)items64;
  const int grid = n_items < 2 * kNumSMs ? n_items : 2 * kNumSMs;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  static SmemOptIn smem_once[2];
  if (hd == 64) {
    if (int rc2 = ensure_dynamic_smem(smem_once[0], attention_pq_kernel<64>, PqSmem<64>::kTotal)) return rc2;
    attention_pq_kernel<64><<<grid, kThreads, PqSmem<64>::kTotal, st>>>(tmMain, tmTail, o, ldo, N, H, n_items, scale_log2);
  } else {
    if (int rc2 = ensure_dynamic_smem(smem_once[1], attention_pq_kernel<72>, PqSmem<72>::kTotal)) return rc2;
    attention_pq_kernel<72><<<grid, kThreads, PqSmem<72>::kTotal, st>>>(tmMain, tmTail, o, ldo, N, H, n_items, scale_log2);
  }
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

}  // namespace dfd
