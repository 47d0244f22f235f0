// Persistent non-causal attention on tcgen05 + TMEM (second generation of attention_tc.cu; same math, same TMEM
// layout, different schedule).
//
// Why a second kernel: ncu on attention_tc (profiles/r01_attention_full.md) showed the four softmax warps of a CTA
// (a) idle ~30 % of the time on the first S tile of every CTA (TMEM alloc + barrier init + Q/K TMA latency, paid
// once per 12 key tiles) and (b) latency bound in between with only two softmax warps per SM sub-partition.
//
//   * persistent CTAs (2 per SM): each loops over (image, head, 128-query tile) work items; the loader and the MMA
//     warp run ahead into the next item (Q', K'0, K'1, S'0, S'1) while the softmax warps finish the current one,
//   * KV/32 softmax warps per TMEM lane quadrant (the same 32 query rows) split each key tile by columns, 32 keys per
//     thread.  The row maximum is exchanged through shared memory with one named barrier per tile; all parts then hold
//     the same running max, keep partial row sums (added once per item) and own a share of the O columns for the
//     (rare) lazy rescale and for the output,
//   * two instantiations: KV = 64 (2 CTAs/SM, 256 TMEM columns each, 8 softmax warps) and KV = 128 (1 CTA/SM, all
//     512 columns, 16 softmax warps).  With everything but the barrier handshakes removed the KV = 64 kernel still
//     takes 80 % of its time: thirteen small MMAs per 64 keys (N = 64 and N = 16) keep the tensor pipe busy at ~45 %
//     efficiency (scripts/ubench_tc.cu: 62 cycles for a 32-cycle M128 N64 K16 SS-mode MMA, 26 for an 8-cycle N = 16
//     one).  KV = 128 halves the MMA count per key (S = Q·Kᵀ as N = 128 MMAs).
//
// Reference semantics: HF:modeling_siglip.py:229-249,293-306 (softmax(q·kᵀ/sqrt(hd)) v, fp32 softmax, no mask).
#include "dfd_common.cuh"

#include <atomic>
#include <mutex>

namespace dfd {

extern std::atomic<int64_t> g_launches;
int make_tmap_qkv_4d(CUtensorMap* out, const void* base, int hd, int heads3, int N, int B, int64_t ld, int box_cols,
                     int box_rows, int swizzle32);

namespace {

constexpr int kQ = 128;            // query rows per work item
constexpr int kStagesKV = 4;
// TMEM: S = two fp32 [128 x KV] buffers at columns 0 and KV (P, bf16 pairs, aliases the first KV/2 columns of its S
// buffer), O = fp32 [128 x 80] at column 2 KV

template <int HD, int KV>
struct WsSmem {
  static constexpr bool kTail = (HD % 64) != 0;
  static constexpr int kParts = KV / 32;                       // softmax warps per lane quadrant
  static constexpr int kThreads = (2 + 4 * kParts) * 32;       // loader, MMA, softmax warps
  static constexpr int kCtasPerSm = KV == 64 ? 2 : 1;
  static constexpr int kTmemCols = KV == 64 ? 256 : 512;
  static constexpr int kMainBytes = KV * 64 * 2;               // KV rows x 128 B, SWIZZLE_128B
  static constexpr int kTailBytes = kTail ? KV * 16 * 2 : 0;   // KV rows x 32 B, SWIZZLE_32B
  static constexpr int kQMain = kQ * 64 * 2, kQTail = kTail ? kQ * 16 * 2 : 0;
  static constexpr int kQBytes = kQMain + kQTail;
  static constexpr int kTileBytes = kMainBytes + kTailBytes;   // one K or V tile
  static constexpr int kXchgBytes = (2 * kParts * kQ + kParts * kQ) * 4;  // row max [parity][part][row], row sum
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = kQBytes + 2 * kStagesKV * kTileBytes + kXchgBytes + kBarBytes + 1024;
  static_assert(kTotal * kCtasPerSm <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
// named barrier 1 + quad over the kThreads softmax threads of one lane quadrant (immediate operands so that ptxas
// reserves 5 barriers, not all 16)
template <int kThreadsPerQuad>
__device__ __forceinline__ void quad_bar_sync(int quad) {
  switch (quad) {
    case 0: asm volatile("bar.sync 1, %0;\n" ::"n"(kThreadsPerQuad) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;\n" ::"n"(kThreadsPerQuad) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;\n" ::"n"(kThreadsPerQuad) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;\n" ::"n"(kThreadsPerQuad) : "memory"); break;
  }
}
__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }  // -> FMNMX3

template <int HD, int KV>
__global__ void __launch_bounds__(WsSmem<HD, KV>::kThreads, WsSmem<HD, KV>::kCtasPerSm)
attention_ws_kernel(const __grid_constant__ CUtensorMap tmMain, const __grid_constant__ CUtensorMap tmTail,
                    __nv_bfloat16* __restrict__ out, int64_t ldo, int N, int H, int n_items, float scale_log2) {
  using S = WsSmem<HD, KV>;
  constexpr bool kTail = S::kTail;
  constexpr int kKV = KV, kParts = S::kParts, kTmemCols = S::kTmemCols;
  constexpr int kColS = 0, kColO = 2 * KV;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + S::kQBytes;                                 // [stage]
  uint8_t* sV = sK + kStagesKV * S::kTileBytes;                    // [stage]
  float* s_max = reinterpret_cast<float*>(sV + kStagesKV * S::kTileBytes);  // [2][kParts][128]
  float* s_sum = s_max + 2 * kParts * kQ;                                    // [kParts][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sum + kParts * kQ);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + kStagesKV;
  uint64_t* s_full = kv_empty + kStagesKV;   // [2]
  uint64_t* p_full = s_full + 2;             // [2]
  uint64_t* o_done = p_full + 2;             // phase g completes when P_g·V_g has retired (g = running tile count)
  uint64_t* o_full = o_done + 1;             // phase i completes when the last P·V of this CTA's i-th item retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = (N + kKV - 1) / kKV;
  const int QT = (N + kQ - 1) / kQ;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmMain);
    if (kTail) tma_prefetch_desc(&tmTail);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
#pragma unroll
    for (int s = 0; s < kStagesKV; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 4 * kParts);
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
    // (elect.sync, not `lane == 0`: only then does ptxas know a single thread is active and emit the TMA / MMA /
    //  commit instructions bare; behind `lane == 0` each one is wrapped in an ELECT..BRA.U.ANY loop that costs
    //  ~100 cycles per issue — scripts/ubench_tc.cu)
    if (elect_one()) {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qt = item % QT, h = (item / QT) % H, b = item / (QT * H);
        mbar_wait(q_empty, (it & 1u) ^ 1u);  // the previous item's last Q·Kᵀ has retired
        mbar_expect_tx(q_full, S::kQBytes);
#pragma unroll
        for (int i = 0; i < kQ / 64; ++i) {  // 64-row TMA boxes
          tma_load_4d(&tmMain, q_full, sQ + i * 8192, 0, h, qt * kQ + i * 64, b);
          if (kTail) tma_load_4d(&tmTail, q_full, sQ + S::kQMain + i * 2048, 64, h, qt * kQ + i * 64, b);
        }
        for (int j = 0; j < T; ++j, ++g) {
          const uint32_t st = g % kStagesKV;
          mbar_wait(&kv_empty[st], ((g / kStagesKV) & 1u) ^ 1u);
          mbar_expect_tx(&kv_full[st], 2 * S::kTileBytes);
          uint8_t* k = sK + st * S::kTileBytes;
          uint8_t* v = sV + st * S::kTileBytes;
#pragma unroll
          for (int i = 0; i < kKV / 64; ++i) {  // 64-row TMA boxes
            tma_load_4d(&tmMain, &kv_full[st], k + i * 8192, 0, H + h, j * kKV + i * 64, b);
            tma_load_4d(&tmMain, &kv_full[st], v + i * 8192, 0, 2 * H + h, j * kKV + i * 64, b);
            if (kTail) {
              tma_load_4d(&tmTail, &kv_full[st], k + S::kMainBytes + i * 2048, 64, H + h, j * kKV + i * 64, b);
              tma_load_4d(&tmTail, &kv_full[st], v + S::kMainBytes + i * 2048, 64, 2 * H + h, j * kKV + i * 64, b);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (elect_one()) {
      const uint32_t tO = tmem_base + kColO;
      const uint64_t dQ = umma_desc(smem_u32(sQ), 16, 1024, 2);
      const uint64_t dQt = umma_desc(smem_u32(sQ + S::kQMain), 16, 256, 6);
      uint32_t g0 = 0, it = 0;  // g0 = running tile count at the start of the item
      // S_g = Q · K_jᵀ into S buffer g&1 (issued up to two tiles ahead of the softmax)
      auto issue_qk = [&](int j) {
        const uint32_t g = g0 + static_cast<uint32_t>(j);
        const uint32_t st = g % kStagesKV;
        const int n16 = (min(kKV, N - j * kKV) + 15) & ~15;
        mbar_wait(&kv_full[st], (g / kStagesKV) & 1u);
        tc_fence_after();
        const uint32_t tS = tmem_base + kColS + (g & 1u) * kKV;
        const uint32_t kaddr = smem_u32(sK + st * S::kTileBytes);
        const uint64_t dK = umma_desc(kaddr, 16, 1024, 2);
        const uint32_t idesc_qk = umma_idesc_bf16_major(kQ, n16, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, dQ + static_cast<uint64_t>(2 * k), dK + static_cast<uint64_t>(2 * k), idesc_qk, k != 0);
        if (kTail) {
          const uint64_t dKt = umma_desc(kaddr + S::kMainBytes, 16, 256, 6);
          umma_bf16_ss(tS, dQt, dKt, idesc_qk, 1u);
        }
        umma_commit(&s_full[g & 1u]);
        if (j == T - 1) umma_commit(q_empty);  // Q may be overwritten once every Q·Kᵀ of this item has retired
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it, g0 += T) {
        mbar_wait(q_full, it & 1u);
        tc_fence_after();
        issue_qk(0);
        if (T > 1) issue_qk(1);
        for (int j = 0; j < T; ++j) {
          const uint32_t g = g0 + static_cast<uint32_t>(j);
          const uint32_t st = g % kStagesKV;
          const int n16 = (min(kKV, N - j * kKV) + 15) & ~15;
          // ---- O (+)= P_g · V_j ----  (P written by the softmax warps into the S_g columns; the first P of an item
          // is only signalled after those warps have read the previous item's O out of TMEM)
          mbar_wait(&p_full[g & 1u], (g >> 1) & 1u);
          tc_fence_after();
          const uint32_t tP = tmem_base + kColS + (g & 1u) * kKV;
          const uint32_t vaddr = smem_u32(sV + st * S::kTileBytes);
          const uint64_t dV = umma_desc(vaddr, 16, 1024, 2);  // MN-major, 8-key groups 1024 B apart
          const uint64_t dVt = umma_desc(vaddr + S::kMainBytes, 16, 256, 6);
          constexpr uint32_t idesc_pv = umma_idesc_bf16_major(kQ, 64, 0, 1);
          constexpr uint32_t idesc_pvt = umma_idesc_bf16_major(kQ, 16, 0, 1);
          const int ksteps = n16 >> 4;
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint32_t acc = (j | kk) != 0 ? 1u : 0u;
            umma_bf16_ts(tO, tP + static_cast<uint32_t>(8 * kk), dV + static_cast<uint64_t>(kk * 128), idesc_pv, acc);
            if (kTail)
              umma_bf16_ts(tO + 64, tP + static_cast<uint32_t>(8 * kk), dVt + static_cast<uint64_t>(kk * 32),
                           idesc_pvt, acc);
          }
          umma_commit(&kv_empty[st]);  // K/V stage back to the loader once these MMAs retire
          umma_commit(o_done);
          if (j == T - 1) umma_commit(o_full);
          // the tensor pipe executes in issue order, so S_g / P_g are free for tile g+2 right after P_g·V_g
          if (j + 2 < T) issue_qk(j + 2);
        }
      }
    }
  } else {
    // ------------------------------------ softmax / output ------------------------------------
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int part = (warp - 2) >> 2;      // which 32 of the tile's keys / which share of the O columns
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + kColO;
    constexpr int kGroups = HD / 8;                      // 8-column groups of O that carry data
    constexpr int kGPer = (kGroups + kParts - 1) / kParts;  // groups [part kGPer, (part + 1) kGPer) belong to `part`
    const int gbeg = min(part * kGPer, kGroups), gend = min(gbeg + kGPer, kGroups);
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qt = item % QT, h = (item / QT) % H, b = item / (QT * H);
      const int grow = qt * kQ + row;
      float m = -INFINITY, l = 0.f;        // m in log2 units (already multiplied by scale_log2); l = partial row sum
      for (int j = 0; j < T; ++j, ++g) {
        const int valid = min(kKV, N - j * kKV) - 32 * part;   // valid keys among this part's 32 (may be <= 0)
        const uint32_t tS = tmem_base + lane_off + kColS + (g & 1u) * kKV;
        mbar_wait(&s_full[g & 1u], (g >> 1) & 1u);
        tc_fence_after();
        uint32_t s[32];
        tmem_ld_32x32b_x32(tS + 32 * part, s);
        tmem_ld_wait();
        if (valid < 32) {  // last tile: keys past the sequence end (zero-filled K rows) never win
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c >= valid) s[c] = __float_as_uint(-INFINITY);
        }
        float mx4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) mx4[c] = fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 4]));
#pragma unroll
        for (int c = 8; c < 32; c += 8)
#pragma unroll
          for (int i = 0; i < 4; ++i) mx4[i] = fmax3(mx4[i], __uint_as_float(s[c + i]), __uint_as_float(s[c + 4 + i]));
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        // exchange with the warps that hold the other keys of the same rows
        float* xm = s_max + (g & 1u) * kParts * kQ;
        xm[part * kQ + row] = mx;
        quad_bar_sync<32 * kParts>(quad);
#pragma unroll
        for (int q = 0; q < kParts; ++q) mx = fmaxf(mx, xm[q * kQ + row]);
        mx *= scale_log2;  // scale > 0
        // lazy rescale: keep the old reference max unless the new one is more than 2^8 larger
        const float m_new = (mx > m + 8.0f) ? mx : m;
        const bool moved = m_new != m;
        const float alpha = (j == 0) ? 0.f : fast_exp2(m - m_new);
        if (j > 0 && __any_sync(0xffffffffu, moved)) {
          mbar_wait(o_done, (g - 1) & 1u);  // P_{g-1}·V has retired: O is stable
          tc_fence_after();
          for (int c = gbeg; c < gend; ++c) {
            uint32_t o[8];
            tmem_ld_32x32b_x8(tO + static_cast<uint32_t>(8 * c), o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32b_x8(tO + static_cast<uint32_t>(8 * c), o);
          }
        }
        m = m_new;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        const float neg_m = -m;
        uint32_t p[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s[2 * c]), scale_log2, neg_m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s[2 * c + 1]), scale_log2, neg_m));
          sum4[(2 * c) & 3] += p0;
          sum4[(2 * c + 1) & 3] += p1;
          p[c] = pack_bf16x2(p0, p1);
        }
        l = l * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
        // every part has loaded its S columns (all passed the named barrier): P may overwrite the first KV/2 columns
        tmem_st_32x32b_x16(tS + 16 * part, p);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g & 1u]);
      }
      // ---- O / l -> bf16 -> global ----
      s_sum[part * kQ + row] = l;
      mbar_wait(o_full, it & 1u);
      tc_fence_after();
      quad_bar_sync<32 * kParts>(quad);
      float lsum = 0.f;
#pragma unroll
      for (int q = 0; q < kParts; ++q) lsum += s_sum[q * kQ + row];
      const float inv = 1.0f / lsum;
      __nv_bfloat16* orow = out + ((int64_t)b * N + grow) * ldo + h * HD;
#pragma unroll
      for (int c = 0; c < kGPer; ++c) {
        const int gi = gbeg + c;
        if (gi < gend) {
          uint32_t o[8];
          tmem_ld_32x32b_x8(tO + static_cast<uint32_t>(8 * gi), o);
          tmem_ld_wait();
          if (grow < N) {
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
            v.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
            v.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
            v.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
            *reinterpret_cast<uint4*>(orow + 8 * gi) = v;
          }
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

template <int HD, int KV>
static int launch_ws(const void* qkv, int64_t ldqkv, __nv_bfloat16* out, int64_t ldo, int B, int N, int H, int n_items,
                     float scale_log2, cudaStream_t st) {
  using S = WsSmem<HD, KV>;
  CUtensorMap tmMain, tmTail;
  int rc = make_tmap_qkv_4d(&tmMain, qkv, HD, 3 * H, N, B, ldqkv, 64, 64, 0);
  if (rc != DFD_OK) return rc;
  tmTail = tmMain;
  if (S::kTail) {
    rc = make_tmap_qkv_4d(&tmTail, qkv, HD, 3 * H, N, B, ldqkv, 16, 64, 1);
    if (rc != DFD_OK) return rc;
  }
  static SmemOptIn smem_once;
  if (int rc2 = ensure_dynamic_smem(smem_once, attention_ws_kernel<HD, KV>, S::kTotal)) return rc2;
  const int slots = S::kCtasPerSm * kNumSMs;
  const int grid = n_items < slots ? n_items : slots;
  attention_ws_kernel<HD, KV><<<grid, S::kThreads, S::kTotal, st>>>(tmMain, tmTail, out, ldo, N, H, n_items, scale_log2);
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

}  // namespace

// 64-key tiles (the 128-key instantiation of the template was a round-1 experiment: never faster, no longer compiled)
int attention_ws_bf16(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int B, int N, int H, int hd,
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
  const float scale_log2 = scale * 1.4426950408889634f;
  const int n_items = (int)items64;
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (hd == 64) return launch_ws<64, 64>(qkv, ldqkv, o, ldo, B, N, H, n_items, scale_log2, st);
  return launch_ws<72, 64>(qkv, ldqkv, o, ldo, B, N, H, n_items, scale_log2, st);
}

}  // namespace dfd
