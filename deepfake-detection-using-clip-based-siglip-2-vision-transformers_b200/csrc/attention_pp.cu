// Non-causal attention on tcgen05 + TMEM, "ping-pong" schedule (third generation; product path).
//
// What the profiles of the first two kernels showed (profiles/r01_attention_*.md): with 64-key tiles and one
// softmax iteration per tile, every warp pays ~1200 cycles of fixed latency per tile (barrier waits, TMEM load,
// max, TMEM store, arrive) next to ~500 cycles of MUFU-bound exponentials, so the MUFU pipe — the real floor for
// head dims 64/72, 16 ex2/clk/SM — sat at ~55 %.  This kernel amortises the fixed part over four times more
// elements per thread and lets two query tiles hide each other's latencies:
//
//   * one persistent CTA per SM, all 512 TMEM columns: two SLOTS, each a 128-query tile of the same (image, head)
//     with its own S [128 x 128 fp32] and O [128 x 80 fp32]; both slots share the K/V tiles in shared memory,
//   * 128-key tiles: S = Q·Kᵀ is M128 x N128 x K16 MMAs (5 per tile), O += P·V is 8 k-steps (N=64 + N=16 tail),
//   * eight softmax warps per slot (warps 2..9 slot 0, 10..17 slot 1): warps w and w+4 share a TMEM lane quadrant
//     (the same 32 query rows) and split the 128-key tile by columns, 64 keys per thread held in registers; the row
//     maximum is exchanged through shared memory with ONE 64-thread named barrier per 128 keys.  (One thread per whole
//     row — 4 warps per slot — was tried first: with two softmax warps per SM sub-partition the per-warp IPC of ~0.3
//     made the softmax phase 3x its MUFU floor.)  P (bf16 pairs) overwrites the first 64 S columns,
//   * the MMA warp alternates slots:  P0·V, Q0·Kᵀ(next) | P1·V, Q1·Kᵀ(next) | ...  so that while slot 0's softmax
//     runs, slot 1's MMAs run and vice versa; within a slot the tensor pipe's in-order execution makes P·V(j) finish
//     before Q·Kᵀ(j+1) overwrites the aliased columns, and S(j+1) complete implies P·V(j) retired (O stable for the
//     lazy rescale),
//   * loader / MMA warps run ahead into the next work item while the softmax warps write the current one out.
//
// Reference semantics: HF:modeling_siglip.py:229-249,293-306 (softmax(q·kᵀ/sqrt(hd)) v, fp32 softmax, no mask).
#include "dfd_common.cuh"

#include <atomic>

namespace dfd {

extern std::atomic<int64_t> g_launches;
int make_tmap_qkv_4d(CUtensorMap* out, const void* base, int hd, int heads3, int N, int B, int64_t ld, int box_cols,
                     int box_rows, int swizzle32);

namespace {

constexpr int kQ = 128;            // query rows per slot
constexpr int kKV = 128;           // keys per tile
constexpr int kThreads = 576;      // loader, MMA, 2 x 8 softmax warps
constexpr int kStagesKV = 4;
constexpr int kTmemCols = 512;
constexpr int kSlotCols = 256;     // slot s: S at column 256 s (128 wide; P aliases its first 64), O at 256 s + 128

template <int HD>
struct PpSmem {
  static constexpr bool kTail = (HD % 64) != 0;
  static constexpr int kMainBytes = 128 * 64 * 2;              // 128 rows x 128 B, SWIZZLE_128B
  static constexpr int kTailBytes = kTail ? 128 * 16 * 2 : 0;  // 128 rows x 32 B, SWIZZLE_32B
  static constexpr int kTileBytes = kMainBytes + kTailBytes;   // one Q, K or V tile (128 rows)
  static constexpr int kXchgBytes = (2 * 2 * 2 * kQ + 2 * 2 * kQ) * 4;  // row max [slot][parity][half][row], row sum
  static constexpr int kBarBytes = 256;
  static constexpr int kTotal = 2 * kTileBytes + 2 * kStagesKV * kTileBytes + kXchgBytes + kBarBytes + 1024;
  static_assert(kTotal <= 227 * 1024, "shared memory budget");
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
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

template <int HD>
__global__ void __launch_bounds__(kThreads, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmMain, const __grid_constant__ CUtensorMap tmTail,
                    __nv_bfloat16* __restrict__ out, int64_t ldo, int N, int H, int n_items, float scale_log2) {
  using S = PpSmem<HD>;
  constexpr bool kTail = S::kTail;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sQ = smem;                                              // [slot]
  uint8_t* sK = smem + 2 * S::kTileBytes;                          // [stage]
  uint8_t* sV = sK + kStagesKV * S::kTileBytes;                    // [stage]
  float* s_max = reinterpret_cast<float*>(sV + kStagesKV * S::kTileBytes);  // [slot][parity][half][128]
  float* s_sum = s_max + 2 * 2 * 2 * kQ;                                     // [slot][half][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_sum + 2 * 2 * kQ);
  uint64_t* q_full = bars;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;
  uint64_t* kv_empty = kv_full + kStagesKV;
  uint64_t* s_full = kv_empty + kStagesKV;   // [slot] phase n completes when Q·Kᵀ of the slot's n-th tile retired
  uint64_t* p_full = s_full + 2;             // [slot] 8 softmax warps have written P of the n-th tile
  uint64_t* o_full = p_full + 2;             // [slot] phase i completes when the last P·V of the i-th item retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = (N + kKV - 1) / kKV;
  const int QT = (N + kQ - 1) / kQ;
  const int QP = (QT + 1) / 2;               // query-tile pairs per (image, head)

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
      mbar_init(&p_full[s], 8);
      mbar_init(&o_full[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------ TMA loader ------------------------------------
    // (elect.sync, not `lane == 0`: see gemm_tcgen05.cu)
    if (elect_one()) {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int qp = item % QP, h = (item / QP) % H, b = item / (QP * H);
        mbar_wait(q_empty, (it & 1u) ^ 1u);  // the previous item's last Q·Kᵀ has retired
        mbar_expect_tx(q_full, 2 * S::kTileBytes);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          uint8_t* q = sQ + s * S::kTileBytes;
          tma_load_4d(&tmMain, q_full, q, 0, h, (2 * qp + s) * kQ, b);
          if (kTail) tma_load_4d(&tmTail, q_full, q + S::kMainBytes, 64, h, (2 * qp + s) * kQ, b);
        }
        for (int j = 0; j < T; ++j, ++g) {
          const uint32_t st = g % kStagesKV;
          mbar_wait(&kv_empty[st], ((g / kStagesKV) & 1u) ^ 1u);
          mbar_expect_tx(&kv_full[st], 2 * S::kTileBytes);
          uint8_t* k = sK + st * S::kTileBytes;
          uint8_t* v = sV + st * S::kTileBytes;
          tma_load_4d(&tmMain, &kv_full[st], k, 0, H + h, j * kKV, b);
          if (kTail) tma_load_4d(&tmTail, &kv_full[st], k + S::kMainBytes, 64, H + h, j * kKV, b);
          tma_load_4d(&tmMain, &kv_full[st], v, 0, 2 * H + h, j * kKV, b);
          if (kTail) tma_load_4d(&tmTail, &kv_full[st], v + S::kMainBytes, 64, 2 * H + h, j * kKV, b);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------ MMA issuer ------------------------------------
    if (elect_one()) {
      uint32_t g0 = 0, n0 = 0, it = 0;  // running K/V tile count and per-slot tile count at the start of the item
      // S_slot = Q_slot · K_jᵀ
      auto issue_qk = [&](int slot, int j) {
        const uint32_t st = (g0 + (uint32_t)j) % kStagesKV;
        const int n16 = (min(kKV, N - j * kKV) + 15) & ~15;
        const uint32_t tS = tmem_base + (uint32_t)(slot * kSlotCols);
        const uint32_t qaddr = smem_u32(sQ + slot * S::kTileBytes);
        const uint32_t kaddr = smem_u32(sK + st * S::kTileBytes);
        const uint64_t dQ = umma_desc(qaddr, 16, 1024, 2);
        const uint64_t dK = umma_desc(kaddr, 16, 1024, 2);
        const uint32_t idesc_qk = umma_idesc_bf16_major(kQ, n16, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16_ss(tS, dQ + static_cast<uint64_t>(2 * k), dK + static_cast<uint64_t>(2 * k), idesc_qk, k != 0);
        if (kTail) {
          const uint64_t dQt = umma_desc(qaddr + S::kMainBytes, 16, 256, 6);
          const uint64_t dKt = umma_desc(kaddr + S::kMainBytes, 16, 256, 6);
          umma_bf16_ss(tS, dQt, dKt, idesc_qk, 1u);
        }
        umma_commit(&s_full[slot]);
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it, g0 += T, n0 += T) {
        mbar_wait(q_full, it & 1u);
        mbar_wait(&kv_full[g0 % kStagesKV], (g0 / kStagesKV) & 1u);
        tc_fence_after();
        issue_qk(0, 0);
        issue_qk(1, 0);
        if (T == 1) umma_commit(q_empty);
        for (int j = 0; j < T; ++j) {
          const uint32_t g = g0 + (uint32_t)j, n = n0 + (uint32_t)j;
          const uint32_t st = g % kStagesKV;
          const int n16 = (min(kKV, N - j * kKV) + 15) & ~15;
          const int ksteps = n16 >> 4;
          const uint32_t vaddr = smem_u32(sV + st * S::kTileBytes);
          const uint64_t dV = umma_desc(vaddr, 16, 1024, 2);  // MN-major, 8-key groups 1024 B apart
          const uint64_t dVt = umma_desc(vaddr + S::kMainBytes, 16, 256, 6);
          constexpr uint32_t idesc_pv = umma_idesc_bf16_major(kQ, 64, 0, 1);
          constexpr uint32_t idesc_pvt = umma_idesc_bf16_major(kQ, 16, 0, 1);
          if (j + 1 < T) {  // next K tile (needed by both slots' next Q·Kᵀ)
            mbar_wait(&kv_full[(g + 1) % kStagesKV], ((g + 1) / kStagesKV) & 1u);
          }
#pragma unroll
          for (int slot = 0; slot < 2; ++slot) {
            const uint32_t tP = tmem_base + (uint32_t)(slot * kSlotCols);
            const uint32_t tO = tP + 128;
            // ---- O_slot (+)= P · V_j ----  (the first P of an item is only signalled after the slot's softmax
            // warps have read the previous item's O out of TMEM)
            mbar_wait(&p_full[slot], n & 1u);
            tc_fence_after();
            for (int kk = 0; kk < ksteps; ++kk) {
              const uint32_t acc = (j | kk) != 0 ? 1u : 0u;
              umma_bf16_ts(tO, tP + static_cast<uint32_t>(8 * kk), dV + static_cast<uint64_t>(kk * 128), idesc_pv, acc);
              if (kTail)
                umma_bf16_ts(tO + 64, tP + static_cast<uint32_t>(8 * kk), dVt + static_cast<uint64_t>(kk * 32),
                             idesc_pvt, acc);
            }
            if (j == T - 1) umma_commit(&o_full[slot]);
            if (slot == 1) umma_commit(&kv_empty[st]);  // K_j / V_j back to the loader once everything so far retired
            if (j + 1 < T) {
              issue_qk(slot, j + 1);
              if (slot == 1 && j + 2 == T) umma_commit(q_empty);  // every Q·Kᵀ of this item has been issued
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------ softmax / output ------------------------------------
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int slot = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;  // which 64 of the tile's 128 keys / which half of the O columns
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + (uint32_t)(slot * kSlotCols);
    const uint32_t tO = tS + 128;
    constexpr int kGroups = HD / 8;                      // 8-column groups of O that carry data
    constexpr int kG0 = (kGroups + 1) / 2;               // groups [0,kG0) -> half 0, [kG0,kGroups) -> half 1
    const int gbeg = half ? kG0 : 0, gend = half ? kGroups : kG0;
    const int bar_id = 1 + slot * 4 + quad;              // named barrier shared by the two warps of a row block
    float* xmax = s_max + slot * (2 * 2 * kQ);
    float* xsum = s_sum + slot * (2 * kQ);
    const float2 scale2 = make_float2(scale_log2, scale_log2);
    uint32_t n = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int qp = item % QP, h = (item / QP) % H, b = item / (QP * H);
      const int grow = (2 * qp + slot) * kQ + row;
      float m = -INFINITY, l = 0.f;        // m in log2 units (already multiplied by scale_log2); l = partial row sum
      for (int j = 0; j < T; ++j, ++n) {
        const int valid = min(kKV, N - j * kKV) - 64 * half;   // valid keys among this half's 64 (may be <= 0)
        mbar_wait(&s_full[slot], n & 1u);
        tc_fence_after();
        uint32_t s[64];
        tmem_ld_32x32b_x32(tS + 64 * half, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld_32x32b_x32(tS + 64 * half + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        tmem_ld_wait();
        if (valid < 64) {  // last tile: keys past the sequence end (zero-filled K rows) never win
#pragma unroll
          for (int c = 0; c < 64; ++c)
            if (c >= valid) s[c] = __float_as_uint(-INFINITY);
        }
        float mx8[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) mx8[c] = fmaxf(__uint_as_float(s[c]), __uint_as_float(s[c + 8]));
#pragma unroll
        for (int c = 16; c < 64; c += 16)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            mx8[i] = fmaxf(mx8[i], fmaxf(__uint_as_float(s[c + i]), __uint_as_float(s[c + 8 + i])));
        float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])),
                         fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        // exchange with the warp that holds the other 64 keys of the same rows
        float* xm = xmax + (n & 1u) * (2 * kQ);
        xm[half * kQ + row] = mx;
        asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
        mx = fmaxf(mx, xm[(half ^ 1) * kQ + row]) * scale_log2;  // scale > 0
        // lazy rescale: keep the old reference max unless the new one is more than 2^8 larger
        const float m_new = (mx > m + 8.0f) ? mx : m;
        const bool moved = m_new != m;
        const float alpha = (j == 0) ? 0.f : fast_exp2(m - m_new);
        if (j > 0 && __any_sync(0xffffffffu, moved)) {
          // S_j complete => P·V_{j-1} (issued before Q·Kᵀ_j on the in-order tensor pipe) has retired: O is stable
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
        const float2 neg_m2 = make_float2(-m, -m);
        float2 sum2[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float2 x = ffma2(make_float2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), scale2, neg_m2);
          float2 p;
          p.x = fast_exp2(x.x);
          p.y = fast_exp2(x.y);
          sum2[c & 3] = fadd2(sum2[c & 3], p);
          s[c] = pack_bf16x2(p.x, p.y);  // in place: s[2c], s[2c+1] (indices >= c) are consumed first
        }
        const float2 t = fadd2(fadd2(sum2[0], sum2[1]), fadd2(sum2[2], sum2[3]));
        l = l * alpha + (t.x + t.y);
        // the partner has loaded its S columns (it passed the named barrier), so half 1 may overwrite columns 32..63
        tmem_st_32x32b_x32(tS + 32 * half, *reinterpret_cast<const uint32_t(*)[32]>(&s[0]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[slot]);
      }
      // ---- O / l -> bf16 -> global ----
      xsum[half * kQ + row] = l;
      mbar_wait(&o_full[slot], it & 1u);
      tc_fence_after();
      asm volatile("bar.sync %0, 64;\n" ::"r"(bar_id) : "memory");
      const float inv = 1.0f / (l + xsum[(half ^ 1) * kQ + row]);
      __nv_bfloat16* orow = out + ((int64_t)b * N + grow) * ldo + h * HD;
#pragma unroll
      for (int c = 0; c < kG0; ++c) {
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

}  // namespace

int attention_pp_bf16(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, int B, int N, int H, int hd,
                      float scale, cudaStream_t st) {
  DFD_REQUIRE(qkv && out, DFD_ERR_BAD_ARG, "attention: null pointer");
  DFD_REQUIRE(B > 0 && N > 0 && H > 0, DFD_ERR_SHAPE, "attention: B, N, H must be positive");
  DFD_REQUIRE(hd == 64 || hd == 72, DFD_ERR_UNSUPPORTED, "attention: head dim %d not supported (64, 72)", hd);
  DFD_REQUIRE(ldqkv % 8 == 0 && ldqkv >= 3 * H * hd && ldo % 8 == 0 && ldo >= H * hd, DFD_ERR_SHAPE,
              "attention: bad leading dimensions");
  DFD_REQUIRE(((uintptr_t)qkv % 16 == 0) && ((uintptr_t)out % 16 == 0), DFD_ERR_BAD_ARG,
              "attention: pointers must be 16-byte aligned");
  const int QT = (N + kQ - 1) / kQ;
  const int64_t items64 = (int64_t)((QT + 1) / 2) * H * B;
  DFD_REQUIRE(items64 < (1ll << 31), DFD_ERR_SHAPE, "attention: too many work items");
  CUtensorMap tmMain, tmTail;
  int rc = make_tmap_qkv_4d(&tmMain, qkv, hd, 3 * H, N, B, ldqkv, 64, 128, 0);
  if (rc != DFD_OK) return rc;
  tmTail = tmMain;
  if (hd == 72) {
    rc = make_tmap_qkv_4d(&tmTail, qkv, hd, 3 * H, N, B, ldqkv, 16, 128, 1);
    if (rc != DFD_OK) return rc;
  }
  const float scale_log2 = scale * 1.4426950408889634f;
  const int n_items = (int)items64;
  const int grid = n_items < kNumSMs ? n_items : kNumSMs;
  static bool attr[2] = {false, false};
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
  if (hd == 64) {
    if (!attr[0]) {
      DFD_CUDA(cudaFuncSetAttribute(attention_pp_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    PpSmem<64>::kTotal));
      attr[0] = true;
    }
    attention_pp_kernel<64><<<grid, kThreads, PpSmem<64>::kTotal, st>>>(tmMain, tmTail, o, ldo, N, H, n_items,
                                                                       scale_log2);
  } else {
    if (!attr[1]) {
      DFD_CUDA(cudaFuncSetAttribute(attention_pp_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    PpSmem<72>::kTotal));
      attr[1] = true;
    }
    attention_pp_kernel<72><<<grid, kThreads, PpSmem<72>::kTotal, st>>>(tmMain, tmTail, o, ldo, N, H, n_items,
                                                                       scale_log2);
  }
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

}  // namespace dfd
