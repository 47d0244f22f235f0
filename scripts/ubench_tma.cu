// TMA delivery micro-benchmark for the attention kernel's K/V ring (B200, sm_100a).  Development aid, not product code.
// Written at the end of round 1 from the in-kernel experiments of profiles/r01_attention_full.md §B (handshake-only
// pipeline: 0.162 ms without K/V loads, 0.180 with the two 128-byte boxes, 0.225 / 0.257 with in-bounds / out-of-bounds
// 32-byte tail boxes); NOT YET RUN on hardware — first thing to run in round 2:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/ubench_tma.bin scripts/ubench_tma.cu && scripts/ubench_tma.bin
// One elected thread per CTA streams the K/V tiles of (image, head, query-tile) items exactly as the attention loader does
// (4-D tensor map (head-dim, head, token, image) over qkv [B*N, 3*H*hd], 64-key tiles, 4-stage ring) and nothing consumes
// them: the time is the TMA / L2 delivery ceiling for that box mix, per SM, with 1 or 2 CTAs per SM.
#include <cstdio>
#include <vector>

#include "../deepfake-detection-using-clip-based-siglip-2-vision-transformers_b200/csrc/dfd_common.cuh"

using namespace dfd;

constexpr int kMaxStages = 8;

struct Mode {
  const char* name;
  int main_boxes;   // 0: none, 1: K only, 2: K and V   (64 rows x 128 B, SWIZZLE_128B, column 0)
  int tail_boxes;   // 0: none, 1: K only, 2: K and V
  int tail_col;     // first column of the tail box (64 = crosses the tensor edge at hd 72, 56 = in bounds)
  int tail_wide;    // 0: 16 columns SWIZZLE_32B, 1: 64 columns SWIZZLE_128B
};

__global__ void __launch_bounds__(64) tma_bench(const __grid_constant__ CUtensorMap tmMain, const __grid_constant__ CUtensorMap tmTail,
                                                Mode m, int N, int H, int n_items, int QT, int kKV, int kStages, int kStageBytes,
                                                long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  __shared__ uint64_t full[kMaxStages];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x < 32 && elect_one()) {
    const int T = (N + kKV - 1) / kKV;
    const uint32_t tail_bytes = m.tail_wide ? kKV * 128 : kKV * 32;
    const uint32_t tx = m.main_boxes * kKV * 128 + m.tail_boxes * tail_bytes;
    uint32_t uses[kMaxStages] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t0 = clock64();
    int step = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int h = (item / QT) % H, b = item / (QT * H);
      for (int j = 0; j < T; ++j, ++step) {
        const int st = step % kStages;
        if (uses[st] > 0) mbar_wait(&full[st], (uses[st] - 1) & 1u);  // the stage's previous tile has landed ("consumed")
        ++uses[st];
        uint8_t* base = smem + st * kStageBytes;
        mbar_expect_tx(&full[st], tx);
        if (m.main_boxes >= 1) tma_load_4d(&tmMain, &full[st], base, 0, H + h, j * kKV, b);
        if (m.main_boxes >= 2) tma_load_4d(&tmMain, &full[st], base + kKV * 128, 0, 2 * H + h, j * kKV, b);
        const CUtensorMap* tt = m.tail_wide ? &tmMain : &tmTail;
        if (m.tail_boxes >= 1) tma_load_4d(tt, &full[st], base + 2 * kKV * 128, m.tail_col, H + h, j * kKV, b);
        if (m.tail_boxes >= 2) tma_load_4d(tt, &full[st], base + 2 * kKV * 128 + tail_bytes, m.tail_col, 2 * H + h, j * kKV, b);
      }
    }
    for (int s = 0; s < kStages; ++s)
      if (uses[s] > 0) mbar_wait(&full[s], (uses[s] - 1) & 1u);
    cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* out, void* base, int hd, int heads3, int N, int B, int64_t ld, int box_cols, CUtensorMapSwizzle sw, int kKV) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return false;
  cuuint64_t gdim[4] = {(cuuint64_t)hd, (cuuint64_t)heads3, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)hd * 2, (cuuint64_t)ld * 2, (cuuint64_t)N * (cuuint64_t)ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_cols, 1, (cuuint32_t)kKV, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return reinterpret_cast<PFN_encodeTiled>(p)(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, gdim, gstr, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int run(int hd, int tail_col_inb, int kKV, int kStages);

int main() {
  // hd 72 as shipped (144-byte heads, tails over 56..71 in bounds / 64..79 out of bounds), then the same tokens with every head
  // padded to 80 columns (160-byte heads at 32-byte alignment, tail box over 64..79 in bounds)
  for (int hd : {72, 80}) {
    const int col = hd == 72 ? 56 : 64;
    if (int rc = run(hd, col, 64, 4)) return rc;
    if (int rc = run(hd, col, 64, 8)) return rc;    // deeper ring: latency or throughput?
    if (int rc = run(hd, col, 128, 4)) return rc;   // 128-key tiles: half as many, twice as tall boxes
  }
  return 0;
}

static int run(int hd, int tail_col_inb, int kKV, int kStages) {
  const int B = 64, N = 729, H = 16, QT = (N + 127) / 128;
  printf("---- head dim stride %d, %d-key tiles, %d stages ----\n", hd, kKV, kStages);
  const int64_t ld = 3 * H * hd;
  void* qkv = nullptr;
  if (cudaMalloc(&qkv, (size_t)B * N * ld * 2) != cudaSuccess) { printf("no device\n"); return 1; }
  cudaMemset(qkv, 0, (size_t)B * N * ld * 2);
  long long* cyc = nullptr;
  cudaMalloc(&cyc, 4096 * sizeof(long long));
  CUtensorMap tmMain, tmTail;
  if (!make_map(&tmMain, qkv, hd, 3 * H, N, B, ld, 64, CU_TENSOR_MAP_SWIZZLE_128B, kKV) ||
      !make_map(&tmTail, qkv, hd, 3 * H, N, B, ld, 16, CU_TENSOR_MAP_SWIZZLE_32B, kKV)) { printf("tensor map failed\n"); return 1; }
  const Mode modes[] = {
      {"K+V main boxes only", 2, 0, 0, 0},
      {"K main only", 1, 0, 0, 0},
      {"main + 32B tails over 64..79 (out of bounds)", 2, 2, 64, 0},
      {"main + 32B tails, in bounds", 2, 2, tail_col_inb, 0},
      {"32B tails only, out of bounds", 0, 2, 64, 0},
      {"32B tails only, in bounds", 0, 2, tail_col_inb, 0},
      {"main + 128B tails over 16..79 (out of bounds)", 2, 2, 16, 1},
      {"main + 128B tails over 8..71 (in bounds)", 2, 2, 8, 1},
  };
  const int n_items = QT * H * B;
  cudaFuncSetAttribute(tma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int dev_clock_khz = 0;
  cudaDeviceGetAttribute(&dev_clock_khz, cudaDevAttrClockRate, 0);
  for (int per_sm = 1; per_sm <= 2; ++per_sm) {
    for (const Mode& m : modes) {
      const int kStageBytes = (2 * kKV * 128 + 2 * (m.tail_wide ? kKV * 128 : kKV * 32) + 1023) & ~1023;
      const int smem = kStages * kStageBytes + 1024;
      if (smem * per_sm > 220 * 1024 || smem > 200 * 1024) continue;
      if ((kKV != 64 || kStages != 4) &&
          !(m.main_boxes == 2 && !m.tail_wide && (m.tail_boxes == 0 || m.tail_col == tail_col_inb)))
        continue;  // the ring-depth / tile-height variants only run the two mixes that matter
      const int grid = 148 * per_sm;
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      tma_bench<<<grid, 64, smem>>>(tmMain, tmTail, m, N, H, n_items, QT, kKV, kStages, kStageBytes, cyc);  // warm-up
      cudaEventRecord(e0);
      tma_bench<<<grid, 64, smem>>>(tmMain, tmTail, m, N, H, n_items, QT, kKV, kStages, kStageBytes, cyc);
      cudaEventRecord(e1);
      if (cudaEventSynchronize(e1) != cudaSuccess) { printf("%s: launch failed: %s\n", m.name, cudaGetErrorString(cudaGetLastError())); return 1; }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      const int T = (N + kKV - 1) / kKV;
      const double tiles = (double)n_items * T;
      const double bytes = tiles * (m.main_boxes * kKV * 128.0 + m.tail_boxes * (m.tail_wide ? kKV * 128.0 : kKV * 32.0));
      const double boxes = tiles * (m.main_boxes + m.tail_boxes);
      printf("%d CTA/SM  %-48s %7.3f ms  %6.2f TB/s into smem  %6.1f Mboxes/s/SM  %7.1f ns per tile per SM\n", per_sm, m.name, ms,
             bytes / ms / 1e9, boxes / ms / 1e3 / 148, ms * 1e6 / (tiles / 148));
    }
  }
  cudaFree(qkv);
  cudaFree(cyc);
  return 0;
}
