// 24-d frequency feature vector of a gray 256x256 fp32 image (FreqMLP input).
// Restates train_fusion_head_only.py:150-226 (= "FreqMLP trainer.py":91-177; app copy
// deepfake-detector-v2/app.py:752-846) as two HBM-streaming kernels:
//
//   freq_rows_kernel   grid (8 row bands, B).  A band of 32 rows (+1 halo row each side) is staged in shared
//                      memory with 16-byte loads.  From it: (a) the three SRM high-pass stencils' raw moments
//                      (Σy..Σy⁴, double), (b) the 2-level Haar energies, (c) 256-point row FFTs — two real
//                      rows packed into one complex Stockham radix-4 FFT per warp — written as the
//                      129-column half spectrum to scratch.
//   freq_cols_kernel   grid (P, B), P = 1..8 CTAs per image so that small batches still fill the SMs (one CTA per
//                      image left a 256-image batch latency bound at < 2 CTAs per SM); the CTAs interleave the 129
//                      columns, write their partial accumulators to scratch, and the last one to finish (ticket) adds
//                      them in a fixed order.  Each warp runs 256-point column FFTs over its stored columns and folds
//                      |F|, log|F| and angle(F) of every bin TOGETHER WITH its Hermitian mirror into the band /
//                      log-radius / sector / phase-histogram accumulators (lane-private shared-memory bins) through
//                      host-built LUTs of the (fft-shifted) 256² grid; thread 0 then finishes the 24 features.
//
// Per image: 262 144 B read + 264 192 B scratch write + read (L2 resident at these sizes) + 96 B out.
#include "dfd_common.cuh"

#include <atomic>
#include <cstddef>

namespace dfd {

extern std::atomic<int64_t> g_launches;

namespace {

constexpr int kN = 256;
constexpr int kHalf = 129;
constexpr int kBand = 32;
constexpr int kThreads = 256;
constexpr int kAccDoubles = 32;  // per-image spatial accumulators: 8 srm moments (2 stencils x 4) + 8 haar + pad
constexpr int64_t kSpecBytes = (int64_t)kN * kHalf * 8;
constexpr int kMaxColParts = 8;          // CTAs that may share one image's column pass
constexpr int kColPartBytes = 2048;      // slot of one CTA's partial accumulators (ColPart, 424 B)
// per image: half spectrum | spatial accumulators (zeroed per call) | ticket counter (zeroed) + pad | column partials
constexpr int64_t kZeroBytes = kAccDoubles * 8 + 16;
constexpr int64_t kImgScratch = kSpecBytes + kZeroBytes + kMaxColParts * kColPartBytes;
constexpr int kRowsBytes = (kBand + 2) * kN * 4;
constexpr int kFftBytes = (kThreads / 32) * 2 * kN * 8;
constexpr int kRowsSmem = kRowsBytes + kFftBytes + kN * 8;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// 256-point forward FFT (e^{-2πi kn/N}), Stockham autosort radix-4, one warp, data in shared memory.
// Input in `a`; after the 4 passes the natural-order result is back in `a`. tw[t] = e^{-2πi t/256}.
__device__ __forceinline__ void fft256_warp(float2* a, float2* b, const float2* tw, int lane) {
  float2* in = a;
  float2* out = b;
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int Ns = 1 << (2 * pass);
    const int tstep = 64 >> (2 * pass);  // 256 / (Ns * 4)
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = lane + 32 * jj;
      const int k = j & (Ns - 1);
      const int t = k * tstep;
      const float2 v0 = in[j];
      const float2 v1 = cmul(in[j + 64], tw[t]);
      const float2 v2 = cmul(in[j + 128], tw[2 * t]);
      const float2 v3 = cmul(in[j + 192], tw[3 * t]);
      const float2 t0 = make_float2(v0.x + v2.x, v0.y + v2.y);
      const float2 t1 = make_float2(v0.x - v2.x, v0.y - v2.y);
      const float2 t2 = make_float2(v1.x + v3.x, v1.y + v3.y);
      const float2 d = make_float2(v1.x - v3.x, v1.y - v3.y);
      const float2 t3 = make_float2(d.y, -d.x);  // -i * (v1 - v3)
      const int j0 = ((j - k) << 2) + k;
      out[j0] = make_float2(t0.x + t2.x, t0.y + t2.y);
      out[j0 + Ns] = make_float2(t1.x + t3.x, t1.y + t3.y);
      out[j0 + 2 * Ns] = make_float2(t0.x - t2.x, t0.y - t2.y);
      out[j0 + 3 * Ns] = make_float2(t1.x - t3.x, t1.y - t3.y);
    }
    __syncwarp();
    float2* tmp = in; in = out; out = tmp;
  }
}

__device__ __forceinline__ void fill_twiddles(float2* tw) {
  for (int t = threadIdx.x; t < kN; t += blockDim.x) {
    float s, c;
    sincospif((float)t * (1.0f / 128.0f), &s, &c);
    tw[t] = make_float2(c, -s);
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
freq_rows_kernel(const float* __restrict__ gray, uint8_t* __restrict__ scratch) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  float (*rows)[kN] = reinterpret_cast<float (*)[kN]>(dyn_smem);
  float2 (*fbuf)[2][kN] = reinterpret_cast<float2 (*)[2][kN]>(dyn_smem + kRowsBytes);
  float2* tw = reinterpret_cast<float2*>(dyn_smem + kRowsBytes + kFftBytes);
  __shared__ double red[kThreads / 32][16];

  const int band = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* img = gray + (int64_t)b * kN * kN;
  const int r0 = band * kBand;

  fill_twiddles(tw);
  for (int i = threadIdx.x; i < (kBand + 2) * (kN / 4); i += kThreads) {
    const int rr = i / (kN / 4), c4 = i % (kN / 4);
    const int gr = r0 - 1 + rr;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // zero padding of conv2d(padding="same")
    if (gr >= 0 && gr < kN) v = __ldg(reinterpret_cast<const float4*>(img + (int64_t)gr * kN) + c4);
    *reinterpret_cast<float4*>(&rows[rr][c4 * 4]) = v;
  }
  __syncthreads();

  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0;

  // ---- (a) SRM stencils: thread = column, walks the band's rows -----------------------------------
  {
    const int x = threadIdx.x;
    const bool hl = x > 0, hr = x < kN - 1;
    for (int r = 1; r <= kBand; ++r) {
      const float a00 = hl ? rows[r - 1][x - 1] : 0.f, a01 = rows[r - 1][x], a02 = hr ? rows[r - 1][x + 1] : 0.f;
      const float a10 = hl ? rows[r][x - 1] : 0.f, a11 = rows[r][x], a12 = hr ? rows[r][x + 1] : 0.f;
      const float a20 = hl ? rows[r + 1][x - 1] : 0.f, a21 = rows[r + 1][x], a22 = hr ? rows[r + 1][x + 1] : 0.f;
      // k = k2d / (sum|k2d| + 1e-8): sums are 16 and 8, exact in fp32
      float y1 = -0.0625f * a00;
      y1 = fmaf(0.125f, a01, y1); y1 = fmaf(-0.0625f, a02, y1);
      y1 = fmaf(0.125f, a10, y1); y1 = fmaf(-0.25f, a11, y1); y1 = fmaf(0.125f, a12, y1);
      y1 = fmaf(-0.0625f, a20, y1); y1 = fmaf(0.125f, a21, y1); y1 = fmaf(-0.0625f, a22, y1);
      float y2 = -0.125f * a01;
      y2 = fmaf(-0.125f, a10, y2); y2 = fmaf(0.5f, a11, y2); y2 = fmaf(-0.125f, a12, y2);
      y2 = fmaf(-0.125f, a21, y2);
      const double d1 = (double)y1, d2 = (double)y2;
      const double s1 = d1 * d1, s2 = d2 * d2;
      acc[0] += d1; acc[1] += s1; acc[2] += s1 * d1; acc[3] += s1 * s1;
      acc[4] += d2; acc[5] += s2; acc[6] += s2 * d2; acc[7] += s2 * s2;
    }
  }
  // ---- (b) Haar energies: thread = 4x4 pixel blocks (2 per thread) ---------------------------------
  for (int blk = threadIdx.x; blk < (kBand / 4) * (kN / 4); blk += kThreads) {
    const int by = blk / (kN / 4), bx = blk % (kN / 4);
    float ca[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int rr = 1 + by * 4 + 2 * i, cc = bx * 4 + 2 * j;
        const float p = rows[rr][cc], q = rows[rr][cc + 1], r = rows[rr + 1][cc], s = rows[rr + 1][cc + 1];
        const float cA = 0.5f * ((p + q) + (r + s)), cH = 0.5f * ((p + q) - (r + s));
        const float cV = 0.5f * ((p - q) + (r - s)), cD = 0.5f * ((p - q) - (r - s));
        ca[i][j] = cA;
        acc[8] += (double)cA * cA; acc[9] += (double)cH * cH;
        acc[10] += (double)cV * cV; acc[11] += (double)cD * cD;
      }
    }
    const float p = ca[0][0], q = ca[0][1], r = ca[1][0], s = ca[1][1];
    const float cA = 0.5f * ((p + q) + (r + s)), cH = 0.5f * ((p + q) - (r + s));
    const float cV = 0.5f * ((p - q) + (r - s)), cD = 0.5f * ((p - q) - (r - s));
    acc[12] += (double)cA * cA; acc[13] += (double)cH * cH;
    acc[14] += (double)cV * cV; acc[15] += (double)cD * cD;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const double v = warp_sum_d(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }

  // ---- (c) row FFTs: each warp transforms 2 row pairs ------------------------------------------------
  float2* spec = reinterpret_cast<float2*>(scratch + (int64_t)b * kImgScratch);
  float2* fa = fbuf[warp][0];
  float2* fb = fbuf[warp][1];
  for (int pr = warp; pr < kBand / 2; pr += kThreads / 32) {
    const int lr = 1 + 2 * pr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = lane + 32 * i;
      fa[n] = make_float2(rows[lr][n], rows[lr + 1][n]);
    }
    __syncwarp();
    fft256_warp(fa, fb, tw, lane);
    float2* o0 = spec + (int64_t)(r0 + 2 * pr) * kHalf;
    float2* o1 = o0 + kHalf;
    for (int k = lane; k < kHalf; k += 32) {
      const float2 z = fa[k];
      const float2 zc = fa[(kN - k) & (kN - 1)];
      // A[k] = (Z[k] + conj(Z[N-k]))/2 ; B[k] = (Z[k] - conj(Z[N-k]))/(2i)
      o0[k] = make_float2(0.5f * (z.x + zc.x), 0.5f * (z.y - zc.y));
      o1[k] = make_float2(0.5f * (z.y + zc.y), 0.5f * (zc.x - z.x));
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w][threadIdx.x];
    double* accg = reinterpret_cast<double*>(scratch + (int64_t)b * kImgScratch + kSpecBytes);
    atomicAdd(accg + threadIdx.x, t);
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int kColWarps = 4;
constexpr int kColThreads = kColWarps * 32;
constexpr int kBins = 48;                  // 40 log-radius bins (39 used) + 8 sectors
// one CTA's contribution to an image, as published to the last CTA of the image
struct ColPart {
  float sums[kBins];                       // log-magnitude sums per log-radius bin, then magnitude sums per sector
  int hist[50];                            // phase histogram
  int pad[2];
  double E[3];                             // band energies
};
static_assert(sizeof(ColPart) <= kColPartBytes && offsetof(ColPart, E) % 8 == 0, "column-pass partials must fit their slot");

// lut word of the shifted grid position (sy, sx), stored TRANSPOSED (index sx*256 + sy: the lanes of a warp walk sy, so a
// warp reads 128 contiguous bytes): bits 0-7 band, 8-15 log-radius bin (int8, -1 = not counted), 16-23 sector (int8).
//
// A bin F(ky, kx) and its Hermitian mirror F(-ky, -kx) = conj(F) are folded TOGETHER (mirrored = true): |F|, log|F| and the
// angle are computed once — hypotf / logf / atan2f were evaluated twice per stored bin before — the mirror's angle is the
// negated one, and because the mirror sits at the same radius it shares the band and the log-radius bin
// (tests/test_oracle_cpu.py checks that symmetry of the host-built table), so those sums take the doubled value in one step.
// Only the sector (an angle of the grid POSITION) differs: it is read from the mirror's own table word.
//
// Accumulators: every LANE owns a private copy of the 48 float bins in shared memory, laid out [bin][lane] (bank = lane: no
// conflicts, no atomics, no shuffles — a bin update is one load, one add, one store, and the order of a lane's additions is
// fixed, so the result is bit-reproducible).  The first version reduced every update across the warp with shuffles (lanes that
// share a key summed, one lane adds): 6-12 shuffles per update, three updates per bin — more instructions than the
// transcendental functions.  The lanes' copies are summed once per CTA, in a fixed tree.
__device__ __forceinline__ void phase_count(int* whist, float ph) {
  // torch.histc(bins=50, min=-pi, max=pi): pos = (int)((x - min) / (max - min) * bins), x == max -> last bin
  const float minv = -3.14159274101257324f, maxv = 3.14159274101257324f;
  if (ph >= minv && ph <= maxv) {
    int pos = (int)((ph - minv) / (maxv - minv) * 50.0f);
    if (pos > 49) pos = 49;
    atomicAdd(&whist[pos], 1);
  }
}
__device__ __forceinline__ void fold_bin(float re, float im, int w, int wm, bool mirrored, float* bins /* [kBins][32] + lane */,
                                         int* whist, double (&eb)[3]) {
  const float mag = hypotf(re, im);
  const float ph = atan2f(im, re);
  const float lg = logf(mag + 1e-6f);
  const int band = w & 0xff;
  const double dm = mirrored ? (double)mag + (double)mag : (double)mag;
  eb[0] += band == 0 ? dm : 0.0;
  eb[1] += band == 1 ? dm : 0.0;
  eb[2] += band == 2 ? dm : 0.0;
  const int rb = (int)(int8_t)(w >> 8), sc = (int)(int8_t)(w >> 16);
  if (rb >= 0) bins[rb * 32] += mirrored ? lg + lg : lg;
  if (sc >= 0) bins[(40 + sc) * 32] += mag;
  phase_count(whist, ph);
  if (mirrored) {   // warp-uniform
    const int scm = (int)(int8_t)(wm >> 16);
    if (scm >= 0) bins[(40 + scm) * 32] += mag;
    phase_count(whist, -ph);   // angle(conj F): atan2f is odd in its first argument
  }
}

// five CTAs per SM (<= 102 registers, 45 KB of shared memory each): the kernel is latency bound — ncu at four CTAs: issue slots
// 31 % busy, 3.4 warps per scheduler — so resident warps, not instructions, set its speed
__global__ void __launch_bounds__(kColThreads, 5)
freq_cols_kernel(uint8_t* __restrict__ scratch, const int32_t* __restrict__ lut, float eps, int zscore,
                 float* __restrict__ feats) {
  __shared__ __align__(16) float2 fbuf[kColWarps][2][kN];
  __shared__ float2 tw[kN];
  __shared__ float lanebins[kColWarps][kBins][32];
  __shared__ int whist[kColWarps][50];       // phase histogram, one copy per warp
  // per-warp sums of the final reduction reuse the FFT buffers (free after the column loop's last barrier): the CTA stays under
  // 227 KB / 5 of shared memory
  float (*wsum)[kBins] = reinterpret_cast<float (*)[kBins]>(&fbuf[0][0][0]);
  double (*ered)[3] = reinterpret_cast<double (*)[3]>(&fbuf[1][0][0]);
  __shared__ ColPart A;                      // this CTA's totals, then (last CTA of the image) the image's

  const int part = blockIdx.x, nparts = gridDim.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* img_scratch = scratch + (int64_t)b * kImgScratch;
  const float2* spec = reinterpret_cast<const float2*>(img_scratch);

  fill_twiddles(tw);
  for (int i = threadIdx.x; i < kColWarps * kBins * 32; i += kColThreads) (&lanebins[0][0][0])[i] = 0.f;
  for (int i = threadIdx.x; i < kColWarps * 50; i += kColThreads) (&whist[0][0])[i] = 0;
  __syncthreads();

  double eb[3] = {0.0, 0.0, 0.0};
  float2* fa = fbuf[warp][0];
  float2* fb = fbuf[warp][1];
  float* bins = &lanebins[warp][0][lane];
  // the CTA takes kColWarps adjacent columns at a time (one per warp): loaded together, every row contributes 32 contiguous bytes
  for (int kx0 = part * kColWarps; kx0 < kHalf; kx0 += nparts * kColWarps) {
    {
      const int c = threadIdx.x & (kColWarps - 1), r_in = threadIdx.x / kColWarps;
      if (kx0 + c < kHalf) {
#pragma unroll
        for (int i = 0; i < kN / (kColThreads / kColWarps); ++i) {
          const int r = r_in + (kColThreads / kColWarps) * i;
          fbuf[c][0][r] = __ldg(spec + (int64_t)r * kHalf + kx0 + c);
        }
      }
    }
    __syncthreads();
    const int kx = kx0 + warp;
    if (kx < kHalf) {
      const bool self_col = (kx == 0) || (kx == kN / 2);
      const int sx = (kx + kN / 2) & (kN - 1);
      const int sxm = ((kN - kx) + kN / 2) & (kN - 1);
      // the column's table words (its own and its mirror's) are fetched ahead of the FFT: their latency ends under it
      int w[8], wm[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ky = lane + 32 * i;
        w[i] = __ldg(lut + sx * kN + ((ky + kN / 2) & (kN - 1)));
        wm[i] = __ldg(lut + sxm * kN + (((kN - ky) + kN / 2) & (kN - 1)));
      }
      fft256_warp(fa, fb, tw, lane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ky = lane + 32 * i;
        float2 v = fa[ky];
        // the 4 self-conjugate bins of a real image have an exactly-zero imaginary part (torch yields +0)
        if (self_col && (ky == 0 || ky == kN / 2)) v.y = 0.0f;
        fold_bin(v.x, v.y, w[i], wm[i], !self_col, bins, whist[warp], eb);
      }
    }
    __syncthreads();
  }
  // lanes' private bins -> one sum per warp (fixed shuffle tree) -> one per CTA (warps in order)
  for (int k = 0; k < kBins; ++k) {
    float v = bins[k * 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) wsum[warp][k] = v;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double v = warp_sum_d(eb[i]);
    if (lane == 0) ered[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < kBins) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kColWarps; ++w) t += wsum[w][threadIdx.x];
    A.sums[threadIdx.x] = t;
  } else if (threadIdx.x < kBins + 50) {
    const int i = threadIdx.x - kBins;
    int t = 0;
#pragma unroll
    for (int w = 0; w < kColWarps; ++w) t += whist[w][i];
    A.hist[i] = t;
  } else if (threadIdx.x < kBins + 50 + 3) {
    const int i = threadIdx.x - kBins - 50;
    double t = 0.0;
    for (int w = 0; w < kColWarps; ++w) t += ered[w][i];
    A.E[i] = t;
  }
  __syncthreads();
  if (nparts > 1) {
    // publish this CTA's totals, take a ticket; the last CTA of the image adds all parts in part order
    __shared__ int s_ticket;
    uint8_t* parts = img_scratch + kSpecBytes + kZeroBytes;
    constexpr int kWords = (int)(sizeof(ColPart) / 4);
    int* mine = reinterpret_cast<int*>(parts + part * kColPartBytes);
    for (int i = threadIdx.x; i < kWords; i += kColThreads) mine[i] = reinterpret_cast<const int*>(&A)[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(reinterpret_cast<int*>(img_scratch + kSpecBytes + kAccDoubles * 8), 1);
    __syncthreads();
    if (s_ticket != nparts - 1) return;
    __threadfence();
    if (threadIdx.x < kBins) {
      float t = 0.f;
      for (int q = 0; q < nparts; ++q) t += __ldcg(reinterpret_cast<const float*>(parts + q * kColPartBytes) + threadIdx.x);
      A.sums[threadIdx.x] = t;
    } else if (threadIdx.x < kBins + 50) {
      const int i = threadIdx.x - kBins;
      int t = 0;
      for (int q = 0; q < nparts; ++q)
        t += __ldcg(reinterpret_cast<const int*>(parts + q * kColPartBytes + offsetof(ColPart, hist)) + i);
      A.hist[i] = t;
    } else if (threadIdx.x < kBins + 50 + 3) {
      const int i = threadIdx.x - kBins - 50;
      double t = 0.0;
      for (int q = 0; q < nparts; ++q)
        t += __ldcg(reinterpret_cast<const double*>(parts + q * kColPartBytes + offsetof(ColPart, E)) + i);
      A.E[i] = t;
    }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  const double E[3] = {A.E[0], A.E[1], A.E[2]};

  // ---- finish the 24 features (double scalars, like the reference's python floats) ----------------
  const double EPS = (double)eps;
  const double Et = E[0] + E[1] + E[2] + EPS;
  double f[24];
  f[0] = E[0] / Et;
  f[1] = E[1] / Et;
  f[2] = E[2] / Et;
  f[3] = (E[2] + EPS) / (E[0] + EPS);
  {  // slope of mean log-magnitude over the 39 log-radius bins (np.polyfit deg 1, closed form)
    double mu[39], mbar = 0.0;
    for (int i = 0; i < 39; ++i) {
      const double s = (double)A.sums[i];
      const int cnt = __ldg(lut + kN * kN + i);  // bins per log-radius ring: geometry only, counted by the host
      mu[i] = cnt > 0 ? (double)(float)(s / (double)cnt) : 0.0;
      mbar += mu[i];
    }
    mbar /= 39.0;
    double num = 0.0, den = 0.0;
    for (int i = 0; i < 39; ++i) {
      const double dx = (double)i - 19.0;
      num += dx * (mu[i] - mbar);
      den += dx * dx;
    }
    f[4] = num / den;
  }
  {  // anisotropy: population variance of the 8 sector means
    double sm[8], mbar = 0.0;
    for (int k = 0; k < 8; ++k) {
      const double s = (double)A.sums[40 + k];
      const int cnt = __ldg(lut + kN * kN + 40 + k);
      sm[k] = cnt > 0 ? (double)(float)(s / (double)cnt) : 0.0;
      mbar += sm[k];
    }
    mbar /= 8.0;
    double v = 0.0;
    for (int k = 0; k < 8; ++k) v += (sm[k] - mbar) * (sm[k] - mbar);
    f[5] = v / 8.0;
  }
  {  // phase entropy, fp32 like the reference's tensor ops
    float tot = 0.f;
    for (int i = 0; i < 50; ++i) tot += (float)A.hist[i];
    const float denom = tot + eps;
    float ent = 0.f;
    for (int i = 0; i < 50; ++i) {
      const float p = (float)A.hist[i] / denom;
      ent += p * logf(p + eps);
    }
    f[6] = (double)(-ent);
  }
  const double* accg = reinterpret_cast<const double*>(img_scratch + kSpecBytes);
  for (int i = 0; i < 4; ++i) f[7 + i] = accg[8 + i] / (128.0 * 128.0);
  for (int i = 0; i < 4; ++i) f[11 + i] = accg[12 + i] / (64.0 * 64.0);
  for (int st = 0; st < 2; ++st) {
    const double n = 65536.0;
    const double m1 = accg[4 * st] / n, m2 = accg[4 * st + 1] / n, m3 = accg[4 * st + 2] / n,
                 m4 = accg[4 * st + 3] / n;
    const double mean = (double)(float)m1;
    const double var = (double)(float)(m2 - m1 * m1);
    const double c4 = m4 - 4.0 * m1 * m3 + 6.0 * m1 * m1 * m2 - 3.0 * m1 * m1 * m1 * m1;
    const double kurt = (double)(float)c4 / ((var + EPS) * (var + EPS));
    if (st == 0) {  // SRM_K[0] (3x3 stencil embedded in 5x5) and SRM_K[1] are the same filter
      f[15] = mean; f[16] = var; f[17] = kurt;
      f[18] = mean; f[19] = var; f[20] = kurt;
    } else {
      f[21] = mean; f[22] = var; f[23] = kurt;
    }
  }
  float o[24];
  for (int i = 0; i < 24; ++i) o[i] = (float)f[i];
  if (zscore) {  // deepfake-detector-v2/app.py:840-846 (torch fp32 mean / unbiased std)
    float m = 0.f;
    for (int i = 0; i < 24; ++i) m += o[i];
    m /= 24.f;
    float v = 0.f;
    for (int i = 0; i < 24; ++i) v += (o[i] - m) * (o[i] - m);
    const float sd = sqrtf(v / 23.f);
    for (int i = 0; i < 24; ++i) o[i] = (sd < 1e-6f) ? 0.f : (o[i] - m) / (sd + 1e-6f);
  }
  for (int i = 0; i < 24; ++i) feats[(int64_t)b * 24 + i] = o[i];
}

}  // namespace

int64_t freq_scratch_bytes(int B) { return B > 0 ? (int64_t)B * kImgScratch : 0; }

int freq_features(const float* gray256, int B, const int32_t* lut, float eps, int zscore, void* scratch, float* feats,
                  cudaStream_t st) {
  DFD_REQUIRE(gray256 && lut && scratch && feats, DFD_ERR_BAD_ARG, "freq_features: null pointer");
  DFD_REQUIRE(B > 0 && B <= 65535, DFD_ERR_SHAPE, "freq_features: B must be in 1..65535");
  DFD_REQUIRE(((uintptr_t)gray256 % 16 == 0) && ((uintptr_t)scratch % 16 == 0), DFD_ERR_BAD_ARG,
              "freq_features: gray256 and scratch must be 16-byte aligned");
  uint8_t* sc = reinterpret_cast<uint8_t*>(scratch);
  // zero the per-image spatial accumulators (they sit behind each image's half spectrum)
  DFD_CUDA(cudaMemset2DAsync(sc + kSpecBytes, kImgScratch, 0, kZeroBytes, B, st));
  static SmemOptIn smem_once;
  if (int rc = ensure_dynamic_smem(smem_once, freq_rows_kernel, kRowsSmem)) return rc;
  freq_rows_kernel<<<dim3(kN / kBand, B), kThreads, kRowsSmem, st>>>(gray256, sc);
  DFD_LAUNCH_CHECK();
  // An image's column pass is 33 groups of 4 columns, dealt round-robin to `parts` CTAs; 5 CTAs fit an SM (launch bounds, 45 KB of shared
  // memory).  Take the split with the fewest (waves of CTAs) x (groups per CTA): e.g. 256 images -> 2 parts (1 wave x 17 groups).
  int parts = 1;
  {
    const int64_t slots = (int64_t)kNumSMs * 5;
    const int groups = (kHalf + kColWarps - 1) / kColWarps;
    int64_t best = -1;
    for (int p = 1; p <= kMaxColParts; ++p) {
      const int64_t cost = (((int64_t)B * p + slots - 1) / slots) * ((groups + p - 1) / p);
      if (best < 0 || cost < best) { best = cost; parts = p; }
    }
  }
  freq_cols_kernel<<<dim3(parts, B), kColThreads, 0, st>>>(sc, lut, eps, zscore, feats);
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(2, std::memory_order_relaxed);
  return DFD_OK;
}

}  // namespace dfd

extern "C" DFD_API int dfd_freq_features(const float* gray256, int B, const int32_t* lut, float eps, int zscore,
                                         void* scratch, float* feats, void* stream) {
  return dfd::freq_features(gray256, B, lut, eps, zscore, scratch, feats, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" DFD_API int64_t dfd_freq_scratch_bytes(int B) { return dfd::freq_scratch_bytes(B); }
