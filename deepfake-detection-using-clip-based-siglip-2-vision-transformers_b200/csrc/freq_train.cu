// FreqMLP (generation 2) forward + analytic backward of mean BCE-with-logits: the training step of
// "FreqMLP trainer.py":330-396 (model: :218-301 = train_fusion_head_only.py:230-301)
//
//   x0 = (f - mean) / (std + 1e-6)            FeatureNormalizer (buffers, not trained)
//   c  = tanh(alpha * x0 + beta)              ContrastScaler
//   x  = c * sigmoid(gates[k / 6])            BandGating, 4 bands of 6 features
//   2x: x = x + drop(fc2(gelu_erf(fc1(LayerNorm(x)))))      ResidualMLPBlock (24 -> 64 -> 24, LayerNorm eps 1e-5)
//   o  = (head_w . x + head_b) / (T + 1e-6)   Linear head + TemperatureScaler
//
// One warp per sample, the 6 494 parameters and the CTA's gradient accumulator live in shared memory; lanes 0..23 own one
// feature each, every lane owns hidden units `lane` and `lane + 32`.  Flat parameter order = state-dict order without the
// two buffers: contrast.alpha[24] | contrast.beta[24] | band.gates[4] | 2 x { norm.weight[24] | norm.bias[24] |
// fc1.weight[64,24] | fc1.bias[64] | fc2.weight[24,64] | fc2.bias[24] } | head.weight[24] | head.bias | temp.T.
// Dropout (p = 0.05 in the reference's train mode) uses a counter-based hash of (seed, sample, block, feature): the
// reference's mask comes from torch's Philox stream, so runs agree in distribution only; p = 0 is deterministic and is
// what the parity tests compare with autograd.
#include "dfd_common.cuh"

#include <atomic>

namespace dfd {

extern std::atomic<int64_t> g_launches;

namespace {

constexpr int kF = 24, kHid = 64, kBands = 4;
constexpr int kOffAlpha = 0, kOffBeta = 24, kOffGates = 48, kOffBlock0 = 52;
constexpr int kBlkNw = 0, kBlkNb = 24, kBlkW1 = 48, kBlkB1 = 48 + kHid * kF, kBlkW2 = kBlkB1 + kHid,
              kBlkB2 = kBlkW2 + kF * kHid, kBlkSize = kBlkB2 + kF;  // 3208
constexpr int kOffHeadW = kOffBlock0 + 2 * kBlkSize, kOffHeadB = kOffHeadW + kF, kOffT = kOffHeadB + 1;
constexpr int kNumParams = kOffT + 1;  // 6494
static_assert(kNumParams == 6494, "FreqMLP G2 parameter count");
constexpr int kThreads = 256, kWarps = kThreads / 32;
constexpr int kScratch = 24 + 64 + 24 + 64;  // per warp: n[24], a[64], dy[24], dh[64]

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(kThreads)
freqmlp_fwd_bwd_kernel(const float* __restrict__ prm_g, const float* __restrict__ mean, const float* __restrict__ stdv,
                       const float* __restrict__ feats, const float* __restrict__ y, int B, float inv_gb,
                       float drop_p, uint32_t seed, float* __restrict__ loss_sum, float* __restrict__ grads,
                       float* __restrict__ logits) {
  extern __shared__ float sm[];
  float* prm = sm;                          // [6494]
  float* g = sm + kNumParams + 2;           // [6494] gradient accumulator of this CTA (+ loss at [6494])
  float* scr = g + kNumParams + 2;          // [warps][kScratch]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool train = grads != nullptr;
  for (int i = threadIdx.x; i < kNumParams; i += kThreads) {
    prm[i] = __ldg(prm_g + i);
    g[i] = 0.f;
  }
  if (threadIdx.x == 0) g[kNumParams] = 0.f;
  __syncthreads();

  float* n_s = scr + warp * kScratch;       // LayerNorm output of the current block
  float* a_s = n_s + 24;                    // gelu(fc1) of the current block
  float* dy_s = a_s + 64;                   // d loss / d (fc2 output)
  float* dh_s = dy_s + 24;                  // d loss / d (fc1 pre-activation)
  const bool kf = lane < kF;                // this lane owns feature `lane`
  const float Tq = prm[kOffT] + 1e-6f;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float lsum = 0.f;

  for (int s = blockIdx.x * kWarps + warp; s < B; s += gridDim.x * kWarps) {
    // ---------------- forward ----------------
    float x0 = 0.f, c = 0.f, gs = 0.f, x = 0.f;
    if (kf) {
      x0 = (__ldg(feats + (int64_t)s * kF + lane) - __ldg(mean + lane)) / (__ldg(stdv + lane) + 1e-6f);
      c = tanhf(prm[kOffAlpha + lane] * x0 + prm[kOffBeta + lane]);
      gs = 1.f / (1.f + expf(-prm[kOffGates + lane / (kF / kBands)]));
      x = c * gs;
    }
    float xh[2], rstd[2], nv[2], h[2][2], a[2][2], msk[2];
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const float* P = prm + kOffBlock0 + b * kBlkSize;
      const float mu = warp_sum(kf ? x : 0.f) * (1.f / kF);
      const float d = kf ? x - mu : 0.f;
      const float var = warp_sum(d * d) * (1.f / kF);
      rstd[b] = rsqrtf(var + 1e-5f);
      xh[b] = d * rstd[b];
      nv[b] = kf ? xh[b] * P[kBlkNw + lane] + P[kBlkNb + lane] : 0.f;
      __syncwarp();
      if (kf) n_s[lane] = nv[b];
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int j = lane + 32 * q;
        float acc = P[kBlkB1 + j];
#pragma unroll
        for (int k = 0; k < kF; ++k) acc += P[kBlkW1 + j * kF + k] * n_s[k];
        h[b][q] = acc;
        a[b][q] = gelu_erf(acc);
        a_s[j] = a[b][q];
      }
      __syncwarp();
      float yv = 0.f;
      msk[b] = 1.f;
      if (kf) {
        yv = P[kBlkB2 + lane];
#pragma unroll 8
        for (int j = 0; j < kHid; ++j) yv += P[kBlkW2 + lane * kHid + j] * a_s[j];
        if (train && drop_p > 0.f) {
          const uint32_t r = hash32(seed ^ hash32((uint32_t)s * 48u + (uint32_t)(b * kF + lane) + 0x9e3779b9u));
          msk[b] = ((r >> 8) * (1.0f / 16777216.0f) >= drop_p) ? keep_scale : 0.f;
        }
        x = x + yv * msk[b];
      }
    }
    const float logit = warp_sum(kf ? prm[kOffHeadW + lane] * x : 0.f) + prm[kOffHeadB];
    const float o = logit / Tq;
    if (logits != nullptr && lane == 0) logits[s] = o;
    if (!train) continue;
    const float yy = __ldg(y + s);
    lsum += fmaxf(o, 0.f) - o * yy + log1pf(expf(-fabsf(o)));   // BCEWithLogits (lane-uniform)

    // ---------------- backward ----------------
    const float dout = (1.f / (1.f + expf(-o)) - yy) * inv_gb;
    const float dlogit = dout / Tq;
    if (lane == 0) {
      atomicAdd(&g[kOffT], -dout * logit / (Tq * Tq));
      atomicAdd(&g[kOffHeadB], dlogit);
    }
    float dx = 0.f;
    if (kf) {
      atomicAdd(&g[kOffHeadW + lane], dlogit * x);
      dx = dlogit * prm[kOffHeadW + lane];
    }
#pragma unroll
    for (int b = 1; b >= 0; --b) {
      const float* P = prm + kOffBlock0 + b * kBlkSize;
      float* G = g + kOffBlock0 + b * kBlkSize;
      const float dy = kf ? dx * msk[b] : 0.f;
      __syncwarp();
      if (kf) {
        dy_s[lane] = dy;
        atomicAdd(&G[kBlkB2 + lane], dy);
      }
      a_s[lane] = a[b][0];
      a_s[lane + 32] = a[b][1];
      if (kf) n_s[lane] = nv[b];
      __syncwarp();
      if (kf) {
#pragma unroll 8
        for (int j = 0; j < kHid; ++j) atomicAdd(&G[kBlkW2 + lane * kHid + j], dy * a_s[j]);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int j = lane + 32 * q;
        float da = 0.f;
#pragma unroll
        for (int k = 0; k < kF; ++k) da += P[kBlkW2 + k * kHid + j] * dy_s[k];
        const float hp = h[b][q];
        const float cdf = 0.5f * (1.f + erff(hp * 0.70710678118654752440f));
        const float pdf = 0.3989422804014327f * expf(-0.5f * hp * hp);
        const float dh = da * (cdf + hp * pdf);
        dh_s[j] = dh;
        atomicAdd(&G[kBlkB1 + j], dh);
#pragma unroll
        for (int k = 0; k < kF; ++k) atomicAdd(&G[kBlkW1 + j * kF + k], dh * n_s[k]);
      }
      __syncwarp();
      float dn = 0.f;
      if (kf) {
#pragma unroll 8
        for (int j = 0; j < kHid; ++j) dn += P[kBlkW1 + j * kF + lane] * dh_s[j];
        atomicAdd(&G[kBlkNw + lane], dn * xh[b]);
        atomicAdd(&G[kBlkNb + lane], dn);
      }
      const float dxh = kf ? dn * P[kBlkNw + lane] : 0.f;
      const float m1 = warp_sum(dxh) * (1.f / kF);
      const float m2 = warp_sum(dxh * xh[b]) * (1.f / kF);
      if (kf) dx += rstd[b] * (dxh - m1 - xh[b] * m2);
    }
    if (kf) {
      atomicAdd(&g[kOffGates + lane / (kF / kBands)], dx * c * gs * (1.f - gs));
      const float du = dx * gs * (1.f - c * c);
      atomicAdd(&g[kOffAlpha + lane], du * x0);
      atomicAdd(&g[kOffBeta + lane], du);
    }
  }
  if (!train) return;
  if (lane == 0) atomicAdd(&g[kNumParams], lsum * inv_gb);
  __syncthreads();
  for (int i = threadIdx.x; i < kNumParams; i += kThreads)
    if (g[i] != 0.f) atomicAdd(grads + i, g[i]);
  if (threadIdx.x == 0) atomicAdd(loss_sum, g[kNumParams]);
}

}  // namespace
}  // namespace dfd

// loss_sum[1] and grads[6494] are ACCUMULATED (zero them first); this rank's partial sums over its B samples with the
// 1/global_batch factor applied — all-reduce-sum them across ranks.  grads == NULL: forward only (eval mode, no dropout);
// logits may be NULL.  Replaces `logits = model(xb); loss = criterion(logits, yb); loss.backward()`
// ("FreqMLP trainer.py":366-369).
extern "C" DFD_API int dfd_freqmlp_fwd_bwd(const float* params6494, const float* mean24, const float* std24,
                                           const float* feats, const float* y, int B, float inv_global_batch,
                                           float dropout_p, uint32_t seed, float* loss_sum, float* grads, float* logits,
                                           void* stream) {
  using namespace dfd;
  DFD_REQUIRE(params6494 && mean24 && std24 && feats, DFD_ERR_BAD_ARG, "freqmlp_fwd_bwd: null pointer");
  DFD_REQUIRE(B > 0, DFD_ERR_SHAPE, "freqmlp_fwd_bwd: B must be positive");
  DFD_REQUIRE(grads == nullptr || (y != nullptr && loss_sum != nullptr), DFD_ERR_BAD_ARG,
              "freqmlp_fwd_bwd: training needs labels and a loss accumulator");
  DFD_REQUIRE(grads != nullptr || logits != nullptr, DFD_ERR_BAD_ARG, "freqmlp_fwd_bwd: nothing to compute");
  DFD_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, DFD_ERR_BAD_ARG, "freqmlp_fwd_bwd: dropout_p must be in [0, 1)");
  const int smem = (2 * (kNumParams + 2) + kWarps * kScratch) * (int)sizeof(float);
  static SmemOptIn smem_once;
  if (int rc = ensure_dynamic_smem(smem_once, freqmlp_fwd_bwd_kernel, smem)) return rc;
  int grid = (B + kWarps - 1) / kWarps;
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
  freqmlp_fwd_bwd_kernel<<<grid, kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      params6494, mean24, std24, feats, y, B, inv_global_batch, dropout_p, seed, loss_sum, grads, logits);
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}
