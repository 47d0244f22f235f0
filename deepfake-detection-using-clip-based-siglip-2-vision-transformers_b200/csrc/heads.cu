// Scoring-stack kernels (all fp32 arithmetic, tiny working sets, latency/HBM bound):
//
//   head_fwd_kernel        L2-normalise pooled embeddings, optional SE gate + classifier MLP (H-A / H-B),
//                          optional prototype distances        (inference_ai_human_images.py:131-152,288-295;
//                                                               train_fusion_head_only.py:84-109)
//   score_epilogue_kernel  one warp per sample: FreqMLP (G1 or G2) → fusion (G1 linear-on-probabilities or
//                          G2 AdaptiveFusionHead) → temperature → CORAL probabilities / argmax / moments
//                          (deepfake-detector-v2/app.py:601-628,691-709,1265-1297,1355-1396;
//                           train_fusion_head_only.py:230-317)
//   fusion_fwd_bwd_kernel  AdaptiveFusionHead forward + analytic backward of mean BCE-with-logits, one warp
//                          per sample, lane j owns hidden unit j  (train_fusion_head_only.py:303-317,423-425)
#include "dfd_common.cuh"

#include <atomic>

namespace dfd {

extern std::atomic<int64_t> g_launches;

namespace {

// ------------------------------------------------------------------------------------------------
// classifier head
// ------------------------------------------------------------------------------------------------
constexpr int kHeadThreads = 1024;  // 32 warps: the dense chain is a latency chain per output row, so rows in flight are what count
constexpr int kHeadS = 4;  // samples per CTA (weights are read once per CTA)

// out[s][j] = act(b[j] + sum_k W[j][k] * in[s][k]); one warp per output row j, lanes stride k.
// act: 0 none, 1 relu, 2 gelu(erf), 3 sigmoid
__device__ void dense_rows(const float* __restrict__ W, const float* __restrict__ bias, const float* in,
                           int in_stride, float* out, int out_stride, int J, int K, int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int j = warp; j < J; j += nwarp) {
    const float* w = W + (int64_t)j * K;
    float acc[kHeadS];
#pragma unroll
    for (int s = 0; s < kHeadS; ++s) acc[s] = 0.f;
    if ((K & 127) == 0 && (in_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
      for (int k = lane * 4; k < K; k += 128) {  // 16-byte weight loads (rows are 16-byte aligned: K % 4 == 0)
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
#pragma unroll
        for (int s = 0; s < kHeadS; ++s) {
          const float4 x = *reinterpret_cast<const float4*>(in + s * in_stride + k);
          acc[s] += (wv.x * x.x + wv.y * x.y) + (wv.z * x.z + wv.w * x.w);
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float wv = __ldg(w + k);
#pragma unroll
        for (int s = 0; s < kHeadS; ++s) acc[s] += wv * in[s * in_stride + k];
      }
    }
#pragma unroll
    for (int s = 0; s < kHeadS; ++s) acc[s] = warp_sum(acc[s]);
    if (lane == 0) {
      const float bj = bias ? __ldg(bias + j) : 0.f;
#pragma unroll
      for (int s = 0; s < kHeadS; ++s) {
        float v = acc[s] + bj;
        if (act == 1) v = fmaxf(v, 0.f);
        else if (act == 2) v = gelu_erf(v);
        else if (act == 3) v = 1.f / (1.f + expf(-v));
        out[s * out_stride + j] = v;
      }
    }
  }
}

__device__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(dfd_head_weights w, const __nv_bfloat16* __restrict__ pooled, int64_t ldp, int B,
                const float* __restrict__ protos, float* __restrict__ feat_out, float* __restrict__ z_sig,
                float* __restrict__ p_proto) {
  extern __shared__ __align__(16) float sm[];
  const int D = w.dim;
  float* f = sm;                 // [S][D] normalised features
  float* a = f + kHeadS * D;     // [S][D] scratch A
  float* c = a + kHeadS * D;     // [S][D] scratch B
  __shared__ float red[kHeadThreads / 32];
  const int b0 = blockIdx.x * kHeadS;

  // ---- L2 normalise (fp32) --------------------------------------------------------------------
  for (int s = 0; s < kHeadS; ++s) {
    const int b = b0 + s;
    float ss = 0.f;
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      const float v = (b < B) ? __bfloat162float(pooled[(int64_t)b * ldp + k]) : 0.f;
      f[s * D + k] = v;
      ss += v * v;
    }
    const float nrm = sqrtf(block_sum(ss, red)) + w.norm_eps;
    const float inv = (b < B && nrm > 0.f) ? 1.f / nrm : 0.f;
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      const float v = f[s * D + k] * inv;
      f[s * D + k] = v;
      if (feat_out != nullptr && b < B) feat_out[(int64_t)b * D + k] = v;
    }
  }
  __syncthreads();

  // ---- prototype classifier -----------------------------------------------------------------------
  if (protos != nullptr && p_proto != nullptr) {
    for (int s = 0; s < kHeadS; ++s) {
      float dr = 0.f, df = 0.f;
      for (int k = threadIdx.x; k < D; k += blockDim.x) {
        const float v = f[s * D + k];
        const float e0 = v - __ldg(protos + k), e1 = v - __ldg(protos + D + k);
        dr += e0 * e0;
        df += e1 * e1;
      }
      dr = sqrtf(block_sum(dr, red));
      df = sqrtf(block_sum(df, red));
      if (threadIdx.x == 0 && b0 + s < B) {
        // softmax([-d_real, -d_fake])[1] = 1 / (1 + exp(d_fake - d_real))
        p_proto[b0 + s] = 1.f / (1.f + expf(df - dr));
      }
    }
    __syncthreads();
  }
  if (w.kind == 0 || z_sig == nullptr) return;

  const float* x = f;
  if (w.kind == 2) {
    // SE gate: f * sigmoid(W2 relu(W1 f + b1) + b2)
    const int R = D / 16;
    dense_rows(w.se_w1, w.se_b1, f, D, a, D, R, D, 1);
    __syncthreads();
    dense_rows(w.se_w2, w.se_b2, a, D, c, D, D, R, 3);
    __syncthreads();
    for (int i = threadIdx.x; i < kHeadS * D; i += blockDim.x) c[i] *= f[i];
    __syncthreads();
    x = c;
  }
  // LayerNorm (biased variance, eps inside the sqrt) -> a
  for (int s = 0; s < kHeadS; ++s) {
    float sum = 0.f;
    for (int k = threadIdx.x; k < D; k += blockDim.x) sum += x[s * D + k];
    const float mean = block_sum(sum, red) / (float)D;
    float sq = 0.f;
    for (int k = threadIdx.x; k < D; k += blockDim.x) {
      const float d = x[s * D + k] - mean;
      sq += d * d;
    }
    const float rstd = rsqrtf(block_sum(sq, red) / (float)D + w.ln_eps);
    for (int k = threadIdx.x; k < D; k += blockDim.x)
      a[s * D + k] = (x[s * D + k] - mean) * rstd * __ldg(w.ln_g + k) + __ldg(w.ln_b + k);
  }
  __syncthreads();
  // dense chain, ping-ponging between `a` and the buffer that is free (f or c)
  float* cur = a;
  float* nxt = (x == c) ? f : c;
  for (int li = 0; li < w.n_layers; ++li) {
    const dfd_dense_layer& L = w.layers[li];
    dense_rows(L.w, L.b, cur, D, nxt, D, L.out_dim, L.in_dim, L.act);
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
  }
  if (threadIdx.x < kHeadS && b0 + threadIdx.x < B) z_sig[b0 + threadIdx.x] = cur[threadIdx.x * D];
}

// ------------------------------------------------------------------------------------------------
// score epilogue
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// LayerNorm over 24 values held redundantly by every lane (biased variance).
__device__ __forceinline__ void ln24(const float (&x)[24], float (&y)[24], const float* g, const float* b,
                                     float eps) {
  float m = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) m += x[i];
  m *= (1.f / 24.f);
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = x[i] - m; v += d * d; }
  const float r = 1.f / sqrtf(v * (1.f / 24.f) + eps);
#pragma unroll
  for (int i = 0; i < 24; ++i) y[i] = (x[i] - m) * r * __ldg(g + i) + __ldg(b + i);
}

__global__ void __launch_bounds__(128)
score_epilogue_kernel(dfd_score_weights w, const float* __restrict__ z_sig, const float* __restrict__ feats,
                      const float* __restrict__ z_freq_in, int B, dfd_scores out) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;

  float zf;
  if (feats != nullptr) {
    float x[24], y[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) x[i] = __ldg(feats + (int64_t)b * 24 + i);
    if (w.gen == 1) {
      // SafeLayerNorm(24, eps 1e-5) -> Linear(24,64) -> GELU(erf) -> Linear(64,1)
      ln24(x, y, w.g1_ln_w, w.g1_ln_b, 1e-5f);
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = lane + 32 * r;
        float hsum = __ldg(w.g1_b1 + j);
#pragma unroll
        for (int i = 0; i < 24; ++i) hsum += __ldg(w.g1_w1 + j * 24 + i) * y[i];
        acc += __ldg(w.g1_w2 + j) * gelu_erf(hsum);
      }
      zf = warp_sum(acc) + __ldg(w.g1_b2);
    } else {
      // normer -> contrast -> band gating -> 2 residual MLP blocks -> head -> temperature
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        float v = (x[i] - __ldg(w.g2_mean + i)) / (__ldg(w.g2_std + i) + 1e-6f);
        v = tanhf(__ldg(w.g2_alpha + i) * v + __ldg(w.g2_beta + i));
        x[i] = v * sigmoidf_(__ldg(w.g2_gates + i / 6));
      }
#pragma unroll 1
      for (int blk = 0; blk < 2; ++blk) {
        const float* nw = w.g2_blk[blk][0];
        const float* nb = w.g2_blk[blk][1];
        const float* w1 = w.g2_blk[blk][2];
        const float* b1 = w.g2_blk[blk][3];
        const float* w2 = w.g2_blk[blk][4];
        const float* b2 = w.g2_blk[blk][5];
        ln24(x, y, nw, nb, 1e-5f);
        float hv[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int j = lane + 32 * r;
          float hsum = __ldg(b1 + j);
#pragma unroll
          for (int i = 0; i < 24; ++i) hsum += __ldg(w1 + j * 24 + i) * y[i];
          hv[r] = gelu_erf(hsum);
        }
#pragma unroll
        for (int i = 0; i < 24; ++i) {
          float p = __ldg(w2 + i * 64 + lane) * hv[0] + __ldg(w2 + i * 64 + lane + 32) * hv[1];
          p = warp_sum(p);
          x[i] = x[i] + p + __ldg(b2 + i);
        }
      }
      float acc = __ldg(w.g2_head_b);
#pragma unroll
      for (int i = 0; i < 24; ++i) acc += __ldg(w.g2_head_w + i) * x[i];
      zf = acc / (w.g2_temp + 1e-6f);
    }
  } else {
    zf = z_freq_in[b];
  }

  const float zs = z_sig[b];
  float z;
  if (w.gen == 1) {
    const float p_sig = sigmoidf_(zs);
    const float p_freq = sigmoidf_(zf / w.freq_temp);
    z = w.g1_fc_w[0] * p_sig + w.g1_fc_w[1] * p_freq + w.g1_fc_b;
  } else {
    const float xin[3] = {zf, zs, fabsf(zf - zs)};
    // lane j owns hidden unit j
    float hpre = __ldg(w.f2_b0 + lane);
#pragma unroll
    for (int k = 0; k < 3; ++k) hpre += __ldg(w.f2_w0 + lane * 3 + k) * xin[k];
    const float hh = gelu_erf(hpre);
    const float l0 = warp_sum(__ldg(w.f2_w1 + lane) * hh) + __ldg(w.f2_b1);
    const float l1 = warp_sum(__ldg(w.f2_w1 + 32 + lane) * hh) + __ldg(w.f2_b1 + 1);
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float w0 = e0 / (e0 + e1), w1 = e1 / (e0 + e1);
    z = (w0 * zf + w1 * zs) / (w.f2_temp + 1e-6f);
  }

  if (lane != 0) return;
  const float zsc = z / fmaxf(w.coral_temp, 1e-3f);
  const float p_raw = sigmoidf_(zsc);
  float g[4], p[5];
#pragma unroll
  for (int k = 0; k < 4; ++k) g[k] = sigmoidf_(zsc - w.coral_cuts[k]);
  p[0] = 1.f - g[0];
  p[1] = g[0] - g[1];
  p[2] = g[1] - g[2];
  p[3] = g[2] - g[3];
  p[4] = g[3];
  const float psum = (((p[0] + p[1]) + p[2]) + p[3]) + p[4] + 1e-8f;
  int idx = 0;
  float best = -INFINITY, mu = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    p[k] = p[k] / psum;
    if (p[k] > best) { best = p[k]; idx = k; }  // first maximum, like torch.argmax
    mu += (float)k * p[k];
  }
  float var = 0.f, ent = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    var += p[k] * ((float)k - mu) * ((float)k - mu);
    ent -= p[k] * logf(p[k] + 1e-8f);
  }
  const float p_coral = fminf(fmaxf(mu / 4.f + 0.5f * var, 0.f), 1.f);
  const float p_blend = fminf(fmaxf(0.70f * p_raw + 0.30f * p_coral, 0.f), 1.f);
  if (out.z_freq) out.z_freq[b] = zf;
  if (out.z) out.z[b] = z;
  if (out.z_scaled) out.z_scaled[b] = zsc;
  if (out.p_raw) out.p_raw[b] = p_raw;
  if (out.risk_probs) {
#pragma unroll
    for (int k = 0; k < 5; ++k) out.risk_probs[(int64_t)b * 5 + k] = p[k];
  }
  if (out.p_coral) out.p_coral[b] = p_coral;
  if (out.entropy) out.entropy[b] = ent;
  if (out.p_blend) out.p_blend[b] = p_blend;
  if (out.risk_idx) out.risk_idx[b] = idx;
}

// ------------------------------------------------------------------------------------------------
// AdaptiveFusionHead forward + backward
//   params: W0[32,3] | b0[32] | W1[2,32] | b1[2] | T        (offsets 0, 96, 128, 192, 194)
// ------------------------------------------------------------------------------------------------
constexpr int kFusThreads = 256;

__global__ void __launch_bounds__(kFusThreads)
fusion_fwd_bwd_kernel(const float* __restrict__ prm, const float* __restrict__ z_freq,
                      const float* __restrict__ z_sig, const float* __restrict__ y, int B, float inv_gb,
                      float* __restrict__ loss_sum, float* __restrict__ grads, float* __restrict__ logits) {
  __shared__ float sg[kFusThreads / 32][200];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = kFusThreads / 32;
  const float w00 = __ldg(prm + lane * 3), w01 = __ldg(prm + lane * 3 + 1), w02 = __ldg(prm + lane * 3 + 2);
  const float b0 = __ldg(prm + 96 + lane);
  const float w10 = __ldg(prm + 128 + lane), w11 = __ldg(prm + 160 + lane);
  const float b10 = __ldg(prm + 192), b11 = __ldg(prm + 193);
  const float Tq = __ldg(prm + 194) + 1e-6f;

  float gw0[3] = {0.f, 0.f, 0.f}, gb0 = 0.f, gw1[2] = {0.f, 0.f};
  float gb1[2] = {0.f, 0.f}, gT = 0.f, lsum = 0.f;  // lane-uniform accumulators

  for (int s = blockIdx.x * nwarp + warp; s < B; s += gridDim.x * nwarp) {
    const float zf = __ldg(z_freq + s), zs = __ldg(z_sig + s), yy = __ldg(y + s);
    const float x2 = fabsf(zf - zs);
    const float hpre = b0 + w00 * zf + w01 * zs + w02 * x2;
    const float cdf = 0.5f * (1.f + erff(hpre * 0.70710678118654752440f));
    const float hh = hpre * cdf;
    const float l0 = warp_sum(w10 * hh) + b10;
    const float l1 = warp_sum(w11 * hh) + b11;
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
    const float z = p0 * zf + p1 * zs;
    const float o = z / Tq;
    if (logits != nullptr && lane == 0) logits[s] = o;
    // BCEWithLogits: max(o,0) - o*y + log1p(exp(-|o|))
    lsum += fmaxf(o, 0.f) - o * yy + log1pf(expf(-fabsf(o)));
    const float dout = (1.f / (1.f + expf(-o)) - yy) * inv_gb;
    gT += -dout * z / (Tq * Tq);
    const float dz = dout / Tq;
    const float dp0 = dz * zf, dp1 = dz * zs;
    const float dot = p0 * dp0 + p1 * dp1;
    const float dl0 = p0 * (dp0 - dot), dl1 = p1 * (dp1 - dot);
    gb1[0] += dl0;
    gb1[1] += dl1;
    gw1[0] += dl0 * hh;
    gw1[1] += dl1 * hh;
    const float dh = dl0 * w10 + dl1 * w11;
    const float pdf = 0.3989422804014327f * expf(-0.5f * hpre * hpre);
    const float dpre = dh * (cdf + hpre * pdf);
    gw0[0] += dpre * zf;
    gw0[1] += dpre * zs;
    gw0[2] += dpre * x2;
    gb0 += dpre;
  }
  float* g = sg[warp];
  g[lane * 3 + 0] = gw0[0];
  g[lane * 3 + 1] = gw0[1];
  g[lane * 3 + 2] = gw0[2];
  g[96 + lane] = gb0;
  g[128 + lane] = gw1[0];
  g[160 + lane] = gw1[1];
  if (lane == 0) {
    g[192] = gb1[0];
    g[193] = gb1[1];
    g[194] = gT;
    g[195] = lsum * inv_gb;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 196; i += kFusThreads) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < kFusThreads / 32; ++wv) t += sg[wv][i];
    if (i < 195) atomicAdd(grads + i, t);
    else atomicAdd(loss_sum, t);
  }
}

}  // namespace

int head_fwd(const dfd_head_weights* w, const void* pooled, int64_t ldp, int B, const float* protos,
             float* feat_out, float* z_sig, float* p_proto, cudaStream_t st) {
  DFD_REQUIRE(w && pooled, DFD_ERR_BAD_ARG, "head_fwd: null pointer");
  DFD_REQUIRE(B > 0 && w->dim > 0 && w->dim % 16 == 0 && ldp >= w->dim, DFD_ERR_SHAPE,
              "head_fwd: bad shape (B=%d dim=%d ldp=%lld)", B, w->dim, (long long)ldp);
  DFD_REQUIRE(w->kind >= 0 && w->kind <= 2, DFD_ERR_BAD_ARG, "head_fwd: kind must be 0, 1 or 2");
  if (w->kind != 0 && z_sig != nullptr) {
    DFD_REQUIRE(w->ln_g && w->ln_b, DFD_ERR_BAD_ARG, "head_fwd: LayerNorm weights missing");
    DFD_REQUIRE(w->n_layers >= 1 && w->n_layers <= 6, DFD_ERR_BAD_ARG, "head_fwd: n_layers must be 1..6");
    int in_dim = w->dim;
    for (int i = 0; i < w->n_layers; ++i) {
      const dfd_dense_layer& L = w->layers[i];
      DFD_REQUIRE(L.w != nullptr && L.in_dim == in_dim && L.out_dim >= 1 && L.out_dim <= w->dim && L.act >= 0 &&
                      L.act <= 3,
                  DFD_ERR_SHAPE, "head_fwd: layer %d is inconsistent (in %d, expected %d, out %d)", i, L.in_dim, in_dim,
                  L.out_dim);
      in_dim = L.out_dim;
    }
    DFD_REQUIRE(in_dim == 1, DFD_ERR_SHAPE, "head_fwd: the last layer must have out_dim 1");
    DFD_REQUIRE(w->kind != 2 || (w->se_w1 && w->se_b1 && w->se_w2 && w->se_b2), DFD_ERR_BAD_ARG,
                "head_fwd: SE weights missing for kind 2");
  }
  const int smem = 3 * kHeadS * w->dim * (int)sizeof(float);
  DFD_REQUIRE(smem <= 200 * 1024, DFD_ERR_UNSUPPORTED, "head_fwd: dim %d too large", w->dim);
  DFD_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  head_fwd_kernel<<<(B + kHeadS - 1) / kHeadS, kHeadThreads, smem, st>>>(
      *w, reinterpret_cast<const __nv_bfloat16*>(pooled), ldp, B, protos, feat_out, z_sig, p_proto);
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

int score_epilogue(const dfd_score_weights* w, const float* z_sig, const float* feats,
                   const float* z_freq_in, int B, const dfd_scores* out, cudaStream_t st) {
  DFD_REQUIRE(w && z_sig && out, DFD_ERR_BAD_ARG, "score_epilogue: null pointer");
  DFD_REQUIRE(B > 0, DFD_ERR_SHAPE, "score_epilogue: B must be positive");
  DFD_REQUIRE(w->gen == 1 || w->gen == 2, DFD_ERR_BAD_ARG, "score_epilogue: gen must be 1 or 2");
  DFD_REQUIRE(feats != nullptr || z_freq_in != nullptr, DFD_ERR_BAD_ARG,
              "score_epilogue: need feats or z_freq_in");
  if (feats != nullptr) {
    if (w->gen == 1) {
      DFD_REQUIRE(w->g1_ln_w && w->g1_ln_b && w->g1_w1 && w->g1_b1 && w->g1_w2 && w->g1_b2, DFD_ERR_BAD_ARG,
                  "score_epilogue: G1 FreqMLP weights missing");
    } else {
      DFD_REQUIRE(w->g2_mean && w->g2_std && w->g2_alpha && w->g2_beta && w->g2_gates && w->g2_head_w &&
                      w->g2_head_b, DFD_ERR_BAD_ARG, "score_epilogue: G2 FreqMLP weights missing");
      for (int b = 0; b < 2; ++b)
        for (int i = 0; i < 6; ++i)
          DFD_REQUIRE(w->g2_blk[b][i] != nullptr, DFD_ERR_BAD_ARG, "score_epilogue: G2 block weights missing");
    }
  }
  if (w->gen == 2)
    DFD_REQUIRE(w->f2_w0 && w->f2_b0 && w->f2_w1 && w->f2_b1, DFD_ERR_BAD_ARG,
                "score_epilogue: G2 fusion weights missing");
  const int warps = 4;
  score_epilogue_kernel<<<(B + warps - 1) / warps, warps * 32, 0, st>>>(*w, z_sig, feats, z_freq_in, B, *out);
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

int fusion_fwd_bwd(const float* params, const float* z_freq, const float* z_sig, const float* y, int B,
                   float inv_gb, float* loss_sum, float* grads, float* logits, cudaStream_t st) {
  DFD_REQUIRE(params && z_freq && z_sig && y && loss_sum && grads, DFD_ERR_BAD_ARG,
              "fusion_fwd_bwd: null pointer");
  DFD_REQUIRE(B > 0, DFD_ERR_SHAPE, "fusion_fwd_bwd: B must be positive");
  const int per = kFusThreads / 32;
  int blocks = (B + per - 1) / per;
  if (blocks > kNumSMs) blocks = kNumSMs;
  fusion_fwd_bwd_kernel<<<blocks, kFusThreads, 0, st>>>(params, z_freq, z_sig, y, B, inv_gb, loss_sum, grads,
                                                        logits);
  DFD_LAUNCH_CHECK();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DFD_OK;
}

}  // namespace dfd

extern "C" DFD_API int dfd_head_fwd(const dfd_head_weights* w, const void* pooled_bf16, int64_t ldp, int B,
                                    const float* prototypes, float* feat_out, float* z_sig, float* p_proto,
                                    void* stream) {
  return dfd::head_fwd(w, pooled_bf16, ldp, B, prototypes, feat_out, z_sig, p_proto,
                       reinterpret_cast<cudaStream_t>(stream));
}
extern "C" DFD_API int dfd_score_epilogue(const dfd_score_weights* w, const float* z_sig, const float* feats,
                                          const float* z_freq_in, int B, const dfd_scores* out,
                                          void* stream) {
  return dfd::score_epilogue(w, z_sig, feats, z_freq_in, B, out, reinterpret_cast<cudaStream_t>(stream));
}
extern "C" DFD_API int dfd_fusion_fwd_bwd(const float* params195, const float* z_freq, const float* z_sig,
                                          const float* y, int B, float inv_global_batch, float* loss_sum,
                                          float* grads195, float* logits, void* stream) {
  return dfd::fusion_fwd_bwd(params195, z_freq, z_sig, y, B, inv_global_batch, loss_sum, grads195, logits,
                             reinterpret_cast<cudaStream_t>(stream));
}
