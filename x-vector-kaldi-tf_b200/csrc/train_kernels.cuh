// train_kernels.cuh -- the HBM-bound / small kernels of one training minibatch (sm_100a), i.e. everything of
// ``sess.run([optimizer, loss, accuracy])`` (local/tf/models.py:263) that is not a frame-level contraction:
//
//   frame level  blk_col_sums_kernel      per aligned 32-row block: column sums (BatchNorm batch moments,
//                                         tf_block.py:19; per-segment pooling sums, models.py:485; BN backward sums)
//                bn_fwd_finalize_kernel   batch mean / population variance -> scale, shift; moving statistics
//                                         pop*decay + batch*(1-decay) (tf_block.py:20-21)
//                bn_apply_kernel          y = r*scale + shift on segment rows, exact zeros on gap rows
//                bn_bwd_finalize_kernel   d gamma, d beta and the three per-channel coefficients of dL/dz
//                bn_relu_bwd_kernel       dz = (r > 0) * (cA*dy + cB*r + cC)  (BN backward + ReLU backward fused)
//                pool_*                   statistics pooling forward / backward fused with the LAST layer's BN:
//                                         dz4 = (r4 > 0) * (A[seg,c] + G[seg,c] * r4) needs no dy tensor at all
//   segment level sgemm64_kernel (+ splitk_reduce_kernel), seg_relu_bn_{fwd,bwd}_kernel, softmax_ce_kernel
//   optimizer    adam_kernel (tf.train.AdamOptimizer, models.py:518), repack_* (fp32 master -> fp16 operand layouts)
//
// All reductions run in a fixed order (bit-reproducible).  Gradients of the frame level are carried in fp16 scaled
// by the loss scale S; every kernel that emits a parameter gradient multiplies by 1/S.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace trk {

constexpr int BLK_ROWS = 32;
constexpr int COLS_PER_CTA = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __half22float2(h[i]);
    f[2 * i] = v.x; f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return u;
}

// ---------------------------------------------------------------------------------------------------------
// Column sums of `rows_per_cta` rows (a multiple of 32) x 256 channels per CTA -> ONE partial row per CTA.
//   OP 0: partial[part][0][c] = sum x,        partial[part][1][c] = sum x*x
//   OP 1: partial[part][0][c] = sum x (= dy), partial[part][1][c] = sum x*y (= dy * r)
// Gap rows hold exact zeros in every tensor this is applied to, so no row mask is needed.  The frame layers use 128
// rows per CTA; the last layer uses one CTA per segment, so that its partial rows are the pooling sums (models.py:485).
template <int OP>
__global__ void __launch_bounds__(256)
blk_col_sums_kernel(const __half* __restrict__ x, const __half* __restrict__ y, int32_t C, int32_t rows_per_cta, float* __restrict__ partial) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ float red[8][2][COLS_PER_CTA];
  const int part = blockIdx.y, c0 = blockIdx.x * COLS_PER_CTA;      // slab fastest: CTAs sharing rows run together
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  for (int r0 = 0; r0 < rows_per_cta; r0 += 2 * BLK_ROWS) {          // 64 rows per trip: 8 loads in flight per thread
    uint4 va[8], vb[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int row = r0 + w * 8 + rr;
      const int64_t off = (int64_t(part) * rows_per_cta + row) * C + c0 + lane * 8;
      const bool ok = row < rows_per_cta;
      va[rr] = ok ? __ldg(reinterpret_cast<const uint4*>(x + off)) : make_uint4(0u, 0u, 0u, 0u);
      if (OP == 1) vb[rr] = ok ? __ldg(reinterpret_cast<const uint4*>(y + off)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      float a[8], b[8];
      unpack8(va[rr], a);
      if (OP == 1) unpack8(vb[rr], b);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s1[i] += a[i];
        s2[i] = fmaf(a[i], OP == 1 ? b[i] : a[i], s2[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[w][0][lane * 8 + i] = s1[i]; red[w][1][lane * 8 + i] = s2[i]; }
  __syncthreads();
  const int t = threadIdx.x;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { a += red[k][0][t]; b += red[k][1][t]; }
  partial[(int64_t(part) * 2 + 0) * C + c0 + t] = a;
  partial[(int64_t(part) * 2 + 1) * C + c0 + t] = b;
}

// Sum partial[blk][j][c] over blk for RED_X channels per CTA (block RED_X x RED_Y), fp64, fixed order.  Few channels per CTA:
// the partial rows sit in L2 and the reduction is latency bound, so it wants many CTAs (C / 8 = 64 .. 192) more than wide rows.
constexpr int RED_X = 8;
constexpr int RED_Y = 128;
constexpr int SEG_Y = 8;      // row groups of the segment-level (<= 64 rows) kernels
template <int NSUM>
__device__ __forceinline__ void reduce_blocks(const float* __restrict__ partial, int32_t n_blk, int32_t C, int c, double (&out)[NSUM],
                                              double (*sred)[NSUM][RED_X]) {
  double s[NSUM], u[NSUM];
#pragma unroll
  for (int j = 0; j < NSUM; ++j) { s[j] = 0.0; u[j] = 0.0; }
  int b = threadIdx.y;
  for (; b + RED_Y < n_blk; b += 2 * RED_Y) {           // two loads in flight per sum
#pragma unroll
    for (int j = 0; j < NSUM; ++j) {
      const float v0 = partial[(int64_t(b) * NSUM + j) * C + c];
      const float v1 = partial[(int64_t(b + RED_Y) * NSUM + j) * C + c];
      s[j] += double(v0);
      u[j] += double(v1);
    }
  }
  if (b < n_blk) {
#pragma unroll
    for (int j = 0; j < NSUM; ++j) s[j] += double(partial[(int64_t(b) * NSUM + j) * C + c]);
  }
#pragma unroll
  for (int j = 0; j < NSUM; ++j) sred[threadIdx.y][j][threadIdx.x] = s[j] + u[j];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < NSUM; ++j) {
      double t = 0.0;
      for (int k = 0; k < RED_Y; ++k) t += sred[k][j][threadIdx.x];
      out[j] = t;
    }
  }
}

struct BnFwdArgs {
  const float* partial;      // [n_blk][2][C]
  int32_t n_blk, C;
  float n_rows;              // frames in the minibatch (B*T)
  float eps, decay;
  const float* gamma; const float* beta;
  float* moving_mean; float* moving_var;     // updated in place (tf_block.py:20-21)
  float* mean; float* inv;                   // batch mean, rsqrt(var + eps)
  float* scale; float* shift;                // gamma*inv, beta - mean*gamma*inv   (tf.nn.batch_normalization)
};
__global__ void __launch_bounds__(1024) bn_fwd_finalize_kernel(const BnFwdArgs a) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ double sred[RED_Y][2][RED_X];
  const int c = blockIdx.x * RED_X + threadIdx.x;
  double s[2];
  reduce_blocks<2>(a.partial, a.n_blk, a.C, c, s, sred);
  if (threadIdx.y != 0) return;
  const double mean = s[0] / double(a.n_rows);
  double var = s[1] / double(a.n_rows) - mean * mean;       // population variance (tf.nn.moments)
  if (var < 0.0) var = 0.0;
  const double inv = 1.0 / sqrt(var + double(a.eps));
  const double sc = double(a.gamma[c]) * inv;
  a.mean[c] = float(mean);
  a.inv[c] = float(inv);
  a.scale[c] = float(sc);
  a.shift[c] = float(double(a.beta[c]) - mean * sc);
  a.moving_mean[c] = float(double(a.moving_mean[c]) * double(a.decay) + mean * (1.0 - double(a.decay)));
  a.moving_var[c] = float(double(a.moving_var[c]) * double(a.decay) + var * (1.0 - double(a.decay)));
}

// y = valid ? r * scale + shift : 0      (8 channels per thread)
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __half* __restrict__ r, const uint8_t* __restrict__ row_valid, const float* __restrict__ scale,
                const float* __restrict__ shift, int64_t n8, int32_t c8, __half* __restrict__ y, uint32_t* overflow_flag) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  // 4 pieces of 8 channels per thread, loads issued before use
  uint4 in[4];
  uint8_t ok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = (int64_t(blockIdx.x) * 4 + k) * blockDim.x + threadIdx.x;
    ok[k] = 0;
    in[k] = make_uint4(0u, 0u, 0u, 0u);
    if (i < n8) {
      ok[k] = row_valid[i / c8];
      in[k] = __ldg(reinterpret_cast<const uint4*>(r) + i);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t i = (int64_t(blockIdx.x) * 4 + k) * blockDim.x + threadIdx.x;
    if (i >= n8) continue;
    const int c = int(i % c8) * 8;
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (ok[k]) {
      float v[8];
      unpack8(in[k], v);
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
      const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
      v[0] = fmaf(v[0], s0.x, h0.x); v[1] = fmaf(v[1], s0.y, h0.y); v[2] = fmaf(v[2], s0.z, h0.z); v[3] = fmaf(v[3], s0.w, h0.w);
      v[4] = fmaf(v[4], s1.x, h1.x); v[5] = fmaf(v[5], s1.y, h1.y); v[6] = fmaf(v[6], s1.z, h1.z); v[7] = fmaf(v[7], s1.w, h1.w);
      float mx = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) mx = fmaxf(mx, fabsf(v[q]));
      if (!(mx <= 65504.f)) atomicOr(overflow_flag, 1u);
      out = pack8(v);
    }
    reinterpret_cast<uint4*>(y)[i] = out;
  }
}

// evaluation branch: scale = gamma * rsqrt(moving_var + eps), shift = beta - moving_mean * scale  (tf_block.py:26)
__global__ void __launch_bounds__(256)
bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
               const float* __restrict__ var, float eps, int32_t C, float* __restrict__ scale, float* __restrict__ shift) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float inv = (1.0f / sqrtf(var[c] + eps)) * gamma[c];
  scale[c] = inv;
  shift[c] = beta[c] - mean[c] * inv;
}

struct BnBwdArgs {
  const float* partial;      // [n_blk][2][C]: sum dy, sum dy*r   (scaled by S)
  int32_t n_blk, C;
  float n_rows, inv_loss_scale;
  const float* gamma; const float* mean; const float* inv;
  float* cA; float* cB; float* cC;           // dz = (r > 0) * (cA*dy + cB*r + cC)
  float* d_gamma; float* d_beta;             // unscaled parameter gradients
};
__global__ void __launch_bounds__(1024) bn_bwd_finalize_kernel(const BnBwdArgs a) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ double sred[RED_Y][2][RED_X];
  const int c = blockIdx.x * RED_X + threadIdx.x;
  double s[2];
  reduce_blocks<2>(a.partial, a.n_blk, a.C, c, s, sred);
  if (threadIdx.y != 0) return;
  const double mu = a.mean[c], inv = a.inv[c], g = a.gamma[c], n = a.n_rows;
  const double dbeta = s[0];
  const double dgamma = inv * (s[1] - mu * s[0]);            // sum dy * (r - mu) * inv
  const double cA = g * inv;
  const double cB = -g * inv * inv * dgamma / n;
  a.cA[c] = float(cA);
  a.cB[c] = float(cB);
  a.cC[c] = float(-cA * dbeta / n - cB * mu);
  a.d_gamma[c] = float(dgamma * double(a.inv_loss_scale));
  a.d_beta[c] = float(dbeta * double(a.inv_loss_scale));
}

// dz = (r > 0) ? cA*dy + cB*r + cC : 0 for rows_per_cta rows x 256 channels; partial1[part][c] = sum of dz (bias gradient)
__global__ void __launch_bounds__(256)
bn_relu_bwd_kernel(const __half* __restrict__ dy, const __half* __restrict__ r, int32_t C, int32_t rows_per_cta, const float* __restrict__ cA,
                   const float* __restrict__ cB, const float* __restrict__ cC, __half* __restrict__ dz,
                   float* __restrict__ partial1, uint32_t* overflow_flag, float neg_slope, const uint8_t* __restrict__ row_valid) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ float red[8][COLS_PER_CTA];
  const int part = blockIdx.y, c0 = blockIdx.x * COLS_PER_CTA;      // slab fastest: CTAs sharing rows run together
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = c0 + lane * 8;
  float ka[8], kb[8], kc[8], s[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(cA + c)), a1 = __ldg(reinterpret_cast<const float4*>(cA + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(cB + c)), b1 = __ldg(reinterpret_cast<const float4*>(cB + c + 4));
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(cC + c)), g1 = __ldg(reinterpret_cast<const float4*>(cC + c + 4));
    ka[0] = a0.x; ka[1] = a0.y; ka[2] = a0.z; ka[3] = a0.w; ka[4] = a1.x; ka[5] = a1.y; ka[6] = a1.z; ka[7] = a1.w;
    kb[0] = b0.x; kb[1] = b0.y; kb[2] = b0.z; kb[3] = b0.w; kb[4] = b1.x; kb[5] = b1.y; kb[6] = b1.z; kb[7] = b1.w;
    kc[0] = g0.x; kc[1] = g0.y; kc[2] = g0.z; kc[3] = g0.w; kc[4] = g1.x; kc[5] = g1.y; kc[6] = g1.z; kc[7] = g1.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  float mx = 0.f;
  for (int r0 = 0; r0 < rows_per_cta; r0 += BLK_ROWS) {
    uint4 vg[4], va[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {                                     // 8 loads in flight per thread
      const int64_t off = (int64_t(part) * rows_per_cta + r0 + w * 4 + rr) * C + c;
      vg[rr] = __ldg(reinterpret_cast<const uint4*>(dy + off));
      va[rr] = __ldg(reinterpret_cast<const uint4*>(r + off));
    }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int64_t off = (int64_t(part) * rows_per_cta + r0 + w * 4 + rr) * C + c;
      float g[8], a[8], o[8];
      unpack8(vg[rr], g);
      unpack8(va[rr], a);
      // gap rows carry no gradient (r = 0 there: relu masks them by itself, a leaky slope would not)
      const float live = row_valid[int64_t(part) * rows_per_cta + r0 + w * 4 + rr] ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // activation backward: slope 1 where the output is positive, else neg_slope (0 = relu, 0.2 = tf.nn.leaky_relu)
        o[i] = fmaf(ka[i], g[i], fmaf(kb[i], a[i], kc[i])) * (a[i] > 0.f ? live : neg_slope * live);
        s[i] += o[i];
        mx = fmaxf(mx, fabsf(o[i]));
      }
      *reinterpret_cast<uint4*>(dz + off) = pack8(o);
    }
  }
  if (!(mx <= 65504.f)) atomicOr(overflow_flag, 1u);
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = s[i];
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
  partial1[int64_t(part) * C + c0 + threadIdx.x] = t;
}

// out[c] = scale * sum over blocks of partial1[blk][c]
__global__ void __launch_bounds__(1024)
colsum_finalize_kernel(const float* __restrict__ partial1, int32_t n_blk, int32_t C, float scale, float* __restrict__ out) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ double sred[RED_Y][1][RED_X];
  const int c = blockIdx.x * RED_X + threadIdx.x;
  double s[1];
  reduce_blocks<1>(partial1, n_blk, C, c, s, sred);
  if (threadIdx.y == 0) out[c] = float(s[0] * double(scale));
}

// ---------------------------------------------------------------------------------------------------------
// Statistics pooling of a training minibatch (all segments have seg_len rows, seg_stride packed rows each),
// fused with the last frame layer's BatchNorm: y = r*scale + shift per channel, so
//   mean_t(y) = scale*mean_t(r) + shift,  var_t(y) = scale^2 * var_t(r)        (models.py:485-486)
struct PoolFwdArgs {
  const float* partial;      // [n_blk][2][C] block sums of r (the ReLU output of the last frame layer)
  int32_t C, n_seg, blks_per_seg;
  float seg_len, var_eps;
  const float* scale; const float* shift;
  float* m_r; float* v_r;    // [n_seg][C] per-segment mean / population variance of r
  float* h0;                 // [n_seg][2C]  [mean | sqrt(var + 1e-5)]
};
__global__ void __launch_bounds__(256) pool_train_fwd_kernel(const PoolFwdArgs a) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = blockIdx.y;
  if (c >= a.C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int b = 0; b < a.blks_per_seg; ++b) {
    const int64_t blk = int64_t(seg) * a.blks_per_seg + b;
    s1 += double(a.partial[(blk * 2 + 0) * a.C + c]);
    s2 += double(a.partial[(blk * 2 + 1) * a.C + c]);
  }
  const double m = s1 / double(a.seg_len);
  double v = s2 / double(a.seg_len) - m * m;
  if (v < 0.0) v = 0.0;
  const double sc = a.scale[c], sh = a.shift[c];
  a.m_r[int64_t(seg) * a.C + c] = float(m);
  a.v_r[int64_t(seg) * a.C + c] = float(v);
  a.h0[int64_t(seg) * 2 * a.C + c] = float(sc * m + sh);
  a.h0[int64_t(seg) * 2 * a.C + a.C + c] = float(sqrt(sc * sc * v + double(a.var_eps)));
}

// Backward of pooling + the last layer's BatchNorm, per channel.  With dmean, dstd = dL/d[mean_y | std_y] (fp32):
//   dy[b,t,c] = a_bc + g_bc * r,   a = dmean/T - dstd*scale*m_r/(T*std_y),   g = dstd*scale/(T*std_y)
// and BatchNorm backward (batch statistics over N = B*T rows)
//   dr = gamma*inv*(dy - dbeta/N - (r-mu)*inv*dgamma/N)
// everything stays affine in r:   dz = (r > 0) * (A_bc + G_bc * r)   (times the loss scale S).
struct PoolBwdArgs {
  const float* dh0;          // [n_seg][2C]
  const float* m_r; const float* v_r;
  int32_t C, n_seg;
  float seg_len, var_eps, loss_scale;
  const float* gamma; const float* mean; const float* inv; const float* scale;
  float* coefA; float* coefG;        // [n_seg][C]
  float* d_gamma; float* d_beta;     // [C]
};
__global__ void __launch_bounds__(256) pool_bwd_coef_kernel(const PoolBwdArgs p) {       // block (32, SEG_Y)
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ double sred[SEG_Y][2][32];
  __shared__ double s_tot[2][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool live = c < p.C;
  const double T = p.seg_len, N = T * p.n_seg;
  const double sc = live ? p.scale[c] : 0.0, mu = live ? p.mean[c] : 0.0, inv = live ? p.inv[c] : 0.0, gam = live ? p.gamma[c] : 0.0;
  double dbeta = 0.0, dgamma = 0.0;
  if (live) {
    for (int b = threadIdx.y; b < p.n_seg; b += SEG_Y) {
      const double m = p.m_r[int64_t(b) * p.C + c], v = p.v_r[int64_t(b) * p.C + c];
      const double dm = p.dh0[int64_t(b) * 2 * p.C + c], ds = p.dh0[int64_t(b) * 2 * p.C + p.C + c];
      const double std_y = sqrt(sc * sc * v + double(p.var_eps));
      const double g = ds * sc / (T * std_y);
      const double a = dm / T - g * m;
      dbeta += T * a + g * T * m;
      dgamma += a * (T * m - T * mu) + g * (T * (v + m * m) - mu * T * m);
    }
  }
  sred[threadIdx.y][0][threadIdx.x] = dbeta;
  sred[threadIdx.y][1][threadIdx.x] = dgamma;
  __syncthreads();
  if (threadIdx.y == 0) {
    dbeta = 0.0; dgamma = 0.0;
    for (int k = 0; k < SEG_Y; ++k) { dbeta += sred[k][0][threadIdx.x]; dgamma += sred[k][1][threadIdx.x]; }
    s_tot[0][threadIdx.x] = dbeta;
    s_tot[1][threadIdx.x] = dgamma * inv;
  }
  __syncthreads();
  if (!live) return;
  dbeta = s_tot[0][threadIdx.x]; dgamma = s_tot[1][threadIdx.x];
  for (int b = threadIdx.y; b < p.n_seg; b += SEG_Y) {
    const double m = p.m_r[int64_t(b) * p.C + c], v = p.v_r[int64_t(b) * p.C + c];
    const double dm = p.dh0[int64_t(b) * 2 * p.C + c], ds = p.dh0[int64_t(b) * 2 * p.C + p.C + c];
    const double std_y = sqrt(sc * sc * v + double(p.var_eps));
    const double g = ds * sc / (T * std_y);
    const double a = dm / T - g * m;
    p.coefA[int64_t(b) * p.C + c] = float(gam * inv * (a - dbeta / N + mu * inv * dgamma / N) * double(p.loss_scale));
    p.coefG[int64_t(b) * p.C + c] = float(gam * inv * (g - inv * dgamma / N) * double(p.loss_scale));
  }
  if (threadIdx.y == 0) {
    p.d_gamma[c] = float(dgamma);
    p.d_beta[c] = float(dbeta);
  }
}

// dz4 = (r4 > 0) ? A[seg,c] + G[seg,c]*r4 : 0 for one segment (seg_stride packed rows) x 256 channels per CTA;
// partial1[seg][c] = column sums of dz4 (bias gradient, scaled by S)
__global__ void __launch_bounds__(256)
pool_relu_bwd_kernel(const __half* __restrict__ r, int32_t C, int32_t seg_stride, int32_t seg_len, const float* __restrict__ coefA,
                     const float* __restrict__ coefG, __half* __restrict__ dz, float* __restrict__ partial1, uint32_t* overflow_flag,
                     float neg_slope) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ float red[8][COLS_PER_CTA];
  const int seg = blockIdx.y, c0 = blockIdx.x * COLS_PER_CTA;       // slab fastest: CTAs sharing rows run together
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = c0 + lane * 8;
  float ka[8], kg[8], s[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(coefA + int64_t(seg) * C + c)), a1 = __ldg(reinterpret_cast<const float4*>(coefA + int64_t(seg) * C + c + 4));
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(coefG + int64_t(seg) * C + c)), g1 = __ldg(reinterpret_cast<const float4*>(coefG + int64_t(seg) * C + c + 4));
    ka[0] = a0.x; ka[1] = a0.y; ka[2] = a0.z; ka[3] = a0.w; ka[4] = a1.x; ka[5] = a1.y; ka[6] = a1.z; ka[7] = a1.w;
    kg[0] = g0.x; kg[1] = g0.y; kg[2] = g0.z; kg[3] = g0.w; kg[4] = g1.x; kg[5] = g1.y; kg[6] = g1.z; kg[7] = g1.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.f;
  float mx = 0.f;
  for (int r0 = 0; r0 < seg_stride; r0 += 2 * BLK_ROWS) {
    uint4 va[8];
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {                                     // 8 loads in flight per thread
      const int row = r0 + w * 8 + rr;
      va[rr] = row < seg_stride ? __ldg(reinterpret_cast<const uint4*>(r + (int64_t(seg) * seg_stride + row) * C + c)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int rr = 0; rr < 8; ++rr) {
      const int row = r0 + w * 8 + rr;
      if (row >= seg_stride) continue;
      float a[8], o[8];
      unpack8(va[rr], a);
      const float live = row < seg_len ? 1.f : 0.f;             // gap rows carry no gradient (with a leaky slope A would leak into them)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        o[i] = fmaf(kg[i], a[i], ka[i]) * (a[i] > 0.f ? 1.f : neg_slope) * live;
        s[i] += o[i];
        mx = fmaxf(mx, fabsf(o[i]));
      }
      *reinterpret_cast<uint4*>(dz + (int64_t(seg) * seg_stride + row) * C + c) = pack8(o);
    }
  }
  if (!(mx <= 65504.f)) atomicOr(overflow_flag, 1u);
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][lane * 8 + i] = s[i];
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
  partial1[int64_t(seg) * C + c0 + threadIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------------------
// Segment level (64 rows): fp32 SIMT GEMM  C[m,n] = sum_k A(m,k) * B(k,n) (+ bias[n]) with arbitrary element strides,
// 64 x 64 tile, split over K across gridDim.z (partials reduced in a fixed order by splitk_reduce_kernel).
struct SgemmArgs {
  const float* A; const float* B; float* C; const float* bias;
  int32_t M, N, K;
  int64_t sam, sak, sbk, sbn;
  int32_t ldc;
  int32_t k_per_split;         // multiple of 16
  float* partial;              // [gridDim.z][M][N] when gridDim.z > 1
};
template <bool A_KFAST, bool B_NFAST>
__global__ void __launch_bounds__(256) sgemm64_kernel(const SgemmArgs a) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int k0 = blockIdx.z * a.k_per_split, k1 = min(a.K, k0 + a.k_per_split);
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  auto gload = [&](int kb) {                       // this thread's 4 + 4 elements of the K tile starting at kb -> registers
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      {
        const int kk = A_KFAST ? (idx & 15) : (idx >> 6), mm = A_KFAST ? (idx >> 4) : (idx & 63);
        const int gm = m0 + mm, gk = kb + kk;
        ra[e] = (gm < a.M && gk < k1) ? __ldg(a.A + int64_t(gm) * a.sam + int64_t(gk) * a.sak) : 0.f;
      }
      {
        const int kk = B_NFAST ? (idx >> 6) : (idx & 15), nn = B_NFAST ? (idx & 63) : (idx >> 4);
        const int gn = n0 + nn, gk = kb + kk;
        rb[e] = (gn < a.N && gk < k1) ? __ldg(a.B + int64_t(gk) * a.sbk + int64_t(gn) * a.sbn) : 0.f;
      }
    }
  };
  gload(k0);
  for (int kb = k0; kb < k1; kb += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      As[A_KFAST ? (idx & 15) : (idx >> 6)][A_KFAST ? (idx >> 4) : (idx & 63)] = ra[e];
      Bs[B_NFAST ? (idx >> 6) : (idx & 15)][B_NFAST ? (idx & 63) : (idx >> 4)] = rb[e];
    }
    __syncthreads();
    if (kb + 16 < k1) gload(kb + 16);             // the next tile's loads fly during this tile's FMAs
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= a.N) continue;
      if (gridDim.z == 1) a.C[int64_t(gm) * a.ldc + gn] = acc[i][j] + (a.bias ? __ldg(a.bias + gn) : 0.f);
      else a.partial[(int64_t(blockIdx.z) * a.M + gm) * a.N + gn] = acc[i][j];
    }
  }
}
// C = sum of the K-split partials (fixed order) + bias
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int32_t splits, int32_t M, int32_t N, const float* __restrict__ bias,
                     float* __restrict__ C, int32_t ldc) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= int64_t(M) * N) return;
  const int m = int(i / N), n = int(i - int64_t(m) * N);
  float s = partial[i];
  for (int k = 1; k < splits; ++k) s += partial[int64_t(k) * M * N + i];
  C[int64_t(m) * ldc + n] = s + (bias ? __ldg(bias + n) : 0.f);
}

// r = relu(z); BatchNorm training branch over the B rows (tf_block.py:18-23); one thread per channel.
struct SegBnArgs {
  const float* z; int32_t B, C;
  float eps, decay;
  const float* gamma; const float* beta;
  float* moving_mean; float* moving_var;
  float* r; float* y; float* mean; float* inv;
  int32_t training;          // 0: evaluation branch (moving statistics, no update; tf_block.py:25-26)
  float neg_slope;           // 0 = relu, 0.2 = tf.nn.leaky_relu (models.py:912)
};
__global__ void __launch_bounds__(256) seg_relu_bn_fwd_kernel(const SegBnArgs a) {     // block (32, SEG_Y)
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ double sred[SEG_Y][2][32];
  __shared__ float s_sc[32], s_sh[32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool live = c < a.C;
  double s1 = 0.0, s2 = 0.0;
  if (live) {
    for (int b = threadIdx.y; b < a.B; b += SEG_Y) {
      const float z = a.z[int64_t(b) * a.C + c];
      const float r = z > 0.f ? z : a.neg_slope * z;
      a.r[int64_t(b) * a.C + c] = r;
      s1 += r; s2 += double(r) * r;
    }
  }
  sred[threadIdx.y][0][threadIdx.x] = s1;
  sred[threadIdx.y][1][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    float sc, sh;
    if (a.training) {
      s1 = 0.0; s2 = 0.0;
      for (int k = 0; k < SEG_Y; ++k) { s1 += sred[k][0][threadIdx.x]; s2 += sred[k][1][threadIdx.x]; }
      const double mean = s1 / a.B;
      double var = s2 / a.B - mean * mean;
      if (var < 0.0) var = 0.0;
      const double inv = 1.0 / sqrt(var + double(a.eps));
      sc = float(double(a.gamma[c]) * inv);
      sh = float(double(a.beta[c]) - mean * double(a.gamma[c]) * inv);
      a.mean[c] = float(mean);
      a.inv[c] = float(inv);
      a.moving_mean[c] = float(double(a.moving_mean[c]) * double(a.decay) + mean * (1.0 - double(a.decay)));
      a.moving_var[c] = float(double(a.moving_var[c]) * double(a.decay) + var * (1.0 - double(a.decay)));
    } else {                                   // evaluation branch: moving statistics, nothing updated (tf_block.py:25-26)
      sc = a.gamma[c] * (1.0f / sqrtf(a.moving_var[c] + a.eps));
      sh = a.beta[c] - a.moving_mean[c] * sc;
    }
    s_sc[threadIdx.x] = sc;
    s_sh[threadIdx.x] = sh;
  }
  __syncthreads();
  if (!live) return;
  const float sc = s_sc[threadIdx.x], sh = s_sh[threadIdx.x];
  for (int b = threadIdx.y; b < a.B; b += SEG_Y) a.y[int64_t(b) * a.C + c] = fmaf(a.r[int64_t(b) * a.C + c], sc, sh);
}
struct SegBnBwdArgs {
  const float* dy; const float* r; int32_t B, C;
  const float* gamma; const float* mean; const float* inv;
  float* dz; float* d_gamma; float* d_beta; float* d_bias;
  float neg_slope;
};
__global__ void __launch_bounds__(256) seg_relu_bn_bwd_kernel(const SegBnBwdArgs a) {   // block (32, SEG_Y)
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ double sred[SEG_Y][2][32];
  __shared__ double s_tot[2][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool live = c < a.C;
  const double mu = live ? a.mean[c] : 0.0, inv = live ? a.inv[c] : 0.0, g = live ? a.gamma[c] : 0.0;
  double dbeta = 0.0, dgamma = 0.0;
  if (live) {
    for (int b = threadIdx.y; b < a.B; b += SEG_Y) {
      const double dy = a.dy[int64_t(b) * a.C + c], rh = (double(a.r[int64_t(b) * a.C + c]) - mu) * inv;
      dbeta += dy; dgamma += dy * rh;
    }
  }
  sred[threadIdx.y][0][threadIdx.x] = dbeta;
  sred[threadIdx.y][1][threadIdx.x] = dgamma;
  __syncthreads();
  if (threadIdx.y == 0) {
    dbeta = 0.0; dgamma = 0.0;
    for (int k = 0; k < SEG_Y; ++k) { dbeta += sred[k][0][threadIdx.x]; dgamma += sred[k][1][threadIdx.x]; }
    s_tot[0][threadIdx.x] = dbeta;
    s_tot[1][threadIdx.x] = dgamma;
  }
  __syncthreads();
  dbeta = s_tot[0][threadIdx.x]; dgamma = s_tot[1][threadIdx.x];
  double db = 0.0;
  if (live) {
    for (int b = threadIdx.y; b < a.B; b += SEG_Y) {
      const double r = a.r[int64_t(b) * a.C + c];
      const double dy = a.dy[int64_t(b) * a.C + c], rh = (r - mu) * inv;
      const double dr = g * inv * (dy - dbeta / a.B - rh * dgamma / a.B);
      const float dz = float(dr) * (r > 0.0 ? 1.f : a.neg_slope);
      a.dz[int64_t(b) * a.C + c] = dz;
      db += dz;
    }
  }
  __syncthreads();
  sred[threadIdx.y][0][threadIdx.x] = db;
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    db = 0.0;
    for (int k = 0; k < SEG_Y; ++k) db += sred[k][0][threadIdx.x];
    a.d_gamma[c] = float(dgamma);
    a.d_beta[c] = float(dbeta);
    a.d_bias[c] = float(db);
  }
}

// Softmax cross-entropy of one row per CTA (models.py:512): loss_row, correct (argmax == label, first maximum as
// tf.argmax), dlogits = (softmax - onehot) / B  (the gradient of reduce_mean, models.py:514).
__global__ void __launch_bounds__(256)
softmax_ce_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, int32_t n_classes, float inv_batch,
                  float* __restrict__ dlogits, float* __restrict__ loss_row, float* __restrict__ correct) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ float s_val[256];
  __shared__ int s_idx[256];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* row = logits + int64_t(b) * n_classes;
  float mx = -INFINITY; int mi = 0x7fffffff;
  for (int j = t; j < n_classes; j += 256) { const float v = row[j]; if (v > mx) { mx = v; mi = j; } }
  s_val[t] = mx; s_idx[t] = mi;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (t < s) {
      const float v = s_val[t + s]; const int i = s_idx[t + s];
      if (v > s_val[t] || (v == s_val[t] && i < s_idx[t])) { s_val[t] = v; s_idx[t] = i; }
    }
    __syncthreads();
  }
  mx = s_val[0]; mi = s_idx[0];
  __syncthreads();
  float se = 0.f;
  for (int j = t; j < n_classes; j += 256) se += __expf(row[j] - mx);
  s_val[t] = se;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) { if (t < s) s_val[t] += s_val[t + s]; __syncthreads(); }
  se = s_val[0];
  const int lab = labels[b];
  const float inv_se = 1.f / se;
  for (int j = t; j < n_classes; j += 256)
    dlogits[int64_t(b) * n_classes + j] = (__expf(row[j] - mx) * inv_se - (j == lab ? 1.f : 0.f)) * inv_batch;
  if (t == 0) {
    loss_row[b] = logf(se) + mx - row[lab];
    correct[b] = mi == lab ? 1.f : 0.f;
  }
}
__global__ void loss_finalize_kernel(const float* loss_row, const float* correct, int32_t B, float* out) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double l = 0.0, c = 0.0;
  for (int b = 0; b < B; ++b) { l += loss_row[b]; c += correct[b]; }
  out[0] = float(l / B);
  out[1] = float(c / B);
}
// out[n] = sum over the B rows of x[b][n]
__global__ void __launch_bounds__(256) colsum_rows_kernel(const float* __restrict__ x, int32_t B, int32_t N, float* __restrict__ out) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += x[int64_t(b) * N + n];
  out[n] = s;
}

// grad[n_params .. +4) = this step's gradient-overflow flag as a float (0 / 1): the tail the all-reduce carries
__global__ void grad_flag_kernel(const uint32_t* __restrict__ flag, float* __restrict__ tail) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  if (threadIdx.x < 4) tail[threadIdx.x] = (threadIdx.x == 0 && *flag != 0u) ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------------------
// tf.train.AdamOptimizer._apply_dense: m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= lr_t * m / (sqrt(v) + eps)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr_t, float b1, float b2, float eps, float grad_scale, const float* __restrict__ combined_flag,
            uint32_t* __restrict__ skipped) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  // a loss-scaled fp16 gradient overflowed somewhere in this step -- on ANY rank: the flag rides behind the gradient
  // (grad[n_params], written by grad_flag_kernel) and is summed by the same all-reduce, so every replica takes the same
  // decision.  The gradient is not trustworthy: leave variables and slots alone (what a dynamic loss scaler does) and
  // count the skip -- the caller lowers the loss scale.  (NaN != 0 is true: a poisoned flag skips as well.)
  if (*combined_flag != 0.f) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(skipped, 1u);
    return;
  }
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  p[i] -= lr_t * mi / (sqrtf(vi) + eps);
}

// fp32 master conv weights W[taps][c_in][c_out] -> the two fp16 operand copies, one 32 x 32 tile per CTA, all frame
// layers in ONE launch (grid = total tiles, block (32, 8)):
//   forward       wf[c_out][k_total], K index = tap*c_in_pad + ci          (transposed through shared memory)
//   data gradient wd[c_in][tap'*c_out + co] = W[taps-1-tap'][c_in][co]     (the conv of dz with the flipped kernel; may be null)
struct RepackLayer {
  const float* W; __half* wf; __half* wd;
  int32_t taps, c_in, c_out, c_in_pad, k_total;
  int32_t tile_begin;        // first tile index of this layer; tiles = (c_out/32) * ceil(c_in/32) * taps
};
struct RepackTable { RepackLayer layer[8]; int32_t n_layers; };
__global__ void __launch_bounds__(256) repack_kernel(const RepackTable tb) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  __shared__ float tile[32][33];
  int li = 0;
  while (li + 1 < tb.n_layers && int(blockIdx.x) >= tb.layer[li + 1].tile_begin) ++li;
  const RepackLayer& L = tb.layer[li];
  int tix = int(blockIdx.x) - L.tile_begin;
  const int n_o = L.c_out / 32, n_c = (L.c_in + 31) / 32;
  const int o0 = (tix % n_o) * 32;
  tix /= n_o;
  const int c0 = (tix % n_c) * 32, j = tix / n_c;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const float w = c < L.c_in ? L.W[(int64_t(j) * L.c_in + c) * L.c_out + o0 + threadIdx.x] : 0.f;
    tile[i][threadIdx.x] = w;
    if (L.wd != nullptr && c < L.c_in) L.wd[(int64_t(c) * L.taps + (L.taps - 1 - j)) * L.c_out + o0 + threadIdx.x] = __float2half_rn(w);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + threadIdx.x;
    if (c < L.c_in) L.wf[int64_t(o0 + i) * L.k_total + int64_t(j) * L.c_in_pad + c] = __float2half_rn(tile[threadIdx.x][i]);
  }
}

// L2 term of the ModelL2Loss* graphs (models.py:930-951,962: loss += beta * coef * tf.nn.l2_loss(p) = beta*coef*sum(p^2)/2):
// grad += beta*coef*p and loss_acc[0] += beta*coef*sum(p^2)/2.  ONE CTA per tensor (fixed-order reduction).
__global__ void __launch_bounds__(1024)
l2_term_kernel(const float* __restrict__ p, float* __restrict__ g, int64_t n, float coef, float* __restrict__ loss_acc) {
  __shared__ double red[1024];
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float v = p[i];
    g[i] = fmaf(coef, v, g[i]);
    s += double(v) * v;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) { if (int(threadIdx.x) < w) red[threadIdx.x] += red[threadIdx.x + w]; __syncthreads(); }
  if (threadIdx.x == 0) loss_acc[0] += float(0.5 * double(coef) * red[0]);
}

// Segment metadata of a training minibatch (n_seg segments of seg_len rows, seg_stride packed rows each), written on the
// device so that nothing is staged from the host when the minibatch geometry changes (the layout stage_meta builds for the
// extraction path): meta = [row_start(n_seg) | feat_start(n_seg) | len(n_seg) | pad to int4 | blk_info(r_pad/32 x int4)],
// blk_info = {first feature row, first frame of the block in its segment, segment length, valid rows}.
__global__ void __launch_bounds__(256)
train_meta_kernel(int32_t* __restrict__ meta, int32_t n_seg, int32_t seg_len, int32_t seg_stride, int32_t n_blk, int32_t blk_info_off) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_seg) {
    meta[i] = i * seg_stride;
    meta[n_seg + i] = i * seg_len;
    meta[2 * n_seg + i] = seg_len;
  }
  if (i < n_blk) {
    const int bps = seg_stride / 32, seg = i / bps, t0 = (i % bps) * 32;
    int4 e = make_int4(0, 0, 0, 0);
    if (seg < n_seg) e = make_int4(seg * seg_len + t0, t0, seg_len, max(0, min(32, seg_len - t0)));
    reinterpret_cast<int4*>(meta + blk_info_off)[i] = e;
  }
}

// float16 minibatch as stored in the egs archives (examples_io.py:165) -> float32 features, 8 values per thread
__global__ void __launch_bounds__(256) f16_to_f32_kernel(const __half* __restrict__ src, float* __restrict__ dst, int64_t n) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(src + i)), f);
    reinterpret_cast<float4*>(dst + i)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(dst + i)[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    for (int64_t k = i; k < n; ++k) dst[k] = __half2float(src[k]);
  }
}

__global__ void __launch_bounds__(256) fill_kernel(float* p, int64_t n, float v) {
  cudaTriggerProgrammaticLaunchCompletion();   // PDL: dependents may be scheduled; they wait in cudaGridDependencySynchronize()
  cudaGridDependencySynchronize();
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace trk
