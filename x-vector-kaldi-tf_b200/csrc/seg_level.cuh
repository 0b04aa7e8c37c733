// seg_level.cuh -- the whole segment level of one training minibatch in ONE persistent cooperative kernel (sm_100a).
//
// Everything of sess.run([optimizer, loss, accuracy]) (local/tf/models.py:263) between the pooled statistics h0 [B, 2C]
// and their gradient dh0: embed_layer-0 / embed_layer-1 (xw_plus_b -> relu -> BatchNorm training branch,
// models.py:489-499, tf_block.py:18-23), the output layer (models.py:502-508), softmax cross-entropy / accuracy
// (models.py:512-523) and their backward.  B is a minibatch of ~64 rows, so these are ~1 GFLOP of fp32 work spread over
// 11 dependent steps.  train_api.cuh runs them as 22 separate launches chained by programmatic dependent launch (the default);
// this kernel is the alternative (option "seg_fused" = 1): one CTA per SM stays resident and the steps are separated by
// grid-wide barriers (cooperative launch).  Measured on B200 it is NOT faster (0.226 ms against ~0.2 ms): the K loops are
// bound by L2 latency with one tile in flight per CTA.  It is kept as an option and as a cross-check of the default path:
//
//   F1 z5 = h0 W0 (K-split partials)   F2 +b0, relu, BN -> y5        F3 z6 = y5 W1        F4 -> y6
//   F5 logits = y6 Wo                  F6 +bo, softmax CE, dlogits, per-row loss / hit
//   B1 dWo = y6^T dlogits, dbo, dy6 = dlogits Wo^T (partials), loss / accuracy
//   B2 BN + relu backward -> dz6, dgamma, dbeta, db1                  B3 dW1 = y5^T dz6, dy5 = dz6 W1^T (partials)
//   B4 -> dz5, ...                                                    B5 dW0 = h0^T dz5, dh0 = dz5 W0^T
//
// GEMM tiles are 64 x 64 x 16 fp32 SIMT (the body of sgemm64_kernel) handed out round-robin to the CTAs; K-splits are
// added in a fixed order by the step that consumes them, so the result does not depend on scheduling (bit-reproducible).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace segk {

namespace cg = cooperative_groups;

struct Layer {            // xw_plus_b -> relu -> BatchNorm
  const float* W; const float* b; const float* gamma; const float* beta;     // [in, out], [out] x3
  float* moving_mean; float* moving_var;
  float* z; float* r; float* y; float* dy; float* dz;                         // [B, out]
  float* mean; float* inv;                                                    // batch statistics [out]
  float* gW; float* gb; float* ggamma; float* gbeta;                          // gradients
  int32_t in, out;
};

struct Args {
  int32_t B, NC;
  float eps, decay;
  const float* h0; float* dh0;                 // [B, L[0].in]
  Layer L[2];
  const float* Wo; const float* bo; float* gWo; float* gbo;                   // [L[1].out, NC], [NC]
  float* logits; float* dlogits;               // [B, NC]
  const int32_t* labels;
  float* loss_row; float* correct; float* loss_acc;                           // [B], [B], [2]
  float* partial;                              // K-split partials of the step in flight
  int32_t splits_f1, splits_f3, splits_f5, splits_dy6, splits_dy5;            // chosen by the host (sized `partial`)
};

struct Smem {
  float As[16][68];
  float Bs[16][68];
  double red[8][2][32];
  double tot[2][32];
  float sc[32], sh[32];
  float val[256];
  int idx[256];
};

// One 64 x 64 tile of C = A B over K range [k0, k1); A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].
// out: plain store of acc (+ bias) at C[m*ldc + n] for m < M, n < N.
__device__ __forceinline__ void gemm_tile(Smem& s, const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ Bm,
                                          int64_t sbk, int64_t sbn, int M, int N, int m0, int n0, int k0, int k1, float* __restrict__ C,
                                          int64_t ldc, const float* __restrict__ bias) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const bool a_kfast = sak == 1, b_nfast = sbn == 1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  auto gload = [&](int kb) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      {
        const int kk = a_kfast ? (idx & 15) : (idx >> 6), mm = a_kfast ? (idx >> 4) : (idx & 63);
        const int gm = m0 + mm, gk = kb + kk;
        ra[e] = (gm < M && gk < k1) ? __ldcg(A + int64_t(gm) * sam + int64_t(gk) * sak) : 0.f;
      }
      {
        const int kk = b_nfast ? (idx >> 6) : (idx & 15), nn = b_nfast ? (idx & 63) : (idx >> 4);
        const int gn = n0 + nn, gk = kb + kk;
        rb[e] = (gn < N && gk < k1) ? __ldcg(Bm + int64_t(gk) * sbk + int64_t(gn) * sbn) : 0.f;
      }
    }
  };
  gload(k0);
  for (int kb = k0; kb < k1; kb += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      s.As[a_kfast ? (idx & 15) : (idx >> 6)][a_kfast ? (idx >> 4) : (idx & 63)] = ra[e];
      s.Bs[b_nfast ? (idx >> 6) : (idx & 15)][b_nfast ? (idx & 63) : (idx >> 4)] = rb[e];
    }
    __syncthreads();
    if (kb + 16 < k1) gload(kb + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&s.As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&s.Bs[kk][tx * 4]);
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
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) C[int64_t(gm) * ldc + gn] = acc[i][j] + (bias ? __ldg(bias + gn) : 0.f);
    }
  }
}

// All tiles of C[M, N] (+)= A B, K cut into `splits`; split s of the result goes to out + s*M*N (ldc = N) when
// splits > 1, else straight to out (ldc given, bias added).  Items are dealt round-robin starting at CTA `first`.
__device__ __forceinline__ void gemm_phase(Smem& s, int first, const float* A, int64_t sam, int64_t sak, const float* Bm, int64_t sbk,
                                           int64_t sbn, int M, int N, int K, int splits, float* out, int64_t ldc, const float* bias) {
  const int tm = (M + 63) / 64, tn = (N + 63) / 64;
  const int kps = ((K + splits - 1) / splits + 15) / 16 * 16;
  const int n_items = tm * tn * splits;
  const int cta = (int(blockIdx.x) + int(gridDim.x) - first % int(gridDim.x)) % int(gridDim.x);
  for (int item = cta; item < n_items; item += gridDim.x) {
    const int sp = item % splits, t = item / splits;
    const int m0 = (t / tn) * 64, n0 = (t % tn) * 64;
    const int k0 = sp * kps, k1 = min(K, k0 + kps);
    if (splits > 1) gemm_tile(s, A, sam, sak, Bm, sbk, sbn, M, N, m0, n0, k0, k1, out + int64_t(sp) * M * N, N, nullptr);
    else gemm_tile(s, A, sam, sak, Bm, sbk, sbn, M, N, m0, n0, k0, k1, out, ldc, bias);
  }
}

__device__ __forceinline__ float sum_splits(const float* partial, int splits, int64_t mn, int64_t i) {
  float v = __ldcg(partial + i);
  for (int k = 1; k < splits; ++k) v += __ldcg(partial + int64_t(k) * mn + i);
  return v;
}

// z = sum of K-split partials + b; r = relu(z); BatchNorm training branch over the B rows; 32 channels per CTA pass.
__device__ __forceinline__ void bn_forward(Smem& s, const Args& a, const Layer& L, int splits) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t mn = int64_t(a.B) * L.out;
  for (int c0 = blockIdx.x * 32; c0 < L.out; c0 += gridDim.x * 32) {
    const int c = c0 + tx;
    const bool live = c < L.out;
    double s1 = 0.0, s2 = 0.0;
    if (live) {
      const float bias = __ldg(L.b + c);
      for (int b = ty; b < a.B; b += 8) {
        const int64_t i = int64_t(b) * L.out + c;
        const float z = sum_splits(a.partial, splits, mn, i) + bias;
        const float r = fmaxf(z, 0.f);
        L.z[i] = z;
        L.r[i] = r;
        s1 += r; s2 += double(r) * r;
      }
    }
    s.red[ty][0][tx] = s1; s.red[ty][1][tx] = s2;
    __syncthreads();
    if (ty == 0 && live) {
      s1 = 0.0; s2 = 0.0;
      for (int k = 0; k < 8; ++k) { s1 += s.red[k][0][tx]; s2 += s.red[k][1][tx]; }
      const double mean = s1 / a.B;
      double var = s2 / a.B - mean * mean;
      if (var < 0.0) var = 0.0;
      const double inv = 1.0 / sqrt(var + double(a.eps));
      s.sc[tx] = float(double(L.gamma[c]) * inv);
      s.sh[tx] = float(double(L.beta[c]) - mean * double(L.gamma[c]) * inv);
      L.mean[c] = float(mean);
      L.inv[c] = float(inv);
      L.moving_mean[c] = float(double(L.moving_mean[c]) * double(a.decay) + mean * (1.0 - double(a.decay)));
      L.moving_var[c] = float(double(L.moving_var[c]) * double(a.decay) + var * (1.0 - double(a.decay)));
    }
    __syncthreads();
    if (live) {
      const float sc = s.sc[tx], sh = s.sh[tx];
      for (int b = ty; b < a.B; b += 8) {
        const int64_t i = int64_t(b) * L.out + c;
        L.y[i] = fmaf(L.r[i], sc, sh);
      }
    }
    __syncthreads();
  }
}

// dy = sum of K-split partials; BatchNorm + relu backward -> dz, dgamma, dbeta, dbias; 32 channels per CTA pass.
__device__ __forceinline__ void bn_backward(Smem& s, const Args& a, const Layer& L, int splits) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t mn = int64_t(a.B) * L.out;
  for (int c0 = blockIdx.x * 32; c0 < L.out; c0 += gridDim.x * 32) {
    const int c = c0 + tx;
    const bool live = c < L.out;
    const double mu = live ? L.mean[c] : 0.0, inv = live ? L.inv[c] : 0.0, g = live ? L.gamma[c] : 0.0;
    double dbeta = 0.0, dgamma = 0.0;
    if (live) {
      for (int b = ty; b < a.B; b += 8) {
        const int64_t i = int64_t(b) * L.out + c;
        const float dy = sum_splits(a.partial, splits, mn, i);
        L.dy[i] = dy;
        dbeta += dy; dgamma += double(dy) * ((double(L.r[i]) - mu) * inv);
      }
    }
    s.red[ty][0][tx] = dbeta; s.red[ty][1][tx] = dgamma;
    __syncthreads();
    if (ty == 0) {
      dbeta = 0.0; dgamma = 0.0;
      for (int k = 0; k < 8; ++k) { dbeta += s.red[k][0][tx]; dgamma += s.red[k][1][tx]; }
      s.tot[0][tx] = dbeta; s.tot[1][tx] = dgamma;
    }
    __syncthreads();
    dbeta = s.tot[0][tx]; dgamma = s.tot[1][tx];
    double db = 0.0;
    if (live) {
      for (int b = ty; b < a.B; b += 8) {
        const int64_t i = int64_t(b) * L.out + c;
        const double r = L.r[i];
        const double dr = g * inv * (double(L.dy[i]) - dbeta / a.B - (r - mu) * inv * dgamma / a.B);
        const float dz = r > 0.0 ? float(dr) : 0.f;
        L.dz[i] = dz;
        db += dz;
      }
    }
    __syncthreads();
    s.red[ty][0][tx] = db;
    __syncthreads();
    if (ty == 0 && live) {
      db = 0.0;
      for (int k = 0; k < 8; ++k) db += s.red[k][0][tx];
      L.ggamma[c] = float(dgamma);
      L.gbeta[c] = float(dbeta);
      L.gb[c] = float(db);
    }
    __syncthreads();
  }
}

// logits row = partial sums + bo; softmax cross-entropy (models.py:512), first-maximum argmax, dlogits = (p - onehot)/B
__device__ __forceinline__ void softmax_rows(Smem& s, const Args& a, int splits) {
  const int t = threadIdx.x;
  const int64_t mn = int64_t(a.B) * a.NC;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    float* row = a.logits + int64_t(b) * a.NC;
    float mx = -INFINITY; int mi = 0x7fffffff;
    for (int j = t; j < a.NC; j += 256) {
      const float v = sum_splits(a.partial, splits, mn, int64_t(b) * a.NC + j) + __ldg(a.bo + j);
      row[j] = v;
      if (v > mx) { mx = v; mi = j; }
    }
    s.val[t] = mx; s.idx[t] = mi;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
      if (t < w) {
        const float v = s.val[t + w]; const int i = s.idx[t + w];
        if (v > s.val[t] || (v == s.val[t] && i < s.idx[t])) { s.val[t] = v; s.idx[t] = i; }
      }
      __syncthreads();
    }
    mx = s.val[0]; mi = s.idx[0];
    __syncthreads();
    float se = 0.f;
    for (int j = t; j < a.NC; j += 256) se += __expf(row[j] - mx);
    s.val[t] = se;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) { if (t < w) s.val[t] += s.val[t + w]; __syncthreads(); }
    se = s.val[0];
    const int lab = a.labels[b];
    const float inv_se = 1.f / se, inv_b = 1.f / float(a.B);
    for (int j = t; j < a.NC; j += 256)
      a.dlogits[int64_t(b) * a.NC + j] = (__expf(row[j] - mx) * inv_se - (j == lab ? 1.f : 0.f)) * inv_b;
    if (t == 0) {
      a.loss_row[b] = logf(se) + mx - row[lab];
      a.correct[b] = mi == lab ? 1.f : 0.f;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256, 3) seg_level_train_kernel(const Args a) {
  __shared__ Smem s;
  cg::grid_group grid = cg::this_grid();
  const Layer& L0 = a.L[0];
  const Layer& L1 = a.L[1];
  const int B = a.B;
  // ---- forward ----
  gemm_phase(s, 0, a.h0, L0.in, 1, L0.W, L0.out, 1, B, L0.out, L0.in, a.splits_f1, a.partial, L0.out, nullptr);
  grid.sync();
  bn_forward(s, a, L0, a.splits_f1);
  grid.sync();
  gemm_phase(s, 0, L0.y, L1.in, 1, L1.W, L1.out, 1, B, L1.out, L1.in, a.splits_f3, a.partial, L1.out, nullptr);
  grid.sync();
  bn_forward(s, a, L1, a.splits_f3);
  grid.sync();
  gemm_phase(s, 0, L1.y, L1.out, 1, a.Wo, a.NC, 1, B, a.NC, L1.out, a.splits_f5, a.partial, a.NC, nullptr);
  grid.sync();
  softmax_rows(s, a, a.splits_f5);
  grid.sync();
  // ---- backward ----
  {
    // dWo[i, j] = sum_b y6[b, i] dlogits[b, j];  dy6 = dlogits Wo^T (K-split partials);  dbo;  loss / accuracy
    const int tiles_w = ((L1.out + 63) / 64) * ((a.NC + 63) / 64);
    gemm_phase(s, 0, L1.y, 1, L1.out, a.dlogits, a.NC, 1, L1.out, a.NC, B, 1, a.gWo, a.NC, nullptr);
    gemm_phase(s, tiles_w, a.dlogits, a.NC, 1, a.Wo, 1, a.NC, B, L1.out, a.NC, a.splits_dy6, a.partial, L1.out, nullptr);
    for (int j = blockIdx.x * 256 + threadIdx.x; j < a.NC; j += gridDim.x * 256) {
      float v = 0.f;
      for (int b = 0; b < B; ++b) v += __ldcg(a.dlogits + int64_t(b) * a.NC + j);
      a.gbo[j] = v;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
      double l = 0.0, c = 0.0;
      for (int b = 0; b < B; ++b) { l += __ldcg(a.loss_row + b); c += __ldcg(a.correct + b); }
      a.loss_acc[0] = float(l / B);
      a.loss_acc[1] = float(c / B);
    }
  }
  grid.sync();
  bn_backward(s, a, L1, a.splits_dy6);
  grid.sync();
  {
    const int tiles_w = ((L1.in + 63) / 64) * ((L1.out + 63) / 64);
    gemm_phase(s, 0, L0.y, 1, L1.in, L1.dz, L1.out, 1, L1.in, L1.out, B, 1, L1.gW, L1.out, nullptr);
    gemm_phase(s, tiles_w, L1.dz, L1.out, 1, L1.W, 1, L1.out, B, L1.in, L1.out, a.splits_dy5, a.partial, L1.in, nullptr);
  }
  grid.sync();
  bn_backward(s, a, L0, a.splits_dy5);
  grid.sync();
  {
    const int tiles_w = ((L0.in + 63) / 64) * ((L0.out + 63) / 64);
    gemm_phase(s, 0, a.h0, 1, L0.in, L0.dz, L0.out, 1, L0.in, L0.out, B, 1, L0.gW, L0.out, nullptr);
    gemm_phase(s, tiles_w, L0.dz, L0.out, 1, L0.W, 1, L0.out, B, L0.in, L0.out, 1, a.dh0, L0.in, nullptr);
  }
}

}  // namespace segk
