// pack_pool.cuh -- the two HBM-bound kernels either side of the TDNN stack (sm_100a).
//
//   pack_im2col_kernel : fp32 MFCC rows [total_frames, D]  ->  fp16 "packed rows" matrix
//                        [R_pad, K0_pad] holding, per row, the taps0*D spliced input of the first
//                        layer (so layer 0 is a plain K = K0_pad GEMM), zero rows in the gaps,
//                        plus the row_valid map every later epilogue uses.
//   pool_embed_kernel  : statistics pooling (tf.nn.moments over time + sqrt(var + 1e-5), concat;
//                        reference models.py:485-486) fused with the first segment-level affine
//                        layer embed_layer-0 (tf.nn.xw_plus_b, models.py:495) = the x-vector.
//   unpack_rows_kernel : debug/parity only: fp16 packed rows -> fp32 [total_frames, C].
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xvk {

struct SegMeta {
  const int32_t* row_start;    // [n_seg] first packed row of the segment
  const int32_t* feat_start;   // [n_seg] first row of the segment in the caller's feats matrix
  const int32_t* len;          // [n_seg] rows
  int32_t n_seg;
};

// largest s with row_start[s] <= r, or -1
__device__ __forceinline__ int find_segment(const int32_t* __restrict__ row_start, int n_seg, int r) {
  int lo = 0, hi = n_seg;                      // invariant: row_start[lo-1] <= r < row_start[hi]
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(row_start + mid) <= r) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}

constexpr int PACK_ROWS_PER_BLOCK = 16;
constexpr int PACK_THREADS = 256;

struct PackArgs {
  const float* feats;        // [total_frames, D]
  SegMeta seg;
  int32_t r_pad;
  int32_t feat_dim, taps, dilation;
  int32_t k0_pad;            // multiple of 64
  __half* x0;                // [r_pad, k0_pad]
  uint8_t* row_valid;        // [r_pad]
  uint32_t* counters;        // [n_counters] zeroed here for pool_embed_kernel
  int32_t n_counters;
};

__global__ void __launch_bounds__(PACK_THREADS) pack_im2col_kernel(const PackArgs a) {
  __shared__ int s_seg[PACK_ROWS_PER_BLOCK], s_t[PACK_ROWS_PER_BLOCK];
  const int r0 = blockIdx.x * PACK_ROWS_PER_BLOCK;
  const int gtid = blockIdx.x * PACK_THREADS + threadIdx.x;
  if (gtid < a.n_counters) a.counters[gtid] = 0u;
  if (threadIdx.x < PACK_ROWS_PER_BLOCK) {
    const int r = r0 + threadIdx.x;
    int seg = -1, t = 0;
    if (r < a.r_pad) {
      const int s = find_segment(a.seg.row_start, a.seg.n_seg, r);
      if (s >= 0) {
        t = r - __ldg(a.seg.row_start + s);
        if (t < __ldg(a.seg.len + s)) seg = s;
      }
      a.row_valid[r] = seg >= 0 ? 1 : 0;
    }
    s_seg[threadIdx.x] = seg;
    s_t[threadIdx.x] = t;
  }
  __syncthreads();
  const int pairs = a.k0_pad >> 1;
  const int half_ctx = (a.taps - 1) >> 1;
  const int k_real = a.taps * a.feat_dim;
  for (int idx = threadIdx.x; idx < PACK_ROWS_PER_BLOCK * pairs; idx += PACK_THREADS) {
    const int lr = idx / pairs, pr = idx - lr * pairs;
    const int r = r0 + lr;
    if (r >= a.r_pad) break;
    const int seg = s_seg[lr];
    float v[2] = {0.f, 0.f};
    if (seg >= 0) {
      const int t = s_t[lr];
      const int len = __ldg(a.seg.len + seg);
      const int64_t fs = __ldg(a.seg.feat_start + seg);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ch = 2 * pr + e;
        if (ch < k_real) {
          const int j = ch / a.feat_dim, c = ch - j * a.feat_dim;
          const int tt = t + (j - half_ctx) * a.dilation;     // SAME padding: outside the segment -> 0
          if (tt >= 0 && tt < len) v[e] = __ldg(a.feats + (fs + tt) * a.feat_dim + c);
        }
      }
    }
    reinterpret_cast<__half2*>(a.x0 + int64_t(r) * a.k0_pad)[pr] = __floats2half2_rn(v[0], v[1]);
  }
}

// ------------------------------------------------------------------------------------------
constexpr int POOL_THREADS = 256;
constexpr int POOL_SLAB = 128;          // channels per CTA
constexpr int POOL_MAX_G = 8;           // segments per CTA (amortises the W0 slab read)
constexpr int POOL_MAX_EPT = 4;         // embedding outputs per thread (emb_dim <= 1024)

struct PoolArgs {
  const __half* h;            // [r_pad, C] last frame layer output
  SegMeta seg;
  int32_t channels;           // C (multiple of 128)
  int32_t emb_dim;            // E (multiple of 256, <= 1024)
  int32_t group;              // segments per CTA, 1..POOL_MAX_G
  const float* w0;            // [2C, E]  embed_layer-0/w
  const float* b0;            // [E]
  float* partial;             // [n_slabs, n_seg, E]
  uint32_t* counters;         // [n_groups], zero on entry
  float* emb;                 // [n_seg, E]
  float* stats_out;           // optional [n_seg, 2C]
  float var_eps;
};

__global__ void __launch_bounds__(POOL_THREADS) pool_embed_kernel(const PoolArgs a) {
  __shared__ float s_red[2][POOL_THREADS / 32][POOL_SLAB];      // 8 KB: per-warp S1 / S2
  __shared__ float s_stats[POOL_MAX_G][2 * POOL_SLAB];          // 8 KB: mean | std per segment
  __shared__ int s_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g0 = blockIdx.x * a.group;
  const int n_in_group = min(a.group, a.seg.n_seg - g0);
  const int slab = blockIdx.y, n_slabs = gridDim.y;
  const int c0 = slab * POOL_SLAB;
  const int C = a.channels, E = a.emb_dim;

  // ---- statistics pooling: each lane owns 4 channels, each warp strides over rows -----------
  for (int gi = 0; gi < n_in_group; ++gi) {
    const int seg = g0 + gi;
    const int len = __ldg(a.seg.len + seg);
    const __half* base = a.h + int64_t(__ldg(a.seg.row_start + seg)) * C + c0 + lane * 4;
    // shifted sums: k = first frame (a sample of the data) keeps S2 - S1^2/n well conditioned
    const uint2 kraw = __ldg(reinterpret_cast<const uint2*>(base));
    const float2 k01 = __half22float2(*reinterpret_cast<const __half2*>(&kraw.x));
    const float2 k23 = __half22float2(*reinterpret_cast<const __half2*>(&kraw.y));
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    int t = warp;
    for (; t + 24 < len; t += 32) {                            // 4 independent 8-byte loads in flight
      uint2 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = __ldg(reinterpret_cast<const uint2*>(base + int64_t(t + 8 * u) * C));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 x01 = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].x));
        const float2 x23 = __half22float2(*reinterpret_cast<const __half2*>(&raw[u].y));
        const float d0 = x01.x - k01.x, d1 = x01.y - k01.y, d2 = x23.x - k23.x, d3 = x23.y - k23.y;
        s1[0] += d0; s1[1] += d1; s1[2] += d2; s1[3] += d3;
        s2[0] = fmaf(d0, d0, s2[0]); s2[1] = fmaf(d1, d1, s2[1]);
        s2[2] = fmaf(d2, d2, s2[2]); s2[3] = fmaf(d3, d3, s2[3]);
      }
    }
    for (; t < len; t += 8) {
      const uint2 raw = __ldg(reinterpret_cast<const uint2*>(base + int64_t(t) * C));
      const float2 x01 = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
      const float2 x23 = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
      const float d0 = x01.x - k01.x, d1 = x01.y - k01.y, d2 = x23.x - k23.x, d3 = x23.y - k23.y;
      s1[0] += d0; s1[1] += d1; s1[2] += d2; s1[3] += d3;
      s2[0] = fmaf(d0, d0, s2[0]); s2[1] = fmaf(d1, d1, s2[1]);
      s2[2] = fmaf(d2, d2, s2[2]); s2[3] = fmaf(d3, d3, s2[3]);
    }
    __syncthreads();                                            // s_red free (previous segment consumed)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s_red[0][warp][lane * 4 + e] = s1[e];
      s_red[1][warp][lane * 4 + e] = s2[e];
    }
    __syncthreads();
    if (tid < POOL_SLAB) {
      float S1 = 0.f, S2 = 0.f;
#pragma unroll
      for (int w = 0; w < POOL_THREADS / 32; ++w) { S1 += s_red[0][w][tid]; S2 += s_red[1][w][tid]; }
      const float k = __half2float(__ldg(a.h + int64_t(__ldg(a.seg.row_start + seg)) * C + c0 + tid));
      const float inv_n = 1.f / float(len);
      const float dm = S1 * inv_n;
      const float mean = k + dm;
      const float var = fmaxf(S2 * inv_n - dm * dm, 0.f);       // population variance (tf.nn.moments)
      const float sd = sqrtf(var + a.var_eps);                  // models.py:486
      s_stats[gi][tid] = mean;
      s_stats[gi][POOL_SLAB + tid] = sd;
      if (a.stats_out != nullptr) {
        a.stats_out[int64_t(seg) * 2 * C + c0 + tid] = mean;
        a.stats_out[int64_t(seg) * 2 * C + C + c0 + tid] = sd;
      }
    }
  }
  __syncthreads();

  // ---- embed_layer-0 partial product for this channel slab -------------------------------
  const int ept = E / POOL_THREADS;                             // outputs per thread
  float acc[POOL_MAX_G][POOL_MAX_EPT];
#pragma unroll
  for (int gi = 0; gi < POOL_MAX_G; ++gi)
#pragma unroll
    for (int i = 0; i < POOL_MAX_EPT; ++i) acc[gi][i] = 0.f;
#pragma unroll 4
  for (int r = 0; r < 2 * POOL_SLAB; ++r) {
    const int wrow = (r < POOL_SLAB) ? (c0 + r) : (C + c0 + r - POOL_SLAB);     // mean rows, then std rows
    const float* wp = a.w0 + int64_t(wrow) * E + tid;
    float w[POOL_MAX_EPT];
#pragma unroll
    for (int i = 0; i < POOL_MAX_EPT; ++i) w[i] = (i < ept) ? __ldg(wp + i * POOL_THREADS) : 0.f;
#pragma unroll
    for (int gi = 0; gi < POOL_MAX_G; ++gi) {
      if (gi < n_in_group) {
        const float s = s_stats[gi][r];
#pragma unroll
        for (int i = 0; i < POOL_MAX_EPT; ++i) acc[gi][i] = fmaf(s, w[i], acc[gi][i]);
      }
    }
  }
#pragma unroll
  for (int gi = 0; gi < POOL_MAX_G; ++gi) {
    if (gi < n_in_group) {
#pragma unroll
      for (int i = 0; i < POOL_MAX_EPT; ++i)
        if (i < ept) a.partial[(int64_t(slab) * a.seg.n_seg + g0 + gi) * E + tid + i * POOL_THREADS] = acc[gi][i];
    }
  }

  // ---- the last CTA of the group sums the slabs in a fixed order (deterministic) ------------
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.counters + blockIdx.x, 1u) == uint32_t(n_slabs - 1));
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int gi = 0; gi < n_in_group; ++gi) {
      for (int o = tid; o < E; o += POOL_THREADS) {
        float sum = __ldg(a.b0 + o);
        for (int s = 0; s < n_slabs; ++s) sum += __ldcg(a.partial + (int64_t(s) * a.seg.n_seg + g0 + gi) * E + o);
        a.emb[int64_t(g0 + gi) * E + o] = sum;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void unpack_rows_kernel(const __half* __restrict__ h, SegMeta seg, int32_t channels,
                                   float* __restrict__ out) {
  const int s = blockIdx.x;
  const int len = seg.len[s];
  const int64_t src0 = int64_t(seg.row_start[s]) * channels, dst0 = int64_t(seg.feat_start[s]) * channels;
  const int64_t n = int64_t(len) * channels;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[dst0 + i] = __half2float(h[src0 + i]);
}

}  // namespace xvk
