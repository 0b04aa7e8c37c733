// pack_pool.cuh -- the two HBM-bound kernels either side of the TDNN stack (sm_100a).
//
//   pack_im2col_kernel : fp32 MFCC rows [total_frames, D]  ->  fp16 "packed rows" matrix
//                        [R_pad, K0_pad] holding, per row, the taps0*D spliced input of the first
//                        layer (so layer 0 is a plain K = K0_pad GEMM), zero rows in the gaps,
//                        plus the row_valid map every later epilogue uses.
//   pool_stats_kernel  : combines the per-32-row-block partial sums written by
//                        the last frame layer's epilogue (the [frames,1536] activation is never
//                        materialised) into [mean | std] per segment.
//   embed_reduce_kernel: adds the K-splits of the tensor-core embedding GEMM (tdnn_pair_kernel<2>) + bias.
//   embed_fc_kernel    : embed_layer-0 as an fp32 SIMT split-K GEMM (option "fc" = 0; the cross-check of the above).
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

constexpr int PACK_ROWS_PER_BLOCK = 32;      // = one aligned pooling block: rows of ONE segment (or gap)
constexpr int PACK_THREADS = 128;
constexpr int PACK_MAX_STAGE_FLOATS = 2048; // (32 + 2*halo0) * feat_dim floats staged per CTA (8 KB: 16 CTAs per SM)
constexpr int PACK_MAX_K0 = 512;             // k0_pad / 8 = 16-byte pieces per row, at most 64 <= PACK_THREADS

struct PackArgs {
  const float* feats;        // [total_frames, D]
  int32_t r_pad;
  int32_t feat_dim, taps, dilation;
  int32_t k0_pad;            // multiple of 64
  __half* x0;                // [r_pad, k0_pad]
  uint8_t* row_valid;        // [r_pad]
  uint8_t* blk_valid;        // [r_pad / 32] valid rows of each aligned 32-row block
  const int32_t* lut;        // [k0_pad] staged-float offset of spliced column ch: tap * dilation * D + ceps, or -1 (padding)
  const int4* blk_info;      // [r_pad / 32] host-built: {first feature row, first frame in segment, segment length, valid rows}
  uint32_t* counters;        // [n_counters] zeroed here for embed_fc_kernel
  int32_t n_counters;
  int32_t split;             // 1: rows are [hi (k0_pad) | lo (k0_pad)], x = hi + lo (split-precision model)
  int32_t feats_f16;         // 1: `feats` points at float16 values (rounded on the host, xv_submit_host_utts_f16): widening them
                             //    and rounding again below gives back the same bits
};

// One CTA per aligned 32-row block.  The block's feature rows (plus the first layer's context) are
// one contiguous span of the caller's matrix: they are staged in shared memory with coalesced
// loads, out-of-segment rows as zeros, which is TF's SAME padding (models.py:476).  Thread
// (row, piece) then emits 16-byte pieces of the spliced fp16 rows; the (tap, ceps) -> staged
// offset table of its piece lives in registers.
__global__ void __launch_bounds__(PACK_THREADS) pack_im2col_kernel(const PackArgs a) {
  __shared__ float s_feat[PACK_MAX_STAGE_FLOATS];
  cudaTriggerProgrammaticLaunchCompletion();         // programmatic dependent launch: the next kernel may get scheduled
  cudaGridDependencySynchronize();                   // ... and this one touches global memory only after its predecessor
  const int r0 = blockIdx.x * PACK_ROWS_PER_BLOCK;
  const int gtid = blockIdx.x * PACK_THREADS + threadIdx.x;
  if (gtid < a.n_counters) a.counters[gtid] = 0u;
  if (r0 >= a.r_pad) return;
  const int D = a.feat_dim;
  const int halo = ((a.taps - 1) >> 1) * a.dilation;
  const int4 bi = __ldg(a.blk_info + blockIdx.x);    // block-uniform; one hop to everything the block needs
  const int t0 = bi.y, len = bi.z, nv = bi.w;
  if (threadIdx.x == 0) a.blk_valid[blockIdx.x] = uint8_t(nv);
  if (threadIdx.x < PACK_ROWS_PER_BLOCK) a.row_valid[r0 + threadIdx.x] = threadIdx.x < nv ? 1 : 0;
  const int pieces = a.k0_pad >> 3;                                   // 16-byte pieces per row (of one term)
  const int row_w = a.split ? 2 * a.k0_pad : a.k0_pad;                // halfs per stored row
  if (nv == 0) {                                                      // gap / tail block: zero rows
    for (int idx = threadIdx.x; idx < PACK_ROWS_PER_BLOCK * (row_w >> 3); idx += PACK_THREADS)
      reinterpret_cast<uint4*>(a.x0 + int64_t(r0) * row_w)[idx] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  // stage frames t0-halo .. t0+31+halo: staged float i is feats[(fs + t0 - halo) * D + i] when its frame exists
  const int n_stage = (PACK_ROWS_PER_BLOCK + 2 * halo) * D;
  const int lo = max(0, halo - t0) * D, hi = min(PACK_ROWS_PER_BLOCK + 2 * halo, len - t0 + halo) * D;
  if (a.feats_f16) {
    const __half* src0 = reinterpret_cast<const __half*>(a.feats) + (int64_t(bi.x) - halo) * D;
    for (int i = threadIdx.x; i < n_stage; i += PACK_THREADS) s_feat[i] = (i >= lo && i < hi) ? __half2float(__ldg(src0 + i)) : 0.f;
  } else {
    const float* src0 = a.feats + (int64_t(bi.x) - halo) * D;
    for (int i = threadIdx.x; i < n_stage; i += PACK_THREADS) s_feat[i] = (i >= lo && i < hi) ? __ldg(src0 + i) : 0.f;
  }
  const int rows_per_pass = PACK_THREADS / pieces;                    // k0_pad <= 512 -> pieces <= 64
  const int pc = threadIdx.x % pieces, lr0 = threadIdx.x / pieces;
  int lut[8];
  {
    const int4 l0 = __ldg(reinterpret_cast<const int4*>(a.lut) + pc * 2), l1 = __ldg(reinterpret_cast<const int4*>(a.lut) + pc * 2 + 1);
    lut[0] = l0.x; lut[1] = l0.y; lut[2] = l0.z; lut[3] = l0.w; lut[4] = l1.x; lut[5] = l1.y; lut[6] = l1.z; lut[7] = l1.w;
  }
  __syncthreads();
  if (lr0 < rows_per_pass) {
    for (int lr = lr0; lr < PACK_ROWS_PER_BLOCK; lr += rows_per_pass) {
      uint32_t w[4] = {0u, 0u, 0u, 0u}, wl[4] = {0u, 0u, 0u, 0u};
      if (lr < nv) {
        const float* src = s_feat + lr * D;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = lut[2 * e] >= 0 ? src[lut[2 * e]] : 0.f, x1 = lut[2 * e + 1] >= 0 ? src[lut[2 * e + 1]] : 0.f;
          const __half2 h = __floats2half2_rn(x0, x1);
          w[e] = *reinterpret_cast<const uint32_t*>(&h);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          wl[e] = *reinterpret_cast<const uint32_t*>(&l);
        }
      }
      reinterpret_cast<uint4*>(a.x0 + int64_t(r0 + lr) * row_w)[pc] = make_uint4(w[0], w[1], w[2], w[3]);
      if (a.split) reinterpret_cast<uint4*>(a.x0 + int64_t(r0 + lr) * row_w + a.k0_pad)[pc] = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Second half of statistics pooling, fed by the per-block partial sums that the last frame
// layer's epilogue (tdnn_pair_kernel mode 1) wrote:
//   partial[block][0][c] = sum of y over the block's valid rows, partial[block][1][c] = sum of y*y.
// The blocks of a segment are combined in a fixed order in fp64: mean = S1/n, var = S2/n - mean^2
// (population variance, tf.nn.moments), std = sqrt(var + 1e-5) (models.py:485-486); the result is
// stats[seg] = [mean(C) | std(C)] (the tf.concat of models.py:486).
constexpr int STATS_THREADS = 256;

struct StatsArgs {
  const float* partial;       // [r_pad/32][2][C]
  SegMeta seg;
  int32_t channels;           // C
  float* stats;               // [n_seg, 2C]  (may be null)
  __half* split;              // [n_seg, 3 * 2C] fp16 [hi | hi | lo] of the same statistics (may be null): the A operand
                              // of the split-precision embedding GEMM, x = hi + lo to ~2^-22
  float var_eps;
  int32_t weighted;           // 1: the partial sums are already weighted by attention weights that sum to 1 (n := 1)
  float sum_scale;            // the partial sums are of y / sum_scale (the last layer's rescue exponent; 0 = 1)
  float split_scale;          // the split copy holds statistic * split_scale (2^-e: the fp16 range rescue; 0 = 1)
  uint32_t* overflow_flag;    // bit 16 is set when a scaled statistic still leaves the fp16 range (may be null)
};

__device__ __forceinline__ void store_split(__half* row, int K, int k, float x) {
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  row[k] = hi; row[K + k] = hi; row[2 * K + k] = lo;
}

__global__ void __launch_bounds__(STATS_THREADS) pool_stats_kernel(const StatsArgs a) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int seg = blockIdx.y;
  const int c = blockIdx.x * STATS_THREADS + threadIdx.x;
  const int C = a.channels;
  if (c >= C) return;
  const int len = __ldg(a.seg.len + seg);
  const int blk0 = __ldg(a.seg.row_start + seg) >> 5;
  const int n_blk = (len + 31) >> 5;
  const float* p = a.partial + int64_t(blk0) * 2 * C + c;
  double s1 = 0.0, s2 = 0.0;
  int b = 0;
  for (; b + 8 <= n_blk; b += 8) {                              // 16 independent loads in flight, fixed order
    float x[8], y[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x[u] = __ldg(p + int64_t(b + u) * 2 * C);
      y[u] = __ldg(p + int64_t(b + u) * 2 * C + C);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { s1 += double(x[u]); s2 += double(y[u]); }
  }
  for (; b < n_blk; ++b) {
    s1 += double(__ldg(p + int64_t(b) * 2 * C));
    s2 += double(__ldg(p + int64_t(b) * 2 * C + C));
  }
  const double inv_n = a.weighted ? 1.0 : 1.0 / double(len);
  const double up = a.sum_scale != 0.f ? double(a.sum_scale) : 1.0;
  const double mean = s1 * inv_n * up;
  const double var = fmax(s2 * inv_n * up * up - mean * mean, 0.0);
  const float fm = float(mean), fs = float(sqrt(var + double(a.var_eps)));
  if (a.stats != nullptr) {
    a.stats[int64_t(seg) * 2 * C + c] = fm;
    a.stats[int64_t(seg) * 2 * C + C + c] = fs;
  }
  if (a.split != nullptr) {
    __half* row = a.split + int64_t(seg) * 6 * C;
    const float k = a.split_scale != 0.f ? a.split_scale : 1.f;
    const float sm = fm * k, ss = fs * k;
    if (!(fabsf(sm) <= 65504.f && ss <= 65504.f) && a.overflow_flag != nullptr) atomicOr(a.overflow_flag, 1u << 16);
    store_split(row, 2 * C, c, sm);
    store_split(row, 2 * C, C + c, ss);
  }
}

// embed_layer-0 epilogue of the tensor-core path: emb = b0 + sum over K-splits (fixed order) of the raw
// accumulators tdnn_pair_kernel<2> wrote.
struct FcReduceArgs {
  const float* partial;       // [splits][n_seg][E]
  const float* b0;            // [E]
  float* emb;                 // [n_seg, E]
  int32_t n_seg, E, splits;
  float out_scale;            // the K-split sum is multiplied by this before the bias (2^e of the statistics' rescue; 0 = 1)
};

__global__ void __launch_bounds__(64) embed_reduce_kernel(const FcReduceArgs a) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int64_t i4 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;          // one float4 of the output
  const int64_t n4 = int64_t(a.n_seg) * a.E / 4;
  if (i4 >= n4) return;
  const int o4 = int(i4 % (a.E / 4));
  const float4 bias = __ldg(reinterpret_cast<const float4*>(a.b0) + o4);
  const bool scaled = a.out_scale != 0.f && a.out_scale != 1.f;
  float4 sum = scaled ? make_float4(0.f, 0.f, 0.f, 0.f) : bias;       // unscaled: bias first, as before (same bits)
  const float4* p = reinterpret_cast<const float4*>(a.partial) + i4;
  int s = 0;
  for (; s + 12 <= a.splits; s += 12) {                                        // 12 loads in flight, fixed order
    float4 v[12];
#pragma unroll
    for (int u = 0; u < 12; ++u) v[u] = __ldcg(p + int64_t(s + u) * n4);
#pragma unroll
    for (int u = 0; u < 12; ++u) { sum.x += v[u].x; sum.y += v[u].y; sum.z += v[u].z; sum.w += v[u].w; }
  }
  for (; s < a.splits; ++s) {
    const float4 v = __ldcg(p + int64_t(s) * n4);
    sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
  }
  if (scaled) {
    sum.x = fmaf(sum.x, a.out_scale, bias.x); sum.y = fmaf(sum.y, a.out_scale, bias.y);
    sum.z = fmaf(sum.z, a.out_scale, bias.z); sum.w = fmaf(sum.w, a.out_scale, bias.w);
  }
  reinterpret_cast<float4*>(a.emb)[i4] = sum;
}

// ------------------------------------------------------------------------------------------
// Frame-weighted average of the chunk x-vectors of each utterance (make_embedding, models.py:398-421):
//   xvector_avg = 0; for each chunk: xvector_avg += offset * xvector; xvector_avg /= tot_weight
// in the reference's own float32 arithmetic and order (separately rounded multiply / add / divide, no FMA), so the
// result is bit-identical to that loop evaluated on the same chunk x-vectors.  A row goes to out[dst_row[u]] -- which may
// be PEER memory (rank 0's result table mapped over NVLink: the "gather" of a multi-GPU job is this store) -- and / or
// to the contiguous out_local[u] (the copy the host reads).
struct UttAvgArgs {
  const float* seg_emb;       // [n_seg, E] embed_layer-0 scores per segment (chunk)
  const int32_t* first_seg;   // [n_utt + 1] segments [first_seg[u], first_seg[u + 1]) are utterance u's chunks
  const int32_t* seg_len;     // [n_seg] rows of each segment = the reference's `offset`
  const int64_t* dst_row;     // [n_utt] destination row in `out` (null: u)
  float* out;                 // may be null
  float* out_local;           // [n_utt, E], may be null
  int32_t n_utt, E;
};

__global__ void __launch_bounds__(128) utt_average_kernel(const UttAvgArgs a) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int64_t i4 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;          // one float4 of one utterance
  const int e4 = a.E / 4;
  if (i4 >= int64_t(a.n_utt) * e4) return;
  const int u = int(i4 / e4), o4 = int(i4 % e4);
  const int s0 = __ldg(a.first_seg + u), s1 = __ldg(a.first_seg + u + 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  double tot = 0.0;
  for (int s = s0; s < s1; ++s) {
    const int len = __ldg(a.seg_len + s);
    const float w = float(len);
    const float4 x = __ldcg(reinterpret_cast<const float4*>(a.seg_emb + int64_t(s) * a.E) + o4);
    acc.x = __fadd_rn(acc.x, __fmul_rn(w, x.x)); acc.y = __fadd_rn(acc.y, __fmul_rn(w, x.y));
    acc.z = __fadd_rn(acc.z, __fmul_rn(w, x.z)); acc.w = __fadd_rn(acc.w, __fmul_rn(w, x.w));
    tot += double(len);
  }
  const float wt = float(tot);
  const float4 r = make_float4(__fdiv_rn(acc.x, wt), __fdiv_rn(acc.y, wt), __fdiv_rn(acc.z, wt), __fdiv_rn(acc.w, wt));
  if (a.out != nullptr) {
    const int64_t row = a.dst_row != nullptr ? a.dst_row[u] : int64_t(u);
    reinterpret_cast<float4*>(a.out + row * a.E)[o4] = r;
  }
  if (a.out_local != nullptr) reinterpret_cast<float4*>(a.out_local + int64_t(u) * a.E)[o4] = r;
}

// embed_reduce_kernel + utt_average_kernel in one launch (the usual case: all segments of the call in one GEMM group):
// thread = one float4 of one UTTERANCE; per chunk the K-split sum in embed_reduce_kernel's order, then the reference's
// float32 chunk average.  Bit-identical to the two kernels run one after the other.
struct FcReduceUttArgs {
  FcReduceArgs r;             // r.emb is not used
  UttAvgArgs u;               // u.seg_emb is not used
};

__global__ void __launch_bounds__(128) embed_reduce_utt_kernel(const FcReduceUttArgs a) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int64_t i4 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int e4 = a.r.E / 4;
  if (i4 >= int64_t(a.u.n_utt) * e4) return;
  const int u = int(i4 / e4), o4 = int(i4 % e4);
  const int s0 = __ldg(a.u.first_seg + u), s1 = __ldg(a.u.first_seg + u + 1);
  const int64_t n4 = int64_t(a.r.n_seg) * e4;
  const float4 bias = __ldg(reinterpret_cast<const float4*>(a.r.b0) + o4);
  const bool scaled = a.r.out_scale != 0.f && a.r.out_scale != 1.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  double tot = 0.0;
  for (int sg = s0; sg < s1; ++sg) {
    float4 sum = scaled ? make_float4(0.f, 0.f, 0.f, 0.f) : bias;
    const float4* p = reinterpret_cast<const float4*>(a.r.partial) + int64_t(sg) * e4 + o4;
    int s = 0;
    for (; s + 12 <= a.r.splits; s += 12) {
      float4 v[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) v[k] = __ldcg(p + int64_t(s + k) * n4);
#pragma unroll
      for (int k = 0; k < 12; ++k) { sum.x += v[k].x; sum.y += v[k].y; sum.z += v[k].z; sum.w += v[k].w; }
    }
    for (; s < a.r.splits; ++s) {
      const float4 v = __ldcg(p + int64_t(s) * n4);
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
    }
    if (scaled) {
      sum.x = fmaf(sum.x, a.r.out_scale, bias.x); sum.y = fmaf(sum.y, a.r.out_scale, bias.y);
      sum.z = fmaf(sum.z, a.r.out_scale, bias.z); sum.w = fmaf(sum.w, a.r.out_scale, bias.w);
    }
    const int len = __ldg(a.u.seg_len + sg);
    const float w = float(len);
    acc.x = __fadd_rn(acc.x, __fmul_rn(w, sum.x)); acc.y = __fadd_rn(acc.y, __fmul_rn(w, sum.y));
    acc.z = __fadd_rn(acc.z, __fmul_rn(w, sum.z)); acc.w = __fadd_rn(acc.w, __fmul_rn(w, sum.w));
    tot += double(len);
  }
  const float wt = float(tot);
  const float4 res = make_float4(__fdiv_rn(acc.x, wt), __fdiv_rn(acc.y, wt), __fdiv_rn(acc.z, wt), __fdiv_rn(acc.w, wt));
  if (a.u.out != nullptr) {
    const int64_t row = a.u.dst_row != nullptr ? a.u.dst_row[u] : int64_t(u);
    reinterpret_cast<float4*>(a.u.out + row * a.r.E)[o4] = res;
  }
  if (a.u.out_local != nullptr) reinterpret_cast<float4*>(a.u.out_local + int64_t(u) * a.r.E)[o4] = res;
}

// ------------------------------------------------------------------------------------------
// embed_layer-0 (tf.nn.xw_plus_b, models.py:495): emb[n_seg, E] = stats[n_seg, K] @ W0[K, E] + b0,
// the x-vector.  fp32 SIMT GEMM (the 1e-3 parity budget leaves no room for 16-bit statistics):
// split over K; the last CTA of a tile adds the K-splits in a fixed order, so every output is
// summed in one order whatever the batch.
constexpr int FC_BM = 64, FC_BN = 128, FC_BK = 32, FC_THREADS = 256;

struct FcArgs {
  const float* stats;         // [n_seg, K]
  const float* w0;            // [K, E]  embed_layer-0/w
  const float* b0;            // [E]
  float* fc_partial;          // [splits][n_seg][E]
  uint32_t* counters;         // [m_tiles * n_tiles], zero on entry
  float* emb;                 // [n_seg, E]
  int32_t n_seg, K, E, k_per_split;   // k_per_split % FC_BK == 0
};

// 64 x 128 output tile per CTA, 8 segments x 4 outputs per thread (per k: 3 LDS.128 feed 32 FFMA, so
// the FP32 pipe, not shared memory, is the limit).
__global__ void __launch_bounds__(FC_THREADS) embed_fc_kernel(const FcArgs a) {
  __shared__ __align__(16) float As[FC_BK][FC_BM + 4];          // k-major: a thread's 8 segments are two float4
  __shared__ __align__(16) float Bs[FC_BK][FC_BN];
  __shared__ int s_last;
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int seg0 = blockIdx.x * FC_BM, o0 = blockIdx.y * FC_BN;
  const int split = blockIdx.z, n_splits = gridDim.z;
  const int k_begin = split * a.k_per_split, k_end = min(a.K, k_begin + a.k_per_split);
  float4 a_reg[2], b_reg[4];
  auto load = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * FC_THREADS, sg = seg0 + (idx >> 3), ak = (idx & 7) * 4;
      a_reg[i] = (sg < a.n_seg && k0 + ak < k_end) ? __ldg(reinterpret_cast<const float4*>(a.stats + int64_t(sg) * a.K + k0 + ak))
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * FC_THREADS, kr = idx >> 5, c4 = (idx & 31) * 4;
      b_reg[i] = (k0 + kr < k_end) ? __ldg(reinterpret_cast<const float4*>(a.w0 + int64_t(k0 + kr) * a.E + o0 + c4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  load(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += FC_BK) {
    __syncthreads();                                            // previous chunk consumed
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + i * FC_THREADS, ar = idx >> 3, ak = (idx & 7) * 4;
      As[ak + 0][ar] = a_reg[i].x; As[ak + 1][ar] = a_reg[i].y; As[ak + 2][ar] = a_reg[i].z; As[ak + 3][ar] = a_reg[i].w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * FC_THREADS;
      *reinterpret_cast<float4*>(&Bs[idx >> 5][(idx & 31) * 4]) = b_reg[i];
    }
    __syncthreads();
    if (k0 + FC_BK < k_end) load(k0 + FC_BK);                   // prefetch the next chunk into registers
#pragma unroll
    for (int k = 0; k < FC_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float am[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bm[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(am[i], bm[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int sg = seg0 + ty * 8 + i;
    if (sg < a.n_seg)
      *reinterpret_cast<float4*>(a.fc_partial + (int64_t(split) * a.n_seg + sg) * a.E + o0 + tx * 4) =
          make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  // ---- the last CTA of this output tile adds the K-splits in a fixed order ---------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(a.counters + blockIdx.y * gridDim.x + blockIdx.x, 1u) == uint32_t(n_splits - 1));
  __syncthreads();
  if (s_last) {
    __threadfence();
    const float4 bias = __ldg(reinterpret_cast<const float4*>(a.b0 + o0 + tx * 4));
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      const int sg = seg0 + ty * 8 + i;
      if (sg >= a.n_seg) continue;
      float4 sum = bias;
      for (int s = 0; s < n_splits; ++s) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(a.fc_partial + (int64_t(s) * a.n_seg + sg) * a.E + o0 + tx * 4));
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
      *reinterpret_cast<float4*>(a.emb + int64_t(sg) * a.E + o0 + tx * 4) = sum;
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void unpack_rows_kernel(const __half* __restrict__ h, SegMeta seg, int32_t channels,
                                   float* __restrict__ out, float scale, int32_t split = 0) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int s = blockIdx.x;
  const int len = seg.len[s];
  const int64_t dst0 = int64_t(seg.feat_start[s]) * channels;
  const int64_t n = int64_t(len) * channels;
  if (split) {                                       // rows are [hi (channels) | lo (channels)]
    const __half* base = h + int64_t(seg.row_start[s]) * 2 * channels;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const int64_t r = i / channels, c = i % channels;
      out[dst0 + i] = (__half2float(base[r * 2 * channels + c]) + __half2float(base[r * 2 * channels + channels + c])) * scale;
    }
    return;
  }
  const int64_t src0 = int64_t(seg.row_start[s]) * channels;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[dst0 + i] = __half2float(h[src0 + i]) * scale;
}


// ------------------------------------------------------------------------------------------
// Self-attention pooling (ModelL2LossWithoutDropoutLReluAttention, local/tf/models.py:1037-1051):
//   attention = softmax over the frames of one segment of  sum_c v_c * tanh((h1 W)_c + b_c)
//   h_m = sum_t a_t h2[t],  h_s = sum_t a_t h2[t]^2 - h_m^2,  stats = [h_m | sqrt(h_s + 1e-5)]
// The score GEMM runs on the tensor cores (tdnn_pair_kernel mode 3) and leaves n_part partial sums per row.
// attn_softmax_kernel: one CTA per segment; fixed-order reductions.
constexpr int ATTN_THREADS = 256;
__global__ void __launch_bounds__(ATTN_THREADS)
attn_softmax_kernel(const float* __restrict__ score_partial, int32_t n_part, SegMeta seg, float* __restrict__ attn) {
  __shared__ float red[ATTN_THREADS];
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int s = blockIdx.x, t = threadIdx.x;
  const int row0 = __ldg(seg.row_start + s), len = __ldg(seg.len + s);
  float mx = -INFINITY;
  for (int i = t; i < len; i += ATTN_THREADS) {
    const float* p = score_partial + int64_t(row0 + i) * n_part;
    float sc = 0.f;
    for (int k = 0; k < n_part; ++k) sc += p[k];
    attn[row0 + i] = sc;
    mx = fmaxf(mx, sc);
  }
  red[t] = mx;
  __syncthreads();
  for (int w = ATTN_THREADS / 2; w > 0; w >>= 1) { if (t < w) red[t] = fmaxf(red[t], red[t + w]); __syncthreads(); }
  mx = red[0];
  __syncthreads();
  float z = 0.f;
  for (int i = t; i < len; i += ATTN_THREADS) {
    const float e = expf(attn[row0 + i] - mx);
    attn[row0 + i] = e;
    z += e;
  }
  red[t] = z;
  __syncthreads();
  for (int w = ATTN_THREADS / 2; w > 0; w >>= 1) { if (t < w) red[t] += red[t + w]; __syncthreads(); }
  const float inv_z = 1.f / red[0];
  for (int i = t; i < len; i += ATTN_THREADS) attn[row0 + i] *= inv_z;
}

// Per aligned 32-row block x 256 channels: partial[blk][0][c] = sum a_t h2[t,c], partial[blk][1][c] = sum a_t h2[t,c]^2
// (the layout pool_stats_kernel combines; h2 = columns [col0, col0 + C) of the last layer's stored activation).
__global__ void __launch_bounds__(256)
attn_pool_kernel(const __half* __restrict__ h, int32_t row_stride, int32_t col0, int32_t C, const float* __restrict__ attn,
                 const uint8_t* __restrict__ blk_valid, float* __restrict__ partial, float in_scale, int32_t lo_off) {
  __shared__ float red[8][2][256];
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int n_slab = C / 256;                                   // 1-D grid, channel slab fastest: the CTAs that share a
  const int blk = blockIdx.x / n_slab, c0 = (blockIdx.x % n_slab) * 256;   // row block run together (whole DRAM rows)
  const int nv = blk_valid[blk];
  if (nv == 0) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
  uint4 u[4], ul[4];
  float a[4];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {                    // all loads first (rows past nv are gap rows: zeros, weight 0)
    const int r = w * 4 + rr;
    const int64_t row = int64_t(blk) * 32 + r;
    u[rr] = __ldg(reinterpret_cast<const uint4*>(h + row * row_stride + col0 + c0 + lane * 8));
    ul[rr] = lo_off > 0 ? __ldg(reinterpret_cast<const uint4*>(h + row * row_stride + lo_off + col0 + c0 + lane * 8))
                        : make_uint4(0u, 0u, 0u, 0u);  // split-precision rows: the lo terms
    a[rr] = r < nv ? __ldg(attn + row) : 0.f;
  }
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const __half2* hp = reinterpret_cast<const __half2*>(&u[rr]);
    const __half2* lp = reinterpret_cast<const __half2*>(&ul[rr]);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(hp[i]);
      const float2 fl = __half22float2(lp[i]);
      f.x += fl.x; f.y += fl.y;
      f.x *= in_scale; f.y *= in_scale;                 // rows are stored / 2^e (fp16 range rescue); exact
      s1[2 * i] = fmaf(a[rr], f.x, s1[2 * i]);         s2[2 * i] = fmaf(a[rr] * f.x, f.x, s2[2 * i]);
      s1[2 * i + 1] = fmaf(a[rr], f.y, s1[2 * i + 1]); s2[2 * i + 1] = fmaf(a[rr] * f.y, f.y, s2[2 * i + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[w][0][lane * 8 + i] = s1[i]; red[w][1][lane * 8 + i] = s2[i]; }
  __syncthreads();
  const int t = threadIdx.x;
  float x = 0.f, y = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) { x += red[k][0][t]; y += red[k][1][t]; }
  partial[(int64_t(blk) * 2 + 0) * C + c0 + t] = x;
  partial[(int64_t(blk) * 2 + 1) * C + c0 + t] = y;
}

}  // namespace xvk
