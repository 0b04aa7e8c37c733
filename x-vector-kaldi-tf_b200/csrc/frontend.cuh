// frontend.cuh -- the feature front end of an extraction job on the device (sm_100a; HBM-bound byte/float work):
//
//   apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 | select-voiced-frames
//                                                                  (reference local/tf/extract_xvectors.sh:68)
//
//   vad_tile_count_kernel : voiced rows of every tile of every utterance (launched only when an utterance is
//                           longer than 4096 frames; shorter ones count the rows in front of a tile in place).
//   cmvn_select_kernel    : one CTA per tile = a whole utterance of up to 512 frames (longer ones are split evenly).  The
//                           rows every window of the tile touches (the utterance itself, or <= tile + cmn_window rows:
//                           ONE contiguous span of the caller's matrix) are staged in shared memory with coalesced
//                           16-byte loads and summed once per (cepstral bin, 16-row segment) in double; thread (bin,
//                           run of 16 frames) builds its first window from those partial sums and then slides it
//                           (subtract the row that left, add the row that entered), which is the recursion of Kaldi's
//                           SlidingWindowCmnInternal restarted every 16 frames; the normalised row goes straight to its
//                           compacted position  out_row0 + (voiced rows before it)  -- each row is 23 consecutive
//                           floats written by 23 consecutive lanes.
//
// Algorithmic traffic per raw frame: 4*D (read) + 4 (VAD) + 4*D*voiced_fraction (write) bytes; a tile that is a whole
// utterance reads nothing twice, split utterances re-read <= cmn_window rows per tile from L2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xvfe {

constexpr int THREADS = 512;
constexpr int TILE_MAX = THREADS; // most frames per CTA: utterances up to this long are ONE tile (one thread per frame for the
                                  // per-frame steps), longer ones are split evenly into tiles of a multiple of 32 frames
constexpr int RUN = 16;          // consecutive frames per thread (the kernel lengthens runs until one pass covers the tile)
constexpr int SEG = 16;          // fewest slab rows per partial sum (ditto)
constexpr int DIRECT_COUNT_MAX = 4096;      // utterances up to this long: voiced rows in front of a tile counted in place
constexpr int COUNT_THREADS = 128;
constexpr uint32_t ERR_VAD_MISMATCH = 2u;   // bit of the model's sticky device flag (bit 0 = fp16 overflow)

struct UttMeta {
  const int32_t* in_row0;   // [n_utt] first raw row
  const int32_t* len;       // [n_utt] raw rows
  const int32_t* out_row0;  // [n_utt] first output row
  const int32_t* keep;      // [n_utt] selected rows to write
  const int32_t* tile0;     // [n_utt + 1] first tile of the utterance; tile0[n_utt] = number of tiles
  const int32_t* tile_utt;  // [n_tiles] utterance of every tile
  const int32_t* tile_rows; // [n_utt] frames per tile of the utterance
  int32_t n_utt;
};

struct CmvnOpts {
  int32_t cmn_window, min_window, center, normalize_variance;
};

// [ws, we) of frame t: Kaldi's window placement (centred windows are shifted, not shrunk, at the edges).
__host__ __device__ __forceinline__ void window_bounds(int t, int T, const CmvnOpts& o, int& ws, int& we) {
  if (o.center) {
    ws = t - o.cmn_window / 2;
    we = ws + o.cmn_window;
  } else {
    ws = t - o.cmn_window;
    we = t + 1;
  }
  if (ws < 0) {
    we -= ws;
    ws = 0;
  }
  if (!o.center && we > t) we = max(t + 1, o.min_window);
  if (we > T) {
    ws -= we - T;
    we = T;
    if (ws < 0) ws = 0;
  }
}

// The same placement without branches (what the kernel's inner loops use):
//   centred : ws = clamp(t - W/2, 0, max(T - W, 0)),  we = min(T, ws + W)
//   trailing: we' = max(t + 1, min_window), ws' = max(t - W, 0); if we' > T the window is shifted left by we' - T
__device__ __forceinline__ void window_bounds_fast(int t, int T, const CmvnOpts& o, int& ws, int& we) {
  if (o.center) {
    ws = max(0, min(t - (o.cmn_window >> 1), T - o.cmn_window));
    we = min(T, ws + o.cmn_window);
  } else {
    const int e = max(t + 1, o.min_window);
    ws = max(0, max(t - o.cmn_window, 0) - max(e - T, 0));
    we = min(e, T);
  }
}

__global__ void __launch_bounds__(COUNT_THREADS)
vad_tile_count_kernel(UttMeta um, const float* __restrict__ vad, int32_t* __restrict__ tile_cnt) {
  cudaGridDependencySynchronize();
  const int tile = blockIdx.x;
  const int u = __ldg(um.tile_utt + tile);
  const int rp = __ldg(um.tile_rows + u);
  const int T = __ldg(um.len + u);
  const int t0 = (tile - __ldg(um.tile0 + u)) * rp;
  const int t1 = min(T, t0 + rp);
  const float* v = vad + size_t(__ldg(um.in_row0 + u));
  int n = 0;
  for (int t = t0 + threadIdx.x; t < t1; t += COUNT_THREADS) n += __ldg(v + t) != 0.f ? 1 : 0;
  __shared__ int s_n;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n != 0) atomicAdd(&s_n, n);
  __syncthreads();
  if (threadIdx.x == 0) tile_cnt[tile] = s_n;
}

// Dynamic shared memory of cmvn_select_kernel, carved in this order (sizes fixed per launch by the host from the longest
// tile / slab of the batch): slab floats | per-16-row partial sums (doubles; twice with variance) | -1/N per frame
// (doubles) | per-frame table (int4).
struct SmemPlan {
  int32_t slab_rows_cap;   // most slab rows of any tile of the launch
  int32_t rows_cap;        // most frames of any tile (multiple of 32)
};
__host__ __device__ inline size_t slab_floats(int slab_rows_cap, int D) { return (size_t(slab_rows_cap) * D + 4 + 3) & ~size_t(3); }
__host__ __device__ inline int seg_doubles(int slab_rows_cap, int D) { return ((slab_rows_cap / SEG + 2) * D + 1) & ~1; }   // even: keeps 16-byte alignment behind it
inline size_t cmvn_smem_bytes(const SmemPlan& sp, int D, bool norm_vars) {
  return slab_floats(sp.slab_rows_cap, D) * sizeof(float) + size_t(seg_doubles(sp.slab_rows_cap, D)) * sizeof(double) * (norm_vars ? 2 : 1) +
         size_t(sp.rows_cap) * (sizeof(double) + sizeof(int4));
}

// One CTA per tile; a tile is a whole utterance when it has at most TILE_MAX frames (then the slab IS the utterance and
// nothing is summed twice), else an even split of it.  tile_cnt == nullptr: the voiced rows in front of the tile are
// counted here from the VAD track (the host launches vad_tile_count_kernel only when some utterance is longer than
// DIRECT_COUNT_MAX frames).
template <bool NV>                                  // NV: also normalise the variance (sums of squares carried along)
__global__ void __launch_bounds__(THREADS)
cmvn_select_kernel(UttMeta um, CmvnOpts opts, SmemPlan sp, int32_t D, const float* __restrict__ feats,
                   const float* __restrict__ vad, const int32_t* __restrict__ tile_cnt, float* __restrict__ out,
                   uint32_t* __restrict__ err_flag) {
  extern __shared__ __align__(16) float slab[];    // [(we_last - ws_first) * D] (+ up to 3 floats of alignment slack)
  __shared__ int32_t s_warp_cnt[THREADS / 32];
  __shared__ int32_t s_before;
  double* seg_sum = reinterpret_cast<double*>(slab + slab_floats(sp.slab_rows_cap, D));   // [n_segs][D]
  double* seg_sq = seg_sum + (NV ? seg_doubles(sp.slab_rows_cap, D) : 0);                 // [n_segs][D] (variance only)
  double* s_scale = seg_sq + seg_doubles(sp.slab_rows_cap, D);                            // [rows_cap] -1 / window frames
  int4* s_row = reinterpret_cast<int4*>(s_scale + sp.rows_cap);   // [rows_cap] {window start, end (slab rows), output row or -1, -}

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  if (tid == 0) s_before = 0;
  cudaGridDependencySynchronize();
  const int u = __ldg(um.tile_utt + tile);
  const int tile_first = __ldg(um.tile0 + u);
  const int T = __ldg(um.len + u);
  const int keep = __ldg(um.keep + u);
  const int rp = __ldg(um.tile_rows + u);            // frames per tile of this utterance (multiple of 32, <= THREADS)
  const int64_t in_row0 = __ldg(um.in_row0 + u);
  const int64_t out_row0 = __ldg(um.out_row0 + u);
  const int t0 = (tile - tile_first) * rp;
  const int nr = min(rp, T - t0);
  int ws_first, we_first, ws_last, we_last;
  window_bounds_fast(t0, T, opts, ws_first, we_first);
  window_bounds_fast(t0 + nr - 1, T, opts, ws_last, we_last);

  // ---- stage rows [ws_first, we_last) of the utterance: one contiguous span of floats, 16-byte loads in the middle
  const int64_t e_begin = (in_row0 + ws_first) * D, e_end = (in_row0 + we_last) * D;
  const int shift = int(e_begin & 3);              // slab[shift + i] = feats[e_begin + i]: keeps 16-byte alignment
  {
    const int64_t v_begin = (e_begin + 3) & ~int64_t(3), v_end = e_end & ~int64_t(3);
    if (v_begin < v_end) {
      for (int64_t e = e_begin + tid; e < v_begin; e += THREADS) slab[shift + int(e - e_begin)] = __ldg(feats + e);
      const float4* src = reinterpret_cast<const float4*>(feats + v_begin);
      float4* dst = reinterpret_cast<float4*>(slab + shift + int(v_begin - e_begin));
      const int nv = int((v_end - v_begin) >> 2);
      for (int i = tid; i < nv; i += THREADS) dst[i] = __ldg(src + i);
      for (int64_t e = v_end + tid; e < e_end; e += THREADS) slab[shift + int(e - e_begin)] = __ldg(feats + e);
    } else {
      for (int64_t e = e_begin + tid; e < e_end; e += THREADS) slab[shift + int(e - e_begin)] = __ldg(feats + e);
    }
  }

  // ---- where every row of the tile goes: voiced rows before it in the utterance
  const bool voiced = tid < nr && (vad == nullptr || __ldg(vad + in_row0 + t0 + tid) != 0.f);
  int before_part = 0;
  if (vad != nullptr) {
    if (tile_cnt != nullptr) {
      for (int j = tile_first + tid; j < tile; j += THREADS) before_part += __ldg(tile_cnt + j);
    } else {
      const float* v = vad + in_row0;
      for (int j = tid; j < t0; j += THREADS) before_part += __ldg(v + j) != 0.f ? 1 : 0;
    }
  }
  __syncthreads();                                 // s_before is initialised and the slab is complete
  const unsigned ballot = __ballot_sync(0xffffffffu, voiced);
  if ((tid & 31) == 0) s_warp_cnt[tid >> 5] = __popc(ballot);
  const int rank_in_warp = voiced ? __popc(ballot & ((1u << (tid & 31)) - 1u)) : -1;
  if (vad != nullptr && t0 > 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before_part += __shfl_xor_sync(0xffffffffu, before_part, o);
    if ((tid & 31) == 0 && before_part != 0) atomicAdd(&s_before, before_part);
  }

  // ---- partial sums of every `seg` slab rows, per cepstral bin (thread = (bin, segment)).  `seg` (>= SEG) is the
  // shortest segment for which every (bin, segment) pair gets its own thread: one pass, no straggling second one.
  const float* x = slab + shift;                   // x[(t - ws_first) * D + d]
  const int slab_rows = we_last - ws_first;
  int seg = max(SEG, (slab_rows * D + THREADS - 1) / THREADS);
  while (((slab_rows + seg - 1) / seg) * D > THREADS && seg < slab_rows) ++seg;
  const int n_segs = (slab_rows + seg - 1) / seg;
  for (int item = tid; item < n_segs * D; item += THREADS) {
    const int sgm = item / D, d = item - sgm * D;
    const int cnt = min(slab_rows - sgm * seg, seg);
    const float* p = x + sgm * seg * D + d;
    double a0 = 0.0, a1 = 0.0, q0 = 0.0, q1 = 0.0;
    int i = 0;
#pragma unroll 4
    for (; i + 2 <= cnt; i += 2, p += 2 * D) {
      const double v0 = double(p[0]), v1 = double(p[D]);
      a0 += v0; a1 += v1;
      if (NV) { q0 += v0 * v0; q1 += v1 * v1; }
    }
    if (i < cnt) {
      const double v0 = double(p[0]);
      a0 += v0;
      if (NV) q0 += v0 * v0;
    }
    seg_sum[item] = a0 + a1;                        // seg_sum[sgm * D + d]
    if (NV) seg_sq[item] = q0 + q1;
  }
  __syncthreads();
  {
    int before = vad != nullptr ? s_before : t0;
    for (int w = 0; w < (tid >> 5); ++w) before += s_warp_cnt[w];
    const int pos = rank_in_warp < 0 ? -1 : before + rank_in_warp;
    if (tid < nr) {                      // the frame's window, once per tile instead of once per cepstral bin
      int ws, we;
      window_bounds_fast(t0 + tid, T, opts, ws, we);
      s_row[tid] = make_int4(ws - ws_first, we - ws_first, pos < keep ? pos : -1, 0);
      s_scale[tid] = -1.0 / double(we - ws);
    }
    if (tid == 0 && t0 + nr == T) {      // last tile of the utterance: did it hold as many voiced rows as the caller said?
      int total = before;
      for (int w = 0; w < THREADS / 32; ++w) total += s_warp_cnt[w];
      if (total < keep) atomicOr(err_flag, ERR_VAD_MISMATCH);
    }
  }
  __syncthreads();

  // ---- sliding sums: thread = (cepstral bin d, run of RUN frames).  The first window of a run is put together from
  // whole-segment partial sums plus the rows at its two ragged ends; after that the recursion of Kaldi's
  // SlidingWindowCmnInternal (subtract the row that left, add the row that entered).  Sums of floats are exact in
  // double (until their exponents are > 2^20 apart), so the association order does not show in the result.
  int run_len = max(RUN / 2, (nr * D + THREADS - 1) / THREADS);   // as above: one (bin, run) pair per thread, one pass
  while (((nr + run_len - 1) / run_len) * D > THREADS && run_len < nr) ++run_len;
  const int n_runs = (nr + run_len - 1) / run_len;
  for (int item = tid; item < n_runs * D; item += THREADS) {
    const int run = item / D, d = item - run * D;
    const int r0 = run * run_len;
    const int r1 = min(nr, r0 + run_len);
    const float* xd = x + d;                                    // column d of the slab
    int4 row = s_row[r0];                                       // slab rows [row.x, row.y) are frame r0's window
    double sum = 0.0, sumsq = 0.0;
    {
      const int a = row.x, b = row.y;
      const int sa = (a + seg - 1) / seg, sb = b / seg;         // whole segments [sa, sb)
      int head_end = b, tail_begin = b;
      if (sa < sb) {
        head_end = sa * seg;
        tail_begin = sb * seg;
        double c0 = 0.0, c1 = 0.0;
        const double* ps = seg_sum + sa * D + d;
        int k = sa;
        for (; k + 2 <= sb; k += 2, ps += 2 * D) { c0 += ps[0]; c1 += ps[D]; }
        if (k < sb) c0 += ps[0];
        sum = c0 + c1;
        if (NV) {
          const double* pq = seg_sq + sa * D + d;
          for (int k2 = sa; k2 < sb; ++k2, pq += D) sumsq += *pq;
        }
      }
      const float* ph = xd + a * D;
      for (int i = a; i < head_end; ++i, ph += D) {
        const double v = double(*ph);
        sum += v;
        if (NV) sumsq += v * v;
      }
      const float* pt = xd + tail_begin * D;
      for (int i = tail_begin; i < b; ++i, pt += D) {
        const double v = double(*pt);
        sum += v;
        if (NV) sumsq += v * v;
      }
    }
    const float* p_lo = xd + row.x * D;                         // row leaving next / row entering next / current row
    const float* p_hi = xd + row.y * D;
    const float* p_t = xd + (t0 + r0 - ws_first) * D;
    float* p_out = out + out_row0 * D + d;                      // + pos * D: 32-bit offsets inside one utterance
    const double* p_scale = s_scale + r0;
    // one frame: normalise and store.  Kaldi: output_frame.AddVec(-1.0 / window_frames, cur_sum), i.e. a separately
    // rounded product and sum.  No FMA here on purpose: x - mean lands EXACTLY on a float rounding midpoint surprisingly
    // often (the window sum of floats is exact in double), and which way such a tie falls is decided by the last bit
    // of the product.
    auto emit = [&](const int4& rw, const float* pt, double scale) {
      double y = __dadd_rn(double(*pt), __dmul_rn(scale, sum));
      if (NV) {
        const int n = rw.y - rw.x;
        if (n == 1) {
          y = 0.0;
        } else {
          double var = __dadd_rn(__dmul_rn(sumsq, 1.0 / double(n)),
                                 __dmul_rn(-1.0 / (double(n) * double(n)), __dmul_rn(sum, sum)));
          var = fmax(var, 1.0e-10);
          y = __dmul_rn(y, 1.0 / sqrt(var));
        }
      }
      p_out[rw.z * D] = float(y);
    };
    if (row.z >= 0) emit(row, p_t, *p_scale);
    for (int r = r0 + 1; r < r1; ++r) {
      const int4 next = s_row[r];
      p_t += D;
      ++p_scale;
      if (next.x > row.x) {
        const double v = double(*p_lo);
        p_lo += D;
        sum -= v;
        if (NV) sumsq -= v * v;
      }
      if (next.y > row.y) {
        const double v = double(*p_hi);
        p_hi += D;
        sum += v;
        if (NV) sumsq += v * v;
      }
      row = next;
      if (row.z >= 0) emit(row, p_t, *p_scale);
    }
  }
}

}  // namespace xvfe
