// frontend.cuh -- the feature front end of an extraction job on the device (sm_100a; HBM-bound byte/float work):
//
//   apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 | select-voiced-frames
//                                                                  (reference local/tf/extract_xvectors.sh:68)
//
//   vad_tile_count_kernel : voiced rows of every 128-row tile of every utterance.
//   cmvn_select_kernel    : one CTA per tile.  The rows every window of the tile touches (<= 128 + cmn_window of them,
//                           ONE contiguous span of the caller's matrix) are staged in shared memory with coalesced
//                           16-byte loads; thread (cepstral bin, run of 16 frames) sums its first window in double and
//                           then slides it (subtract the row that left, add the row that entered), which is the
//                           recursion of Kaldi's SlidingWindowCmnInternal restarted every 16 frames; the normalised row
//                           goes straight to its compacted position  out_row0 + (voiced rows before it)  -- each row is
//                           23 consecutive floats written by 23 consecutive lanes.
//
// Algorithmic traffic per raw frame: 4*D (read) + 4 (VAD) + 4*D*voiced_fraction (write) bytes; the window re-reads
// (x(1 + W/128)) are served by L1/L2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace xvfe {

constexpr int TILE = 128;        // frames per CTA
constexpr int RUN = 16;          // consecutive frames per thread
constexpr int THREADS = 256;
constexpr int COUNT_THREADS = 128;
constexpr uint32_t ERR_VAD_MISMATCH = 2u;   // bit of the model's sticky device flag (bit 0 = fp16 overflow)

struct UttMeta {
  const int32_t* in_row0;   // [n_utt] first raw row
  const int32_t* len;       // [n_utt] raw rows
  const int32_t* out_row0;  // [n_utt] first output row
  const int32_t* keep;      // [n_utt] selected rows to write
  const int32_t* tile0;     // [n_utt + 1] first tile of the utterance; tile0[n_utt] = number of tiles
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

// utterance of a tile: last u with tile0[u] <= tile (utterances without rows own no tile and are skipped)
__device__ __forceinline__ int find_utt(const int32_t* __restrict__ tile0, int n_utt, int tile) {
  int lo = 0, hi = n_utt;          // invariant: tile0[lo] <= tile < tile0[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(tile0 + mid) <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(COUNT_THREADS)
vad_tile_count_kernel(UttMeta um, const float* __restrict__ vad, int32_t* __restrict__ tile_cnt) {
  cudaGridDependencySynchronize();
  const int tile = blockIdx.x;
  const int u = find_utt(um.tile0, um.n_utt, tile);
  const int t = (tile - __ldg(um.tile0 + u)) * TILE + threadIdx.x;
  const bool voiced = t < __ldg(um.len + u) && __ldg(vad + size_t(__ldg(um.in_row0 + u)) + t) != 0.f;
  const int n = __syncthreads_count(voiced);
  if (threadIdx.x == 0) tile_cnt[tile] = n;
}

__global__ void __launch_bounds__(THREADS)
cmvn_select_kernel(UttMeta um, CmvnOpts opts, int32_t D, const float* __restrict__ feats, const float* __restrict__ vad,
                   const int32_t* __restrict__ tile_cnt, float* __restrict__ out, uint32_t* __restrict__ err_flag) {
  extern __shared__ __align__(16) float slab[];    // [(we_last - ws_first) * D] (+ up to 3 floats of alignment slack)
  __shared__ int32_t s_pos[TILE];                   // output row inside the utterance, or -1 (unvoiced / beyond keep)
  __shared__ int32_t s_warp_cnt[TILE / 32];
  __shared__ int32_t s_before;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int u = find_utt(um.tile0, um.n_utt, tile);
  const int tile_first = __ldg(um.tile0 + u);
  const int T = __ldg(um.len + u);
  const int keep = __ldg(um.keep + u);
  const int64_t in_row0 = __ldg(um.in_row0 + u);
  const int64_t out_row0 = __ldg(um.out_row0 + u);
  const int t0 = (tile - tile_first) * TILE;
  const int nr = min(TILE, T - t0);
  int ws_first, we_first, ws_last, we_last;
  window_bounds(t0, T, opts, ws_first, we_first);
  window_bounds(t0 + nr - 1, T, opts, ws_last, we_last);
  if (tid == 0) s_before = 0;
  cudaGridDependencySynchronize();

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

  __syncthreads();                                 // s_before is initialised (and the slab is complete)

  // ---- where every row of the tile goes: voiced rows before it in the utterance
  if (tid < TILE) {
    const bool voiced = tid < nr && (vad == nullptr || __ldg(vad + in_row0 + t0 + tid) != 0.f);
    const unsigned ballot = __ballot_sync(0xffffffffu, voiced);
    if ((tid & 31) == 0) s_warp_cnt[tid >> 5] = __popc(ballot);
    s_pos[tid] = voiced ? __popc(ballot & ((1u << (tid & 31)) - 1u)) : -1;
  } else if (vad != nullptr) {
    int part = 0;
    for (int j = tile_first + (tid - TILE); j < tile; j += THREADS - TILE) part += __ldg(tile_cnt + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0 && part != 0) atomicAdd(&s_before, part);
  }
  __syncthreads();
  if (tid < TILE) {
    int before = vad != nullptr ? s_before : t0;
    for (int w = 0; w < (tid >> 5); ++w) before += s_warp_cnt[w];
    const int p = s_pos[tid];
    const int pos = p < 0 ? -1 : before + p;
    s_pos[tid] = pos < keep ? pos : -1;
    if (tid == 0 && t0 + nr == T) {      // last tile of the utterance: did it hold as many voiced rows as the caller said?
      int total = before;
      for (int w = 0; w < TILE / 32; ++w) total += s_warp_cnt[w];
      if (total < keep) atomicOr(err_flag, ERR_VAD_MISMATCH);
    }
  }
  __syncthreads();

  // ---- sliding sums: thread = (cepstral bin d, run of RUN frames)
  const float* x = slab + shift;                   // x[(t - ws_first) * D + d]
  const int n_runs = (nr + RUN - 1) / RUN;
  for (int item = tid; item < n_runs * D; item += THREADS) {
    const int d = item % D, r0 = (item / D) * RUN;
    const int r1 = min(nr, r0 + RUN);
    int ws, we;
    window_bounds(t0 + r0, T, opts, ws, we);
    double sum = 0.0, sumsq = 0.0;
    {
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      const float* p = x + (ws - ws_first) * D + d;
      int i = ws;
      for (; i + 4 <= we; i += 4, p += 4 * D) {
        a0 += double(p[0]); a1 += double(p[D]); a2 += double(p[2 * D]); a3 += double(p[3 * D]);
      }
      for (; i < we; ++i, p += D) a0 += double(*p);
      sum = (a0 + a1) + (a2 + a3);
      if (opts.normalize_variance) {
        const float* q = x + (ws - ws_first) * D + d;
        for (int j = ws; j < we; ++j, q += D) sumsq += double(*q) * double(*q);
      }
    }
    for (int r = r0; r < r1; ++r) {
      const int t = t0 + r;
      if (r > r0) {
        int nws, nwe;
        window_bounds(t, T, opts, nws, nwe);
        if (nws > ws) {
          const double v = double(x[(ws - ws_first) * D + d]);
          sum -= v;
          if (opts.normalize_variance) sumsq -= v * v;
        }
        if (nwe > we) {
          const double v = double(x[(we - ws_first) * D + d]);
          sum += v;
          if (opts.normalize_variance) sumsq += v * v;
        }
        ws = nws;
        we = nwe;
      }
      const int pos = s_pos[r];
      if (pos < 0) continue;
      const int n = we - ws;
      double y = double(x[(t - ws_first) * D + d]) + (-1.0 / double(n)) * sum;
      if (opts.normalize_variance) {
        if (n == 1) {
          y = 0.0;
        } else {
          double var = sumsq * (1.0 / double(n)) + (-1.0 / (double(n) * double(n))) * sum * sum;
          var = fmax(var, 1.0e-10);
          y *= 1.0 / sqrt(var);
        }
      }
      out[(out_row0 + pos) * D + d] = float(y);
    }
  }
}

}  // namespace xvfe
