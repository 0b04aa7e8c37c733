// synth.cuh -- counter-based synthetic MFCC rows generated ON the device (include/xvec_job.h: xv_synth_mfcc).
// SURVEY 7 (hard part 7) / 8d: BASELINE configs[3] is 1 M utterances x ~600 frames = 55 GB of features -- more than the
// host can hold or feed; the job is measured with every rank generating its utterances' rows from (seed, utterance id,
// frame, coefficient), and the same integer recipe in numpy (synthetic.counter_mfcc) reproduces any utterance bit for bit
// for the parity spot checks.
//   h1 = splitmix64(seed ^ (utt << 20 | frame) * 0x9E3779B97F4A7C15 + coefficient), h2 = splitmix64(h1)
//   x  = float(lo32(h1) + hi32(h1) + lo32(h2) + hi32(h2) - 2^33) * k[coefficient]      (Irwin-Hall(4): ~normal, exact integers)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace synth {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

constexpr int MAX_DIM = 64;
struct Args {
  float* out;                 // [total_rows, feat_dim]
  const int64_t* utt_id;      // [n_utt]
  const int32_t* row_start;   // [n_utt + 1] prefix sums of the utterance lengths
  int32_t n_utt, feat_dim;
  int64_t total_rows;
  uint64_t seed;
  float k[MAX_DIM];           // per-coefficient scale
};

__global__ void __launch_bounds__(256) synth_mfcc_kernel(const Args a) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;      // one value per thread: coalesced stores
  if (e >= a.total_rows * a.feat_dim) return;
  const int64_t row = e / a.feat_dim;
  const int d = int(e - row * a.feat_dim);
  int lo = 0, hi = a.n_utt;                    // the utterance holding this row: last u with row_start[u] <= row
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a.row_start + mid) <= row) lo = mid; else hi = mid;
  }
  const uint64_t frame = uint64_t(row - __ldg(a.row_start + lo));
  const uint64_t base = a.seed ^ (((uint64_t(__ldg(a.utt_id + lo)) << 20) | frame) * 0x9E3779B97F4A7C15ull);
  const uint64_t h1 = splitmix64(base + uint64_t(d)), h2 = splitmix64(h1);
  const int64_t s = int64_t((h1 & 0xffffffffull) + (h1 >> 32) + (h2 & 0xffffffffull) + (h2 >> 32)) - (int64_t(1) << 33);
  a.out[e] = __fmul_rn(__ll2float_rn(s), a.k[d]);
}

}  // namespace synth
