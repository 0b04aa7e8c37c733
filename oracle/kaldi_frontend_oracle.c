/* kaldi_frontend_oracle.c -- plain-C restatement of the feature pipe of the reference's
 * local/tf/extract_xvectors.sh:68  (apply-cmvn-sliding | select-voiced-frames), third restatement next to the two numpy
 * ones in kaldi_frontend_oracle.py.
 *
 * TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg): the product path computes
 * this on the GPU (x-vector-kaldi-tf_b200/csrc/frontend.cuh) and never links or loads this file.
 *
 * Parity unpinned: Kaldi (third-party dependency of the reference, version unpinned) is neither under /root/reference nor
 * in the build image; what is restated is Kaldi's published algorithm -- SlidingWindowCmnInternal
 * (src/feat/feature-functions.cc) and select-voiced-frames (src/ivectorbin/select-voiced-frames.cc) -- with the options
 * the reference passes.  Build: gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction: Kaldi's AddVec is a
 * separately rounded product and sum).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static void window_bounds(int t, int num_frames, int cmn_window, int center, int min_window, int* ws, int* we) {
  int window_start, window_end;
  if (center) {
    window_start = t - cmn_window / 2;
    window_end = window_start + cmn_window;
  } else {
    window_start = t - cmn_window;
    window_end = t + 1;
  }
  if (window_start < 0) {            /* shift the window right */
    window_end -= window_start;
    window_start = 0;
  }
  if (!center) {
    if (window_end > t) window_end = (t + 1 > min_window) ? t + 1 : min_window;
  }
  if (window_end > num_frames) {     /* shift it left */
    window_start -= window_end - num_frames;
    window_end = num_frames;
    if (window_start < 0) window_start = 0;
  }
  *ws = window_start;
  *we = window_end;
}

/* in, out: [num_frames, dim] float32 row-major.  Returns 0, or -1 when scratch memory cannot be had. */
int kfo_sliding_window_cmn(const float* in, int32_t num_frames, int32_t dim, int32_t cmn_window, int32_t min_window,
                           int32_t center, int32_t normalize_variance, float* out) {
  double* cur_sum = (double*)calloc((size_t)(2 * dim > 0 ? 2 * dim : 1), sizeof(double));
  if (!cur_sum) return -1;
  double* cur_sumsq = cur_sum + dim;
  int last_start = -1, last_end = -1;
  for (int t = 0; t < num_frames; ++t) {
    int ws, we;
    window_bounds(t, num_frames, cmn_window, center, min_window, &ws, &we);
    if (last_start == -1) {
      for (int i = ws; i < we; ++i)
        for (int d = 0; d < dim; ++d) {
          const double v = (double)in[(size_t)i * dim + d];
          cur_sum[d] += v;
          cur_sumsq[d] += v * v;
        }
    } else {
      if (ws > last_start)
        for (int d = 0; d < dim; ++d) {
          const double v = (double)in[(size_t)last_start * dim + d];
          cur_sum[d] -= v;
          cur_sumsq[d] -= v * v;
        }
      if (we > last_end)
        for (int d = 0; d < dim; ++d) {
          const double v = (double)in[(size_t)last_end * dim + d];
          cur_sum[d] += v;
          cur_sumsq[d] += v * v;
        }
    }
    const int n = we - ws;
    last_start = ws;
    last_end = we;
    const double alpha = -1.0 / n;
    for (int d = 0; d < dim; ++d) {
      double y = (double)in[(size_t)t * dim + d] + alpha * cur_sum[d];
      if (normalize_variance) {
        if (n == 1) {
          y = 0.0;
        } else {
          double variance = cur_sumsq[d] * (1.0 / n) + (-1.0 / ((double)n * n)) * (cur_sum[d] * cur_sum[d]);
          if (variance < 1.0e-10) variance = 1.0e-10;
          y *= pow(variance, -0.5);
        }
      }
      out[(size_t)t * dim + d] = (float)y;
    }
  }
  free(cur_sum);
  return 0;
}

/* Rows of `in` whose VAD decision is non-zero, in order; returns how many were written to `out`. */
int32_t kfo_select_voiced_frames(const float* in, const float* vad, int32_t num_frames, int32_t dim, float* out) {
  int32_t kept = 0;
  for (int t = 0; t < num_frames; ++t) {
    if (vad[t] == 0.0f) continue;
    for (int d = 0; d < dim; ++d) out[(size_t)kept * dim + d] = in[(size_t)t * dim + d];
    ++kept;
  }
  return kept;
}
