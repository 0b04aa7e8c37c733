/* xvec_frontend.h -- C ABI of the on-device feature front end of libxvec_b200.so (sm_100a).
 *
 * Replaces the two Kaldi processes the reference pipes in front of every extract_embedding.py job
 *
 *     apply-cmvn-sliding --norm-vars=false --center=true --cmn-window=300 scp:feats.scp ark:- |
 *     select-voiced-frames ark:- scp,s,cs:vad.scp ark:- |         -- reference local/tf/extract_xvectors.sh:68
 *
 * (SURVEY.md section 8 row f3): raw MFCC rows and the VAD track of many utterances go to the GPU once; sliding-window
 * mean (optionally variance) normalisation and the voiced-frame compaction run there and leave the rows exactly where
 * xv_forward (include/xvec.h) expects its `feats_dev`.  The arithmetic follows Kaldi's SlidingWindowCmnInternal
 * (src/feat/feature-functions.cc): window sums and `x + (-1/N) * sum` in double, result narrowed to float.
 *
 * Same conventions as xvec.h: plain C types, XV_OK / negative XV_E* return codes, xv_last_error(), enqueue-only on the
 * caller's stream unless stated.
 */
#ifndef XVEC_B200_FRONTEND_H_
#define XVEC_B200_FRONTEND_H_

#include "xvec.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Options of apply-cmvn-sliding (Kaldi SlidingWindowCmnOptions; the reference passes cmn_window = 300, center = 1,
 * normalize_variance = 0 and leaves min_window at its default of 100, which only matters when center = 0). */
typedef struct xv_cmvn_opts {
  int32_t cmn_window;          /* frames, >= 1                                                     */
  int32_t min_window;          /* used when center = 0                                             */
  int32_t center;              /* 1: window centred on the frame, shifted at the utterance edges   */
  int32_t normalize_variance;  /* 1: also scale to unit variance over the window                   */
} xv_cmvn_opts;

/* Bytes of device scratch xv_frontend needs for `n_utt` utterances with `total_rows` raw rows. */
size_t xv_frontend_workspace_bytes(const xv_model* m, int64_t total_rows, int32_t n_utt);

/* apply-cmvn-sliding | select-voiced-frames for a batch of utterances.
 *   feats_dev      [sum(utt_len), feat_dim] fp32, device: raw rows, utterances concatenated in order
 *   vad_dev        [sum(utt_len)] fp32, device: VAD decision per raw row (non-zero = voiced, Kaldi's vad.scp vectors);
 *                  NULL = keep every row (apply-cmvn-sliding alone)
 *   utt_len_host   [n_utt] int32, HOST: raw rows per utterance (>= 0)
 *   out_keep_host  [n_utt] int32, HOST: how many of the utterance's selected rows to write (the caller knows the voiced
 *                  counts -- it read the VAD vectors -- and may drop a tail that make_embedding's chunking would drop,
 *                  models.py:388-396); NULL = every row when vad_dev is NULL.  Utterance u's rows land at output row
 *                  sum(out_keep_host[0..u)), so the output is the `feats_dev` of an xv_forward call.
 *   out_dev        [sum(out_keep), feat_dim] fp32, device
 * If an utterance has fewer voiced rows than out_keep says, a sticky error bit is set on the device and the next
 * xv_check_overflow / xv_collect returns XV_EINVAL. */
int xv_frontend(xv_model* m, const float* feats_dev, const float* vad_dev, const int32_t* utt_len_host,
                const int32_t* out_keep_host, int32_t n_utt, const xv_cmvn_opts* opts, float* out_dev,
                void* workspace_dev, size_t workspace_bytes, void* stream);

/* Host-buffer, pipelined form: xv_submit_host (xvec.h) with the front end in front of the network.  Copies the raw
 * rows and the VAD track host->device, runs xv_frontend and then the forward over `n_seg` segments that tile the
 * selected rows (sum(seg_len_host) == sum(out_keep_host)), copies the embeddings back; collect with xv_collect.
 * This is "feats.scp + vad.scp in, x-vectors out": the shape of one extract_xvectors.sh job (:63-88) without the pipe. */
int xv_submit_host_raw(xv_model* m, const float* feats_host, const float* vad_host, const int32_t* utt_len_host,
                       const int32_t* out_keep_host, int32_t n_utt, const xv_cmvn_opts* opts,
                       const int32_t* seg_len_host, int32_t n_seg, float* emb_host, int32_t* ticket);

#ifdef __cplusplus
}
#endif
#endif /* XVEC_B200_FRONTEND_H_ */
