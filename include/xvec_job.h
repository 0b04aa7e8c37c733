/* xvec_job.h -- C ABI of the host side of an ark -> x-vector-ark extraction job (libxvec_b200.so).
 *
 * These entry points replace, for a feature archive that lies in a regular file, the per-utterance Python of the
 * reference's extraction loop:
 *
 *     for key, mat in kaldi_io.read_mat_ark(input_stream):            -- local/tf/models.py:373, kaldi_io.py:372-437
 *         <skip rules, chunk plan>                                     -- models.py:377-409
 *         ... sess.run per chunk ...                                   -- models.py:410-419  (include/xvec.h)
 *         kaldi_io.write_vec_flt(output_stream, xvector_avg, key=key)  -- models.py:422, kaldi_io.py:309-343
 *
 * The reader indexes the archive's headers, applies make_embedding's skip and chunk rules, and a pool of threads preads
 * the payloads of whole batches straight into page-locked buffers that xv_submit_host_utts (include/xvec.h) consumes; the
 * formatter emits the exact bytes write_vec_flt would.  Host-only code: nothing here launches a kernel, and with
 * `pinned = 0` it needs no GPU at all (CPU tests).
 *
 * Multi-GPU: a job is STRIPED by byte ranges of the archive -- rank r owns the entries whose binary marker ("\0B", the
 * byte an scp offset points at) lies in [byte_begin, byte_end) -- so every rank reads a contiguous 1/N of the file and the
 * frames balance to within one utterance.  A stripe that does not start at a known entry boundary finds its first entry by
 * pattern search + forward validation; the caller cross-checks it against the previous stripe's chain
 * (xv_ark_index_info.next_marker_off) and calls xv_ark_reader_set_first when they agree (or re-indexes from the true
 * boundary when they do not), which also fixes where the first entry's key starts.
 *
 * Conventions as in xvec.h: XV_OK / negative XV_E*, xv_last_error(), no exceptions across the boundary.
 */
#ifndef XVEC_B200_JOB_H_
#define XVEC_B200_JOB_H_

#include <stddef.h>
#include <stdint.h>

#include "xvec.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xv_ark_reader xv_ark_reader;

typedef struct xv_ark_reader_opts {
  int32_t feat_dim;         /* columns every matrix must have (input_feature_dim)                                  */
  int32_t min_chunk_size;   /* make_embedding's arguments (models.py:356; run_xvector.sh:70,75)                    */
  int32_t chunk_size;       /* -1 = whole utterance                                                                */
  int32_t n_threads;        /* pread workers (>= 1)                                                                */
  int32_t n_slots;          /* batches in the ring (>= 2; 3 = one being filled while two are in flight)            */
  int32_t pinned;           /* 1: cudaHostAlloc'ed batch buffers; 0: plain memory (no CUDA call is made)            */
  int64_t batch_frames;     /* rows per batch (a longer utterance gets a batch of its own)                          */
  int64_t byte_begin;       /* stripe [byte_begin, byte_end) of the file; byte_end < 0 = end of file               */
  int64_t byte_end;
  int32_t begin_is_boundary;/* 1: byte_begin is the first byte of an entry (start of the stream); 0: resynchronise  */
  int32_t feats_f16;        /* 1: batches hold the rows rounded to float16 (IEEE RN-even, what the device's pack kernel does
                               to a feature first anyway: bit-identical x-vectors, half the pinned-buffer / PCIe bytes);
                               consume them with xv_submit_host_utts_f16                                              */
} xv_ark_reader_opts;

/* Why an utterance produced no x-vector (the reference's warnings, models.py:378-387). */
enum { XV_UTT_OK = 0, XV_UTT_ZERO_LENGTH = 1, XV_UTT_TOO_SHORT = 2 };

typedef struct xv_ark_index_info {
  int64_t n_entries;        /* matrices whose marker lies in the stripe                                             */
  int64_t n_ok, n_fail;     /* utterances that yield an x-vector / are skipped                                      */
  int64_t n_segments;       /* chunks (segments of xv_forward) over all ok utterances                               */
  int64_t rows_used;        /* rows fed to the network = the reference's total_segments_len                         */
  int64_t n_batches;
  int64_t first_marker_off; /* offset of the "\0B" marker of the stripe's first entry, -1 if the stripe is empty     */
  int64_t next_marker_off;  /* marker offset of the first entry BEYOND the stripe (file size if there is none) ...   */
  int64_t next_key_off;     /* ... and where that entry's key starts: what the next stripe's reader must be told      */
  int64_t stopped_at;       /* >= 0: offset of an entry this reader cannot parse (text / compressed matrix, malformed
                               header): the caller falls back to its general parser; -1 otherwise                    */
  int64_t key_bytes;        /* total length of the ok utterances' keys                                              */
} xv_ark_index_info;

typedef struct xv_ark_batch {
  int32_t slot;                     /* to hand back with xv_ark_reader_release                                       */
  int32_t n_seg, n_utt;             /* n_utt == 0: end of the stripe                                                 */
  int32_t feats_f16;                /* 1: feats points at float16 values (xv_ark_reader_opts.feats_f16)              */
  int64_t n_rows;
  const float* feats;               /* [n_rows, feat_dim] page-locked: the feats_host of xv_submit_host_utts (float32, or
                                       float16 bit patterns for xv_submit_host_utts_f16)                             */
  const int32_t* seg_len;           /* [n_seg]                                                                       */
  const int32_t* utt_first_seg;     /* [n_utt + 1]                                                                   */
  const int64_t* utt_dst_row;       /* [n_utt] = dst_row_base + index of the utterance among the stripe's ok ones     */
  int64_t first_ok_index;           /* index (within the stripe) of the batch's first utterance                       */
} xv_ark_batch;

/* Opens `path`, maps it for header scanning.  Nothing is read yet. */
int xv_ark_reader_open(xv_ark_reader** out, const char* path, const xv_ark_reader_opts* opts);
/* Indexes the stripe: headers (xv_ark_scan), skip rules, chunk plans, batch plan. */
int xv_ark_reader_index(xv_ark_reader* r, xv_ark_index_info* info);
/* Stripe r > 0 only, after the stripes' infos have been exchanged: `marker_off` / `key_off` are the previous stripe's
 * next_marker_off / next_key_off.  If marker_off equals this stripe's first_marker_off the first key is fixed; if not,
 * the stripe is re-indexed from key_off (the true boundary).  `info` is refreshed either way. */
int xv_ark_reader_set_first(xv_ark_reader* r, int64_t marker_off, int64_t key_off, xv_ark_index_info* info);
/* Keys of the ok utterances, in order: blob[key_off[i] .. key_off[i+1]) (no separators); key_off has n_ok + 1 entries. */
int xv_ark_reader_keys(const xv_ark_reader* r, char* blob, int64_t blob_cap, int64_t* key_off);
/* Skipped utterances in order: reason (XV_UTT_*), rows, and their keys as above.  Arrays of n_fail (+ 1) entries. */
int xv_ark_reader_failures(const xv_ark_reader* r, int32_t* reason, int32_t* rows, char* blob, int64_t blob_cap, int64_t* key_off);
/* Starts the pread workers.  utt_dst_row of every batch = dst_row_base + ok index. */
int xv_ark_reader_start(xv_ark_reader* r, int64_t dst_row_base);
/* Next batch, in order; blocks until its payloads are in memory.  n_utt == 0 after the last one. */
int xv_ark_reader_next(xv_ark_reader* r, xv_ark_batch* batch);
/* The batch's buffers may be refilled (call once the submission that read them has been collected). */
int xv_ark_reader_release(xv_ark_reader* r, int32_t slot);
void xv_ark_reader_close(xv_ark_reader* r);

/* The bytes kaldi_io.write_vec_flt (reference kaldi_io.py:309-343) emits for n float32 vectors of `dim` values, entry i
 * keyed blob[key_off[i] .. key_off[i+1]):  key ' ' '\0' 'B' 'F' 'V' ' ' '\4' <uint32 dim> <dim float32 LE>.
 * `out` needs xv_vec_ark_bytes(...) bytes.  marker_off (may be NULL) receives, per entry, the offset within `out` of its
 * "\0B" marker -- what an scp line points at.  Formats with up to n_threads threads.  Returns the bytes written or < 0. */
int64_t xv_vec_ark_bytes(const int64_t* key_off, int64_t n, int32_t dim);
int64_t xv_vec_ark_format(const char* key_blob, const int64_t* key_off, int64_t n, const float* vecs, int32_t dim,
                          uint8_t* out, int64_t out_cap, int64_t* marker_off, int32_t n_threads);
/* The scp lines Kaldi's table writer puts beside such an ark: "<key> <ark_name>:<base + marker_off[i]>\n".
 * Returns the bytes written (out_cap >= sum(key_len) + n * (strlen(ark_name) + 24) is always enough) or < 0. */
int64_t xv_scp_format(const char* key_blob, const int64_t* key_off, int64_t n, const char* ark_name, int64_t base,
                      const int64_t* marker_off, char* out, int64_t out_cap);

/* xv_submit_host_utts (include/xvec.h) for feature rows that are ALREADY rounded to float16 on the host (a reader opened with
 * feats_f16, or xv_convert_f32_to_f16_host): feats_host_f16 holds [total_frames, feat_dim] float16 bit patterns.  Results are
 * bit-identical to the float32 call -- the device rounds features to fp16 before anything else -- except with the
 * split-precision option, which needs the float32 values (XV_EINVAL). */
int xv_submit_host_utts_f16(xv_model* m, const uint16_t* feats_host_f16, const int32_t* seg_len_host, int32_t n_seg,
                            const int32_t* utt_first_seg_host, const int64_t* utt_dst_row_host, int32_t n_utt, float* out_dev,
                            float* out_host, int32_t* ticket);
/* float32 -> float16, IEEE round-to-nearest-even, on the host (F16C + AVX2 with non-temporal stores when the CPU has them;
 * the _scalar form is the portable routine both are tested against). */
void xv_convert_f32_to_f16_host(const float* src, uint16_t* dst, size_t n);
void xv_convert_f32_to_f16_host_scalar(const float* src, uint16_t* dst, size_t n);

/* Synthetic workload (bench / scale tests only; SURVEY 8d, BASELINE configs[3]): the MFCC rows of n_utt utterances generated
 * on the device from (seed, utterance id, frame, coefficient) -- out_dev [sum(len), feat_dim] fp32, utterances concatenated
 * in order.  utt_id_host / len_host are HOST arrays; the same integer recipe in numpy (synthetic.counter_mfcc) reproduces any
 * utterance bit for bit.  Enqueue only (two small host->device copies ride on `stream`). */
int xv_synth_mfcc(int device, float* out_dev, const int64_t* utt_id_host, const int32_t* len_host, int32_t n_utt, int32_t feat_dim,
                  uint64_t seed, void* stream);

/* xv_submit_host_utts (include/xvec.h) for features that are ALREADY on the device (they must stay untouched until the
 * ticket is collected: an fp16 range rescue re-runs the submission from them).  ready_event: a cudaEvent_t recorded behind
 * the work that produces the features on another stream (the submission waits for it on the device), or NULL. */
int xv_submit_dev_utts(xv_model* m, const float* feats_dev, const int32_t* seg_len_host, int32_t n_seg,
                       const int32_t* utt_first_seg_host, const int64_t* utt_dst_row_host, int32_t n_utt, float* out_dev,
                       float* out_host, void* ready_event, int32_t* ticket);

#ifdef __cplusplus
}
#endif
#endif /* XVEC_B200_JOB_H_ */
