/* xvec.h -- C ABI of libxvec_b200.so: the B200 (sm_100a) x-vector extraction hot path.
 *
 * This is the drop-in boundary for the ONE operator call the reference makes into its
 * runtime on the extraction path:
 *
 *     xvector = sess.run(self.embedding[0],
 *                        feed_dict={input_x: data[1,T,D], dropout_keep_prob: 1.0, phase: False})
 *                                                       -- reference local/tf/models.py:412-415
 *
 * i.e. "features of one chunk in, 512-dim embed_layer-0/scores out", evaluated over the graph
 * that Model.build_model declares (local/tf/models.py:441-534 / :543-639) with the variables
 * that Model.load_model restores by NAME (local/tf/models.py:143-162, :199-210).
 * Each entry point below names the reference interface it replaces.
 *
 * Conventions: plain C types only (no torch / C++ types); every function returns XV_OK (0)
 * or a negative XV_E* code and never throws; xv_last_error() gives a thread-local message;
 * the caller owns every buffer it passes in, the model owns its weights.  One xv_model is
 * bound to one CUDA device; calls on the same model must not run concurrently.
 *
 * "Segment" = one chunk of one utterance as cut by make_embedding (models.py:388-410): the
 * rows of ONE sess.run call.  A call here evaluates many segments at once.
 */
#ifndef XVEC_B200_H_
#define XVEC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XV_MAX_FRAME_LAYERS 8
#define XV_ABI_VERSION 2        /* bumped whenever a struct or a signature below changes */

enum {
  XV_OK = 0,
  XV_EINVAL = -1,      /* bad argument / unknown variable name / shape mismatch            */
  XV_ECUDA = -2,       /* CUDA runtime or driver error (message has the CUDA error string) */
  XV_ENOMEM = -3,      /* workspace too small / allocation failed                          */
  XV_ESTATE = -4,      /* parameters missing at forward time                               */
  XV_EOVERFLOW = -5    /* an activation left the fp16 range and was not rescued (see xv_check_overflow) */
};

/* Frame-layer nonlinearity (between bias_add and BatchNorm):
 *   XV_ACT_RELU   tf.nn.relu                    (models.py:479; Model, ModelWithoutDropout[Tdnn], ...ReluHeInit)
 *   XV_ACT_LRELU  tf.nn.leaky_relu(alpha=0.2)   (models.py:912; ModelL2LossWithoutDropoutLRelu)
 *   XV_ACT_PRELU  prelu(h, shared=False)        (tf_block.py:38-47; per-channel slope "<scope>/prelu/prelu:0") */
enum { XV_ACT_RELU = 0, XV_ACT_LRELU = 1, XV_ACT_PRELU = 2 };

/* Pooling between the frame level and the segment level:
 *   XV_POOL_STATS      mean | sqrt(var + 1e-5) over time                        (models.py:485-486)
 *   XV_POOL_ATTENTION  the last frame layer (width 2C) is split into h1 | h2; attention = softmax over time of
 *                      v . tanh(h1 W + b) ("attention/w:0" [C,C], "attention/b:0", "attention/v:0" [C]); weighted mean and
 *                      sqrt(weighted var + 1e-5) of h2       (ModelL2LossWithoutDropoutLReluAttention, models.py:1037-1051) */
enum { XV_POOL_STATS = 0, XV_POOL_ATTENTION = 1 };

/* Topology = the constants hard-coded in each build_model body
 * (kernel_sizes / dilation_rates / layer_sizes: models.py:443-445, :545-548). */
typedef struct xv_topology {
  int32_t feat_dim;                         /* input_feature_dim (23: conf/mfcc.conf)        */
  int32_t n_frame_layers;                   /* 5                                             */
  int32_t taps[XV_MAX_FRAME_LAYERS];        /* kernel_sizes, odd                             */
  int32_t dilation[XV_MAX_FRAME_LAYERS];    /* dilation_rates (1 for tf.nn.conv1d)           */
  int32_t width[XV_MAX_FRAME_LAYERS];       /* layer_sizes; multiples of 256                 */
  int32_t emb_dim;                          /* embedding_sizes[0] = 512                      */
  int32_t act;                              /* XV_ACT_*                                      */
  float bn_eps;                             /* 1e-3  (tf_block.py:9)                         */
  float var_eps;                            /* 1e-5  (models.py:16 VAR2STD_EPSILON)          */
  int32_t pooling;                          /* XV_POOL_*                                     */
} xv_topology;

typedef struct xv_model xv_model;           /* opaque: device weights, metadata, tensor maps */

/* Replaces graph construction / import (build_model, tf.train.import_meta_graph: models.py:146). */
int xv_create(xv_model** out, int device, const xv_topology* topo);
void xv_destroy(xv_model* m);

/* Replaces saver.restore's per-variable assignment (models.py:147).  `tf_var_name` is the
 * reference's variable name, e.g. "frame_level_info_layer-2/w:0" [k,Cin,Cout],
 * ".../b:0", ".../gamma:0", ".../beta:0", ".../mean:0", ".../variance:0", ".../prelu/prelu:0" (XV_ACT_PRELU),
 * "embed_layer-0/w:0" [2*C,emb], "embed_layer-0/b:0".  `host` is fp32, C-contiguous, and is
 * copied.  Names of variables the extraction path never reads (embed_layer-1/..., output/...)
 * are accepted and ignored. */
int xv_set_param(xv_model* m, const char* tf_var_name, const float* host,
                 const int64_t* shape, int32_t rank);

/* Bytes of device scratch a forward over `n_seg` segments with `total_frames` rows needs. */
size_t xv_workspace_bytes(const xv_model* m, int64_t total_frames, int32_t n_seg);

/* Replaces sess.run(embedding[0], ...) (models.py:414) for a whole batch of segments.
 *   feats_dev    [total_frames, feat_dim] fp32, device, segments concatenated in order
 *   seg_len_host [n_seg] int32, HOST: rows of each segment (each >= 1)
 *   emb_dev      [n_seg, emb_dim] fp32, device: embed_layer-0/scores per segment
 *   workspace_dev / workspace_bytes: device scratch, >= xv_workspace_bytes(...)
 *   stream       cudaStream_t (NULL = legacy default stream)
 * Enqueue only: no host synchronisation, no device allocation. */
int xv_forward(xv_model* m, const float* feats_dev, const int32_t* seg_len_host, int32_t n_seg,
               float* emb_dev, void* workspace_dev, size_t workspace_bytes, void* stream);

/* Debug / parity: as xv_forward, additionally writes each frame layer's output (after
 * ReLU+BatchNorm, i.e. the tensor models.py:480 produces) as fp32 [total_frames, width[i]]
 * into layer_out_dev[i] (NULL entries are skipped) and the pooled statistics
 * [n_seg, 2*width[last]] into stats_out_dev (may be NULL). */
int xv_forward_layers(xv_model* m, const float* feats_dev, const int32_t* seg_len_host,
                      int32_t n_seg, float* emb_dev, void* workspace_dev, size_t workspace_bytes,
                      void* stream, float* const* layer_out_dev, float* stats_out_dev);

/* Host-buffer form of the same call: the shape of the reference's operator boundary (host
 * numpy in, host numpy out; models.py:410-415).  Copies feats host->device, runs the forward,
 * copies embeddings device->host and synchronises.  Buffers should be page-locked for full
 * PCIe speed.  Device scratch is owned (and grown on demand) by the model. */
int xv_extract_host(xv_model* m, const float* feats_host, const int32_t* seg_len_host,
                    int32_t n_seg, float* emb_host);

/* Pipelined form of xv_extract_host.  xv_submit_host enqueues (copy in, forward, copy out) on one
 * of XV_HOST_SLOTS internal streams and returns at once with a ticket; xv_collect waits for that
 * submission and reports XV_EOVERFLOW like xv_extract_host.  With two submissions in flight the
 * host->device copy of one batch overlaps the kernels of the previous one.  Both host buffers of
 * a submission must stay valid (and should be page-locked) until it is collected; tickets must
 * be collected before their slot is reused (XV_ESTATE otherwise).  xv_extract_host is
 * xv_submit_host + xv_collect.  (The reference overlaps nothing: one blocking sess.run per
 * utterance, models.py:401-419.) */
#define XV_HOST_SLOTS 2
int xv_submit_host(xv_model* m, const float* feats_host, const int32_t* seg_len_host,
                   int32_t n_seg, float* emb_host, int32_t* ticket);
int xv_collect(xv_model* m, int32_t ticket);

/* Utterance-level form of xv_forward: the chunk loop of make_embedding (models.py:398-421) finished on the device.
 * Segments [utt_first_seg_host[u], utt_first_seg_host[u+1]) are the chunks of utterance u (utt_first_seg_host: [n_utt+1],
 * HOST, starts at 0, ends at n_seg; NULL = every segment is its own utterance and n_utt is ignored).  Row
 * utt_dst_row_host[u] (HOST int64; NULL = u) of out_dev receives
 *     xvector_avg = sum_c float32(len_c) * xvector_c  (chunk order, float32, separately rounded multiply and add)
 *                   / float32(sum_c len_c)
 * i.e. exactly the reference's float32 arithmetic, so the row is bit-identical to that loop run over the same chunk
 * x-vectors.  out_dev may be PEER device memory (xv_peer_open): in a multi-GPU job every rank stores its utterances
 * straight into rank 0's result table over NVLink and no gather collective is left (replaces the per-job arks +
 * `cat xvector.*.scp` merge of local/tf/extract_xvectors.sh:83-95).  Enqueue only. */
int xv_forward_utts(xv_model* m, const float* feats_dev, const int32_t* seg_len_host, int32_t n_seg,
                    const int32_t* utt_first_seg_host, const int64_t* utt_dst_row_host, int32_t n_utt, float* out_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/* xv_submit_host with utterance-level output: rows go to out_dev[utt_dst_row_host[u]] (device or peer memory; may be
 * NULL) and / or contiguously, in utterance order, to out_host[u] (HOST, should be page-locked; may be NULL).
 * Collected with xv_collect like any other ticket. */
int xv_submit_host_utts(xv_model* m, const float* feats_host, const int32_t* seg_len_host, int32_t n_seg,
                        const int32_t* utt_first_seg_host, const int64_t* utt_dst_row_host, int32_t n_utt, float* out_dev,
                        float* out_host, int32_t* ticket);

/* Peer memory for the result table of a multi-GPU job on one node (one process per GPU).  The owner (rank 0) allocates
 * `bytes` of device memory and gets a 64-byte handle (a cudaIpcMemHandle_t) to hand to the other processes by any means
 * (torch.distributed broadcast, a file); they map it with xv_peer_open and pass the mapped pointer as the out_dev of
 * xv_forward_utts / xv_submit_host_utts.  A writer's rows are visible to the owner once the writer has synchronised its
 * stream (xv_collect does) and the two processes have met at a barrier.  xv_peer_read is a blocking device->host copy for
 * owners that hold no other handle on the memory. */
int xv_peer_alloc(int device, size_t bytes, void** dev_ptr, uint8_t* handle64);
int xv_peer_open(int device, const uint8_t* handle64, void** dev_ptr);
int xv_peer_close(int device, void* dev_ptr);
int xv_peer_free(int device, void* dev_ptr);
int xv_peer_read(int device, void* dst_host, const void* src_dev, size_t bytes);

/* fp16 range.  Activations travel between the frame layers as fp16 (|x| <= 65 504); the reference computes in fp32 and
 * loads any trained model (models.py:476-480).  Every store tracks its largest magnitude, and a model whose activations
 * pass the range is RESCUED rather than refused: layer i's rows are then stored divided by 2^e_i (folded into its
 * BatchNorm scale / shift, undone by the consumer's fp32 accumulator -- exact operations), likewise the pooled
 * statistics on their way into the embedding GEMM.
 *   - xv_collect / xv_extract_host do this by themselves: exponents are raised for what overflowed and the submission is
 *     run again from its device copy of the features; the exponents stay with the model (option "rescue" = 0 turns this
 *     off: XV_EOVERFLOW is returned instead, as in ABI version 1).
 *   - xv_forward / xv_forward_utts only enqueue, so their caller checks: xv_check_overflow returns XV_OK, or
 *     XV_EOVERFLOW if any store overflowed since the last call (synchronises `stream`; clears the flag);
 *     xv_rescue_overflow does the same check and, on overflow, raises the exponents and returns how many it raised
 *     (> 0: enqueue the forward again; 0: nothing overflowed; < 0: XV_E*, e.g. beyond the largest rescue scale 2^96). */
int xv_check_overflow(xv_model* m, void* stream);
int xv_rescue_overflow(xv_model* m, void* stream);

/* Number of kernels the last xv_forward / xv_extract_host / xv_submit_host launched (for bench.py's
 * "gpu_launches" claim). */
int32_t xv_last_launch_count(const xv_model* m);

/* With option "profile" = 1 every kernel launch of a forward is bracketed by a CUDA event pair
 * on the launching stream; this returns the number of launches of the last forward and writes
 * their device durations (ms, launch order: pack, frame layers 0..n-1, pool+embed) into
 * ms_out[0..cap).  Synchronises on the last event.  Negative XV_E* on error.  (The reference has
 * only wall-clock deltas around sess.run, models.py:413-417.) */
int32_t xv_last_kernel_ms(xv_model* m, float* ms_out, int32_t cap);

/* Options of a model (integers; unknown names are refused with XV_EINVAL):
 *   "precision"        1 = split-precision operands (two fp16 terms per activation and weight; default for attention pooling)
 *   "fuse_first"       1 (default) = the input splice and the first frame layer as ONE kernel (tdnn_first.cuh) when the
 *                      topology allows it; 0 = pack_im2col_kernel + the layer kernel.  Same bits either way.
 *   "fuse_tail"        1 (default) = the last two context-free frame layers as ONE kernel (tdnn_tail.cuh); 0 = one launch each
 *   "fc"               1 (default) = embed_layer-0 on the tensor cores (split fp16), 0 = fp32 SIMT GEMM; "fc_max_splits"
 *   "pdl"              programmatic dependent launch between the kernels of a forward (default 1)
 *   "rescue"           1 (default) = xv_collect raises the fp16 range exponents and re-runs a batch that overflowed
 *   "blocking_collect" 1 = xv_collect sleeps on a blocking-sync event instead of spinning
 *   "clusters"         CTA pairs per persistent grid (default: one per pair of SMs); a data-parallel trainer may lower it
 *   "resident", "prefetch"   measured-not-faster schedules of the layer kernel, kept for experiments
 *   "profile"          1 = CUDA events around every launch (xv_last_kernel_ms); "trace_ptr" / "trace_layer": per-tile clock
 *                      stamps of one layer (tools/trace_tiles.py) */
int xv_set_option(xv_model* m, const char* name, int64_t value);

/* Host-only helper of the reader that feeds xv_submit_host: index of the binary float matrices of a Kaldi ark held in
 * memory (an mmap'ed file), replacing the per-utterance header parsing of read_mat_ark / _read_mat_binary
 * (reference local/tf/kaldi_io.py:372-392, :413-437; key rules :120-133).  For entry i < return value:
 * key = buf[key_off[i] .. +key_len[i]), a rows[i] x cols[i] matrix of elem_bytes[i] (4 = 'FM ', 8 = 'DM ') per element
 * at buf + payload_off[i].  Stops at max_entries, at the end of the buffer, or at the first entry of another kind (text,
 * compressed, malformed, truncated); *consumed = offset of the first unparsed byte.  Returns the number of entries, -1 on
 * a null argument. */
int64_t xv_ark_scan(const uint8_t* buf, int64_t len, int64_t max_entries, int64_t* key_off, int32_t* key_len,
                    int32_t* rows, int32_t* cols, int32_t* elem_bytes, int64_t* payload_off, int64_t* consumed);

const char* xv_last_error(void);
const char* xv_version(void);
/* Guards for bindings written against this header: XV_ABI_VERSION of the built library and
 * sizeof(xv_topology) as the library sees it -- a binding whose struct is shorter must refuse to call xv_create. */
int32_t xv_abi_version(void);
size_t xv_topology_size(void);

#ifdef __cplusplus
}
#endif
#endif /* XVEC_B200_H_ */
