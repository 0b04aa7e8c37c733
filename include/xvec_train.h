/* xvec_train.h -- C ABI of the training step in libxvec_b200.so (B200, sm_100a).
 *
 * Drop-in boundary for the ONE operator call the reference makes into its runtime per minibatch:
 *
 *     _, loss, accuracy = sess.run([self.optimizer, self.loss, self.accuracy],
 *                                  feed_dict={input_x: batch[B,T,D], input_y: one_hot[B,C],
 *                                             dropout_keep_prob, learning_rate, phase: True})
 *                                                       -- reference local/tf/models.py:258-263
 *
 * on the graph of ModelWithoutDropout / ModelWithoutDropoutTdnn.build_model (models.py:441-534 / :543-639):
 * frame layers conv -> +b -> relu -> BatchNorm(training branch, tf_block.py:18-23), statistics pooling,
 * two segment layers, softmax cross-entropy over num_classes (models.py:512-514), tf.train.AdamOptimizer
 * (models.py:516-519).  The call is split in two so that a data-parallel caller can all-reduce the flat
 * gradient between the halves (torch.distributed / NCCL on the Python side):
 *
 *     xv_train_forward_backward  ->  [all-reduce grad]  ->  xv_train_apply
 *
 * State lives in flat fp32 device vectors addressed by TF variable name through xv_train_span:
 *     which = XV_TRAIN_PARAMS   trainable variables, in graph order (w, b, gamma, beta per layer; output/w, output/b)
 *             XV_TRAIN_ADAM_M / XV_TRAIN_ADAM_V   the "<var>/Adam" and "<var>/Adam_1" slots, same offsets
 *             XV_TRAIN_MOVING   the non-trainable "<scope>/mean:0", "<scope>/variance:0" moving statistics
 *             XV_TRAIN_GRAD     the gradient of the last xv_train_forward_backward into the trainer's own buffer
 * Conventions as in xvec.h: plain C, 0 / negative XV_E*, xv_last_error().  Topologies: ReLU and leaky-ReLU (XV_ACT_RELU /
 * XV_ACT_LRELU) with statistics pooling; option "l2_beta" adds the L2 term of the ModelL2Loss* graphs (models.py:930-962).
 */
#ifndef XVEC_B200_TRAIN_H_
#define XVEC_B200_TRAIN_H_

#include "xvec.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xv_trainer xv_trainer;

enum { XV_TRAIN_PARAMS = 0, XV_TRAIN_ADAM_M = 1, XV_TRAIN_ADAM_V = 2, XV_TRAIN_MOVING = 3, XV_TRAIN_GRAD = 4 };
#define XV_TRAIN_GRAD_TAIL 4     /* floats behind the gradient: [0] = the step's gradient-overflow flag (see xv_train_apply) */

/* Replaces Model.build_model's graph construction for training (models.py:441-534): the frame-level topology is
 * the xv_model's; emb1_dim = embedding_sizes[1] (512), num_classes = width of "output/w".  The xv_model supplies
 * the device, the kernels' launch state and the overflow flag; it must outlive the trainer. */
int xv_train_create(xv_trainer** out, xv_model* model, int32_t num_classes, int32_t emb1_dim);
void xv_train_destroy(xv_trainer* t);

/* Number of floats in the flat vector `which`. */
int64_t xv_train_size(const xv_trainer* t, int32_t which);
/* Where variable `tf_var_name` (e.g. "frame_level_info_layer-1/w:0", "embed_layer-0/mean:0") lives:
 * *which = XV_TRAIN_PARAMS or XV_TRAIN_MOVING, [*offset, *offset + *count).  Replaces the by-name binding of
 * Saver.restore / graph.get_tensor_by_name (models.py:143-162). */
int xv_train_span(const xv_trainer* t, const char* tf_var_name, int32_t* which, int64_t* offset, int64_t* count);
/* Host <-> device copies of a range of a flat vector (synchronous).  After an upload into XV_TRAIN_PARAMS the
 * fp16 operand copies are rebuilt on the next step. */
int xv_train_upload(xv_trainer* t, int32_t which, const float* host, int64_t offset, int64_t count);
int xv_train_download(xv_trainer* t, int32_t which, float* host, int64_t offset, int64_t count);
/* Adam's step counter t (TF keeps beta1_power = b1^t, beta2_power = b2^t). */
int xv_train_set_step(xv_trainer* t, int64_t step);
int64_t xv_train_get_step(const xv_trainer* t);

/* Forward + backward of one minibatch (the loss/accuracy/gradient half of models.py:263).
 *   feats_dev   [n_seg * seg_len, feat_dim] fp32 on the device (input_x, every segment seg_len rows)
 *   labels_dev  [n_seg] int32 class ids on the device (the argmax of input_y's one-hot rows, models.py:164-169)
 *   grad_dev    [xv_train_size(XV_TRAIN_GRAD)] fp32 out (gradient + XV_TRAIN_GRAD_TAIL floats), or NULL to use the trainer's own buffer (XV_TRAIN_GRAD)
 *   loss_acc_dev [2] fp32 out: mean cross-entropy, accuracy
 * Also applies the moving-statistics update of every BatchNorm (tf_block.py:20-21).  Enqueue only. */
int xv_train_forward_backward(xv_trainer* t, const float* feats_dev, const int32_t* labels_dev, int32_t n_seg,
                              int32_t seg_len, float* grad_dev, float* loss_acc_dev, void* stream);
/* The same step in two halves, so that a data-parallel caller can hide most of its gradient all-reduce (the reference has no
 * such seam: one sess.run per minibatch).  part 1: forward, loss and the SEGMENT-level backward -- when it ends (in stream order)
 * the gradients [xv_train_segment_grad_offset(t), n_params) of embed_layer-*, output/* (60 % of the bytes) are final and their
 * all-reduce can start on another stream; part 2: pooling and frame-level backward, which fills [0, offset) and the overflow
 * flag behind the gradient; part 0 = both (xv_train_forward_backward).  Same arithmetic, bit for bit. */
int xv_train_forward_backward_part(xv_trainer* t, const float* feats_dev, const int32_t* labels_dev, int32_t n_seg,
                                   int32_t seg_len, float* grad_dev, float* loss_acc_dev, void* stream, int32_t part);
int64_t xv_train_segment_grad_offset(const xv_trainer* t);
/* part = XV_TRAIN_PART_FRAME + i: the slice of part 2 that ends with the gradients of frame layer i final -- for the top layer
 * it starts with the pooling backward, for layer 0 it ends with the overflow flag behind the gradient.  Run from the top layer
 * down after part 1 they are part 2, bit for bit; a data-parallel caller all-reduces [offset, offset + count) of
 * xv_train_frame_grad_span(t, i) after slice i, under the backward of the layers below. */
#define XV_TRAIN_PART_FRAME 16
int xv_train_frame_grad_span(const xv_trainer* t, int32_t layer, int64_t* offset, int64_t* count);
/* Loss and accuracy of one minibatch with phase = False (moving statistics, nothing is updated): the
 * sess.run([self.loss, self.accuracy]) of Model.eval (models.py:338-339). */
int xv_train_eval(xv_trainer* t, const float* feats_dev, const int32_t* labels_dev, int32_t n_seg, int32_t seg_len,
                  float* loss_acc_dev, void* stream);
/* Adam update (the optimizer half of models.py:263) with gradient grad_dev * grad_scale (NULL = own buffer;
 * grad_scale = 1/world_size after a sum all-reduce), then refreshes the fp16 operand copies.  Enqueue only.
 * grad_dev holds xv_train_size(XV_TRAIN_GRAD) = n_params + XV_TRAIN_GRAD_TAIL floats: behind the gradient rides the
 * step's gradient-overflow flag (0 / 1, written by xv_train_forward_backward), so that a data-parallel caller's sum
 * all-reduce of the WHOLE buffer combines it over the ranks.  If the combined flag is non-zero -- a loss-scaled fp16
 * gradient left the fp16 range on any rank -- the update is skipped on the device on every rank alike (variables and slots
 * untouched) and counted; the caller reads the count with xv_train_skipped_updates and lowers "loss_scale".  Overflow of
 * a forward ACTIVATION is a different condition (no loss scale can help): it is the xv_model's flag, reported by
 * xv_check_overflow. */
int xv_train_apply(xv_trainer* t, const float* grad_dev, float learning_rate, float grad_scale, void* stream);
/* Updates skipped so far because of gradient overflow.  blocking = 0: never waits -- returns the newest value whose
 * read-back has landed and starts another read-back on `stream` (call once per step); blocking = 1: synchronises. */
int64_t xv_train_skipped_updates(xv_trainer* t, void* stream, int32_t blocking);

/* Copies the current variables and moving statistics into the xv_model, so that xv_forward / xv_extract_host
 * evaluate the trained network: what Model.save_model + load_model do between train and extract. */
int xv_train_sync_model(xv_trainer* t);

/* Parity hook: fp32 copy of an intermediate of the last forward_backward.  Packed fp16 tensors come back as
 * [n_seg * seg_len, C]:  "r<i>" relu output and "y<i>" BatchNorm output of frame layer i, "dz<i>" / "dy<i>"
 * loss-scaled gradients w.r.t. the pre-activation / the BatchNorm output; fp32 arrays as stored: "h0" (pooled
 * statistics), "z5", "y5", "z6", "y6", "logits", "dlogits", "dh0".  Returns the number of floats (or < 0). */
int64_t xv_train_debug_tensor(xv_trainer* t, const char* name, float* host_out, int64_t capacity);
/* float16 -> float32 on the device (the egs archives store minibatches as float16, examples_io.py:165; input_x is float32):
 * lets the host ship half the bytes and skip its own conversion.  Both pointers 16-byte aligned.  Enqueue only. */
int xv_convert_f16_to_f32(const void* src_dev, float* dst_dev, int64_t n, void* stream);
/* Options: "graph" (1: a step's kernels are captured once per (geometry, buffers) into a CUDA graph and replayed),
 * "loss_scale" (0 = automatic: 8 * frames rounded to a power of two), "l2_beta" (0; 0.0002 for ModelL2Loss*),
 * "wgrad_lbo", "wgrad_sbo" (diagnostics), "wgrad_reuse", "fused_stats" (alternative schedules). */
int xv_train_set_option(xv_trainer* t, const char* name, double value);
int32_t xv_train_last_launch_count(const xv_trainer* t);
/* With the xv_model option "profile" on, every launch of a step is bracketed by CUDA events: read the times with
 * xv_last_kernel_ms(model, ...) and the kernel names, ';'-separated and in the same order, here.  Returns the count. */
int64_t xv_train_last_kernel_names(const xv_trainer* t, char* buf, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* XVEC_B200_TRAIN_H_ */
