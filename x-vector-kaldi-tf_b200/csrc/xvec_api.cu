// xvec_api.cu -- host side of libxvec_b200.so: the C ABI declared in include/xvec.h.
// Owns parameter packing (TF variable names -> device layouts), the workspace plan, TMA tensor
// maps and the kernel launch sequence  pack -> 5 x tdnn_layer -> pool_embed.
#include "../../include/xvec.h"

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "pack_pool.cuh"
#include "tdnn_pair.cuh"
#include "tdnn_tail.cuh"
#include "tdnn_first.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define XV_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(XV_ECUDA, std::string(#expr) + ": " + cudaGetErrorName(e_) + ": " + cudaGetErrorString(e_)); \
  } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct FrameLayer {
  int taps = 0, dilation = 1, c_in = 0, c_in_pad = 0, c_out = 0, k_total = 0;   // k_total = packed K (multiple of 64)
  int gemm_taps = 0;           // taps seen by the GEMM kernel (1 for the im2col'ed first layer)
  __half* w_dev = nullptr;     // [c_out, k_total] fp16, K-major
  float* bias_dev = nullptr;   // [c_out]
  float* scale_dev = nullptr;  // [c_out]
  float* shift_dev = nullptr;  // [c_out]
  float* alpha_dev = nullptr;  // [c_out] negative slope (leaky / PReLU topologies), else null
  // fp16 range rescue: this layer's output rows are stored divided by 2^exp_out (scale_dev / shift_dev hold the folded
  // BatchNorm times 2^-exp_out; the next layer's epilogue multiplies its accumulator by 2^exp_out).  Powers of two: exact.
  int exp_out = 0;
  std::vector<float> scale_host, shift_host;     // the unscaled folded BatchNorm
  std::vector<float> bias_host, alpha_host;      // host copies for kernels that take them by value (tdnn_tail.cuh)
};

struct Plan {
  int64_t total_frames = 0;
  int32_t n_seg = 0;
  int64_t r_pad = 0;
  int32_t fc_m_tiles = 0, fc_n_tiles = 0, fc_splits = 1, fc_k_per_split = 0, n_counters = 0;
  int32_t fc_counters = 0;
  int32_t tc_splits = 1;             // K-splits of the tensor-core embedding GEMM
  size_t off_split = 0;
  size_t off_seg_emb = 0;            // [n_seg, E] per-segment embeddings when the caller asks for utterance averages
  size_t off_utt = 0;                // utterance plan: first_seg (n_seg + 1 int32, padded) | dst_row (n_seg int64)
  size_t off_score = 0, off_attn = 0;
  size_t off_meta = 0, off_counters = 0, off_valid = 0, off_blk_valid = 0, off_x0 = 0, off_ha = 0, off_hb = 0,
         off_hlast = 0, off_pool_partial = 0, off_stats = 0, off_partial = 0, bytes = 0;
};

constexpr int META_SLOTS = 4;
constexpr int64_t FC_GROUP = 1024;   // segments per launch of the tensor-core embedding GEMM (bounds its K-split partials)

}  // namespace

struct xv_model {
  xv_topology topo{};
  int device = 0;
  int num_sms = 148;
  int gap = 0;                       // zero rows between segments = max half context of any layer
  int k0_pad = 0;                    // packed width of the first layer's spliced input
  int w_mid = 0;                     // widest of layers 0..n-2 (ping-pong buffers)
  std::map<std::string, std::vector<float>> host_params;
  std::map<std::string, std::vector<int64_t>> host_shapes;
  bool dirty = true;
  std::vector<FrameLayer> layers;
  __half* w0_split_dev = nullptr;    // [E, 3 * 2C] fp16 K-major [hi | lo | hi]: B operand of the split-precision GEMM
  int opt_fc_max_splits = 36;        // cap on the K-splits of the tensor-core embedding GEMM
  int opt_pdl = 1;                   // programmatic dependent launch between the kernels of a forward
  int opt_fc = 1;                    // 1: embed_layer-0 on tensor cores (split fp16), 0: fp32 SIMT GEMM
  int opt_blocking_collect = 0;      // 1: xv_collect sleeps on a blocking-sync event instead of spinning
  int opt_fuse_tail = 1;             // 1: the last two (context-free) frame layers of a statistics-pooling model run as ONE kernel
                                     // that keeps the 512-wide intermediate in shared memory (tdnn_tail.cuh); 0: one launch per layer
  int opt_split = 0;                 // option "precision" = 1: split-precision operands (two fp16 terms per activation and per
                                     // weight, three products per contraction: ~2^-22 relative instead of 2^-11; 3x the MMA work).
                                     // Default for attention pooling, whose softmax over time turns absolute score errors into
                                     // relative weight errors: plain fp16 misses the 1e-3 gate there (tests/test_precision_model.py)
  int c_pool = 0;                    // channels that are pooled: width of the last frame layer (attention pooling: half of it)
  __half* att_w_dev = nullptr;       // attention pooling: [C, C] fp16 K-major (out channel major) copy of "attention/w:0"
  float* att_b_dev = nullptr;        // [C]
  float* att_v_dev = nullptr;        // [C]
  float* w0_dev = nullptr;           // [2C, E]
  float* b0_dev = nullptr;           // [E]
  int32_t* pack_lut_dev = nullptr;   // [k0_pad] spliced column -> staged feature offset (pack kernel)
  bool first_fusable = false;        // the topology fits tdnn_first_kernel (tdnn_first.cuh)
  int opt_fuse_first = 1;            // 1: the input splice and the first frame layer run as ONE kernel; 0: pack_im2col + tdnn_pair
  uint32_t* overflow_dev = nullptr;  // [1 + XV_HOST_SLOTS] flag words: [0] xv_forward / training, [1 + s] submission slot s
  uint32_t* cur_flag = nullptr;      // the word the kernels of the enqueue in progress report to
  int cur_feats_f16 = 0;             // 1 while an enqueue's features are float16 values (xv_submit_host_utts_f16)
  uint32_t* overflow_host = nullptr; // pinned
  int exp_stats = 0;                 // fp16 range rescue of the pooled statistics (see apply_exponents)
  int opt_rescue = 1;                // 1: xv_collect raises exponents and re-runs a batch whose fp16 stores overflowed
  int rescues = 0;                   // batches re-run so far
  EncodeTiledFn encode = nullptr;
  // pinned staging ring for segment metadata
  int32_t* meta_host[META_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t meta_event[META_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  int64_t meta_cap = 0;
  int meta_next = 0;
  // same ring for the utterance plan of xv_forward_utts (first_seg | dst_row)
  int32_t* utt_host[META_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t utt_event[META_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
  int64_t utt_cap = 0;
  int utt_next = 0;
  // xv_extract_host / xv_submit_host state: two independent submission slots (own stream and device
  // buffers) so that the host->device copy of one batch overlaps the kernels of the previous one
  struct HostSlot {
    cudaStream_t stream = nullptr;
    float* feats_dev = nullptr; size_t feats_cap = 0;
    float* emb_dev = nullptr; size_t emb_cap = 0;
    void* ws_dev = nullptr; size_t ws_cap = 0;
    // xv_submit_host_raw (include/xvec_frontend.h): raw rows, VAD track and scratch of the feature front end
    float* raw_dev = nullptr; size_t raw_cap = 0;
    float* vad_dev = nullptr; size_t vad_cap = 0;
    void* fe_ws_dev = nullptr; size_t fe_ws_cap = 0;
    uint32_t* overflow_host = nullptr;   // pinned
    // what a re-run after an fp16 range rescue needs (the caller's arrays may be gone by xv_collect)
    struct Redo {
      std::vector<int32_t> seg_len, first_seg;
      std::vector<int64_t> dst_row;
      bool utt = false, has_first = false, has_dst = false;
      int32_t n_utt = 0;
      float* out_dev = nullptr;
      float* host_out = nullptr;         // emb_host / out_host of the submission (may be null)
      const float* feats_ext = nullptr;  // features the caller keeps on the device (xv_submit_dev_utts), else the slot's copy
      bool feats_f16 = false;            // the slot's copy holds float16 values
    } redo;
    cudaEvent_t done = nullptr;          // recorded behind the submission's last copy; created with cudaEventBlockingSync so that
                                         // xv_collect SLEEPS instead of spinning (a multi-GPU job runs reader threads on those cores)
    bool busy = false;
  } slots[XV_HOST_SLOTS];
  int slot_next = 0;
  int32_t last_launches = 0;
  int32_t last_frontend_launches = 0;
  size_t fe_smem_opted = 0, fe_smem_opted_nv = 0;   // dynamic shared memory cmvn_select_kernel<NV> has been opted in for
  // options
  int num_clusters = 74;             // co-resident CTA pairs of tdnn_pair_kernel
  int opt_resident = 0;              // 1: keep a channel tile's weights resident in shared memory when they fit
                                     // (measured on B200: no faster than streaming 128-wide stages; kept as an option)
  int opt_prefetch = 0;              // L2 prefetch of the next work item's activation boxes
  long long* opt_trace = nullptr;    // diagnostics: per-tile clock stamps of ONE layer (opt_trace_layer)
  int opt_trace_layer = -1;
  int opt_profile = 0;               // 1: bracket every kernel launch with CUDA events (bench / diagnostics)
  std::vector<cudaEvent_t> prof_events;   // 2 per launch, in launch order
  int prof_used = 0;
};

namespace {

Plan make_plan(const xv_model* m, int64_t total_frames, int32_t n_seg) {
  Plan p;
  p.total_frames = total_frames;
  p.n_seg = n_seg;
  // upper bound of the packed rows: every segment starts at a multiple of 32 and is followed by >= gap zero rows
  const int64_t rows = total_frames + int64_t(n_seg) * (m->gap + tdnn2::POOL_BLOCK - 1);
  p.r_pad = round_up(std::max<int64_t>(rows, 1), tdnn2::TILE_ROWS);
  const int c_last = m->topo.width[m->topo.n_frame_layers - 1];
  const int K = 2 * m->c_pool;
  p.fc_m_tiles = (n_seg + xvk::FC_BM - 1) / xvk::FC_BM;
  p.fc_n_tiles = m->topo.emb_dim / xvk::FC_BN;
  p.fc_splits = K / 256;                                       // C_last is a multiple of 128
  p.fc_k_per_split = K / p.fc_splits;
  p.fc_counters = p.fc_m_tiles * p.fc_n_tiles;
  p.n_counters = p.fc_counters;
  {
    // tensor-core embedding GEMM: K' = 3K in 128-wide chunks, always cut into the same number of K-splits (the
    // largest divisor of the chunk count that leaves >= 2 chunks per split) whatever the batch: the summation order
    // of an embedding must not depend on how many other segments are in the call
    const int chunks = 3 * K / 128;
    int best = 1;
    for (int d = 1; d <= chunks / 2; ++d)
      if (chunks % d == 0 && d <= m->opt_fc_max_splits) best = d;
    p.tc_splits = best;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = size_t(round_up(int64_t(off + bytes), 1024)); return o; };
  p.off_meta = take((size_t(4) * n_seg + size_t(4) * size_t(p.r_pad / tdnn2::POOL_BLOCK)) * 4);   // 3 n_seg (+ pad to 16 B) + int4 per block
  p.off_counters = take(size_t(p.n_counters) * 4);
  p.off_valid = take(size_t(p.r_pad));
  p.off_blk_valid = take(size_t(p.r_pad / tdnn2::POOL_BLOCK));
  const size_t terms = m->opt_split ? 2 : 1;                 // split precision: every stored row is [hi | lo]
  p.off_x0 = take(size_t(p.r_pad) * m->k0_pad * 2 * terms);
  p.off_ha = take(size_t(p.r_pad) * m->w_mid * 2 * terms);
  p.off_hb = take(size_t(p.r_pad) * m->w_mid * 2 * terms);
  p.off_hlast = take(size_t(p.r_pad) * c_last * 2 * terms);  // only written when a caller asks for the last layer's activations
  p.off_pool_partial = take(size_t(p.r_pad / tdnn2::POOL_BLOCK) * 2 * c_last * 4);
  p.off_stats = take(size_t(n_seg) * K * 4);
  if (m->topo.pooling == XV_POOL_ATTENTION) {
    p.off_score = take(size_t(p.r_pad) * 2 * (m->c_pool / tdnn2::TILE_CH) * 4);
    p.off_attn = take(size_t(p.r_pad) * 4);
  }
  p.off_split = take(size_t(round_up(n_seg, tdnn2::CTA_ROWS)) * 3 * K * 2);   // rows padded to the TMA box (never read back)
  p.off_partial = take(std::max(size_t(p.fc_splits) * n_seg, size_t(p.tc_splits) * std::min<int64_t>(n_seg, FC_GROUP)) *
                       m->topo.emb_dim * 4);
  p.off_seg_emb = take(size_t(n_seg) * m->topo.emb_dim * 4);
  p.off_utt = take(size_t(round_up(n_seg + 1, 2)) * 4 + size_t(n_seg) * 8);
  p.bytes = off;
  return p;
}

int encode_2d(const xv_model* m, CUtensorMap* map, void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
              uint32_t box_outer, CUtensorMapSwizzle swz, uint64_t row_stride_elems = 0) {
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {(row_stride_elems ? row_stride_elems : inner) * 2};   // bytes, fp16
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = m->encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(XV_ECUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(int(r)));
  return XV_OK;
}

void free_layers(xv_model* m) {
  for (auto& L : m->layers) {
    cudaFree(L.w_dev); cudaFree(L.bias_dev); cudaFree(L.scale_dev); cudaFree(L.shift_dev); cudaFree(L.alpha_dev);
    L.w_dev = nullptr; L.bias_dev = L.scale_dev = L.shift_dev = L.alpha_dev = nullptr;
  }
  cudaFree(m->w0_dev); cudaFree(m->b0_dev); cudaFree(m->w0_split_dev);
  cudaFree(m->att_w_dev); cudaFree(m->att_b_dev); cudaFree(m->att_v_dev);
  m->att_w_dev = nullptr; m->att_b_dev = m->att_v_dev = nullptr;
  m->w0_dev = m->b0_dev = nullptr;
  m->w0_split_dev = nullptr;
}

const std::vector<float>* find_param(const xv_model* m, const std::string& name, std::initializer_list<int64_t> shape) {
  auto it = m->host_params.find(name);
  if (it == m->host_params.end()) return nullptr;
  const auto& s = m->host_shapes.at(name);
  if (s.size() != shape.size() || !std::equal(s.begin(), s.end(), shape.begin())) return nullptr;
  return &it->second;
}

// fp16 range rescue.  Every stored activation tensor is fp16; a trained model whose activations pass 65 504 (the reference
// computes in fp32 and loads any model, models.py:476-480) is handled by storing layer i's rows divided by 2^exp_out[i]:
// y' = (act(a) * scale + shift) * 2^-e is folded into scale / shift, the consumer multiplies its fp32 accumulator by 2^e
// before the bias -- exact operations, so results equal the unscaled arithmetic wherever that one does not overflow
// (only fp16 subnormals, |y| < 6e-5 * 2^e, round differently).  The pooled statistics get the same treatment on their
// way into the split-fp16 embedding GEMM (exp_stats).  Exponents start at 0 and are raised when a store overflows
// (xv_rescue_overflow; xv_collect does it by itself and re-runs the batch).
int apply_exponents(xv_model* m, cudaStream_t stream_or_null) {
  for (auto& L : m->layers) {
    std::vector<float> sc(L.scale_host), sh(L.shift_host);
    const float f = std::ldexp(1.0f, -L.exp_out);
    for (auto& v : sc) v *= f;
    for (auto& v : sh) v *= f;
    if (stream_or_null) XV_CUDA(cudaStreamSynchronize(stream_or_null));
    XV_CUDA(cudaMemcpy(L.scale_dev, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice));
    XV_CUDA(cudaMemcpy(L.shift_dev, sh.data(), sh.size() * 4, cudaMemcpyHostToDevice));
  }
  return XV_OK;
}

// Overflow flag word: bit 0 = an fp16 store of the training path overflowed, bit 1 = front-end row-count mismatch,
// bit 8 + i = a store of frame layer i overflowed, bit 16 = a pooled statistic left the fp16 range on its way into the
// embedding GEMM.  Raises the exponents of what overflowed; returns how many were raised.
constexpr int RESCUE_STEP = 6;        // x 1/64 per attempt
constexpr int RESCUE_MAX_EXP = 96;    // fp32 accumulators hold up to 3e38
int raise_exponents(xv_model* m, uint32_t flag) {
  int raised = 0;
  for (int i = 0; i < int(m->layers.size()); ++i)
    if ((flag >> (8 + i)) & 1u) {
      if (m->layers[i].exp_out + RESCUE_STEP > RESCUE_MAX_EXP) return -1;
      m->layers[i].exp_out += RESCUE_STEP;
      ++raised;
    }
  if ((flag >> 16) & 1u) {
    if (m->exp_stats + RESCUE_STEP > RESCUE_MAX_EXP) return -1;
    m->exp_stats += RESCUE_STEP;
    ++raised;
  }
  return raised;
}

// Pack host parameters into device layouts.  BatchNorm (evaluation branch, tf_block.py:25-26)
// is folded to scale = gamma * rsqrt(var + eps), shift = beta - mean * scale, in fp32.
int finalize_params(xv_model* m) {
  if (!m->dirty) return XV_OK;
  XV_CUDA(cudaSetDevice(m->device));
  free_layers(m);
  const xv_topology& t = m->topo;
  for (int i = 0; i < t.n_frame_layers; ++i) {
    FrameLayer& L = m->layers[i];
    const std::string s = "frame_level_info_layer-" + std::to_string(i) + "/";
    const auto* w = find_param(m, s + "w:0", {L.taps, L.c_in, L.c_out});
    const auto* b = find_param(m, s + "b:0", {L.c_out});
    const auto* gamma = find_param(m, s + "gamma:0", {L.c_out});
    const auto* beta = find_param(m, s + "beta:0", {L.c_out});
    const auto* mean = find_param(m, s + "mean:0", {L.c_out});
    const auto* var = find_param(m, s + "variance:0", {L.c_out});
    if (!w || !b || !gamma || !beta || !mean || !var)
      return fail(XV_ESTATE, "missing or mis-shaped parameter(s) under scope '" + s + "' (need w,b,gamma,beta,mean,variance)");
    // weights: TF [k, Cin, Cout] -> [Cout, K] fp16 (round to nearest), K index = tap * c_in_pad + c
    // (first layer: c_in_pad == c_in, i.e. densely spliced, zero padded up to k_total)
    // split precision: per tap (first layer: per spliced row) the K run is [w_hi | w_lo | w_hi], against the activations'
    // virtual [x_hi | x_hi | x_lo]
    const int split = m->opt_split ? 1 : 0;
    const int run = (i == 0) ? L.k_total : L.c_in_pad;             // K columns of one tap's run (one term)
    const int kw = (split ? 3 : 1) * L.k_total;                    // packed row length
    std::vector<__half> wt(size_t(L.c_out) * kw, __float2half_rn(0.f));
    for (int j = 0; j < L.taps; ++j)
      for (int c = 0; c < L.c_in; ++c) {
        const float* src = w->data() + (size_t(j) * L.c_in + c) * L.c_out;
        const size_t kidx = size_t(j) * L.c_in_pad + c;            // index within the unsplit layout
        const size_t tap = kidx / run, within = kidx % run;
        for (int o = 0; o < L.c_out; ++o) {
          const __half hi = __float2half_rn(src[o]);
          if (!split) { wt[size_t(o) * kw + kidx] = hi; continue; }
          const __half lo = __float2half_rn(src[o] - __half2float(hi));
          __half* row = wt.data() + size_t(o) * kw + tap * 3 * run;
          row[within] = hi; row[run + within] = lo; row[2 * run + within] = hi;
        }
      }
    std::vector<float> scale(L.c_out), shift(L.c_out);
    for (int o = 0; o < L.c_out; ++o) {
      const float inv = (1.0f / sqrtf((*var)[o] + t.bn_eps)) * (*gamma)[o];
      scale[o] = inv;
      shift[o] = (*beta)[o] - (*mean)[o] * inv;
    }
    XV_CUDA(cudaMalloc(&L.w_dev, wt.size() * sizeof(__half)));
    XV_CUDA(cudaMalloc(&L.bias_dev, L.c_out * 4));
    XV_CUDA(cudaMalloc(&L.scale_dev, L.c_out * 4));
    XV_CUDA(cudaMalloc(&L.shift_dev, L.c_out * 4));
    XV_CUDA(cudaMemcpy(L.w_dev, wt.data(), wt.size() * sizeof(__half), cudaMemcpyHostToDevice));
    XV_CUDA(cudaMemcpy(L.bias_dev, b->data(), L.c_out * 4, cudaMemcpyHostToDevice));
    L.scale_host = scale;
    L.shift_host = shift;
    L.bias_host = *b;
    L.alpha_host.clear();
    L.exp_out = 0;
    if (t.act != XV_ACT_RELU) {
      std::vector<float> alpha(L.c_out, 0.2f);                       // tf.nn.leaky_relu(h, alpha=0.2)  (models.py:912)
      if (t.act == XV_ACT_PRELU) {                                   // prelu(h, shared=False)  (tf_block.py:38-47)
        const auto* al = find_param(m, s + "prelu/prelu:0", {L.c_out});
        if (!al) return fail(XV_ESTATE, "missing or mis-shaped parameter '" + s + "prelu/prelu:0'");
        alpha = *al;
      }
      XV_CUDA(cudaMalloc(&L.alpha_dev, L.c_out * 4));
      XV_CUDA(cudaMemcpy(L.alpha_dev, alpha.data(), L.c_out * 4, cudaMemcpyHostToDevice));
      L.alpha_host = alpha;
    }
  }
  const int c_last = m->c_pool;
  if (t.pooling == XV_POOL_ATTENTION) {
    const int C = m->c_pool;
    const auto* aw = find_param(m, "attention/w:0", {C, C});
    const auto* ab = find_param(m, "attention/b:0", {C});
    const auto* av = find_param(m, "attention/v:0", {C});
    if (!aw || !ab || !av) return fail(XV_ESTATE, "missing or mis-shaped attention/w:0, attention/b:0 or attention/v:0");
    const int kw = m->opt_split ? 3 * C : C;
    std::vector<__half> wt(size_t(C) * kw);                      // einsum('ijk,kl->ijl'): [k, l] -> [l][k], K-major
    for (int k = 0; k < C; ++k)
      for (int l = 0; l < C; ++l) {
        const float x = (*aw)[size_t(k) * C + l];
        const __half hi = __float2half_rn(x);
        __half* row = wt.data() + size_t(l) * kw;
        row[k] = hi;
        if (m->opt_split) { row[C + k] = __float2half_rn(x - __half2float(hi)); row[2 * C + k] = hi; }
      }
    XV_CUDA(cudaMalloc(&m->att_w_dev, wt.size() * sizeof(__half)));
    XV_CUDA(cudaMalloc(&m->att_b_dev, C * 4));
    XV_CUDA(cudaMalloc(&m->att_v_dev, C * 4));
    XV_CUDA(cudaMemcpy(m->att_w_dev, wt.data(), wt.size() * sizeof(__half), cudaMemcpyHostToDevice));
    XV_CUDA(cudaMemcpy(m->att_b_dev, ab->data(), C * 4, cudaMemcpyHostToDevice));
    XV_CUDA(cudaMemcpy(m->att_v_dev, av->data(), C * 4, cudaMemcpyHostToDevice));
  }
  const auto* w0 = find_param(m, "embed_layer-0/w:0", {2 * c_last, t.emb_dim});
  const auto* b0 = find_param(m, "embed_layer-0/b:0", {t.emb_dim});
  if (!w0 || !b0) return fail(XV_ESTATE, "missing or mis-shaped embed_layer-0/w:0 or embed_layer-0/b:0");
  XV_CUDA(cudaMalloc(&m->w0_dev, w0->size() * 4));
  XV_CUDA(cudaMalloc(&m->b0_dev, b0->size() * 4));
  XV_CUDA(cudaMemcpy(m->w0_dev, w0->data(), w0->size() * 4, cudaMemcpyHostToDevice));
  XV_CUDA(cudaMemcpy(m->b0_dev, b0->data(), b0->size() * 4, cudaMemcpyHostToDevice));
  {
    // split-precision copy for the tensor-core embedding GEMM: w = hi + lo (fp16 each, |w - hi - lo| ~ 2^-22 |w|);
    // emb = s_hi*w_hi + s_hi*w_lo + s_lo*w_hi with K' = 3 * 2C, operands [s_hi | s_hi | s_lo] x [w_hi | w_lo | w_hi]
    const int K = 2 * c_last, E = t.emb_dim;
    std::vector<__half> ws(size_t(E) * 3 * K);
    for (int k = 0; k < K; ++k)
      for (int o = 0; o < E; ++o) {
        const float w = (*w0)[size_t(k) * E + o];
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        __half* row = ws.data() + size_t(o) * 3 * K;
        row[k] = hi; row[K + k] = lo; row[2 * K + k] = hi;
      }
    XV_CUDA(cudaMalloc(&m->w0_split_dev, ws.size() * sizeof(__half)));
    XV_CUDA(cudaMemcpy(m->w0_split_dev, ws.data(), ws.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  m->exp_stats = 0;
  m->dirty = false;
  return apply_exponents(m, nullptr);
}

// The model's sticky device flag: bit 0 = an fp16 store overflowed, bit 1 = the feature front end found fewer voiced rows
// in an utterance than the caller's out_keep said (frontend.cuh).
int sticky_flag_error(uint32_t flag) {
  if (flag & 2u)
    return fail(XV_EINVAL, "feature front end: an utterance has fewer voiced rows than out_keep says (VAD track and row "
                           "counts disagree); results are not trustworthy");
  return fail(XV_EOVERFLOW, "an activation exceeded the fp16 range (|x| > 65504); results are not trustworthy");
}

int ensure_meta_capacity(xv_model* m, int64_t n_ints) {
  if (n_ints <= m->meta_cap) return XV_OK;
  const int64_t cap = std::max<int64_t>(n_ints * 2, 4096);
  for (int s = 0; s < META_SLOTS; ++s) {
    if (m->meta_event[s]) XV_CUDA(cudaEventSynchronize(m->meta_event[s]));
    if (m->meta_host[s]) XV_CUDA(cudaFreeHost(m->meta_host[s]));
    m->meta_host[s] = nullptr;
    XV_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&m->meta_host[s]), size_t(cap) * 4, cudaHostAllocDefault));
    if (!m->meta_event[s]) XV_CUDA(cudaEventCreateWithFlags(&m->meta_event[s], cudaEventDisableTiming));
  }
  m->meta_cap = cap;
  return XV_OK;
}

int ensure_utt_capacity(xv_model* m, int64_t n_ints) {
  if (n_ints <= m->utt_cap) return XV_OK;
  const int64_t cap = std::max<int64_t>(n_ints * 2, 4096);
  for (int s = 0; s < META_SLOTS; ++s) {
    if (m->utt_event[s]) XV_CUDA(cudaEventSynchronize(m->utt_event[s]));
    if (m->utt_host[s]) XV_CUDA(cudaFreeHost(m->utt_host[s]));
    m->utt_host[s] = nullptr;
    XV_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&m->utt_host[s]), size_t(cap) * 4, cudaHostAllocDefault));
    if (!m->utt_event[s]) XV_CUDA(cudaEventCreateWithFlags(&m->utt_event[s], cudaEventDisableTiming));
  }
  m->utt_cap = cap;
  return XV_OK;
}

// Profiling (opt_profile): an event pair around every launch, on the launching stream.
int prof_mark(xv_model* m, cudaStream_t stream) {
  if (!m->opt_profile) return XV_OK;
  if (m->prof_used == int(m->prof_events.size())) {
    cudaEvent_t e;
    XV_CUDA(cudaEventCreate(&e));
    m->prof_events.push_back(e);
  }
  XV_CUDA(cudaEventRecord(m->prof_events[m->prof_used++], stream));
  return XV_OK;
}


// Every kernel is launched with the programmatic-stream-serialization attribute (PDL): the kernels call
// cudaGridDependencySynchronize() before touching global memory, so a kernel's prologue overlaps its
// predecessor's tail.  opt_pdl = 0 falls back to plain stream order.
template <typename... KArgs, typename... Args>
cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// Segment metadata of one call, built on the host in a pinned staging slot and copied to meta_dev:
// [row_start(n_seg) | feat_start(n_seg) | len(n_seg) | pad to int4 | blk_info(r_pad/32 x int4)].
struct StagedMeta {
  int64_t r_pad = 0;
  xvk::SegMeta seg{};
  const int4* blk_info_dev = nullptr;
};
int stage_meta(xv_model* m, const int32_t* seg_len_host, int32_t n_seg, int64_t plan_r_pad, int32_t* meta_dev,
               cudaStream_t stream, StagedMeta* out) {
  int rc = ensure_meta_capacity(m, int64_t(4) * n_seg + 4 * (plan_r_pad / tdnn2::POOL_BLOCK) + 4);
  if (rc != XV_OK) return rc;
  const int slot = m->meta_next;
  m->meta_next = (m->meta_next + 1) % META_SLOTS;
  XV_CUDA(cudaEventSynchronize(m->meta_event[slot]));      // previous copy out of this slot finished
  int32_t* mh = m->meta_host[slot];
  int64_t rows_used = 0;
  const int64_t blk_info_off = round_up(int64_t(3) * n_seg, 4);   // int4-aligned
  {
    int64_t row = 0, fs = 0;
    // per aligned 32-row block: {first feature row, first frame of the block in its segment, segment length, valid rows}
    int32_t* blk_info = mh + blk_info_off;
    for (int i = 0; i < n_seg; ++i) {
      const int32_t len = seg_len_host[i];
      mh[i] = int32_t(row);
      mh[n_seg + i] = int32_t(fs);
      mh[2 * n_seg + i] = len;
      const int64_t next = row + round_up(int64_t(len) + m->gap, tdnn2::POOL_BLOCK);
      int32_t t0 = 0;
      for (int64_t b = row / tdnn2::POOL_BLOCK; b < next / tdnn2::POOL_BLOCK; ++b, t0 += tdnn2::POOL_BLOCK) {
        int32_t* e = blk_info + 4 * b;
        const int32_t nv = std::max(0, std::min(tdnn2::POOL_BLOCK, len - t0));
        e[0] = int32_t(fs + t0); e[1] = t0; e[2] = len; e[3] = nv;
      }
      row = next;
      fs += len;
    }
    rows_used = row;
  }
  const int64_t r_pad = round_up(rows_used, tdnn2::TILE_ROWS);
  const int64_t n_blocks = r_pad / tdnn2::POOL_BLOCK;
  for (int64_t b = rows_used / tdnn2::POOL_BLOCK; b < n_blocks; ++b) {
    int32_t* e = mh + blk_info_off + 4 * b;
    e[0] = e[1] = e[2] = e[3] = 0;
  }
  XV_CUDA(cudaMemcpyAsync(meta_dev, mh, (size_t(blk_info_off) + size_t(4) * n_blocks) * 4, cudaMemcpyHostToDevice, stream));
  XV_CUDA(cudaEventRecord(m->meta_event[slot], stream));
  out->r_pad = r_pad;
  out->seg = xvk::SegMeta{meta_dev, meta_dev + n_seg, meta_dev + 2 * n_seg, n_seg};
  out->blk_info_dev = reinterpret_cast<const int4*>(meta_dev + blk_info_off);
  return XV_OK;
}

// Utterance-level output of a forward (xv_forward_utts): the frame-weighted chunk average of make_embedding
// (models.py:398-421), written to scattered rows of `out_dev` (possibly peer memory) and / or contiguously to `out_local_dev`.
struct UttOut {
  const int32_t* first_seg_host = nullptr;   // [n_utt + 1], null: every segment is its own utterance
  const int64_t* dst_row_host = nullptr;     // [n_utt], null: row u
  int32_t n_utt = 0;
  float* out_dev = nullptr;
  float* out_local_dev = nullptr;
};

int forward_impl(xv_model* m, const float* feats_dev, const int32_t* seg_len_host, int32_t n_seg, float* emb_dev,
                 void* workspace_dev, size_t workspace_bytes, cudaStream_t stream, float* const* layer_out_dev,
                 float* stats_out_dev, const UttOut* utt = nullptr) {
  if (!m || !feats_dev || !seg_len_host || (!emb_dev && !utt) || !workspace_dev) return fail(XV_EINVAL, "null argument");
  if (n_seg <= 0) return fail(XV_EINVAL, "n_seg must be >= 1");
  if (utt) {
    if (!utt->out_dev && !utt->out_local_dev) return fail(XV_EINVAL, "utterance output: no destination");
    const int32_t nu = utt->first_seg_host ? utt->n_utt : n_seg;
    if (nu <= 0 || nu > n_seg) return fail(XV_EINVAL, "n_utt must be in [1, n_seg]");
    if (utt->first_seg_host) {
      if (utt->first_seg_host[0] != 0 || utt->first_seg_host[nu] != n_seg)
        return fail(XV_EINVAL, "utt_first_seg must start at 0 and end at n_seg");
      for (int u = 0; u < nu; ++u)
        if (utt->first_seg_host[u + 1] <= utt->first_seg_host[u]) return fail(XV_EINVAL, "utt_first_seg must be strictly increasing");
    }
    if (utt->dst_row_host)
      for (int u = 0; u < nu; ++u)
        if (utt->dst_row_host[u] < 0) return fail(XV_EINVAL, "negative destination row");
  }
  XV_CUDA(cudaSetDevice(m->device));
  int rc = finalize_params(m);
  if (rc != XV_OK) return rc;
  int64_t total = 0;
  for (int i = 0; i < n_seg; ++i) {
    if (seg_len_host[i] <= 0) return fail(XV_EINVAL, "segment " + std::to_string(i) + " has non-positive length");
    total += seg_len_host[i];
  }
  if (total + int64_t(n_seg) * (m->gap + 32) > (int64_t(1) << 31) - 4096)
    return fail(XV_EINVAL, "batch too large: packed rows exceed int32 range");
  const Plan p = make_plan(m, total, n_seg);
  if (workspace_bytes < p.bytes)
    return fail(XV_ENOMEM, "workspace too small: need " + std::to_string(p.bytes) + " bytes, got " + std::to_string(workspace_bytes));
  if (reinterpret_cast<uintptr_t>(workspace_dev) % 1024 != 0) return fail(XV_EINVAL, "workspace must be 1024-byte aligned");
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);

  // ---- segment metadata: packed row starts (multiples of 32), feature row starts, lengths ----
  StagedMeta sm;
  rc = stage_meta(m, seg_len_host, n_seg, p.r_pad, reinterpret_cast<int32_t*>(ws + p.off_meta), stream, &sm);
  if (rc != XV_OK) return rc;
  const int64_t r_pad = sm.r_pad;                                   // <= p.r_pad (the plan's upper bound)
  const xvk::SegMeta seg = sm.seg;
  if (utt && !emb_dev) emb_dev = reinterpret_cast<float*>(ws + p.off_seg_emb);
  const int32_t* utt_first_dev = nullptr;
  const int64_t* utt_dst_dev = nullptr;
  const int32_t n_utt = utt ? (utt->first_seg_host ? utt->n_utt : n_seg) : 0;
  if (utt) {
    // the utterance plan rides in its own pinned slot (same ring discipline as the segment metadata)
    const int64_t first_ints = round_up(int64_t(n_seg) + 1, 2);
    rc = ensure_utt_capacity(m, first_ints + 2 * int64_t(n_seg));
    if (rc != XV_OK) return rc;
    const int slot = m->utt_next;
    m->utt_next = (m->utt_next + 1) % META_SLOTS;
    XV_CUDA(cudaEventSynchronize(m->utt_event[slot]));
    int32_t* uh = m->utt_host[slot];
    if (utt->first_seg_host) std::memcpy(uh, utt->first_seg_host, size_t(n_utt + 1) * 4);
    else for (int i = 0; i <= n_seg; ++i) uh[i] = i;
    int64_t* dh = reinterpret_cast<int64_t*>(uh + first_ints);
    if (utt->dst_row_host) std::memcpy(dh, utt->dst_row_host, size_t(n_utt) * 8);
    uint8_t* ud = ws + p.off_utt;
    XV_CUDA(cudaMemcpyAsync(ud, uh, size_t(first_ints) * 4 + (utt->dst_row_host ? size_t(n_utt) * 8 : 0), cudaMemcpyHostToDevice, stream));
    XV_CUDA(cudaEventRecord(m->utt_event[slot], stream));
    utt_first_dev = reinterpret_cast<const int32_t*>(ud);
    if (utt->dst_row_host) utt_dst_dev = reinterpret_cast<const int64_t*>(ud + size_t(first_ints) * 4);
  }
  const int4* blk_info_dev = sm.blk_info_dev;

  uint8_t* row_valid = ws + p.off_valid;
  uint8_t* blk_valid = ws + p.off_blk_valid;
  uint32_t* counters = reinterpret_cast<uint32_t*>(ws + p.off_counters);
  __half* x0 = reinterpret_cast<__half*>(ws + p.off_x0);
  __half* ha = reinterpret_cast<__half*>(ws + p.off_ha);
  __half* hb = reinterpret_cast<__half*>(ws + p.off_hb);
  __half* hlast = reinterpret_cast<__half*>(ws + p.off_hlast);
  float* pool_partial = reinterpret_cast<float*>(ws + p.off_pool_partial);
  int launches = 0;
  const bool pdl = m->opt_pdl != 0 && !m->opt_profile;
  m->prof_used = 0;
#define XV_PROF() do { int prc_ = prof_mark(m, stream); if (prc_ != XV_OK) return prc_; } while (0)

  // the input splice inside the first layer's kernel (tdnn_first.cuh) unless the topology or an option asks for the two launches
  const bool fuse_first = m->opt_fuse_first && m->first_fusable && !m->opt_split && !m->opt_resident &&
                          reinterpret_cast<uintptr_t>(feats_dev) % 16 == 0;
  // ---- pack: fp32 features -> spliced fp16 packed rows + row / block maps --------------------
  if (!fuse_first) {
    xvk::PackArgs a{};
    a.feats = feats_dev;
    a.r_pad = int32_t(r_pad);
    a.feat_dim = m->topo.feat_dim;
    a.taps = m->layers[0].taps;
    a.dilation = m->layers[0].dilation;
    a.k0_pad = m->k0_pad;
    a.x0 = x0;
    a.row_valid = row_valid;
    a.blk_valid = blk_valid;
    a.blk_info = blk_info_dev;
    a.lut = m->pack_lut_dev;
    a.counters = counters;
    a.n_counters = p.n_counters;
    a.split = m->opt_split ? 1 : 0;
    a.feats_f16 = m->cur_feats_f16;
    const int blocks = int(r_pad / xvk::PACK_ROWS_PER_BLOCK);      // >= n_seg: covers n_counters with 256 threads each
    XV_PROF();
    XV_CUDA(launch_k(pdl, xvk::pack_im2col_kernel, dim3(blocks), dim3(xvk::PACK_THREADS), 0, stream, a));
    XV_PROF();
    XV_CUDA(cudaGetLastError());
    ++launches;
  }

  // ---- frame-level TDNN stack: one fused tcgen05 kernel per layer --------------------------
  const int nl = m->topo.n_frame_layers;
  const bool attention = m->topo.pooling == XV_POOL_ATTENTION;
  const bool want_last = attention || (layer_out_dev && layer_out_dev[nl - 1]);   // attention pooling reads the stored activation
  const __half* in = x0;
  // the last two frame layers as one kernel when both are context-free, the intermediate is 512 wide and nobody asks for
  // their stored activations (statistics pooling, plain fp16 operands)
  const bool fuse_tail = m->opt_fuse_tail && !attention && !m->opt_split && !layer_out_dev && nl >= 3 && !m->opt_resident &&
                         m->layers[nl - 2].gemm_taps == 1 && m->layers[nl - 1].gemm_taps == 1 &&
                         m->layers[nl - 2].c_out == tdnn2::FT_MID_CH && m->layers[nl - 2].c_in_pad % tdnn2::BLOCK_K == 0 &&
                         m->layers[nl - 2].c_in_pad <= tdnn2::FT_MID_CH && m->layers[nl - 1].c_out >= 2 * tdnn2::TILE_CH &&
                         (m->opt_trace_layer < 0 || m->opt_trace_layer == nl - 2);
  for (int i = 0; i < nl; ++i) {
    if (fuse_tail && i == nl - 2) {
      const FrameLayer& L3 = m->layers[nl - 2];
      const FrameLayer& L4 = m->layers[nl - 1];
      CUtensorMap tx, tw3, tw4;
      rc = encode_2d(m, &tx, const_cast<__half*>(in), uint64_t(L3.c_in_pad), uint64_t(r_pad), tdnn2::BLOCK_K, tdnn2::ACT_BOX_ROWS_PLAIN,
                     CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != XV_OK) return rc;
      rc = encode_2d(m, &tw3, L3.w_dev, uint64_t(L3.k_total), uint64_t(L3.c_out), tdnn2::BLOCK_K, tdnn2::CTA_CH, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != XV_OK) return rc;
      rc = encode_2d(m, &tw4, L4.w_dev, uint64_t(L4.k_total), uint64_t(L4.c_out), tdnn2::BLOCK_K, tdnn2::CTA_CH, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != XV_OK) return rc;
      tdnn2::FusedTailArgs a{};
      a.n_row_tiles = int32_t(r_pad / tdnn2::TILE_ROWS);
      a.k_atoms_in = L3.c_in_pad / tdnn2::BLOCK_K;
      a.n_ch_tiles = L4.c_out / tdnn2::TILE_CH;
      a.c_out = L4.c_out;
      a.bias3 = L3.bias_dev; a.scale3 = L3.scale_dev; a.shift3 = L3.shift_dev; a.alpha3 = L3.alpha_dev;
      a.acc_scale3 = std::ldexp(1.0f, m->layers[nl - 3].exp_out);
      a.bias4 = L4.bias_dev; a.scale4 = L4.scale_dev; a.shift4 = L4.shift_dev; a.alpha4 = L4.alpha_dev;
      a.acc_scale4 = std::ldexp(1.0f, L3.exp_out);
      a.row_valid = row_valid;
      a.blk_valid = blk_valid;
      a.partial = pool_partial;
      a.overflow_flag = m->cur_flag;
      a.overflow_bit3 = 1u << (8 + nl - 2);
      a.trace = (m->opt_trace_layer == nl - 2) ? m->opt_trace : nullptr;
      const int grid = 2 * int(std::min<int64_t>(a.n_row_tiles, m->num_clusters));
      XV_PROF();
      if (L3.alpha_dev != nullptr)
        XV_CUDA(launch_k(pdl, tdnn2::tdnn_tail_fused_kernel<true>, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::FT_SMEM_BYTES, stream, tx, tw3, tw4, a));
      else
        XV_CUDA(launch_k(pdl, tdnn2::tdnn_tail_fused_kernel<false>, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::FT_SMEM_BYTES, stream, tx, tw3, tw4, a));
      XV_PROF();
      XV_CUDA(cudaGetLastError());
      ++launches;
      break;
    }
    const FrameLayer& L = m->layers[i];
    const bool last = i == nl - 1;
    __half* out = last ? hlast : ((i & 1) ? hb : ha);
    if (fuse_first && i == 0) {
      CUtensorMap tw, tc;
      rc = encode_2d(m, &tw, L.w_dev, uint64_t(L.k_total), uint64_t(L.c_out), tdnn2::BLOCK_K, tdnn2::CTA_CH, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != XV_OK) return rc;
      rc = encode_2d(m, &tc, out, uint64_t(L.c_out), uint64_t(r_pad), tdnn2::C_CHUNK, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != XV_OK) return rc;
      tdnn2::FirstArgs a{};
      a.n_row_tiles = int32_t(r_pad / tdnn2::TILE_ROWS);
      a.n_ch_tiles = L.c_out / tdnn2::TILE_CH;
      a.c_out = L.c_out;
      a.feat_dim = m->topo.feat_dim;
      a.halo = (L.taps - 1) / 2 * L.dilation;
      a.feats = feats_dev;
      a.feats_f16 = m->cur_feats_f16;
      a.n_feat_bytes = total * m->topo.feat_dim * (m->cur_feats_f16 ? 2 : 4);
      a.taps = L.taps;
      a.blk_info = blk_info_dev;
      a.bias = L.bias_dev; a.scale = L.scale_dev; a.shift = L.shift_dev; a.alpha = L.alpha_dev;
      a.row_valid = row_valid;
      a.blk_valid = blk_valid;
      a.counters = counters;
      a.n_counters = p.n_counters;
      a.overflow_flag = m->cur_flag;
      a.overflow_bit = 1u << 8;
      a.trace = (m->opt_trace_layer == 0) ? m->opt_trace : nullptr;
      const int grid = 2 * int(std::min<int64_t>(a.n_row_tiles, m->num_clusters));
      XV_PROF();
      if (L.alpha_dev != nullptr)
        XV_CUDA(launch_k(pdl, tdnn2::tdnn_first_kernel<true>, dim3(grid), dim3(tdnn2::F1_THREADS), tdnn2::F1_SMEM_BYTES, stream, tw, tc, a));
      else
        XV_CUDA(launch_k(pdl, tdnn2::tdnn_first_kernel<false>, dim3(grid), dim3(tdnn2::F1_THREADS), tdnn2::F1_SMEM_BYTES, stream, tw, tc, a));
      XV_PROF();
      XV_CUDA(cudaGetLastError());
      ++launches;
      if (layer_out_dev && layer_out_dev[i]) {
        XV_CUDA(launch_k(pdl, xvk::unpack_rows_kernel, dim3(n_seg), dim3(256), 0, stream, static_cast<const __half*>(out), seg, int32_t(L.c_out), layer_out_dev[i],
                         std::ldexp(1.0f, L.exp_out), int32_t(0)));
        XV_CUDA(cudaGetLastError());
        ++launches;
      }
      in = out;
      continue;
    }
    const int halo = (L.gemm_taps - 1) / 2 * L.dilation;
    const int c_in_real = (i == 0) ? L.k_total : L.c_in_pad;       // values per input row
    const int split = m->opt_split ? 1 : 0;
    // split precision: rows hold [hi | lo] (2 x c_in_real halfs), the K loop walks the virtual [hi | hi | lo]
    const int c_in_gemm = split ? 3 * c_in_real : c_in_real;
    {
      const bool reuse = L.gemm_taps > 1 && halo <= tdnn2::MAX_REUSE_HALO;
      CUtensorMap ta, tw, tc;
      rc = encode_2d(m, &ta, const_cast<__half*>(in), uint64_t((split ? 2 : 1) * c_in_real), uint64_t(r_pad), tdnn2::BLOCK_K,
                     reuse ? tdnn2::ACT_BOX_ROWS_REUSE : tdnn2::ACT_BOX_ROWS_PLAIN, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != XV_OK) return rc;
      rc = encode_2d(m, &tw, L.w_dev, uint64_t((split ? 3 : 1) * L.k_total), uint64_t(L.c_out), tdnn2::BLOCK_K, tdnn2::CTA_CH,
                     CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != XV_OK) return rc;
      rc = encode_2d(m, &tc, out, uint64_t((split ? 2 : 1) * L.c_out), uint64_t(r_pad), tdnn2::C_CHUNK, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc != XV_OK) return rc;
      // Weights of one channel tile fit in shared memory next to an activation ring when K_total <= 512
      // (every k=1 layer and the spliced first layer): keep them resident, stream activations in 64-wide
      // stages.  Otherwise stream both operands in 128-wide stages.
      const int64_t ring_cap = tdnn2::RING_BYTES;
      const int n_ch_tiles = L.c_out / tdnn2::TILE_CH;
      const int k_atoms = L.gemm_taps * (c_in_gemm / tdnn2::BLOCK_K);
      bool resident = m->opt_resident && !split && L.gemm_taps == 1 && k_atoms % 2 == 0 && k_atoms <= tdnn2::MAX_STAGES &&
                      m->num_clusters >= n_ch_tiles;
      tdnn2::PairArgs a{};
      a.n_row_tiles = int32_t(r_pad / tdnn2::TILE_ROWS);
      a.n_ch_tiles = n_ch_tiles;
      a.taps = L.gemm_taps;
      a.dilation = L.dilation;
      a.c_in_pad = c_in_gemm;
      a.reuse = reuse ? 1 : 0;
      a.c_out = L.c_out;
      a.bias = L.bias_dev;
      a.scale = L.scale_dev;
      a.shift = L.shift_dev;
      a.alpha = L.alpha_dev;
      a.row_valid = row_valid;
      a.blk_valid = blk_valid;
      a.partial = pool_partial;
      a.overflow_flag = m->cur_flag;
      a.overflow_bit = 1u << (8 + i);
      a.acc_scale = i > 0 ? std::ldexp(1.0f, m->layers[i - 1].exp_out) : 1.0f;      // the input rows are stored / 2^exp
      a.split_c = split ? c_in_real : 0;
      a.split_lo_off = c_in_real;
      a.trace = (i == m->opt_trace_layer) ? m->opt_trace : nullptr;
      a.wgt_resident = resident ? 1 : 0;
      a.prefetch = m->opt_prefetch;
      const int64_t tiles = int64_t(a.n_row_tiles) * a.n_ch_tiles;
      const int n_cl = resident ? (m->num_clusters / n_ch_tiles) * n_ch_tiles : int(std::min<int64_t>(tiles, m->num_clusters));
      const int grid = 2 * n_cl;
      // last layer: pooled partial sums (mode 1); its activation is only stored when a caller asks for it
      for (int mode = (last && !attention) ? 1 : 0; mode >= 0; --mode) {
        if (last && mode == 0 && !want_last) break;
        a.mode = mode;
        const int64_t cap = ring_cap + (mode == 1 ? 16384 : 0);          // the pooled mode has no output staging
        const int64_t act_atom = reuse ? tdnn2::ACT_ATOM_BYTES : tdnn2::ACT_BOX_ROWS_PLAIN * 128;
        // resident weights: 64-wide stages (opt_resident == 1) or 128-wide stages when two of them still fit (== 2)
        int atoms = 2;
        if (resident) {
          atoms = (m->opt_resident == 2 && cap - int64_t(k_atoms) * tdnn2::WGT_ATOM_BYTES >= 4 * act_atom) ? 2 : 1;
          a.n_wgt_stages = k_atoms / atoms;
          a.n_act_stages = int(std::min<int64_t>(tdnn2::MAX_STAGES, (cap - int64_t(k_atoms) * tdnn2::WGT_ATOM_BYTES) / (atoms * act_atom)));
        } else if (reuse) {                                              // one activation slab feeds `taps` weight stages
          a.n_act_stages = 2;
          a.n_wgt_stages = int(std::min<int64_t>(tdnn2::MAX_STAGES, (cap - 2 * 2 * act_atom) / (2 * tdnn2::WGT_ATOM_BYTES)));
        } else {                                                         // k = 1: activations and weights advance together
          a.n_act_stages = a.n_wgt_stages = int(std::min<int64_t>(tdnn2::MAX_STAGES, cap / (2 * (act_atom + tdnn2::WGT_ATOM_BYTES))));
        }
        a.c_chunks = c_in_gemm / (atoms * tdnn2::BLOCK_K);
        XV_PROF();
        {
          void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, tdnn2::PairArgs) = nullptr;
          const bool leaky = L.alpha_dev != nullptr;
#define XV_PICK(MO, AT) (leaky ? tdnn2::tdnn_pair_kernel<MO, AT, true> : tdnn2::tdnn_pair_kernel<MO, AT, false>)
          if (mode == 1) kern = atoms == 1 ? XV_PICK(1, 1) : XV_PICK(1, 2);
          else if (split) kern = leaky ? tdnn2::tdnn_pair_kernel<0, 2, true, false, true> : tdnn2::tdnn_pair_kernel<0, 2, false, false, true>;
          else kern = atoms == 1 ? XV_PICK(0, 1) : XV_PICK(0, 2);
#undef XV_PICK
          XV_CUDA(launch_k(pdl, kern, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::SMEM_BYTES, stream, ta, tw, tc, a));
        }
        XV_PROF();
        XV_CUDA(cudaGetLastError());
        ++launches;
      }
    }
    if (layer_out_dev && layer_out_dev[i]) {
      XV_CUDA(launch_k(pdl, xvk::unpack_rows_kernel, dim3(n_seg), dim3(256), 0, stream, static_cast<const __half*>(out), seg, int32_t(L.c_out), layer_out_dev[i],
                       std::ldexp(1.0f, L.exp_out), int32_t(split)));
      XV_CUDA(cudaGetLastError());
      ++launches;
    }
    in = out;
  }

  // ---- statistics pooling + embed_layer-0 -----------------------------------------------
  bool utt_done = false;
  {
    if (attention) {
      // ---- self-attention pooling (models.py:1037-1051): score GEMM on the tensor cores, softmax over time, weighted sums
      const int C = m->c_pool, W = m->topo.width[nl - 1];
      float* score = reinterpret_cast<float*>(ws + p.off_score);
      float* attn = reinterpret_cast<float*>(ws + p.off_attn);
      CUtensorMap ta, tw;
      const int split = m->opt_split ? 1 : 0;
      // h1 = the first C channels of the stored last layer; split precision: their lo terms sit W columns further on
      rc = encode_2d(m, &ta, hlast, uint64_t(split ? W + C : C), uint64_t(r_pad), tdnn2::BLOCK_K, tdnn2::ACT_BOX_ROWS_PLAIN,
                     CU_TENSOR_MAP_SWIZZLE_128B, uint64_t((split ? 2 : 1) * W));
      if (rc != XV_OK) return rc;
      rc = encode_2d(m, &tw, m->att_w_dev, uint64_t((split ? 3 : 1) * C), uint64_t(C), tdnn2::BLOCK_K, tdnn2::CTA_CH, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != XV_OK) return rc;
      tdnn2::PairArgs a{};
      a.n_row_tiles = int32_t(r_pad / tdnn2::TILE_ROWS);
      a.n_ch_tiles = C / tdnn2::TILE_CH;
      a.taps = 1; a.dilation = 1; a.c_in_pad = (split ? 3 : 1) * C; a.reuse = 0;
      a.split_c = split ? C : 0;
      a.split_lo_off = W;
      a.c_out = C;
      a.bias = m->att_b_dev; a.scale = m->att_v_dev;
      a.partial = score;
      a.overflow_flag = m->cur_flag;
      a.acc_scale = std::ldexp(1.0f, m->layers[nl - 1].exp_out);
      a.mode = 3;
      a.n_act_stages = a.n_wgt_stages = int(std::min<int64_t>(tdnn2::MAX_STAGES, tdnn2::RING_BYTES / (2 * (tdnn2::ACT_BOX_ROWS_PLAIN * 128 + tdnn2::WGT_ATOM_BYTES))));
      a.c_chunks = (split ? 3 : 1) * C / (2 * tdnn2::BLOCK_K);
      const int64_t tiles = int64_t(a.n_row_tiles) * a.n_ch_tiles;
      const int grid = 2 * int(std::min<int64_t>(tiles, m->num_clusters));
      XV_PROF();
      XV_CUDA(launch_k(pdl, tdnn2::tdnn_pair_kernel<3, 2, false>, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::SMEM_BYTES, stream, ta, tw, tw, a));
      XV_PROF();
      XV_CUDA(cudaGetLastError());
      ++launches;
      XV_PROF();
      XV_CUDA(launch_k(pdl, xvk::attn_softmax_kernel, dim3(n_seg), dim3(xvk::ATTN_THREADS), 0, stream, static_cast<const float*>(score),
                       int32_t(2 * a.n_ch_tiles), seg, attn));
      XV_PROF();
      XV_CUDA(cudaGetLastError());
      ++launches;
      XV_PROF();
      XV_CUDA(launch_k(pdl, xvk::attn_pool_kernel, dim3(unsigned(r_pad / 32) * (C / 256)), dim3(256), 0, stream, static_cast<const __half*>(hlast),
                       int32_t((split ? 2 : 1) * W), int32_t(C), int32_t(C), static_cast<const float*>(attn), static_cast<const uint8_t*>(blk_valid), pool_partial,
                       std::ldexp(1.0f, m->layers[nl - 1].exp_out), int32_t(split ? W : 0)));
      XV_PROF();
      XV_CUDA(cudaGetLastError());
      ++launches;
    }
    const bool fc_tc = m->opt_fc == 1 && (2 * m->c_pool) % 128 == 0 && m->topo.emb_dim % tdnn2::TILE_CH == 0;
    float* stats = stats_out_dev ? stats_out_dev : (fc_tc ? nullptr : reinterpret_cast<float*>(ws + p.off_stats));
    __half* split = fc_tc ? reinterpret_cast<__half*>(ws + p.off_split) : nullptr;
    float* fc_partial = reinterpret_cast<float*>(ws + p.off_partial);
    const int K = 2 * m->c_pool, E = m->topo.emb_dim;
    {
      xvk::StatsArgs a{};
      a.partial = pool_partial;
      a.seg = seg;
      a.channels = m->c_pool;
      a.weighted = attention ? 1 : 0;
      a.stats = stats;
      a.split = split;
      a.var_eps = m->topo.var_eps;
      a.sum_scale = attention ? 1.0f : std::ldexp(1.0f, m->layers[nl - 1].exp_out);   // (attn_pool_kernel already rescaled its rows)
      a.split_scale = std::ldexp(1.0f, -m->exp_stats);
      a.overflow_flag = m->cur_flag;
      dim3 grid((a.channels + xvk::STATS_THREADS - 1) / xvk::STATS_THREADS, n_seg);
      XV_PROF();
      XV_CUDA(launch_k(pdl, xvk::pool_stats_kernel, grid, dim3(xvk::STATS_THREADS), 0, stream, a));
      XV_PROF();
      XV_CUDA(cudaGetLastError());
      ++launches;
    }
    if (fc_tc) {
      // embed_layer-0 (models.py:495) on the tensor cores: [n_seg, 3K] x [E, 3K]^T, split over K across the CTA pairs;
      // segments are taken FC_GROUP at a time so that the K-split partials stay small (one group for usual batches)
      CUtensorMap tw;
      rc = encode_2d(m, &tw, m->w0_split_dev, uint64_t(3 * K), uint64_t(E), tdnn2::BLOCK_K, tdnn2::CTA_CH, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != XV_OK) return rc;
      for (int64_t g0 = 0; g0 < n_seg; g0 += FC_GROUP) {
        const int32_t gn = int32_t(std::min<int64_t>(FC_GROUP, n_seg - g0));
        CUtensorMap ta;
        rc = encode_2d(m, &ta, split + g0 * 3 * K, uint64_t(3 * K), uint64_t(round_up(gn, tdnn2::CTA_ROWS)), tdnn2::BLOCK_K,
                       tdnn2::ACT_BOX_ROWS_PLAIN, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != XV_OK) return rc;
        tdnn2::PairArgs a{};
        a.n_row_tiles = (gn + tdnn2::TILE_ROWS - 1) / tdnn2::TILE_ROWS;
        a.n_ch_tiles = E / tdnn2::TILE_CH;
        a.k_splits = p.tc_splits;
        a.c_chunks = (3 * K / 128) / p.tc_splits;
        a.taps = 1;
        a.dilation = 1;
        a.c_in_pad = 3 * K;
        a.reuse = 0;
        a.n_act_stages = a.n_wgt_stages = 3;
        a.mode = 2;
        a.c_out = E;
        a.n_rows = gn;
        a.out_f32 = fc_partial;
        a.overflow_flag = m->cur_flag;
        const int64_t tiles = int64_t(a.n_row_tiles) * a.n_ch_tiles * a.k_splits;
        const int grid = 2 * int(std::min<int64_t>(tiles, m->num_clusters));
        XV_PROF();
        XV_CUDA(launch_k(pdl, tdnn2::tdnn_pair_kernel<2, 2, false>, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::SMEM_BYTES, stream, ta, tw, tw, a));
        XV_PROF();
        XV_CUDA(cudaGetLastError());
        ++launches;
        xvk::FcReduceArgs r{};
        r.partial = fc_partial;
        r.b0 = m->b0_dev;
        r.emb = emb_dev + g0 * E;
        r.n_seg = gn;
        r.E = E;
        r.splits = p.tc_splits;
        r.out_scale = std::ldexp(1.0f, m->exp_stats);
        if (utt && n_seg <= FC_GROUP) {
          // one group holds every segment: K-split reduction and make_embedding's chunk average in ONE launch
          xvk::FcReduceUttArgs ru{};
          ru.r = r;
          ru.u.first_seg = utt_first_dev;
          ru.u.seg_len = seg.len;
          ru.u.dst_row = utt_dst_dev;
          ru.u.out = utt->out_dev;
          ru.u.out_local = utt->out_local_dev;
          ru.u.n_utt = n_utt;
          ru.u.E = E;
          const int64_t u4 = int64_t(n_utt) * E / 4;
          XV_PROF();
          XV_CUDA(launch_k(pdl, xvk::embed_reduce_utt_kernel, dim3(unsigned((u4 + 127) / 128)), dim3(128), 0, stream, ru));
          XV_PROF();
          XV_CUDA(cudaGetLastError());
          ++launches;
          utt_done = true;
          break;
        }
        const int64_t n4 = int64_t(gn) * E / 4;
        XV_PROF();
        XV_CUDA(launch_k(pdl, xvk::embed_reduce_kernel, dim3(unsigned((n4 + 63) / 64)), dim3(64), 0, stream, r));
        XV_PROF();
        XV_CUDA(cudaGetLastError());
        ++launches;
      }
    } else {
      xvk::FcArgs a{};
      a.stats = stats;
      a.w0 = m->w0_dev;
      a.b0 = m->b0_dev;
      a.fc_partial = fc_partial;
      a.counters = counters;
      a.emb = emb_dev;
      a.n_seg = n_seg;
      a.K = K;
      a.E = E;
      a.k_per_split = p.fc_k_per_split;
      dim3 grid(p.fc_m_tiles, p.fc_n_tiles, p.fc_splits);
      XV_PROF();
      XV_CUDA(launch_k(pdl, xvk::embed_fc_kernel, grid, dim3(xvk::FC_THREADS), 0, stream, a));
      XV_PROF();
      XV_CUDA(cudaGetLastError());
      ++launches;
    }
  }
  if (utt && !utt_done) {
    xvk::UttAvgArgs a{};
    a.seg_emb = emb_dev;
    a.first_seg = utt_first_dev;
    a.seg_len = seg.len;
    a.dst_row = utt_dst_dev;
    a.out = utt->out_dev;
    a.out_local = utt->out_local_dev;
    a.n_utt = n_utt;
    a.E = m->topo.emb_dim;
    const int64_t n4 = int64_t(n_utt) * a.E / 4;
    XV_PROF();
    XV_CUDA(launch_k(pdl, xvk::utt_average_kernel, dim3(unsigned((n4 + 127) / 128)), dim3(128), 0, stream, a));
    XV_PROF();
    XV_CUDA(cudaGetLastError());
    ++launches;
  }
  m->last_launches = launches;
#undef XV_PROF
  return XV_OK;
}

}  // namespace

extern "C" {

const char* xv_last_error(void) { return g_err.c_str(); }
const char* xv_version(void) { return "xvec_b200 0.2 (sm_100a; tcgen05 + TMA)"; }
int32_t xv_abi_version(void) { return XV_ABI_VERSION; }
size_t xv_topology_size(void) { return sizeof(xv_topology); }

int xv_create(xv_model** out, int device, const xv_topology* topo) {
  if (!out || !topo) return fail(XV_EINVAL, "null argument");
  *out = nullptr;
  const xv_topology& t = *topo;
  if (t.n_frame_layers < 2 || t.n_frame_layers > XV_MAX_FRAME_LAYERS) return fail(XV_EINVAL, "n_frame_layers out of range");
  if (t.feat_dim <= 0) return fail(XV_EINVAL, "feat_dim must be positive");
  if (t.act != XV_ACT_RELU && t.act != XV_ACT_LRELU && t.act != XV_ACT_PRELU) return fail(XV_EINVAL, "unknown activation");
  if (t.pooling != XV_POOL_STATS && t.pooling != XV_POOL_ATTENTION) return fail(XV_EINVAL, "unknown pooling");
  if (t.pooling == XV_POOL_ATTENTION && t.width[t.n_frame_layers - 1] % (2 * tdnn2::TILE_CH) != 0)
    return fail(XV_EINVAL, "attention pooling needs a last frame layer whose width is a multiple of 512");
  if (t.emb_dim <= 0 || t.emb_dim % tdnn2::TILE_CH != 0 || t.emb_dim > 4096)
    return fail(XV_EINVAL, "emb_dim must be a multiple of 256 and <= 4096");
  for (int i = 0; i < t.n_frame_layers; ++i) {
    if (t.taps[i] < 1 || t.taps[i] % 2 == 0) return fail(XV_EINVAL, "taps must be odd and >= 1");
    if (t.dilation[i] < 1) return fail(XV_EINVAL, "dilation must be >= 1");
    if (t.width[i] <= 0 || t.width[i] % tdnn2::TILE_CH != 0) return fail(XV_EINVAL, "layer widths must be multiples of 256");
  }
  {
    const int halo0 = (t.taps[0] - 1) / 2 * t.dilation[0];
    const int64_t k0 = round_up(int64_t(t.taps[0]) * t.feat_dim, (2 * tdnn2::BLOCK_K));
    if ((xvk::PACK_ROWS_PER_BLOCK + 2 * halo0) * int64_t(t.feat_dim) > xvk::PACK_MAX_STAGE_FLOATS || k0 > xvk::PACK_MAX_K0 ||
        int64_t(t.taps[0]) * t.dilation[0] * t.feat_dim > 32000)
      return fail(XV_EINVAL, "first layer too wide for the pack kernel (taps*feat_dim must be <= 512)");
  }
  int n_dev = 0;
  XV_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) return fail(XV_EINVAL, "no such CUDA device");
  XV_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  XV_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(XV_ECUDA, std::string("this library is built for sm_100a (B200) only; device is sm_") +
                              std::to_string(prop.major) + std::to_string(prop.minor));
  xv_model* m = new xv_model();
  m->topo = t;
  m->device = device;
  m->num_sms = prop.multiProcessorCount;
  m->layers.resize(t.n_frame_layers);
  int prev = t.feat_dim;
  for (int i = 0; i < t.n_frame_layers; ++i) {
    FrameLayer& L = m->layers[i];
    L.taps = t.taps[i];
    L.dilation = t.dilation[i];
    L.c_in = prev;
    L.c_out = t.width[i];
    if (i == 0) {                          // spliced (im2col) input: dense K, padded to a multiple of 64
      L.c_in_pad = prev;
      L.k_total = int(round_up(int64_t(L.taps) * prev, (2 * tdnn2::BLOCK_K)));
      L.gemm_taps = 1;
      m->k0_pad = L.k_total;
    } else {
      L.c_in_pad = prev;                   // widths are multiples of 256, hence of 64
      L.k_total = L.taps * prev;
      L.gemm_taps = L.taps;
      m->gap = std::max(m->gap, (L.taps - 1) / 2 * L.dilation);
    }
    if (i < t.n_frame_layers - 1) m->w_mid = std::max(m->w_mid, L.c_out);
    prev = L.c_out;
  }
  m->gap = std::max(m->gap, 1);
  m->c_pool = t.pooling == XV_POOL_ATTENTION ? prev / 2 : prev;
  m->opt_split = t.pooling == XV_POOL_ATTENTION ? 1 : 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    delete m;
    return fail(XV_ECUDA, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  }
  m->encode = reinterpret_cast<EncodeTiledFn>(fn);
  e = cudaSuccess;
  {
    void (*kernels[])(CUtensorMap, CUtensorMap, CUtensorMap, tdnn2::PairArgs) = {
        tdnn2::tdnn_pair_kernel<0, 1, false>, tdnn2::tdnn_pair_kernel<0, 2, false>, tdnn2::tdnn_pair_kernel<1, 1, false>,
        tdnn2::tdnn_pair_kernel<1, 2, false>, tdnn2::tdnn_pair_kernel<0, 1, true>,  tdnn2::tdnn_pair_kernel<0, 2, true>,
        tdnn2::tdnn_pair_kernel<1, 1, true>,  tdnn2::tdnn_pair_kernel<1, 2, true>,  tdnn2::tdnn_pair_kernel<2, 2, false>,
        tdnn2::tdnn_pair_kernel<3, 2, false>, tdnn2::tdnn_pair_kernel<0, 2, false, false, true>,
        tdnn2::tdnn_pair_kernel<0, 2, true, false, true>};
    for (auto k : kernels)
      if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, tdnn2::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdnn2::tdnn_tail_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdnn2::FT_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdnn2::tdnn_tail_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdnn2::FT_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdnn2::tdnn_first_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdnn2::F1_SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tdnn2::tdnn_first_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdnn2::F1_SMEM_BYTES);
  }
  if (e == cudaSuccess) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (prop.multiProcessorCount / 2));
    cfg.blockDim = dim3(tdnn2::NUM_THREADS);
    cfg.dynamicSmemBytes = tdnn2::SMEM_BYTES;
    int n_clusters = 0;
    // the kernel carries __cluster_dims__(2,1,1); ask how many pairs can be co-resident
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaError_t qe = cudaOccupancyMaxActiveClusters(&n_clusters, tdnn2::tdnn_pair_kernel<0, 2, false>, &cfg);
    if (qe != cudaSuccess || n_clusters <= 0) { (void)cudaGetLastError(); n_clusters = prop.multiProcessorCount / 2; }
    m->num_clusters = std::min(n_clusters, prop.multiProcessorCount / 2);
  }
  if (e == cudaSuccess) {
    std::vector<int32_t> lut(m->k0_pad, -1);
    const FrameLayer& L0 = m->layers[0];
    for (int ch = 0; ch < L0.taps * t.feat_dim; ++ch) lut[ch] = (ch / t.feat_dim) * L0.dilation * t.feat_dim + ch % t.feat_dim;
    const int halo0 = (L0.taps - 1) / 2 * L0.dilation;
    m->first_fusable = m->k0_pad == tdnn2::F1_K && L0.dilation == 1 && L0.c_out <= tdnn2::F1_MAX_CT * tdnn2::TILE_CH &&
                       (32 + 2 * halo0) * t.feat_dim * 4 + 30 <= tdnn2::F1_STAGE_BYTES;
    e = cudaMalloc(&m->pack_lut_dev, lut.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(m->pack_lut_dev, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMalloc(&m->overflow_dev, 4 * (1 + XV_HOST_SLOTS));
  if (e == cudaSuccess) e = cudaMemset(m->overflow_dev, 0, 4 * (1 + XV_HOST_SLOTS));
  m->cur_flag = m->overflow_dev;
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&m->overflow_host), 4, cudaHostAllocDefault);
  for (int i = 0; i < XV_HOST_SLOTS && e == cudaSuccess; ++i) {
    e = cudaStreamCreateWithFlags(&m->slots[i].stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&m->slots[i].overflow_host), 4, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->slots[i].done, cudaEventBlockingSync | cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    std::string msg = std::string("xv_create: ") + cudaGetErrorName(e) + ": " + cudaGetErrorString(e);
    xv_destroy(m);
    return fail(XV_ECUDA, msg);
  }
  *out = m;
  return XV_OK;
}

void xv_destroy(xv_model* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  free_layers(m);
  for (int s = 0; s < META_SLOTS; ++s) {
    if (m->meta_host[s]) cudaFreeHost(m->meta_host[s]);
    if (m->meta_event[s]) cudaEventDestroy(m->meta_event[s]);
    if (m->utt_host[s]) cudaFreeHost(m->utt_host[s]);
    if (m->utt_event[s]) cudaEventDestroy(m->utt_event[s]);
  }
  for (cudaEvent_t e : m->prof_events) cudaEventDestroy(e);
  cudaFree(m->overflow_dev);
  cudaFree(m->pack_lut_dev);
  if (m->overflow_host) cudaFreeHost(m->overflow_host);
  for (auto& sl : m->slots) {
    cudaFree(sl.feats_dev); cudaFree(sl.emb_dev); cudaFree(sl.ws_dev);
    cudaFree(sl.raw_dev); cudaFree(sl.vad_dev); cudaFree(sl.fe_ws_dev);
    if (sl.overflow_host) cudaFreeHost(sl.overflow_host);
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.stream) cudaStreamDestroy(sl.stream);
  }
  delete m;
}

int xv_set_param(xv_model* m, const char* tf_var_name, const float* host, const int64_t* shape, int32_t rank) {
  if (!m || !tf_var_name || !host || !shape || rank < 1 || rank > 3) return fail(XV_EINVAL, "bad argument");
  const std::string name(tf_var_name);
  const xv_topology& t = m->topo;
  std::vector<int64_t> want;
  bool used = false, known = false;
  auto scope_of = [&](const std::string& prefix, int& idx, std::string& leaf) {
    if (name.compare(0, prefix.size(), prefix) != 0) return false;
    const size_t slash = name.find('/', prefix.size());
    if (slash == std::string::npos) return false;
    try { idx = std::stoi(name.substr(prefix.size(), slash - prefix.size())); } catch (...) { return false; }
    leaf = name.substr(slash + 1);
    return true;
  };
  int idx = -1;
  std::string leaf;
  if (scope_of("frame_level_info_layer-", idx, leaf)) {
    if (idx < 0 || idx >= t.n_frame_layers) return fail(XV_EINVAL, "no such frame layer: " + name);
    const FrameLayer& L = m->layers[idx];
    known = used = true;
    if (leaf == "w:0") want = {L.taps, L.c_in, L.c_out};
    else if (leaf == "b:0" || leaf == "gamma:0" || leaf == "beta:0" || leaf == "mean:0" || leaf == "variance:0" ||
             leaf == "prelu/prelu:0") want = {L.c_out};
    else known = used = false;
  } else if (scope_of("embed_layer-", idx, leaf)) {
    known = true;                            // embed_layer-*/prelu/prelu:0 etc.: training-only, ignored
    if (idx == 0 && leaf == "w:0") { used = true; want = {2 * m->c_pool, t.emb_dim}; }
    else if (idx == 0 && leaf == "b:0") { used = true; want = {t.emb_dim}; }
  } else if (name.compare(0, 10, "attention/") == 0 && t.pooling == XV_POOL_ATTENTION) {
    known = used = true;
    if (name == "attention/w:0") want = {m->c_pool, m->c_pool};
    else if (name == "attention/b:0" || name == "attention/v:0") want = {m->c_pool};
    else known = used = false;
  } else if (name.compare(0, 7, "output/") == 0 || name.compare(0, 10, "attention/") == 0) {
    known = true;
  }
  if (!known) return fail(XV_EINVAL, "unknown variable name: " + name);
  if (!used) return XV_OK;                   // training-only variable: not read by the extraction path
  if (int(want.size()) != rank || !std::equal(want.begin(), want.end(), shape)) {
    std::string got = "[", exp = "[";
    for (int i = 0; i < rank; ++i) got += std::to_string(shape[i]) + (i + 1 < rank ? "," : "]");
    for (size_t i = 0; i < want.size(); ++i) exp += std::to_string(want[i]) + (i + 1 < want.size() ? "," : "]");
    return fail(XV_EINVAL, "shape mismatch for " + name + ": got " + got + ", expected " + exp);
  }
  size_t n = 1;
  for (int i = 0; i < rank; ++i) n *= size_t(shape[i]);
  for (size_t i = 0; i < n; ++i)
    if (!(host[i] == host[i]) || host[i] > 3.0e38f || host[i] < -3.0e38f) return fail(XV_EINVAL, "non-finite value in " + name);
  m->host_params[name].assign(host, host + n);
  m->host_shapes[name].assign(shape, shape + rank);
  m->dirty = true;
  return XV_OK;
}

size_t xv_workspace_bytes(const xv_model* m, int64_t total_frames, int32_t n_seg) {
  if (!m || total_frames <= 0 || n_seg <= 0) return 0;
  return make_plan(m, total_frames, n_seg).bytes;
}

int xv_forward(xv_model* m, const float* feats_dev, const int32_t* seg_len_host, int32_t n_seg, float* emb_dev,
               void* workspace_dev, size_t workspace_bytes, void* stream) {
  return forward_impl(m, feats_dev, seg_len_host, n_seg, emb_dev, workspace_dev, workspace_bytes,
                      static_cast<cudaStream_t>(stream), nullptr, nullptr);
}

int xv_forward_layers(xv_model* m, const float* feats_dev, const int32_t* seg_len_host, int32_t n_seg, float* emb_dev,
                      void* workspace_dev, size_t workspace_bytes, void* stream, float* const* layer_out_dev,
                      float* stats_out_dev) {
  return forward_impl(m, feats_dev, seg_len_host, n_seg, emb_dev, workspace_dev, workspace_bytes,
                      static_cast<cudaStream_t>(stream), layer_out_dev, stats_out_dev);
}

namespace {
// Forward of the submission remembered in slot `si` (features already in the slot's device buffer) + the copies back.
int enqueue_slot(xv_model* m, int si) {
  xv_model::HostSlot& sl = m->slots[si];
  const xv_model::HostSlot::Redo& rd = sl.redo;
  const int32_t n_seg = int32_t(rd.seg_len.size());
  m->cur_flag = m->overflow_dev + 1 + si;
  m->cur_feats_f16 = rd.feats_f16 ? 1 : 0;
  struct Restore { xv_model* m; ~Restore() { m->cur_feats_f16 = 0; } } restore{m};
  int rc;
  if (rd.utt) {
    UttOut u;
    u.first_seg_host = rd.has_first ? rd.first_seg.data() : nullptr;
    u.dst_row_host = rd.has_dst ? rd.dst_row.data() : nullptr;
    u.n_utt = rd.n_utt;
    u.out_dev = rd.out_dev;
    u.out_local_dev = rd.host_out ? sl.emb_dev : nullptr;        // the slot's buffer holds the rows the host reads
    rc = forward_impl(m, rd.feats_ext ? rd.feats_ext : sl.feats_dev, rd.seg_len.data(), n_seg, nullptr, sl.ws_dev, sl.ws_cap, sl.stream,
                      nullptr, nullptr, &u);
    m->cur_flag = m->overflow_dev;
    if (rc != XV_OK) return rc;
    if (rd.host_out)
      XV_CUDA(cudaMemcpyAsync(rd.host_out, sl.emb_dev, size_t(rd.n_utt) * m->topo.emb_dim * 4, cudaMemcpyDeviceToHost, sl.stream));
  } else {
    rc = forward_impl(m, rd.feats_ext ? rd.feats_ext : sl.feats_dev, rd.seg_len.data(), n_seg, sl.emb_dev, sl.ws_dev, sl.ws_cap, sl.stream,
                      nullptr, nullptr);
    m->cur_flag = m->overflow_dev;
    if (rc != XV_OK) return rc;
    XV_CUDA(cudaMemcpyAsync(rd.host_out, sl.emb_dev, size_t(n_seg) * m->topo.emb_dim * 4, cudaMemcpyDeviceToHost, sl.stream));
  }
  XV_CUDA(cudaMemcpyAsync(sl.overflow_host, m->overflow_dev + 1 + si, 4, cudaMemcpyDeviceToHost, sl.stream));
  XV_CUDA(cudaEventRecord(sl.done, sl.stream));
  return XV_OK;
}

// Common body of xv_submit_host / xv_submit_host_utts.
int submit_impl(xv_model* m, const float* feats_host, const int32_t* seg_len_host, int32_t n_seg, float* emb_host,
                const UttOut* utt_in, float* utt_host_out, int32_t* ticket, const float* feats_dev_ext = nullptr,
                bool feats_f16 = false) {
  if (!m || (!feats_host && !feats_dev_ext) || !seg_len_host || !ticket) return fail(XV_EINVAL, "null argument");
  if (n_seg <= 0) return fail(XV_EINVAL, "n_seg must be >= 1");
  XV_CUDA(cudaSetDevice(m->device));
  int64_t total = 0;
  for (int i = 0; i < n_seg; ++i) {
    if (seg_len_host[i] <= 0) return fail(XV_EINVAL, "segment " + std::to_string(i) + " has non-positive length");
    total += seg_len_host[i];
  }
  const int si = m->slot_next;
  xv_model::HostSlot& sl = m->slots[si];
  if (sl.busy) return fail(XV_ESTATE, "all submission slots are in flight: xv_collect the oldest ticket first");
  if (feats_f16 && m->opt_split)
    return fail(XV_EINVAL, "float16 features cannot feed the split-precision model: it needs the float32 values (hi + lo terms)");
  const size_t feat_bytes = size_t(total) * m->topo.feat_dim * (feats_f16 ? 2 : 4);
  const size_t emb_bytes = size_t(n_seg) * m->topo.emb_dim * 4;
  const size_t ws_bytes = xv_workspace_bytes(m, total, n_seg);
  auto grow = [&](void** p, size_t* cap, size_t need) -> cudaError_t {
    if (need <= *cap) return cudaSuccess;
    cudaStreamSynchronize(sl.stream);
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    const size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(p, want);
    if (e == cudaSuccess) *cap = want;
    return e;
  };
  sl.redo.feats_f16 = feats_f16;
  if (!feats_dev_ext) XV_CUDA(grow(reinterpret_cast<void**>(&sl.feats_dev), &sl.feats_cap, feat_bytes));
  XV_CUDA(grow(reinterpret_cast<void**>(&sl.emb_dev), &sl.emb_cap, emb_bytes));
  XV_CUDA(grow(&sl.ws_dev, &sl.ws_cap, ws_bytes));
  if (!feats_dev_ext) XV_CUDA(cudaMemcpyAsync(sl.feats_dev, feats_host, feat_bytes, cudaMemcpyHostToDevice, sl.stream));
  sl.redo.feats_ext = feats_dev_ext;
  // remember the submission: an fp16 range rescue in xv_collect re-runs it from the slot's device copy of the features
  xv_model::HostSlot::Redo& rd = sl.redo;
  rd.seg_len.assign(seg_len_host, seg_len_host + n_seg);
  rd.utt = utt_in != nullptr;
  rd.has_first = rd.has_dst = false;
  rd.n_utt = 0;
  rd.out_dev = nullptr;
  rd.host_out = utt_in ? utt_host_out : emb_host;
  if (utt_in) {
    rd.n_utt = utt_in->first_seg_host ? utt_in->n_utt : n_seg;
    rd.out_dev = utt_in->out_dev;
    if (utt_in->first_seg_host) { rd.has_first = true; rd.first_seg.assign(utt_in->first_seg_host, utt_in->first_seg_host + rd.n_utt + 1); }
    if (utt_in->dst_row_host) { rd.has_dst = true; rd.dst_row.assign(utt_in->dst_row_host, utt_in->dst_row_host + rd.n_utt); }
  }
  int rc = enqueue_slot(m, si);
  if (rc != XV_OK) return rc;
  sl.busy = true;
  m->slot_next = (si + 1) % XV_HOST_SLOTS;
  *ticket = si;
  return XV_OK;
}
}  // namespace

int xv_submit_host(xv_model* m, const float* feats_host, const int32_t* seg_len_host, int32_t n_seg, float* emb_host,
                   int32_t* ticket) {
  if (!emb_host) return fail(XV_EINVAL, "null argument");
  return submit_impl(m, feats_host, seg_len_host, n_seg, emb_host, nullptr, nullptr, ticket);
}

int xv_forward_utts(xv_model* m, const float* feats_dev, const int32_t* seg_len_host, int32_t n_seg,
                    const int32_t* utt_first_seg_host, const int64_t* utt_dst_row_host, int32_t n_utt, float* out_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream) {
  if (!out_dev) return fail(XV_EINVAL, "null argument");
  UttOut u;
  u.first_seg_host = utt_first_seg_host;
  u.dst_row_host = utt_dst_row_host;
  u.n_utt = n_utt;
  u.out_dev = out_dev;
  return forward_impl(m, feats_dev, seg_len_host, n_seg, nullptr, workspace_dev, workspace_bytes,
                      static_cast<cudaStream_t>(stream), nullptr, nullptr, &u);
}

int xv_submit_host_utts(xv_model* m, const float* feats_host, const int32_t* seg_len_host, int32_t n_seg,
                        const int32_t* utt_first_seg_host, const int64_t* utt_dst_row_host, int32_t n_utt, float* out_dev,
                        float* out_host, int32_t* ticket) {
  if (!out_dev && !out_host) return fail(XV_EINVAL, "xv_submit_host_utts: no destination (out_dev and out_host are both null)");
  UttOut u;
  u.first_seg_host = utt_first_seg_host;
  u.dst_row_host = utt_dst_row_host;
  u.n_utt = n_utt;
  u.out_dev = out_dev;
  return submit_impl(m, feats_host, seg_len_host, n_seg, nullptr, &u, out_host, ticket);
}

int xv_submit_host_utts_f16(xv_model* m, const uint16_t* feats_host_f16, const int32_t* seg_len_host, int32_t n_seg,
                            const int32_t* utt_first_seg_host, const int64_t* utt_dst_row_host, int32_t n_utt, float* out_dev,
                            float* out_host, int32_t* ticket) {
  if (!out_dev && !out_host) return fail(XV_EINVAL, "xv_submit_host_utts_f16: no destination (out_dev and out_host are both null)");
  UttOut u;
  u.first_seg_host = utt_first_seg_host;
  u.dst_row_host = utt_dst_row_host;
  u.n_utt = n_utt;
  u.out_dev = out_dev;
  return submit_impl(m, reinterpret_cast<const float*>(feats_host_f16), seg_len_host, n_seg, nullptr, &u, out_host, ticket, nullptr, true);
}

int xv_submit_dev_utts(xv_model* m, const float* feats_dev, const int32_t* seg_len_host, int32_t n_seg,
                       const int32_t* utt_first_seg_host, const int64_t* utt_dst_row_host, int32_t n_utt, float* out_dev,
                       float* out_host, void* ready_event, int32_t* ticket) {
  if (!m || !feats_dev) return fail(XV_EINVAL, "null argument");
  if (ready_event) {                        // the producer of the features runs on another stream: order this slot's stream behind it
    XV_CUDA(cudaSetDevice(m->device));
    XV_CUDA(cudaStreamWaitEvent(m->slots[m->slot_next].stream, static_cast<cudaEvent_t>(ready_event), 0));
  }
  if (!out_dev && !out_host) return fail(XV_EINVAL, "xv_submit_dev_utts: no destination (out_dev and out_host are both null)");
  UttOut u;
  u.first_seg_host = utt_first_seg_host;
  u.dst_row_host = utt_dst_row_host;
  u.n_utt = n_utt;
  u.out_dev = out_dev;
  return submit_impl(m, nullptr, seg_len_host, n_seg, nullptr, &u, out_host, ticket, feats_dev);
}

// ---- peer memory (one node, one process per GPU): rank 0's result table mapped into every rank ----
int xv_peer_alloc(int device, size_t bytes, void** dev_ptr, uint8_t* handle64) {
  if (!dev_ptr || !handle64 || bytes == 0) return fail(XV_EINVAL, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  XV_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  XV_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(XV_ECUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  std::memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return XV_OK;
}

int xv_peer_open(int device, const uint8_t* handle64, void** dev_ptr) {
  if (!dev_ptr || !handle64) return fail(XV_EINVAL, "bad argument");
  XV_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  XV_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return XV_OK;
}

int xv_peer_close(int device, void* dev_ptr) {
  if (!dev_ptr) return XV_OK;
  XV_CUDA(cudaSetDevice(device));
  XV_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return XV_OK;
}

int xv_peer_free(int device, void* dev_ptr) {
  if (!dev_ptr) return XV_OK;
  XV_CUDA(cudaSetDevice(device));
  XV_CUDA(cudaFree(dev_ptr));
  return XV_OK;
}

int xv_peer_read(int device, void* dst_host, const void* src_dev, size_t bytes) {
  if (!dst_host || !src_dev) return fail(XV_EINVAL, "null argument");
  XV_CUDA(cudaSetDevice(device));
  XV_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return XV_OK;
}

int xv_collect(xv_model* m, int32_t ticket) {
  if (!m || ticket < 0 || ticket >= XV_HOST_SLOTS) return fail(XV_EINVAL, "bad ticket");
  xv_model::HostSlot& sl = m->slots[ticket];
  if (!sl.busy) return fail(XV_ESTATE, "ticket is not in flight");
  XV_CUDA(cudaSetDevice(m->device));
  sl.busy = false;
  // option "blocking_collect": sleep on the slot's blocking-sync event instead of spinning in the driver (frees a core per
  // process at the price of a wake-up latency per collect; measured slower for 0.5 ms steps, so off by default)
  if (m->opt_blocking_collect) XV_CUDA(cudaEventSynchronize(sl.done));
  else XV_CUDA(cudaStreamSynchronize(sl.stream));
  uint32_t flag = *sl.overflow_host;
  // fp16 range rescue: a store of a frame layer (bits 8..) or a pooled statistic (bit 16) overflowed.  Raise the exponents
  // of what overflowed and run the submission again from the slot's device copy of its features, until it fits.
  for (int attempt = 0; flag != 0 && (flag & 3u) == 0 && m->opt_rescue && attempt < 20; ++attempt) {
    if (raise_exponents(m, flag) <= 0) break;
    for (auto& other : m->slots) XV_CUDA(cudaStreamSynchronize(other.stream));   // nobody reads the old scale / shift any more
    int rc = apply_exponents(m, nullptr);
    if (rc != XV_OK) return rc;
    XV_CUDA(cudaMemsetAsync(m->overflow_dev + 1 + ticket, 0, 4, sl.stream));
    rc = enqueue_slot(m, ticket);
    if (rc != XV_OK) return rc;
    XV_CUDA(cudaStreamSynchronize(sl.stream));
    flag = *sl.overflow_host;
    ++m->rescues;
  }
  if (flag != 0) {
    XV_CUDA(cudaMemsetAsync(m->overflow_dev + 1 + ticket, 0, 4, sl.stream));
    return sticky_flag_error(flag);
  }
  return XV_OK;
}

int xv_rescue_overflow(xv_model* m, void* stream) {
  if (!m) return fail(XV_EINVAL, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  XV_CUDA(cudaSetDevice(m->device));
  XV_CUDA(cudaMemcpyAsync(m->overflow_host, m->overflow_dev, 4, cudaMemcpyDeviceToHost, s));
  XV_CUDA(cudaStreamSynchronize(s));
  const uint32_t flag = *m->overflow_host;
  if (flag == 0) return 0;
  XV_CUDA(cudaMemsetAsync(m->overflow_dev, 0, 4, s));
  if (flag & 3u) return sticky_flag_error(flag);
  const int raised = raise_exponents(m, flag);
  if (raised <= 0) return fail(XV_EOVERFLOW, "activations exceed the fp16 range even after the largest rescue scale");
  for (auto& sl : m->slots) XV_CUDA(cudaStreamSynchronize(sl.stream));
  int rc = apply_exponents(m, s);
  if (rc != XV_OK) return rc;
  ++m->rescues;
  return raised;
}

int xv_extract_host(xv_model* m, const float* feats_host, const int32_t* seg_len_host, int32_t n_seg, float* emb_host) {
  int32_t ticket = -1;
  int rc = xv_submit_host(m, feats_host, seg_len_host, n_seg, emb_host, &ticket);
  if (rc != XV_OK) return rc;
  return xv_collect(m, ticket);
}

int xv_check_overflow(xv_model* m, void* stream) {
  if (!m) return fail(XV_EINVAL, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  XV_CUDA(cudaSetDevice(m->device));
  XV_CUDA(cudaMemcpyAsync(m->overflow_host, m->overflow_dev, 4, cudaMemcpyDeviceToHost, s));
  XV_CUDA(cudaStreamSynchronize(s));
  if (*m->overflow_host != 0) {
    XV_CUDA(cudaMemsetAsync(m->overflow_dev, 0, 4, s));
    return sticky_flag_error(*m->overflow_host);
  }
  return XV_OK;
}

int32_t xv_last_launch_count(const xv_model* m) { return m ? m->last_launches : 0; }

int32_t xv_last_kernel_ms(xv_model* m, float* ms_out, int32_t cap) {
  if (!m || !ms_out || cap < 0) return fail(XV_EINVAL, "bad argument");
  if (!m->opt_profile) return fail(XV_ESTATE, "option 'profile' is off");
  const int n = m->prof_used / 2;
  if (n == 0) return 0;
  XV_CUDA(cudaSetDevice(m->device));
  XV_CUDA(cudaEventSynchronize(m->prof_events[m->prof_used - 1]));
  for (int i = 0; i < n && i < cap; ++i) XV_CUDA(cudaEventElapsedTime(ms_out + i, m->prof_events[2 * i], m->prof_events[2 * i + 1]));
  return n;
}

int xv_set_option(xv_model* m, const char* name, int64_t value) {
  if (!m || !name) return fail(XV_EINVAL, "null argument");
  const std::string n(name);
  if (n == "profile") m->opt_profile = value != 0;
  else if (n == "resident") m->opt_resident = int(value);
  else if (n == "prefetch") m->opt_prefetch = value != 0;
  else if (n == "clusters") {
    // CTA pairs per persistent grid (default: all that are co-resident, one per pair of SMs).  A data-parallel trainer lowers
    // it to leave SMs to NCCL's kernels: a collective that has to wait for an SM of a persistent grid holds its late CTAs back
    if (value < 1 || value > m->num_sms / 2) return fail(XV_EINVAL, "clusters out of range");
    m->num_clusters = int(value);
  }
  else if (n == "fc") m->opt_fc = int(value);
  else if (n == "pdl") m->opt_pdl = value != 0;
  else if (n == "blocking_collect") m->opt_blocking_collect = value != 0;
  else if (n == "rescue") m->opt_rescue = value != 0;
  else if (n == "fuse_tail") m->opt_fuse_tail = value != 0;
  else if (n == "fuse_first") m->opt_fuse_first = value != 0;
  else if (n == "precision") { m->opt_split = value != 0; m->dirty = true; }
  else if (n == "fc_max_splits") m->opt_fc_max_splits = std::max(1, int(value));
  else if (n == "trace_ptr") m->opt_trace = reinterpret_cast<long long*>(static_cast<intptr_t>(value));   // device buffer
  else if (n == "trace_layer") m->opt_trace_layer = int(value);
  else return fail(XV_EINVAL, "unknown option: " + n);
  return XV_OK;
}

}  // extern "C"

// ---- training step (include/xvec_train.h) ----
#include "train_api.cuh"

// ---- feature front end: sliding-window CMVN + voiced-frame selection (include/xvec_frontend.h) ----
#include "frontend_api.cuh"

// ---- host-only: ark index for the extractor's reader (xv_ark_scan, include/xvec.h) ----
#include "ark_scan.cuh"

// ---- host-only: striped ark reader + vector-ark formatter of an extraction job (include/xvec_job.h) ----
#include "ark_job.cuh"
