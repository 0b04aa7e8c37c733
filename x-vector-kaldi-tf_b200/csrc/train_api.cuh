// train_api.cuh -- host side of the training step (include/xvec_train.h); included at the end of xvec_api.cu so that
// it shares the xv_model internals (tensor-map encoder, metadata staging, profiling events, overflow flag).
//
// One minibatch = what sess.run([optimizer, loss, accuracy]) evaluates (local/tf/models.py:263):
//   forward   pack -> per frame layer: tdnn_pair_kernel<0> (conv+b+relu -> r) -> block column sums -> batch-norm
//             moments -> y = BN(r);  last layer: pooled statistics straight from the block sums of r
//             segment level (fp32 SIMT): embed-0, embed-1 (xw_plus_b -> relu -> BN), output, softmax cross-entropy
//   backward  segment level (fp32), pooling+BN backward folded into per-(segment, channel) coefficients,
//             per frame layer: dz (elementwise) -> wgrad_pair_kernel (dW on tcgen05, MN-major operands) ->
//             tdnn_pair_kernel<0> with the flipped/transposed kernel (dy of the layer below)
//   update    adam_kernel over the flat parameter vector, repack of the fp16 operand copies
#pragma once
#include "../../include/xvec_train.h"
#include "train_kernels.cuh"
#include "wgrad_pair.cuh"

namespace {

struct TrFrame {
  int taps = 0, dil = 1, c_in = 0, c_out = 0, c_in_gemm = 0, k_total = 0, gemm_taps = 0, wg_rows = 0;
  int64_t off_w = 0, off_b = 0, off_gamma = 0, off_beta = 0, off_mov = 0;
  __half* wf = nullptr;       // [c_out, k_total]           forward operand
  __half* wd = nullptr;       // [c_in, taps * c_out]       data-gradient operand (layers >= 1)
  float* bn = nullptr;        // [4][c_out]: batch mean | inv | scale | shift
  __half *r = nullptr, *y = nullptr, *dy = nullptr, *dz = nullptr;    // workspace, [r_pad, c_out]
};
struct TrSeg {
  int in = 0, out = 0;
  int64_t off_w = 0, off_b = 0, off_gamma = 0, off_beta = 0, off_mov = 0;
  float* bn = nullptr;        // [2][out]: batch mean | inv
  float *z = nullptr, *r = nullptr, *y = nullptr, *dy = nullptr, *dz = nullptr;   // workspace, [n_seg, out]
};
struct TrSpan { int32_t which; int64_t offset, count; std::vector<int64_t> shape; };
struct TrDebug { const void* ptr; int32_t cols; int32_t kind; int64_t count; };   // kind 0: packed fp16 rows, 1: fp32 array

constexpr float ADAM_B1 = 0.9f, ADAM_B2 = 0.999f, ADAM_EPS = 1e-8f;   // tf.train.AdamOptimizer defaults
constexpr float BN_DECAY = 0.95f;                                      // models.py:480,497

struct SgemmPlan { int splits, k_per_split; };
SgemmPlan sgemm_plan(int M, int N, int K, int num_sms) {
  const int tiles = ((M + 63) / 64) * ((N + 63) / 64);
  int splits = std::max(1, std::min((2 * num_sms + tiles - 1) / tiles, std::max(1, K / 32)));
  int kps = int(round_up((K + splits - 1) / splits, 16));
  splits = (K + kps - 1) / kps;
  return {splits, kps};
}

}  // namespace

struct xv_trainer {
  xv_model* m = nullptr;
  int32_t num_classes = 0;
  int64_t n_params = 0, n_moving = 0;
  std::vector<TrFrame> frames;
  TrSeg seg[2];
  int64_t off_wo = 0, off_bo = 0;
  std::map<std::string, TrSpan> spans;
  std::vector<std::string> span_order;
  float *params = nullptr, *adam_m = nullptr, *adam_v = nullptr, *moving = nullptr, *grad = nullptr;
  float *ones = nullptr, *zeros = nullptr;        // [max width]
  int64_t step = 0;
  bool operands_dirty = true;
  double opt_loss_scale = 0.0;
  int opt_wgrad_lbo = wgrad::BOX_BYTES, opt_wgrad_sbo = 1024;
  float neg_slope = 0.f;             // activation: 0 = relu, 0.2 = tf.nn.leaky_relu(alpha=0.2) (models.py:912), from the xv_model's topology
  float* slope_dev = nullptr;        // [max width] filled with neg_slope (the LEAKY layer kernel's per-channel slope)
  double opt_l2_beta = 0.0;          // ModelL2Loss*: loss += beta * (0.1 l2(embed-0 w,b) + l2(embed-1 w,b) + l2(output w,b))
  int opt_wgrad_reuse = 0;           // 1: tap-reuse weight-gradient kernel for layers with temporal context (wgrad_reuse_kernel):
                                     // halves the operand bytes per FLOP but needs UMMA 256x128x16 (three taps share the 512 TMEM
                                     // columns) whose A-operand reads per FLOP double -- measured 83 vs 75 us (tdnn), 142 vs 129 us
                                     // (dense) for the two layers it applies to: not faster, kept as a tested option
  int opt_fused_stats = 1;           // 1: BatchNorm / pooling column sums come out of the layer kernel's epilogue (STATS instantiation)
  float last_loss_scale = 0.f;
  // geometry-dependent workspace
  int32_t n_seg = 0, seg_len = 0, seg_stride = 0;
  int64_t r_pad = 0;
  uint8_t* ws = nullptr;
  int32_t* meta = nullptr;
  uint8_t *row_valid = nullptr, *blk_valid = nullptr;
  __half* x0 = nullptr;
  float *partial = nullptr, *partial1 = nullptr, *cA = nullptr, *cB = nullptr, *cC = nullptr;
  float *m_r = nullptr, *v_r = nullptr, *coefA = nullptr, *coefG = nullptr, *h0 = nullptr, *dh0 = nullptr;
  float *logits = nullptr, *dlogits = nullptr, *loss_row = nullptr, *correct = nullptr;
  float *wg_partial = nullptr, *sg_partial = nullptr;
  size_t wg_partial_floats = 0, sg_partial_floats = 0;
  xvk::SegMeta seg_meta{};
  const int4* blk_info_dev = nullptr;           // written by train_meta_kernel at the start of every step
  int32_t blk_info_off = 0;
  size_t ws_cap = 0;
  std::vector<std::pair<int32_t, int32_t>> seen_geometries;   // a step is captured into a graph the second time its geometry shows up
  // CUDA graph of one forward_backward (74 launches): captured once per (geometry, buffer pointers), replayed afterwards
  struct StepGraph { const void* feats; const void* labels; const void* grad; const void* loss; int32_t n_seg, seg_len;
                     cudaGraphExec_t exec; int32_t launches; int32_t part; };
  std::vector<StepGraph> graphs;
  cudaStream_t gstream = nullptr;
  cudaEvent_t g_in = nullptr, g_out = nullptr;
  int opt_graph = 1;                 // 0: plain stream launches
  std::map<std::string, TrDebug> debug;
  int32_t last_launches = 0;
  std::vector<std::string> prof_names;
  // gradient-overflow bookkeeping (separate from the xv_model's forward-activation flag): gflag[0] = a loss-scaled fp16
  // gradient of THIS step left the fp16 range (zeroed at the start of every step, published as grad[n_params] so that a
  // data-parallel all-reduce combines it over the ranks), gflag[1] = updates skipped so far (adam_kernel)
  bool emit = true;                  // false while tr_step_body walks the half of a step that a split call does NOT enqueue
  uint32_t* gflag = nullptr;
  uint32_t* gflag_host = nullptr;    // pinned copy of gflag[1] for xv_train_skipped_updates
  cudaEvent_t gflag_event = nullptr;
  bool gflag_pending = false;
  uint32_t gflag_seen = 0;
};

namespace {

#define TR_CUDA XV_CUDA

int tr_launch_check(xv_trainer* t) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(XV_ECUDA, std::string("kernel launch: ") + cudaGetErrorName(e) + ": " + cudaGetErrorString(e));
  ++t->last_launches;
  return XV_OK;
}
// bracket a launch with profiling events (option "profile" of the xv_model) and count it
#define TR_BEGIN(name)                                                     \
  do {                                                                     \
    if (!t->emit) break;                                                   \
    if (t->m->opt_profile) t->prof_names.push_back(name);                  \
    int prc_ = prof_mark(t->m, stream);                                    \
    if (prc_ != XV_OK) return prc_;                                        \
  } while (0)
// a stream operation of a training step: skipped while the step's other half is being walked (xv_train_forward_backward_part)
#define TR_EMIT(expr)                                                      \
  do {                                                                     \
    if (t->emit) TR_CUDA(expr);                                            \
  } while (0)
// launch + count; programmatic dependent launch (every kernel starts with cudaGridDependencySynchronize()) unless profiling
#define TR_LAUNCH(name, kernel, grid, block, smem, ...)                                                            \
  do {                                                                                                             \
    TR_BEGIN(name);                                                                                                \
    TR_EMIT(launch_k(t->m->opt_pdl != 0 && !t->m->opt_profile, kernel, grid, block, smem, stream, __VA_ARGS__));   \
    TR_END();                                                                                                      \
  } while (0)
#define TR_END()                                                           \
  do {                                                                     \
    if (!t->emit) break;                                                   \
    int prc_ = prof_mark(t->m, stream);                                    \
    if (prc_ != XV_OK) return prc_;                                        \
    prc_ = tr_launch_check(t);                                             \
    if (prc_ != XV_OK) return prc_;                                        \
  } while (0)

// Work decomposition of one layer's weight gradient.  Layers with temporal context use the tap-reuse kernel (groups of <= 3
// taps whose row span fits the 136-row x slab); k = 1 layers and the spliced first layer use one item per 256 x 256 tile.
struct WgradPlan { bool reuse; int group, n_groups, splits, cps; };
WgradPlan tr_wgrad_plan(const xv_trainer* t, const TrFrame& L, int64_t r_pad) {
  WgradPlan p{};
  const int n_chunks = int(r_pad / wgrad::STAGE_ROWS);
  const int n_mt = (L.c_in_gemm + wgrad::TILE - 1) / wgrad::TILE;
  p.reuse = t->opt_wgrad_reuse && L.gemm_taps > 1 && L.c_out % wgrad::reuse::TILE_N == 0;
  int tiles;
  if (p.reuse) {
    p.group = std::min(L.gemm_taps, wgrad::reuse::MAX_GROUP);
    while (p.group > 1 && (p.group - 1) * L.dil > 8) --p.group;
    p.n_groups = (L.gemm_taps + p.group - 1) / p.group;
    tiles = p.n_groups * n_mt * (L.c_out / wgrad::reuse::TILE_N);
  } else {
    p.group = 1;
    p.n_groups = L.gemm_taps;
    tiles = L.gemm_taps * n_mt * (L.c_out / wgrad::TILE);
  }
  p.splits = std::max(1, std::min(t->m->num_clusters / tiles, n_chunks));
  p.cps = (n_chunks + p.splits - 1) / p.splits;
  p.splits = (n_chunks + p.cps - 1) / p.cps;
  return p;
}

void tr_drop_graphs(xv_trainer* t) {
  for (auto& g : t->graphs) cudaGraphExecDestroy(g.exec);
  t->graphs.clear();
}

// Lays the workspace out for a minibatch geometry.  Egs archives mix minibatch lengths (200..400 frames), so a change of
// geometry must be cheap: the buffer is only re-allocated when it has to grow, nothing is staged from the host (the segment
// metadata is written by train_meta_kernel inside every step) and no buffer relies on having been zeroed.
int tr_ensure_workspace(xv_trainer* t, int32_t n_seg, int32_t seg_len) {
  if (t->ws && t->n_seg == n_seg && t->seg_len == seg_len) return XV_OK;
  xv_model* m = t->m;
  const int32_t stride = int32_t(round_up(int64_t(seg_len) + m->gap, tdnn2::POOL_BLOCK));
  const int64_t r_pad = round_up(int64_t(n_seg) * stride, tdnn2::TILE_ROWS);
  if (r_pad > (int64_t(1) << 31) - 4096) return fail(XV_EINVAL, "minibatch too large");
  const int64_t n_blk = r_pad / 32;
  const int nl = int(t->frames.size());
  const int c_last = t->frames[nl - 1].c_out;
  int c_max = 0;
  for (auto& L : t->frames) c_max = std::max(c_max, L.c_out);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = size_t(round_up(int64_t(off + bytes), 1024)); return o; };
  std::vector<std::pair<void**, size_t>> carve;
  auto want = [&](void** p, size_t bytes) { carve.push_back({p, take(bytes)}); };
  want(reinterpret_cast<void**>(&t->meta), (size_t(4) * n_seg + size_t(4) * n_blk + 8) * 4);
  want(reinterpret_cast<void**>(&t->row_valid), size_t(r_pad));
  want(reinterpret_cast<void**>(&t->blk_valid), size_t(n_blk));
  want(reinterpret_cast<void**>(&t->x0), size_t(r_pad) * m->k0_pad * 2);
  for (int i = 0; i < nl; ++i) {
    TrFrame& L = t->frames[i];
    want(reinterpret_cast<void**>(&L.r), size_t(r_pad) * L.c_out * 2);
    want(reinterpret_cast<void**>(&L.dz), size_t(r_pad) * L.c_out * 2);
    if (i < nl - 1) {
      want(reinterpret_cast<void**>(&L.y), size_t(r_pad) * L.c_out * 2);
      want(reinterpret_cast<void**>(&L.dy), size_t(r_pad) * L.c_out * 2);
    }
  }
  want(reinterpret_cast<void**>(&t->partial), size_t(n_blk) * 2 * c_max * 4);
  want(reinterpret_cast<void**>(&t->partial1), size_t(n_blk) * c_max * 4);
  want(reinterpret_cast<void**>(&t->cA), size_t(c_max) * 4);
  want(reinterpret_cast<void**>(&t->cB), size_t(c_max) * 4);
  want(reinterpret_cast<void**>(&t->cC), size_t(c_max) * 4);
  want(reinterpret_cast<void**>(&t->m_r), size_t(n_seg) * c_last * 4);
  want(reinterpret_cast<void**>(&t->v_r), size_t(n_seg) * c_last * 4);
  want(reinterpret_cast<void**>(&t->coefA), size_t(n_seg) * c_last * 4);
  want(reinterpret_cast<void**>(&t->coefG), size_t(n_seg) * c_last * 4);
  want(reinterpret_cast<void**>(&t->h0), size_t(n_seg) * 2 * c_last * 4);
  want(reinterpret_cast<void**>(&t->dh0), size_t(n_seg) * 2 * c_last * 4);
  for (auto& S : t->seg) {
    for (float** p : {&S.z, &S.r, &S.y, &S.dy, &S.dz}) want(reinterpret_cast<void**>(p), size_t(n_seg) * S.out * 4);
  }
  want(reinterpret_cast<void**>(&t->logits), size_t(n_seg) * t->num_classes * 4);
  want(reinterpret_cast<void**>(&t->dlogits), size_t(n_seg) * t->num_classes * 4);
  want(reinterpret_cast<void**>(&t->loss_row), size_t(n_seg) * 4);
  want(reinterpret_cast<void**>(&t->correct), size_t(n_seg) * 4);
  size_t wg = 0;
  for (auto& L : t->frames) {
    const WgradPlan wp = tr_wgrad_plan(t, L, r_pad);
    wg = std::max(wg, size_t(wp.splits) * L.gemm_taps * L.wg_rows * L.c_out);
  }
  t->wg_partial_floats = wg;
  want(reinterpret_cast<void**>(&t->wg_partial), wg * 4);
  size_t sg = 0;
  {
    const int B = n_seg, K0 = 2 * c_last, E0 = t->seg[0].out, E1 = t->seg[1].out, NC = t->num_classes;
    const int dims[9][3] = {{B, E0, K0}, {B, E1, E0}, {B, NC, E1}, {E1, NC, B}, {B, E1, NC}, {E0, E1, B}, {B, E0, E1}, {K0, E0, B}, {B, K0, E0}};
    for (auto& d : dims) {
      const SgemmPlan sp = sgemm_plan(d[0], d[1], d[2], m->num_sms);
      if (sp.splits > 1) sg = std::max(sg, size_t(sp.splits) * d[0] * d[1]);
    }
  }
  t->sg_partial_floats = sg;
  want(reinterpret_cast<void**>(&t->sg_partial), std::max<size_t>(sg, 1) * 4);
  if (off > t->ws_cap) {
    TR_CUDA(cudaDeviceSynchronize());
    tr_drop_graphs(t);                        // captured steps hold pointers into the old buffer
    cudaFree(t->ws);
    t->ws = nullptr;
    t->ws_cap = 0;
    const size_t want_bytes = off + off / 8;
    TR_CUDA(cudaMalloc(reinterpret_cast<void**>(&t->ws), want_bytes));
    t->ws_cap = want_bytes;
  }
  for (auto& c : carve) *c.first = t->ws + c.second;
  t->n_seg = n_seg;
  t->seg_len = seg_len;
  t->seg_stride = stride;
  t->r_pad = r_pad;
  t->debug.clear();
  {
    const int64_t blk_info_off = round_up(int64_t(3) * n_seg, 4);
    t->blk_info_off = int32_t(blk_info_off);
    t->seg_meta = xvk::SegMeta{t->meta, t->meta + n_seg, t->meta + 2 * n_seg, n_seg};
    t->blk_info_dev = reinterpret_cast<const int4*>(t->meta + blk_info_off);
  }
  return XV_OK;
}

// One frame-level contraction through tdnn_pair_kernel<0, 2, LEAKY> (store mode): out = act(in (*) w + bias)*scale + shift
int tr_pair_layer(xv_trainer* t, cudaStream_t stream, const char* name, const __half* in, int c_in_gemm, __half* out, int c_out,
                  const __half* w, int k_total, int gemm_taps, int dilation, const float* bias, const float* scale,
                  const float* shift, const float* alpha, float* col_partial = nullptr, uint32_t* overflow_flag = nullptr) {
  xv_model* m = t->m;
  const int64_t r_pad = t->r_pad;
  const int halo = (gemm_taps - 1) / 2 * dilation;
  if (halo > tdnn2::TILE_ROWS) return fail(XV_EINVAL, "temporal context wider than one row tile");
  if (c_in_gemm % (2 * tdnn2::BLOCK_K) != 0 || c_out % tdnn2::TILE_CH != 0) return fail(XV_EINVAL, "layer width not supported by the pair kernel");
  const bool reuse = gemm_taps > 1 && halo <= tdnn2::MAX_REUSE_HALO;
  CUtensorMap ta, tw, tc;
  int rc = encode_2d(m, &ta, const_cast<__half*>(in), uint64_t(c_in_gemm), uint64_t(r_pad), tdnn2::BLOCK_K,
                     reuse ? tdnn2::ACT_BOX_ROWS_REUSE : tdnn2::ACT_BOX_ROWS_PLAIN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != XV_OK) return rc;
  rc = encode_2d(m, &tw, const_cast<__half*>(w), uint64_t(k_total), uint64_t(c_out), tdnn2::BLOCK_K, tdnn2::CTA_CH, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != XV_OK) return rc;
  rc = encode_2d(m, &tc, out, uint64_t(c_out), uint64_t(r_pad), tdnn2::C_CHUNK, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc != XV_OK) return rc;
  tdnn2::PairArgs a{};
  a.n_row_tiles = int32_t(r_pad / tdnn2::TILE_ROWS);
  a.n_ch_tiles = c_out / tdnn2::TILE_CH;
  a.taps = gemm_taps;
  a.dilation = dilation;
  a.c_in_pad = c_in_gemm;
  a.reuse = reuse ? 1 : 0;
  a.c_out = c_out;
  a.bias = bias; a.scale = scale; a.shift = shift; a.alpha = alpha;
  a.row_valid = t->row_valid;
  a.blk_valid = t->blk_valid;
  a.partial = col_partial;           // STATS instantiation: per 32-row block column sums of the stored output
  a.overflow_flag = overflow_flag ? overflow_flag : m->overflow_dev;
  a.mode = 0;
  const int64_t cap = tdnn2::RING_BYTES;
  const int64_t act_atom = reuse ? tdnn2::ACT_ATOM_BYTES : tdnn2::ACT_BOX_ROWS_PLAIN * 128;
  if (reuse) {
    a.n_act_stages = 2;
    a.n_wgt_stages = int(std::min<int64_t>(tdnn2::MAX_STAGES, (cap - 2 * 2 * act_atom) / (2 * tdnn2::WGT_ATOM_BYTES)));
  } else {
    a.n_act_stages = a.n_wgt_stages = int(std::min<int64_t>(tdnn2::MAX_STAGES, cap / (2 * (act_atom + tdnn2::WGT_ATOM_BYTES))));
  }
  a.c_chunks = c_in_gemm / (2 * tdnn2::BLOCK_K);
  const int64_t tiles = int64_t(a.n_row_tiles) * a.n_ch_tiles;
  const int grid = 2 * int(std::min<int64_t>(tiles, m->num_clusters));
  const bool pdl = m->opt_pdl != 0 && !m->opt_profile;
  TR_BEGIN(name);
  if (col_partial && alpha) TR_EMIT(launch_k(pdl, tdnn2::tdnn_pair_kernel<0, 2, true, true>, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::SMEM_BYTES, stream, ta, tw, tc, a));
  else if (col_partial) TR_EMIT(launch_k(pdl, tdnn2::tdnn_pair_kernel<0, 2, false, true>, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::SMEM_BYTES, stream, ta, tw, tc, a));
  else if (alpha) TR_EMIT(launch_k(pdl, tdnn2::tdnn_pair_kernel<0, 2, true>, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::SMEM_BYTES, stream, ta, tw, tc, a));
  else TR_EMIT(launch_k(pdl, tdnn2::tdnn_pair_kernel<0, 2, false>, dim3(grid), dim3(tdnn2::NUM_THREADS), tdnn2::SMEM_BYTES, stream, ta, tw, tc, a));
  TR_END();
  return XV_OK;
}

// dW of one frame layer: wgrad_pair_kernel (K-split partials) + fixed-order reduction into grad_w (x 1/S)
int tr_wgrad(xv_trainer* t, cudaStream_t stream, const char* name, const TrFrame& L, const __half* x, const __half* dz, float* grad_w,
             float inv_loss_scale) {
  xv_model* m = t->m;
  const WgradPlan wp = tr_wgrad_plan(t, L, t->r_pad);
  const bool pdl = m->opt_pdl != 0 && !m->opt_profile;
  CUtensorMap tx, tz;
  int rc = encode_2d(m, &tx, const_cast<__half*>(x), uint64_t(L.c_in_gemm), uint64_t(t->r_pad), wgrad::BOX_CH,
                     wp.reuse ? wgrad::reuse::X_BOX_ROWS : wgrad::STAGE_ROWS, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != XV_OK) return rc;
  rc = encode_2d(m, &tz, const_cast<__half*>(dz), uint64_t(L.c_out), uint64_t(t->r_pad), wgrad::BOX_CH, wgrad::STAGE_ROWS, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != XV_OK) return rc;
  struct { int32_t taps, c_in, c_out, k_splits; } a{L.gemm_taps, L.wg_rows, L.c_out, wp.splits};
  const size_t need = size_t(a.k_splits) * a.taps * a.c_in * a.c_out;
  if (need > t->wg_partial_floats) return fail(XV_ESTATE, "wgrad partial buffer too small");
  if (wp.reuse) {
    wgrad::WgradReuseArgs r{};
    r.n_chunks = int32_t(t->r_pad / wgrad::STAGE_ROWS);
    r.chunks_per_split = wp.cps;
    r.k_splits = wp.splits;
    r.taps = L.gemm_taps;
    r.dilation = L.dil;
    r.n_groups = wp.n_groups;
    r.group = wp.group;
    r.n_mt = (L.c_in_gemm + wgrad::TILE - 1) / wgrad::TILE;
    r.n_nt = L.c_out / wgrad::reuse::TILE_N;
    r.c_in = L.wg_rows;
    r.c_out = L.c_out;
    r.partial = t->wg_partial;
    const int items = r.k_splits * r.n_groups * r.n_mt * r.n_nt;
    const int grid = 2 * std::min(items, m->num_clusters);
    TR_BEGIN("wgrad_reuse_kernel");
    TR_EMIT(launch_k(pdl, wgrad::wgrad_reuse_kernel, dim3(grid), dim3(wgrad::NUM_THREADS), wgrad::reuse::SMEM_BYTES, stream, tx, tz, r));
    TR_END();
  } else {
    wgrad::WgradArgs w{};
    w.n_chunks = int32_t(t->r_pad / wgrad::STAGE_ROWS);
    w.k_splits = wp.splits;
    w.chunks_per_split = wp.cps;
    w.taps = L.gemm_taps;
    w.dilation = L.dil;
    w.n_mt = (L.c_in_gemm + wgrad::TILE - 1) / wgrad::TILE;
    w.n_nt = L.c_out / wgrad::TILE;
    w.c_in = L.wg_rows;
    w.c_out = L.c_out;
    w.lbo_bytes = t->opt_wgrad_lbo;
    w.sbo_bytes = t->opt_wgrad_sbo;
    w.partial = t->wg_partial;
    const int items = w.k_splits * w.taps * w.n_mt * w.n_nt;
    const int grid = 2 * std::min(items, m->num_clusters);
    TR_BEGIN(name);
    TR_EMIT(launch_k(pdl, wgrad::wgrad_pair_kernel, dim3(grid), dim3(wgrad::NUM_THREADS), wgrad::SMEM_BYTES, stream, tx, tz, w));
    TR_END();
  }
  const int64_t n4 = int64_t(a.taps) * a.c_in * a.c_out / 4;
  TR_LAUNCH("wgrad_reduce_kernel", wgrad::wgrad_reduce_kernel, dim3(unsigned((n4 + 255) / 256)), dim3(256), 0, t->wg_partial, grad_w, n4, a.k_splits, inv_loss_scale);
  return XV_OK;
}

int tr_sgemm(xv_trainer* t, cudaStream_t stream, const char* name, const float* A, const float* B, float* C, const float* bias,
             int M, int N, int K, int64_t sam, int64_t sak, int64_t sbk, int64_t sbn, int ldc) {
  const SgemmPlan sp = sgemm_plan(M, N, K, t->m->num_sms);
  if (sp.splits > 1 && size_t(sp.splits) * M * N > t->sg_partial_floats) return fail(XV_ESTATE, "sgemm partial buffer too small");
  trk::SgemmArgs a{};
  a.A = A; a.B = B; a.C = C; a.bias = bias;
  a.M = M; a.N = N; a.K = K;
  a.sam = sam; a.sak = sak; a.sbk = sbk; a.sbn = sbn;
  a.ldc = ldc;
  a.k_per_split = sp.k_per_split;
  a.partial = t->sg_partial;
  const dim3 grid((N + 63) / 64, (M + 63) / 64, sp.splits);
  const bool ak = sak == 1, bn = sbn == 1;
  const bool pdl = t->m->opt_pdl != 0 && !t->m->opt_profile;
  TR_BEGIN(name);
  if (ak && bn) TR_EMIT(launch_k(pdl, trk::sgemm64_kernel<true, true>, grid, dim3(256), 0, stream, a));
  else if (!ak && bn) TR_EMIT(launch_k(pdl, trk::sgemm64_kernel<false, true>, grid, dim3(256), 0, stream, a));
  else if (ak && !bn) TR_EMIT(launch_k(pdl, trk::sgemm64_kernel<true, false>, grid, dim3(256), 0, stream, a));
  else return fail(XV_EINVAL, "sgemm: unsupported operand strides");
  TR_END();
  if (sp.splits > 1) {
    const int64_t n = int64_t(M) * N;
    TR_LAUNCH("splitk_reduce_kernel", trk::splitk_reduce_kernel, dim3(unsigned((n + 255) / 256)), dim3(256), 0,
              static_cast<const float*>(t->sg_partial), sp.splits, M, N, bias, C, ldc);
  }
  return XV_OK;
}

// fp32 master weights -> fp16 operand copies of every frame layer
int tr_repack(xv_trainer* t, cudaStream_t stream) {
  trk::RepackTable tb{};
  int tiles = 0;
  tb.n_layers = int(t->frames.size());
  for (size_t i = 0; i < t->frames.size(); ++i) {
    TrFrame& L = t->frames[i];
    trk::RepackLayer& R = tb.layer[i];
    R.W = t->params + L.off_w; R.wf = L.wf; R.wd = i > 0 ? L.wd : nullptr;
    R.taps = L.taps; R.c_in = L.c_in; R.c_out = L.c_out; R.k_total = L.k_total;
    R.c_in_pad = (i == 0) ? L.c_in : L.c_in_gemm;          // first layer: densely spliced K index tap*D + c
    R.tile_begin = tiles;
    tiles += (L.c_out / 32) * ((L.c_in + 31) / 32) * L.taps;
  }
  TR_LAUNCH("repack_kernel", trk::repack_kernel, dim3(tiles), dim3(32, 8), 0, tb);
  t->operands_dirty = false;
  return XV_OK;
}

void tr_add_span(xv_trainer* t, const std::string& name, int32_t which, int64_t offset, std::vector<int64_t> shape) {
  int64_t n = 1;
  for (auto d : shape) n *= d;
  t->spans[name] = TrSpan{which, offset, n, shape};
  t->span_order.push_back(name);
}

float* tr_vec(xv_trainer* t, int32_t which) {
  switch (which) {
    case XV_TRAIN_PARAMS: return t->params;
    case XV_TRAIN_ADAM_M: return t->adam_m;
    case XV_TRAIN_ADAM_V: return t->adam_v;
    case XV_TRAIN_MOVING: return t->moving;
    case XV_TRAIN_GRAD: return t->grad;
    default: return nullptr;
  }
}

}  // namespace

extern "C" {

int xv_train_create(xv_trainer** out, xv_model* model, int32_t num_classes, int32_t emb1_dim) {
  if (!out || !model) return fail(XV_EINVAL, "null argument");
  *out = nullptr;
  if (num_classes < 2 || emb1_dim < 1) return fail(XV_EINVAL, "num_classes must be >= 2 and emb1_dim >= 1");
  if ((model->topo.act != XV_ACT_RELU && model->topo.act != XV_ACT_LRELU) || model->topo.pooling != XV_POOL_STATS)
    return fail(XV_EINVAL, "the training step supports the ReLU / leaky-ReLU topologies with statistics pooling only");
  XV_CUDA(cudaSetDevice(model->device));
  xv_trainer* t = new xv_trainer();
  t->m = model;
  t->num_classes = num_classes;
  const xv_topology& tp = model->topo;
  const int nl = tp.n_frame_layers;
  t->frames.resize(nl);
  int64_t po = 0, mo = 0;
  int c_max = 0;
  for (int i = 0; i < nl; ++i) {
    const FrameLayer& F = model->layers[i];
    TrFrame& L = t->frames[i];
    L.taps = F.taps; L.dil = F.dilation; L.c_in = F.c_in; L.c_out = F.c_out;
    L.k_total = F.k_total; L.gemm_taps = F.gemm_taps;
    L.c_in_gemm = (i == 0) ? F.k_total : F.c_in_pad;
    L.wg_rows = (i == 0) ? F.taps * F.c_in : F.c_in;
    const std::string s = "frame_level_info_layer-" + std::to_string(i) + "/";
    L.off_w = po; tr_add_span(t, s + "w:0", XV_TRAIN_PARAMS, po, {L.taps, L.c_in, L.c_out}); po += int64_t(L.taps) * L.c_in * L.c_out;
    po = round_up(po, 4);
    L.off_b = po; tr_add_span(t, s + "b:0", XV_TRAIN_PARAMS, po, {L.c_out}); po += L.c_out;
    L.off_gamma = po; tr_add_span(t, s + "gamma:0", XV_TRAIN_PARAMS, po, {L.c_out}); po += L.c_out;
    L.off_beta = po; tr_add_span(t, s + "beta:0", XV_TRAIN_PARAMS, po, {L.c_out}); po += L.c_out;
    L.off_mov = mo;
    tr_add_span(t, s + "mean:0", XV_TRAIN_MOVING, mo, {L.c_out}); mo += L.c_out;
    tr_add_span(t, s + "variance:0", XV_TRAIN_MOVING, mo, {L.c_out}); mo += L.c_out;
    c_max = std::max(c_max, L.c_out);
  }
  int prev = 2 * t->frames[nl - 1].c_out;
  const int outs[2] = {tp.emb_dim, emb1_dim};
  for (int i = 0; i < 2; ++i) {
    TrSeg& S = t->seg[i];
    S.in = prev; S.out = outs[i];
    const std::string s = "embed_layer-" + std::to_string(i) + "/";
    S.off_w = po; tr_add_span(t, s + "w:0", XV_TRAIN_PARAMS, po, {S.in, S.out}); po += int64_t(S.in) * S.out;
    po = round_up(po, 4);
    S.off_b = po; tr_add_span(t, s + "b:0", XV_TRAIN_PARAMS, po, {S.out}); po += S.out;
    S.off_gamma = po; tr_add_span(t, s + "gamma:0", XV_TRAIN_PARAMS, po, {S.out}); po += S.out;
    S.off_beta = po; tr_add_span(t, s + "beta:0", XV_TRAIN_PARAMS, po, {S.out}); po += S.out;
    po = round_up(po, 4);
    S.off_mov = mo;
    tr_add_span(t, s + "mean:0", XV_TRAIN_MOVING, mo, {S.out}); mo += S.out;
    tr_add_span(t, s + "variance:0", XV_TRAIN_MOVING, mo, {S.out}); mo += S.out;
    prev = S.out;
    c_max = std::max(c_max, S.out);
  }
  t->off_wo = po; tr_add_span(t, "output/w:0", XV_TRAIN_PARAMS, po, {prev, num_classes}); po += int64_t(prev) * num_classes;
  po = round_up(po, 4);
  t->off_bo = po; tr_add_span(t, "output/b:0", XV_TRAIN_PARAMS, po, {num_classes}); po += num_classes;
  t->n_params = round_up(po, 4);
  t->n_moving = mo;
  cudaError_t e = cudaSuccess;
  auto alloc0 = [&](float** p, int64_t n) {
    if (e != cudaSuccess) return;
    e = cudaMalloc(reinterpret_cast<void**>(p), size_t(n) * 4);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, size_t(n) * 4);
  };
  alloc0(&t->params, t->n_params);
  alloc0(&t->adam_m, t->n_params);
  alloc0(&t->adam_v, t->n_params);
  alloc0(&t->grad, t->n_params + XV_TRAIN_GRAD_TAIL);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&t->gflag), 16);
  if (e == cudaSuccess) e = cudaMemset(t->gflag, 0, 16);
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&t->gflag_host), 4, cudaHostAllocDefault);
  if (e == cudaSuccess) { *t->gflag_host = 0; e = cudaEventCreateWithFlags(&t->gflag_event, cudaEventDisableTiming); }
  alloc0(&t->moving, t->n_moving);
  alloc0(&t->zeros, c_max);
  alloc0(&t->ones, c_max);
  t->neg_slope = model->topo.act == XV_ACT_LRELU ? 0.2f : 0.f;
  alloc0(&t->slope_dev, c_max);
  if (e == cudaSuccess && t->neg_slope != 0.f) {
    trk::fill_kernel<<<(c_max + 255) / 256, 256>>>(t->slope_dev, c_max, t->neg_slope);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) {
    trk::fill_kernel<<<(c_max + 255) / 256, 256>>>(t->ones, c_max, 1.f);
    e = cudaGetLastError();
  }
  for (int i = 0; i < nl && e == cudaSuccess; ++i) {
    TrFrame& L = t->frames[i];
    e = cudaMalloc(reinterpret_cast<void**>(&L.wf), size_t(L.c_out) * L.k_total * 2);
    if (e == cudaSuccess) e = cudaMemset(L.wf, 0, size_t(L.c_out) * L.k_total * 2);     // K padding of the first layer stays zero
    if (e == cudaSuccess && i > 0) e = cudaMalloc(reinterpret_cast<void**>(&L.wd), size_t(L.c_in) * L.taps * L.c_out * 2);
    alloc0(&L.bn, 4 * int64_t(L.c_out));
  }
  for (auto& S : t->seg) alloc0(&S.bn, 2 * int64_t(S.out));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad::wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad::wgrad_reuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wgrad::reuse::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tdnn2::tdnn_pair_kernel<0, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdnn2::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tdnn2::tdnn_pair_kernel<0, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tdnn2::SMEM_BYTES);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    std::string msg = std::string("xv_train_create: ") + cudaGetErrorName(e) + ": " + cudaGetErrorString(e);
    xv_train_destroy(t);
    return fail(XV_ECUDA, msg);
  }
  *out = t;
  return XV_OK;
}

void xv_train_destroy(xv_trainer* t) {
  if (!t) return;
  cudaSetDevice(t->m->device);
  cudaDeviceSynchronize();
  tr_drop_graphs(t);
  if (t->gstream) cudaStreamDestroy(t->gstream);
  if (t->g_in) cudaEventDestroy(t->g_in);
  if (t->g_out) cudaEventDestroy(t->g_out);
  for (auto& L : t->frames) { cudaFree(L.wf); cudaFree(L.wd); cudaFree(L.bn); }
  for (auto& S : t->seg) cudaFree(S.bn);
  cudaFree(t->params); cudaFree(t->adam_m); cudaFree(t->adam_v); cudaFree(t->grad); cudaFree(t->moving);
  cudaFree(t->gflag);
  if (t->gflag_host) cudaFreeHost(t->gflag_host);
  if (t->gflag_event) cudaEventDestroy(t->gflag_event);
  cudaFree(t->ones); cudaFree(t->zeros); cudaFree(t->ws); cudaFree(t->slope_dev);
  delete t;
}

int64_t xv_train_size(const xv_trainer* t, int32_t which) {
  if (!t) return 0;
  if (which == XV_TRAIN_GRAD) return t->n_params + XV_TRAIN_GRAD_TAIL;   // + the combined overflow flag behind the gradient
  return which == XV_TRAIN_MOVING ? t->n_moving : (which >= XV_TRAIN_PARAMS && which < XV_TRAIN_GRAD ? t->n_params : 0);
}

int xv_train_span(const xv_trainer* t, const char* tf_var_name, int32_t* which, int64_t* offset, int64_t* count) {
  if (!t || !tf_var_name || !which || !offset || !count) return fail(XV_EINVAL, "null argument");
  auto it = t->spans.find(tf_var_name);
  if (it == t->spans.end()) return fail(XV_EINVAL, std::string("unknown variable name: ") + tf_var_name);
  *which = it->second.which; *offset = it->second.offset; *count = it->second.count;
  return XV_OK;
}

int xv_train_upload(xv_trainer* t, int32_t which, const float* host, int64_t offset, int64_t count) {
  if (!t || !host) return fail(XV_EINVAL, "null argument");
  float* v = tr_vec(t, which);
  if (!v || offset < 0 || count < 0 || offset + count > xv_train_size(t, which)) return fail(XV_EINVAL, "range outside the flat vector");
  XV_CUDA(cudaSetDevice(t->m->device));
  XV_CUDA(cudaDeviceSynchronize());
  XV_CUDA(cudaMemcpy(v + offset, host, size_t(count) * 4, cudaMemcpyHostToDevice));
  if (which == XV_TRAIN_PARAMS) t->operands_dirty = true;
  return XV_OK;
}

int xv_train_download(xv_trainer* t, int32_t which, float* host, int64_t offset, int64_t count) {
  if (!t || !host) return fail(XV_EINVAL, "null argument");
  float* v = tr_vec(t, which);
  if (!v || offset < 0 || count < 0 || offset + count > xv_train_size(t, which)) return fail(XV_EINVAL, "range outside the flat vector");
  XV_CUDA(cudaSetDevice(t->m->device));
  XV_CUDA(cudaDeviceSynchronize());
  XV_CUDA(cudaMemcpy(host, v + offset, size_t(count) * 4, cudaMemcpyDeviceToHost));
  return XV_OK;
}

int xv_train_set_step(xv_trainer* t, int64_t step) {
  if (!t || step < 0) return fail(XV_EINVAL, "bad argument");
  t->step = step;
  return XV_OK;
}
int64_t xv_train_get_step(const xv_trainer* t) { return t ? t->step : 0; }
int32_t xv_train_last_launch_count(const xv_trainer* t) { return t ? t->last_launches : 0; }

int xv_train_set_option(xv_trainer* t, const char* name, double value) {
  if (!t || !name) return fail(XV_EINVAL, "null argument");
  const std::string n(name);
  tr_drop_graphs(t);                              // options are baked into a captured step
  if (n == "graph") t->opt_graph = value != 0.0;
  else if (n == "loss_scale") t->opt_loss_scale = value;
  else if (n == "wgrad_lbo") t->opt_wgrad_lbo = int(value);
  else if (n == "wgrad_sbo") t->opt_wgrad_sbo = int(value);
  else if (n == "fused_stats") t->opt_fused_stats = value != 0.0;
  else if (n == "l2_beta") t->opt_l2_beta = value;
  else if (n == "wgrad_reuse") { t->opt_wgrad_reuse = value != 0.0; t->n_seg = 0; }     // the partial buffer is re-planned
  else return fail(XV_EINVAL, "unknown option: " + n);
  return XV_OK;
}

}  // extern "C"

namespace {

// training = true: forward (batch statistics, moving-statistics update) + backward.
// training = false: forward only with the moving statistics (phase: False), loss and accuracy (Model.eval, models.py:307-354).
// part: 0 = the whole step; 1 = forward, loss and the segment-level backward (the gradients of embed_layer-*, output/* are final
// when it ends); 2 = pooling and frame-level backward (+ the overflow flag behind the gradient); XV_TRAIN_PART_FRAME + i = the
// slice of 2 that ends with frame layer i's gradients final (the top layer's slice starts with the pooling backward, layer 0's
// ends with the overflow flag).  A data-parallel caller runs 1, starts the all-reduce of the segment-level gradients on another
// stream, then the frame slices from the top layer down, each layer's all-reduce under the backward of the layers below it
// (xv_train_forward_backward_part).
int tr_step_body(xv_trainer* t, const float* feats_dev, const int32_t* labels_dev, int32_t n_seg, int32_t seg_len,
                 float* grad_dev, float* loss_acc_dev, cudaStream_t stream, bool training, int part = 0) {
  xv_model* m = t->m;
  int rc = XV_OK;
  t->last_launches = 0;
  m->prof_used = 0;
  t->prof_names.clear();
  float* grad = grad_dev ? grad_dev : t->grad;
  if (t->operands_dirty) { rc = tr_repack(t, stream); if (rc != XV_OK) return rc; }
  struct EmitGuard { xv_trainer* t; ~EmitGuard() { t->emit = true; } } emit_guard{t};
  const int nl = int(t->frames.size());
  const bool frame_slice = part >= XV_TRAIN_PART_FRAME;
  auto emits = [&](int layer) { return part == 0 || part == 2 || part == XV_TRAIN_PART_FRAME + layer; };   // second half, by frame layer
  t->emit = part != 2 && !frame_slice;
  const int64_t r_pad = t->r_pad;
  constexpr int32_t ROWS_PER_PART = 128;                              // frame layers: rows per partial-sum CTA
  const int32_t n_part = int32_t(r_pad / ROWS_PER_PART);
  const float n_rows = float(int64_t(n_seg) * seg_len);
  double S = t->opt_loss_scale;
  if (S <= 0.0) { S = 1.0; while (S < 8.0 * double(n_rows)) S *= 2.0; }
  t->last_loss_scale = float(S);
  const float inv_S = float(1.0 / S);

  // ---- metadata + pack --------------------------------------------------------------------------
  {
    const int32_t n_blk_all = int32_t(r_pad / 32);
    TR_LAUNCH("train_meta_kernel", trk::train_meta_kernel, dim3((std::max(n_blk_all, n_seg) + 255) / 256), dim3(256), 0, t->meta, n_seg, seg_len,
              t->seg_stride, n_blk_all, t->blk_info_off);
    if (training) {
      // rows past the last segment of the last layer's dz are written by no kernel (pool_relu_bwd works per segment) but read
      // by the weight / data gradient kernels: exact zeros
      const int64_t tail0 = int64_t(n_seg) * t->seg_stride;
      TrFrame& LZ = t->frames[nl - 1];
      if (r_pad > tail0) TR_EMIT(cudaMemsetAsync(LZ.dz + tail0 * LZ.c_out, 0, size_t(r_pad - tail0) * LZ.c_out * 2, stream));
      TR_EMIT(cudaMemsetAsync(t->gflag, 0, 4, stream));          // this step's gradient-overflow flag
    }
    xvk::PackArgs a{};
    a.feats = feats_dev;
    a.r_pad = int32_t(r_pad);
    a.feat_dim = m->topo.feat_dim;
    a.taps = m->layers[0].taps;
    a.dilation = m->layers[0].dilation;
    a.k0_pad = m->k0_pad;
    a.x0 = t->x0;
    a.row_valid = t->row_valid;
    a.blk_valid = t->blk_valid;
    a.blk_info = t->blk_info_dev;
    a.lut = m->pack_lut_dev;
    a.counters = nullptr;
    a.n_counters = 0;
    TR_LAUNCH("pack_im2col_kernel", xvk::pack_im2col_kernel, dim3(unsigned(r_pad / xvk::PACK_ROWS_PER_BLOCK)), dim3(xvk::PACK_THREADS), 0, a);
  }

  // ---- frame layers, training branch of BatchNorm ---------------------------------------------------
  const float* act_alpha = t->neg_slope != 0.f ? t->slope_dev : nullptr;       // leaky topologies: LEAKY layer kernel
  const __half* in = t->x0;
  for (int i = 0; i < nl && !training; ++i) {
    // evaluation branch of BatchNorm (tf_block.py:25-26) folded into the layer kernel's epilogue, as on the extraction path
    TrFrame& L = t->frames[i];
    float* scale = L.bn + 2 * L.c_out;
    float* shift = L.bn + 3 * L.c_out;
    TR_LAUNCH("bn_fold_kernel", trk::bn_fold_kernel, dim3((L.c_out + 255) / 256), dim3(256), 0, t->params + L.off_gamma, t->params + L.off_beta, t->moving + L.off_mov, t->moving + L.off_mov + L.c_out, m->topo.bn_eps, L.c_out, scale, shift);
    __half* out = (i < nl - 1) ? L.y : L.r;
    rc = tr_pair_layer(t, stream, "tdnn_pair_kernel[fwd]", in, L.c_in_gemm, out, L.c_out, L.wf, L.k_total, L.gemm_taps, L.dil,
                       t->params + L.off_b, scale, shift, act_alpha);
    if (rc != XV_OK) return rc;
    in = out;
    if (i == nl - 1) {
      TR_LAUNCH("blk_col_sums_kernel<0>", trk::blk_col_sums_kernel<0>, dim3(L.c_out / trk::COLS_PER_CTA, n_seg), dim3(256), 0, static_cast<const __half*>(out), static_cast<const __half*>(nullptr), L.c_out, t->seg_stride, t->partial);
    }
  }
  for (int i = 0; i < nl && training; ++i) {
    TrFrame& L = t->frames[i];
    // the layer kernel's epilogue leaves the per-32-row-block column sums of r in t->partial (no second pass over HBM)
    rc = tr_pair_layer(t, stream, "tdnn_pair_kernel[fwd]", in, L.c_in_gemm, L.r, L.c_out, L.wf, L.k_total, L.gemm_taps, L.dil,
                       t->params + L.off_b, t->ones, t->zeros, act_alpha, t->opt_fused_stats ? t->partial : nullptr);
    if (rc != XV_OK) return rc;
    const bool last = i == nl - 1;
    int32_t parts = int32_t(r_pad / 32);
    if (!t->opt_fused_stats) {
      // separate pass: the last layer sums per segment (= the pooling sums), the others per 128 rows
      parts = last ? n_seg : n_part;
      TR_LAUNCH("blk_col_sums_kernel<0>", trk::blk_col_sums_kernel<0>, dim3(L.c_out / trk::COLS_PER_CTA, parts), dim3(256), 0, static_cast<const __half*>(L.r), static_cast<const __half*>(nullptr), L.c_out, last ? t->seg_stride : ROWS_PER_PART, t->partial);
    }
    trk::BnFwdArgs b{};
    b.partial = t->partial; b.n_blk = parts; b.C = L.c_out; b.n_rows = n_rows;
    b.eps = m->topo.bn_eps; b.decay = BN_DECAY;
    b.gamma = t->params + L.off_gamma; b.beta = t->params + L.off_beta;
    b.moving_mean = t->moving + L.off_mov; b.moving_var = t->moving + L.off_mov + L.c_out;
    b.mean = L.bn; b.inv = L.bn + L.c_out; b.scale = L.bn + 2 * L.c_out; b.shift = L.bn + 3 * L.c_out;
    TR_LAUNCH("bn_fwd_finalize_kernel", trk::bn_fwd_finalize_kernel, dim3(L.c_out / trk::RED_X), dim3(trk::RED_X, trk::RED_Y), 0, b);
    if (i < nl - 1) {
      const int64_t n8 = r_pad * L.c_out / 8;
      TR_LAUNCH("bn_apply_kernel", trk::bn_apply_kernel, dim3(unsigned((n8 + 1023) / 1024)), dim3(256), 0, L.r, t->row_valid, b.scale, b.shift, n8, L.c_out / 8, L.y, m->overflow_dev);
      in = L.y;
    }
  }
  TrFrame& LL = t->frames[nl - 1];
  const int C = LL.c_out;
  {
    trk::PoolFwdArgs a{};
    a.partial = t->partial; a.C = C; a.n_seg = n_seg; a.blks_per_seg = (training && t->opt_fused_stats) ? t->seg_stride / 32 : 1;
    a.seg_len = float(seg_len); a.var_eps = m->topo.var_eps;
    a.scale = training ? LL.bn + 2 * C : t->ones;           // evaluation: the block sums are already those of y
    a.shift = training ? LL.bn + 3 * C : t->zeros;
    a.m_r = t->m_r; a.v_r = t->v_r; a.h0 = t->h0;
    TR_LAUNCH("pool_train_fwd_kernel", trk::pool_train_fwd_kernel, dim3(dim3((C + 255) / 256, n_seg)), dim3(256), 0, a);
  }

  const int NC = t->num_classes, E1 = t->seg[1].out;
  {
  // ---- segment level forward (fp32) ---------------------------------------------------------------
  const float* hin = t->h0;
  for (int i = 0; i < 2; ++i) {
    TrSeg& Sg = t->seg[i];
    rc = tr_sgemm(t, stream, "sgemm64_kernel[fwd]", hin, t->params + Sg.off_w, Sg.z, t->params + Sg.off_b, n_seg, Sg.out, Sg.in,
                  Sg.in, 1, Sg.out, 1, Sg.out);
    if (rc != XV_OK) return rc;
    trk::SegBnArgs a{};
    a.z = Sg.z; a.B = n_seg; a.C = Sg.out; a.eps = m->topo.bn_eps; a.decay = BN_DECAY;
    a.gamma = t->params + Sg.off_gamma; a.beta = t->params + Sg.off_beta;
    a.moving_mean = t->moving + Sg.off_mov; a.moving_var = t->moving + Sg.off_mov + Sg.out;
    a.r = Sg.r; a.y = Sg.y; a.mean = Sg.bn; a.inv = Sg.bn + Sg.out;
    a.training = training ? 1 : 0;
    a.neg_slope = t->neg_slope;
    TR_LAUNCH("seg_relu_bn_fwd_kernel", trk::seg_relu_bn_fwd_kernel, dim3((Sg.out + 31) / 32), dim3(dim3(32, trk::SEG_Y)), 0, a);
    hin = Sg.y;
  }
  rc = tr_sgemm(t, stream, "sgemm64_kernel[fwd]", t->seg[1].y, t->params + t->off_wo, t->logits, t->params + t->off_bo, n_seg, NC, E1,
                E1, 1, NC, 1, NC);
  if (rc != XV_OK) return rc;
  TR_LAUNCH("softmax_ce_kernel", trk::softmax_ce_kernel, dim3(n_seg), dim3(256), 0, t->logits, labels_dev, NC, 1.f / float(n_seg), t->dlogits, t->loss_row, t->correct);
  TR_LAUNCH("loss_finalize_kernel", trk::loss_finalize_kernel, dim3(1), dim3(32), 0, t->loss_row, t->correct, n_seg, loss_acc_dev);
  if (!training) { t->debug.clear(); return XV_OK; }

  // ---- segment level backward -----------------------------------------------------------------------
  // output layer: dWo = y6^T dlogits, dbo = colsum(dlogits), dy6 = dlogits Wo^T
  rc = tr_sgemm(t, stream, "sgemm64_kernel[dW]", t->seg[1].y, t->dlogits, grad + t->off_wo, nullptr, E1, NC, n_seg, 1, E1, NC, 1, NC);
  if (rc != XV_OK) return rc;
  TR_LAUNCH("colsum_rows_kernel", trk::colsum_rows_kernel, dim3((NC + 255) / 256), dim3(256), 0, t->dlogits, n_seg, NC, grad + t->off_bo);
  rc = tr_sgemm(t, stream, "sgemm64_kernel[dX]", t->dlogits, t->params + t->off_wo, t->seg[1].dy, nullptr, n_seg, E1, NC, NC, 1, 1, NC, E1);
  if (rc != XV_OK) return rc;
  for (int i = 1; i >= 0; --i) {
    TrSeg& Sg = t->seg[i];
    trk::SegBnBwdArgs a{};
    a.dy = Sg.dy; a.r = Sg.r; a.B = n_seg; a.C = Sg.out;
    a.gamma = t->params + Sg.off_gamma; a.mean = Sg.bn; a.inv = Sg.bn + Sg.out;
    a.dz = Sg.dz; a.d_gamma = grad + Sg.off_gamma; a.d_beta = grad + Sg.off_beta; a.d_bias = grad + Sg.off_b;
    a.neg_slope = t->neg_slope;
    TR_LAUNCH("seg_relu_bn_bwd_kernel", trk::seg_relu_bn_bwd_kernel, dim3((Sg.out + 31) / 32), dim3(dim3(32, trk::SEG_Y)), 0, a);
    const float* xin = (i == 1) ? t->seg[0].y : t->h0;
    float* dxin = (i == 1) ? t->seg[0].dy : t->dh0;
    rc = tr_sgemm(t, stream, "sgemm64_kernel[dW]", xin, Sg.dz, grad + Sg.off_w, nullptr, Sg.in, Sg.out, n_seg, 1, Sg.in, Sg.out, 1, Sg.out);
    if (rc != XV_OK) return rc;
    rc = tr_sgemm(t, stream, "sgemm64_kernel[dX]", Sg.dz, t->params + Sg.off_w, dxin, nullptr, n_seg, Sg.in, Sg.out, Sg.out, 1, 1, Sg.out, Sg.in);
    if (rc != XV_OK) return rc;
  }
  }

  if (t->opt_l2_beta != 0.0) {
    // ModelL2Loss* (models.py:930-951,962): loss += beta*(0.1 l2(embed-0 w, b) + l2(embed-1 w, b) + l2(output w, b)), l2 = sum(p^2)/2
    const float beta = float(t->opt_l2_beta);
    struct { int64_t off, n; float coef; } terms[6] = {
        {t->seg[0].off_w, int64_t(t->seg[0].in) * t->seg[0].out, 0.1f * beta}, {t->seg[0].off_b, t->seg[0].out, 0.1f * beta},
        {t->seg[1].off_w, int64_t(t->seg[1].in) * t->seg[1].out, beta},        {t->seg[1].off_b, t->seg[1].out, beta},
        {t->off_wo, int64_t(t->seg[1].out) * t->num_classes, beta},            {t->off_bo, t->num_classes, beta}};
    for (auto& tm : terms)
      TR_LAUNCH("l2_term_kernel", trk::l2_term_kernel, dim3(1), dim3(1024), 0, static_cast<const float*>(t->params + tm.off), grad + tm.off, tm.n, tm.coef, loss_acc_dev);
  }

  if (part == 1) return XV_OK;
  t->emit = emits(nl - 1);

  // ---- pooling + last layer's BatchNorm + ReLU backward ---------------------------------------------------
  {
    trk::PoolBwdArgs a{};
    a.dh0 = t->dh0; a.m_r = t->m_r; a.v_r = t->v_r; a.C = C; a.n_seg = n_seg;
    a.seg_len = float(seg_len); a.var_eps = m->topo.var_eps; a.loss_scale = float(S);
    a.gamma = t->params + LL.off_gamma; a.mean = LL.bn; a.inv = LL.bn + C; a.scale = LL.bn + 2 * C;
    a.coefA = t->coefA; a.coefG = t->coefG; a.d_gamma = grad + LL.off_gamma; a.d_beta = grad + LL.off_beta;
    TR_LAUNCH("pool_bwd_coef_kernel", trk::pool_bwd_coef_kernel, dim3((C + 31) / 32), dim3(dim3(32, trk::SEG_Y)), 0, a);
    TR_LAUNCH("pool_relu_bwd_kernel", trk::pool_relu_bwd_kernel, dim3(C / trk::COLS_PER_CTA, n_seg), dim3(256), 0, static_cast<const __half*>(LL.r), C, t->seg_stride, seg_len, static_cast<const float*>(t->coefA), static_cast<const float*>(t->coefG), LL.dz, t->partial1, t->gflag, t->neg_slope);
    TR_LAUNCH("colsum_finalize_kernel", trk::colsum_finalize_kernel, dim3(C / trk::RED_X), dim3(trk::RED_X, trk::RED_Y), 0, static_cast<const float*>(t->partial1), n_seg, C, inv_S, grad + LL.off_b);
  }

  // ---- frame layers backward ---------------------------------------------------------------------------
  for (int i = nl - 1; i >= 0; --i) {
    TrFrame& L = t->frames[i];
    t->emit = emits(i);
    if (i < nl - 1) {
      // dy_i (written by the data gradient of layer i+1) -> BatchNorm + ReLU backward -> dz_i
      TR_LAUNCH("blk_col_sums_kernel<1>", trk::blk_col_sums_kernel<1>, dim3(L.c_out / trk::COLS_PER_CTA, n_part), dim3(256), 0, static_cast<const __half*>(L.dy), static_cast<const __half*>(L.r), L.c_out, ROWS_PER_PART, t->partial);
      trk::BnBwdArgs b{};
      b.partial = t->partial; b.n_blk = n_part; b.C = L.c_out; b.n_rows = n_rows; b.inv_loss_scale = inv_S;
      b.gamma = t->params + L.off_gamma; b.mean = L.bn; b.inv = L.bn + L.c_out;
      b.cA = t->cA; b.cB = t->cB; b.cC = t->cC;
      b.d_gamma = grad + L.off_gamma; b.d_beta = grad + L.off_beta;
      TR_LAUNCH("bn_bwd_finalize_kernel", trk::bn_bwd_finalize_kernel, dim3(L.c_out / trk::RED_X), dim3(trk::RED_X, trk::RED_Y), 0, b);
      TR_LAUNCH("bn_relu_bwd_kernel", trk::bn_relu_bwd_kernel, dim3(L.c_out / trk::COLS_PER_CTA, n_part), dim3(256), 0, static_cast<const __half*>(L.dy), static_cast<const __half*>(L.r), L.c_out, ROWS_PER_PART, static_cast<const float*>(t->cA), static_cast<const float*>(t->cB), static_cast<const float*>(t->cC), L.dz, t->partial1, t->gflag, t->neg_slope, static_cast<const uint8_t*>(t->row_valid));
      TR_LAUNCH("colsum_finalize_kernel", trk::colsum_finalize_kernel, dim3(L.c_out / trk::RED_X), dim3(trk::RED_X, trk::RED_Y), 0, static_cast<const float*>(t->partial1), n_part, L.c_out, inv_S, grad + L.off_b);
    }
    const __half* x = (i == 0) ? t->x0 : t->frames[i - 1].y;
    rc = tr_wgrad(t, stream, "wgrad_pair_kernel", L, x, L.dz, grad + L.off_w, inv_S);
    if (rc != XV_OK) return rc;
    if (i > 0) {
      TrFrame& P = t->frames[i - 1];
      rc = tr_pair_layer(t, stream, "tdnn_pair_kernel[dgrad]", L.dz, L.c_out, P.dy, L.c_in, L.wd, L.taps * L.c_out, L.taps, L.dil,
                         t->zeros, t->ones, t->zeros, t->ones, nullptr, t->gflag);
      if (rc != XV_OK) return rc;
    }
  }
  // the step's gradient-overflow flag rides behind the gradient (grad[n_params]): summed with it by the all-reduce
  t->emit = emits(0);
  TR_LAUNCH("grad_flag_kernel", trk::grad_flag_kernel, dim3(1), dim3(32), 0, static_cast<const uint32_t*>(t->gflag), grad + t->n_params);

  // parity hooks
  t->debug.clear();
  for (int i = 0; i < nl; ++i) {
    TrFrame& L = t->frames[i];
    const std::string k = std::to_string(i);
    t->debug["r" + k] = TrDebug{L.r, L.c_out, 0, 0};
    t->debug["dz" + k] = TrDebug{L.dz, L.c_out, 0, 0};
    if (i < nl - 1) { t->debug["y" + k] = TrDebug{L.y, L.c_out, 0, 0}; t->debug["dy" + k] = TrDebug{L.dy, L.c_out, 0, 0}; }
  }
  t->debug["x0"] = TrDebug{t->x0, m->k0_pad, 0, 0};
  t->debug["h0"] = TrDebug{t->h0, 0, 1, int64_t(n_seg) * 2 * C};
  t->debug["dh0"] = TrDebug{t->dh0, 0, 1, int64_t(n_seg) * 2 * C};
  t->debug["z5"] = TrDebug{t->seg[0].z, 0, 1, int64_t(n_seg) * t->seg[0].out};
  t->debug["y5"] = TrDebug{t->seg[0].y, 0, 1, int64_t(n_seg) * t->seg[0].out};
  t->debug["z6"] = TrDebug{t->seg[1].z, 0, 1, int64_t(n_seg) * t->seg[1].out};
  t->debug["y6"] = TrDebug{t->seg[1].y, 0, 1, int64_t(n_seg) * t->seg[1].out};
  t->debug["logits"] = TrDebug{t->logits, 0, 1, int64_t(n_seg) * NC};
  t->debug["dlogits"] = TrDebug{t->dlogits, 0, 1, int64_t(n_seg) * NC};
  return XV_OK;
}


// One step: plain stream launches, or (training, option "graph", not profiling) a CUDA graph captured on the first call for
// this (geometry, buffers) and replayed afterwards -- the host then enqueues one graph instead of 74 kernels (0.87 ms of host
// time per step otherwise, about what the GPU needs for the step itself).
int tr_step(xv_trainer* t, const float* feats_dev, const int32_t* labels_dev, int32_t n_seg, int32_t seg_len,
            float* grad_dev, float* loss_acc_dev, void* stream_, bool training, int part = 0) {
  if (!t || !feats_dev || !labels_dev || !loss_acc_dev) return fail(XV_EINVAL, "null argument");
  if (n_seg < 1 || seg_len < 1) return fail(XV_EINVAL, "need n_seg >= 1 and seg_len >= 1");
  xv_model* m = t->m;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XV_CUDA(cudaSetDevice(m->device));
  int rc = tr_ensure_workspace(t, n_seg, seg_len);
  if (rc != XV_OK) return rc;
  if (!training || !t->opt_graph || m->opt_profile)
    return tr_step_body(t, feats_dev, labels_dev, n_seg, seg_len, grad_dev, loss_acc_dev, stream, training, part);
  if (t->operands_dirty) { rc = tr_repack(t, stream); if (rc != XV_OK) return rc; }        // never part of the graph
  if (!t->gstream) {
    XV_CUDA(cudaStreamCreateWithFlags(&t->gstream, cudaStreamNonBlocking));
    XV_CUDA(cudaEventCreateWithFlags(&t->g_in, cudaEventDisableTiming));
    XV_CUDA(cudaEventCreateWithFlags(&t->g_out, cudaEventDisableTiming));
  }
  const void* gkey = grad_dev ? static_cast<const void*>(grad_dev) : static_cast<const void*>(t->grad);
  xv_trainer::StepGraph* found = nullptr;
  for (auto& g : t->graphs)
    if (g.feats == feats_dev && g.labels == labels_dev && g.grad == gkey && g.loss == loss_acc_dev && g.n_seg == n_seg && g.seg_len == seg_len &&
        g.part == part)
      found = &g;
  if (!found) {
    // egs archives mix minibatch lengths: capturing costs about two steps, so only a geometry that comes back is captured
    const std::pair<int32_t, int32_t> geo(n_seg, seg_len * 64 + part);
    if (std::find(t->seen_geometries.begin(), t->seen_geometries.end(), geo) == t->seen_geometries.end()) {
      if (t->seen_geometries.size() >= 4096) t->seen_geometries.clear();
      t->seen_geometries.push_back(geo);
      return tr_step_body(t, feats_dev, labels_dev, n_seg, seg_len, grad_dev, loss_acc_dev, stream, true, part);
    }
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(t->gstream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      rc = tr_step_body(t, feats_dev, labels_dev, n_seg, seg_len, grad_dev, loss_acc_dev, t->gstream, true, part);
      e = cudaStreamEndCapture(t->gstream, &graph);
    }
    cudaGraphExec_t exec = nullptr;
    if (e == cudaSuccess && rc == XV_OK) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || rc != XV_OK || !exec) {
      (void)cudaGetLastError();                  // capture not possible here: fall back to plain launches for good
      t->opt_graph = 0;
      return tr_step_body(t, feats_dev, labels_dev, n_seg, seg_len, grad_dev, loss_acc_dev, stream, true, part);
    }
    if (t->graphs.size() >= 64) tr_drop_graphs(t);
    t->graphs.push_back(xv_trainer::StepGraph{feats_dev, labels_dev, gkey, loss_acc_dev, n_seg, seg_len, exec, t->last_launches, part});
    found = &t->graphs.back();
  }
  // the caller's stream order is kept: its earlier work -> graph -> its later work
  XV_CUDA(cudaEventRecord(t->g_in, stream));
  XV_CUDA(cudaStreamWaitEvent(t->gstream, t->g_in, 0));
  XV_CUDA(cudaGraphLaunch(found->exec, t->gstream));
  XV_CUDA(cudaEventRecord(t->g_out, t->gstream));
  XV_CUDA(cudaStreamWaitEvent(stream, t->g_out, 0));
  t->last_launches = found->launches;
  return XV_OK;
}
}  // namespace

extern "C" {

int xv_train_forward_backward(xv_trainer* t, const float* feats_dev, const int32_t* labels_dev, int32_t n_seg, int32_t seg_len,
                              float* grad_dev, float* loss_acc_dev, void* stream) {
  return tr_step(t, feats_dev, labels_dev, n_seg, seg_len, grad_dev, loss_acc_dev, stream, true);
}

int xv_train_forward_backward_part(xv_trainer* t, const float* feats_dev, const int32_t* labels_dev, int32_t n_seg, int32_t seg_len,
                                   float* grad_dev, float* loss_acc_dev, void* stream, int32_t part) {
  if (!t) return fail(XV_EINVAL, "null argument");
  if (!(part >= 0 && part <= 2) && !(part >= XV_TRAIN_PART_FRAME && part < XV_TRAIN_PART_FRAME + int32_t(t->frames.size())))
    return fail(XV_EINVAL, "part must be 0, 1, 2 or XV_TRAIN_PART_FRAME + frame layer");
  return tr_step(t, feats_dev, labels_dev, n_seg, seg_len, grad_dev, loss_acc_dev, stream, true, part);
}

int64_t xv_train_segment_grad_offset(const xv_trainer* t) { return t ? t->seg[0].off_w : 0; }

int xv_train_frame_grad_span(const xv_trainer* t, int32_t layer, int64_t* offset, int64_t* count) {
  if (!t || !offset || !count) return fail(XV_EINVAL, "null argument");
  if (layer < 0 || layer >= int32_t(t->frames.size())) return fail(XV_EINVAL, "no such frame layer");
  const TrFrame& L = t->frames[layer];
  *offset = L.off_w;                                    // w | b | gamma | beta of a layer lie behind one another
  *count = L.off_beta + L.c_out - L.off_w;
  return XV_OK;
}

int xv_train_eval(xv_trainer* t, const float* feats_dev, const int32_t* labels_dev, int32_t n_seg, int32_t seg_len,
                  float* loss_acc_dev, void* stream) {
  return tr_step(t, feats_dev, labels_dev, n_seg, seg_len, nullptr, loss_acc_dev, stream, false);
}

int xv_train_apply(xv_trainer* t, const float* grad_dev, float learning_rate, float grad_scale, void* stream_) {
  if (!t) return fail(XV_EINVAL, "null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XV_CUDA(cudaSetDevice(t->m->device));
  const float* g = grad_dev ? grad_dev : t->grad;
  t->step += 1;
  const double b1t = std::pow(double(ADAM_B1), double(t->step)), b2t = std::pow(double(ADAM_B2), double(t->step));
  const float lr_t = float(double(learning_rate) * std::sqrt(1.0 - b2t) / (1.0 - b1t));
  TR_LAUNCH("adam_kernel", trk::adam_kernel, dim3(unsigned((t->n_params + 255) / 256)), dim3(256), 0, t->params, g, t->adam_m, t->adam_v, t->n_params, lr_t, ADAM_B1, ADAM_B2, ADAM_EPS, grad_scale,
            g + t->n_params, t->gflag + 1);
  return tr_repack(t, stream);
}

int64_t xv_train_skipped_updates(xv_trainer* t, void* stream_, int32_t blocking) {
  if (!t) return fail(XV_EINVAL, "null argument");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  XV_CUDA(cudaSetDevice(t->m->device));
  if (blocking) {
    XV_CUDA(cudaMemcpyAsync(t->gflag_host, t->gflag + 1, 4, cudaMemcpyDeviceToHost, stream));
    XV_CUDA(cudaStreamSynchronize(stream));
    t->gflag_pending = false;
    t->gflag_seen = *t->gflag_host;
    return int64_t(t->gflag_seen);
  }
  // non-blocking poll: harvest the previous read-back if it has landed, then start the next one
  if (t->gflag_pending) {
    cudaError_t q = cudaEventQuery(t->gflag_event);
    if (q == cudaErrorNotReady) return int64_t(t->gflag_seen);
    XV_CUDA(q);
    t->gflag_seen = *t->gflag_host;
    t->gflag_pending = false;
  }
  XV_CUDA(cudaMemcpyAsync(t->gflag_host, t->gflag + 1, 4, cudaMemcpyDeviceToHost, stream));
  XV_CUDA(cudaEventRecord(t->gflag_event, stream));
  t->gflag_pending = true;
  return int64_t(t->gflag_seen);
}

int xv_train_sync_model(xv_trainer* t) {
  if (!t) return fail(XV_EINVAL, "null argument");
  XV_CUDA(cudaSetDevice(t->m->device));
  XV_CUDA(cudaDeviceSynchronize());
  std::vector<float> p(size_t(t->n_params)), mv(size_t(t->n_moving));
  XV_CUDA(cudaMemcpy(p.data(), t->params, p.size() * 4, cudaMemcpyDeviceToHost));
  XV_CUDA(cudaMemcpy(mv.data(), t->moving, mv.size() * 4, cudaMemcpyDeviceToHost));
  for (const auto& name : t->span_order) {
    const TrSpan& s = t->spans.at(name);
    const float* src = (s.which == XV_TRAIN_MOVING ? mv.data() : p.data()) + s.offset;
    int rc = xv_set_param(t->m, name.c_str(), src, s.shape.data(), int32_t(s.shape.size()));
    if (rc != XV_OK) return rc;
  }
  return XV_OK;
}

int xv_convert_f16_to_f32(const void* src_dev, float* dst_dev, int64_t n, void* stream_) {
  if (!src_dev || !dst_dev || n < 0) return fail(XV_EINVAL, "bad argument");
  if ((reinterpret_cast<uintptr_t>(src_dev) & 15) || (reinterpret_cast<uintptr_t>(dst_dev) & 15)) return fail(XV_EINVAL, "buffers must be 16-byte aligned");
  if (n == 0) return XV_OK;
  trk::f16_to_f32_kernel<<<unsigned((n / 8 + 256) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<const __half*>(src_dev), dst_dev, n);
  XV_CUDA(cudaGetLastError());
  return XV_OK;
}

int64_t xv_train_last_kernel_names(const xv_trainer* t, char* buf, int64_t capacity) {
  if (!t || !buf || capacity < 1) return fail(XV_EINVAL, "bad argument");
  std::string all;
  for (const auto& n : t->prof_names) { all += n; all += ';'; }
  if (int64_t(all.size()) + 1 > capacity) return fail(XV_ENOMEM, "name buffer too small");
  std::memcpy(buf, all.c_str(), all.size() + 1);
  return int64_t(t->prof_names.size());
}

int64_t xv_train_debug_tensor(xv_trainer* t, const char* name, float* host_out, int64_t capacity) {
  if (!t || !name || !host_out) return fail(XV_EINVAL, "null argument");
  auto it = t->debug.find(name);
  if (it == t->debug.end()) return fail(XV_EINVAL, std::string("no such intermediate: ") + name);
  XV_CUDA(cudaSetDevice(t->m->device));
  XV_CUDA(cudaDeviceSynchronize());
  const TrDebug& d = it->second;
  if (d.kind == 1) {
    if (capacity < d.count) return fail(XV_ENOMEM, "host buffer too small");
    XV_CUDA(cudaMemcpy(host_out, d.ptr, size_t(d.count) * 4, cudaMemcpyDeviceToHost));
    return d.count;
  }
  const int64_t n = int64_t(t->n_seg) * t->seg_len * d.cols;
  if (capacity < n) return fail(XV_ENOMEM, "host buffer too small");
  float* tmp = nullptr;
  XV_CUDA(cudaMalloc(reinterpret_cast<void**>(&tmp), size_t(n) * 4));
  xvk::unpack_rows_kernel<<<t->n_seg, 256>>>(static_cast<const __half*>(d.ptr), t->seg_meta, d.cols, tmp, 1.0f);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(host_out, tmp, size_t(n) * 4, cudaMemcpyDeviceToHost);
  cudaFree(tmp);
  if (e != cudaSuccess) return fail(XV_ECUDA, std::string("debug tensor: ") + cudaGetErrorString(e));
  return n;
}

}  // extern "C"
