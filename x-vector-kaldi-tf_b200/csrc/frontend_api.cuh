// frontend_api.cuh -- host side of include/xvec_frontend.h (included at the end of xvec_api.cu: shares xv_model, the
// pinned metadata ring, launch_k and forward_impl).
#pragma once
#include "../../include/xvec_frontend.h"
#include "frontend.cuh"

namespace {

struct FrontendPlan {
  int64_t total_rows = 0, total_keep = 0;
  int32_t n_tiles = 0;
  size_t off_meta = 0, off_tile_cnt = 0, bytes = 0;
};

inline int64_t fe_tiles_upper_bound(int64_t total_rows, int32_t n_utt) {
  return total_rows / xvfe::TILE_MAX + n_utt;  // ceil(len / TILE_MAX) tiles per utterance
}

inline FrontendPlan fe_plan(int64_t total_rows, int32_t n_utt) {
  FrontendPlan p;
  p.total_rows = total_rows;
  p.off_meta = 0;
  p.off_tile_cnt = size_t(round_up((int64_t(6) * n_utt + 1 + fe_tiles_upper_bound(total_rows, n_utt)) * 4, 1024));
  p.bytes = p.off_tile_cnt + size_t(round_up(fe_tiles_upper_bound(total_rows, n_utt) * 4, 1024));
  return p;
}

int fe_check_opts(const xv_model* m, const xv_cmvn_opts* o) {
  if (!o) return fail(XV_EINVAL, "null cmvn options");
  if (o->cmn_window < 1) return fail(XV_EINVAL, "cmn_window must be >= 1");
  if (o->min_window < 0 || o->min_window > o->cmn_window) return fail(XV_EINVAL, "min_window must be in [0, cmn_window]");
  (void)m;
  return XV_OK;
}

// Frames per tile of an utterance: the whole utterance when it fits one CTA, else an even split; multiple of 32.
inline int32_t fe_tile_rows(int32_t len) {
  if (len <= 0) return 32;
  const int32_t nt = (len + xvfe::TILE_MAX - 1) / xvfe::TILE_MAX;
  return int32_t(round_up((len + nt - 1) / nt, 32));
}

int frontend_impl(xv_model* m, const float* feats_dev, const float* vad_dev, const int32_t* utt_len_host,
                  const int32_t* out_keep_host, int32_t n_utt, const xv_cmvn_opts* opts, float* out_dev, void* workspace_dev,
                  size_t workspace_bytes, cudaStream_t stream, int64_t* total_keep_out) {
  if (!m || !feats_dev || !utt_len_host || !out_dev || !workspace_dev) return fail(XV_EINVAL, "null argument");
  if (n_utt <= 0) return fail(XV_EINVAL, "n_utt must be >= 1");
  if (vad_dev && !out_keep_host) return fail(XV_EINVAL, "out_keep_host is required when a VAD track is given");
  int rc = fe_check_opts(m, opts);
  if (rc != XV_OK) return rc;
  XV_CUDA(cudaSetDevice(m->device));
  int64_t total_rows = 0;
  for (int i = 0; i < n_utt; ++i) {
    if (utt_len_host[i] < 0) return fail(XV_EINVAL, "utterance " + std::to_string(i) + " has negative length");
    const int32_t keep = out_keep_host ? out_keep_host[i] : utt_len_host[i];
    if (keep < 0 || keep > utt_len_host[i])
      return fail(XV_EINVAL, "utterance " + std::to_string(i) + ": out_keep must be in [0, utt_len]");
    total_rows += utt_len_host[i];
  }
  if (total_rows >= (int64_t(1) << 31) - 4096)
    return fail(XV_EINVAL, "batch too large: raw rows exceed int32 range");
  const FrontendPlan p = fe_plan(total_rows, n_utt);
  if (workspace_bytes < p.bytes)
    return fail(XV_ENOMEM, "frontend workspace too small: need " + std::to_string(p.bytes) + " bytes, got " + std::to_string(workspace_bytes));
  if (reinterpret_cast<uintptr_t>(workspace_dev) % 16 != 0) return fail(XV_EINVAL, "frontend workspace must be 16-byte aligned");

  // ---- utterance table, built in a pinned staging slot:
  //      [in_row0 | len | out_row0 | keep | tile0 (n_utt + 1) | tile_rows | tile_utt (n_tiles)] ----
  rc = ensure_meta_capacity(m, int64_t(6) * n_utt + 1 + fe_tiles_upper_bound(total_rows, n_utt));
  if (rc != XV_OK) return rc;
  const int slot = m->meta_next;
  m->meta_next = (m->meta_next + 1) % META_SLOTS;
  XV_CUDA(cudaEventSynchronize(m->meta_event[slot]));
  int32_t* mh = m->meta_host[slot];
  int64_t in_row = 0, out_row = 0, tile = 0;
  int32_t longest = 0;
  xvfe::SmemPlan sp{1, 32};
  int32_t* tile_rows = mh + 5 * int64_t(n_utt) + 1;
  int32_t* tile_utt = mh + 6 * int64_t(n_utt) + 1;
  for (int i = 0; i < n_utt; ++i) {
    const int32_t len = utt_len_host[i];
    const int32_t keep = out_keep_host ? out_keep_host[i] : len;
    const int32_t rp = fe_tile_rows(len);
    const int32_t nt = (len + rp - 1) / rp;
    longest = std::max(longest, len);
    for (int32_t k = 0; k < nt; ++k) tile_utt[tile + k] = i;
    if (len > 0) {
      sp.rows_cap = std::max(sp.rows_cap, rp);
      sp.slab_rows_cap = std::max(sp.slab_rows_cap, std::min(len, rp + opts->cmn_window));   // every window lies inside the utterance
    }
    mh[i] = int32_t(in_row);
    mh[n_utt + i] = len;
    mh[2 * n_utt + i] = int32_t(out_row);
    mh[3 * n_utt + i] = keep;
    mh[4 * n_utt + i] = int32_t(tile);
    tile_rows[i] = rp;
    in_row += len;
    out_row += keep;
    tile += nt;
  }
  mh[5 * n_utt] = int32_t(tile);
  if (total_keep_out) *total_keep_out = out_row;
  m->last_frontend_launches = 0;
  if (tile == 0) return XV_OK;                                    // nothing but empty utterances
  uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
  int32_t* meta_dev = reinterpret_cast<int32_t*>(ws + p.off_meta);
  int32_t* tile_cnt = reinterpret_cast<int32_t*>(ws + p.off_tile_cnt);
  const size_t smem = xvfe::cmvn_smem_bytes(sp, m->topo.feat_dim, opts->normalize_variance != 0);
  if (smem > size_t(226) * 1024)
    return fail(XV_EINVAL, "cmn_window * feat_dim too large for one CTA's shared memory (" + std::to_string(smem) + " bytes)");
  XV_CUDA(cudaMemcpyAsync(meta_dev, mh, (size_t(6) * n_utt + 1 + size_t(tile)) * 4, cudaMemcpyHostToDevice, stream));
  XV_CUDA(cudaEventRecord(m->meta_event[slot], stream));

  xvfe::UttMeta um{meta_dev, meta_dev + n_utt, meta_dev + 2 * n_utt, meta_dev + 3 * n_utt, meta_dev + 4 * n_utt,
                   meta_dev + 6 * int64_t(n_utt) + 1, meta_dev + 5 * int64_t(n_utt) + 1, n_utt};
  const bool count_pass = vad_dev != nullptr && longest > xvfe::DIRECT_COUNT_MAX;
  const xvfe::CmvnOpts o{opts->cmn_window, opts->min_window, opts->center != 0, opts->normalize_variance != 0};
  const bool pdl = m->opt_pdl != 0;
  if (count_pass) {
    XV_CUDA(launch_k(pdl, xvfe::vad_tile_count_kernel, dim3(unsigned(tile)), dim3(xvfe::COUNT_THREADS), 0, stream, um, vad_dev, tile_cnt));
    ++m->last_frontend_launches;
  }
  auto kernel = o.normalize_variance ? xvfe::cmvn_select_kernel<true> : xvfe::cmvn_select_kernel<false>;
  size_t& opted = o.normalize_variance ? m->fe_smem_opted_nv : m->fe_smem_opted;
  if (smem > size_t(48) * 1024 && smem > opted) {
    XV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    opted = smem;
  }
  XV_CUDA(launch_k(pdl, kernel, dim3(unsigned(tile)), dim3(xvfe::THREADS), smem, stream, um, o, sp, int32_t(m->topo.feat_dim),
                   feats_dev, vad_dev, static_cast<const int32_t*>(count_pass ? tile_cnt : nullptr), out_dev, m->cur_flag));
  ++m->last_frontend_launches;
  XV_CUDA(cudaGetLastError());
  return XV_OK;
}

}  // namespace

extern "C" {

size_t xv_frontend_workspace_bytes(const xv_model* m, int64_t total_rows, int32_t n_utt) {
  if (!m || total_rows < 0 || n_utt <= 0) return 0;
  return fe_plan(total_rows, n_utt).bytes;
}

int xv_frontend(xv_model* m, const float* feats_dev, const float* vad_dev, const int32_t* utt_len_host,
                const int32_t* out_keep_host, int32_t n_utt, const xv_cmvn_opts* opts, float* out_dev, void* workspace_dev,
                size_t workspace_bytes, void* stream) {
  return frontend_impl(m, feats_dev, vad_dev, utt_len_host, out_keep_host, n_utt, opts, out_dev, workspace_dev,
                       workspace_bytes, static_cast<cudaStream_t>(stream), nullptr);
}

int xv_submit_host_raw(xv_model* m, const float* feats_host, const float* vad_host, const int32_t* utt_len_host,
                       const int32_t* out_keep_host, int32_t n_utt, const xv_cmvn_opts* opts, const int32_t* seg_len_host,
                       int32_t n_seg, float* emb_host, int32_t* ticket) {
  if (!m || !feats_host || !utt_len_host || !seg_len_host || !emb_host || !ticket) return fail(XV_EINVAL, "null argument");
  if (vad_host && !out_keep_host) return fail(XV_EINVAL, "out_keep_host is required when a VAD track is given");
  if (n_utt <= 0 || n_seg <= 0) return fail(XV_EINVAL, "n_utt and n_seg must be >= 1");
  XV_CUDA(cudaSetDevice(m->device));
  int64_t raw_rows = 0, kept = 0, seg_rows = 0;
  for (int i = 0; i < n_utt; ++i) {
    if (utt_len_host[i] < 0) return fail(XV_EINVAL, "utterance " + std::to_string(i) + " has negative length");
    raw_rows += utt_len_host[i];
    kept += out_keep_host ? out_keep_host[i] : utt_len_host[i];
  }
  for (int i = 0; i < n_seg; ++i) {
    if (seg_len_host[i] <= 0) return fail(XV_EINVAL, "segment " + std::to_string(i) + " has non-positive length");
    seg_rows += seg_len_host[i];
  }
  if (seg_rows != kept)
    return fail(XV_EINVAL, "segments cover " + std::to_string(seg_rows) + " rows but the front end keeps " + std::to_string(kept));
  const int si = m->slot_next;
  xv_model::HostSlot& sl = m->slots[si];
  if (sl.busy) return fail(XV_ESTATE, "all submission slots are in flight: xv_collect the oldest ticket first");
  const int D = m->topo.feat_dim;
  const size_t raw_bytes = size_t(raw_rows) * D * 4, vad_bytes = size_t(raw_rows) * 4;
  const size_t feat_bytes = size_t(kept) * D * 4;
  const size_t emb_bytes = size_t(n_seg) * m->topo.emb_dim * 4;
  const size_t ws_bytes = xv_workspace_bytes(m, kept, n_seg);
  const size_t fe_bytes = xv_frontend_workspace_bytes(m, raw_rows, n_utt);
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
  XV_CUDA(grow(reinterpret_cast<void**>(&sl.raw_dev), &sl.raw_cap, raw_bytes));
  if (vad_host) XV_CUDA(grow(reinterpret_cast<void**>(&sl.vad_dev), &sl.vad_cap, vad_bytes));
  XV_CUDA(grow(&sl.fe_ws_dev, &sl.fe_ws_cap, fe_bytes));
  XV_CUDA(grow(reinterpret_cast<void**>(&sl.feats_dev), &sl.feats_cap, feat_bytes));
  XV_CUDA(grow(reinterpret_cast<void**>(&sl.emb_dev), &sl.emb_cap, emb_bytes));
  XV_CUDA(grow(&sl.ws_dev, &sl.ws_cap, ws_bytes));
  XV_CUDA(cudaMemcpyAsync(sl.raw_dev, feats_host, raw_bytes, cudaMemcpyHostToDevice, sl.stream));
  if (vad_host) XV_CUDA(cudaMemcpyAsync(sl.vad_dev, vad_host, vad_bytes, cudaMemcpyHostToDevice, sl.stream));
  m->cur_flag = m->overflow_dev + 1 + si;          // this submission's flag word (front-end mismatch bit, fp16 store overflow bits)
  int rc = frontend_impl(m, sl.raw_dev, vad_host ? sl.vad_dev : nullptr, utt_len_host, out_keep_host, n_utt, opts, sl.feats_dev,
                         sl.fe_ws_dev, sl.fe_ws_cap, sl.stream, nullptr);
  m->cur_flag = m->overflow_dev;
  if (rc != XV_OK) return rc;
  const int fe_launches = m->last_frontend_launches;
  // the network part is remembered like any other submission: an fp16 range rescue re-runs it from sl.feats_dev
  xv_model::HostSlot::Redo& rd = sl.redo;
  rd.seg_len.assign(seg_len_host, seg_len_host + n_seg);
  rd.utt = false; rd.has_first = rd.has_dst = false; rd.n_utt = 0; rd.out_dev = nullptr;
  rd.host_out = emb_host;
  rd.feats_ext = nullptr;
  rd.feats_f16 = false;
  rc = enqueue_slot(m, si);
  if (rc != XV_OK) return rc;
  m->last_launches += fe_launches;
  sl.busy = true;
  m->slot_next = (si + 1) % XV_HOST_SLOTS;
  *ticket = si;
  return XV_OK;
}

}  // extern "C"
