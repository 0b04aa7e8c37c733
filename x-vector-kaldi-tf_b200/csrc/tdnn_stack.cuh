// tdnn_stack.cuh -- the whole frame-level TDNN stack (all five fused layers) as ONE persistent launch.
//
// Same tiles, rings, warp roles and epilogues as tdnn_pair_kernel (tdnn_pair.cuh, which stays the per-layer /
// debug form and the bit-exact cross-check of this one).  What changes is what happens BETWEEN layers.  With one
// launch per layer every boundary costs a pipeline drain, the idle tail of the last partial round of tiles
// (832 tiles on 74 CTA pairs = 11.24 rounds), a launch and a pipeline fill: 14-25 k cycles per layer, measured.
// Here a CTA pair that has finished its tiles of layer i goes straight on to its tiles of layer i+1; a tile of
// layer i+1 only needs the rows it reads, so it is gated by per-row-tile completion counters of layer i
// (done[i][row tile], release/acquire at gpu scope), not by a grid-wide barrier:
//   * epilogue warps: after the TMA stores of a tile have COMPLETED (cp.async.bulk.wait_group, deferred by one tile
//     so nobody stalls on it) each warp adds 1 to done[layer][row tile]; a row tile is complete at
//     16 * n_ch_tiles arrivals (8 epilogue warps x 2 CTAs x channel tiles);
//   * TMA producer: before the first activation load of a tile of layer i+1 it waits for row tiles t-1, t, t+1 of
//     layer i (the halo never exceeds one tile).  Dependencies only point to lower layers and every pair walks its
//     tiles in order, so there is no cycle; all pairs are co-resident (persistent grid), so spinning is safe.
// The same counters make the ping-pong activation buffers safe: the tiles of layer i that READ rows of row tile t
// are exactly t-1, t, t+1, which are complete before layer i+1 overwrites those rows two layers later.
// Inside a CTA the operand rings are re-carved per layer (stage sizes differ), so the producer waits on a
// "layer drained" barrier that the MMA thread commits after the last MMA of a layer.  mbarrier phases are tracked
// per stage in bit masks, which keeps them valid across ring sizes.
#pragma once
#include "tdnn_pair.cuh"

namespace tdnn2 {

constexpr int MAX_STACK_LAYERS = 8;

struct StackLayer {
  CUtensorMap tmap_act, tmap_wgt, tmap_out;
  int32_t n_ch_tiles, c_chunks, taps, dilation, c_in_pad, reuse, n_act_stages, n_wgt_stages, mode, c_out;
  const float* bias;
  const float* scale;
  const float* shift;
  const float* alpha;
};

struct StackArgs {
  StackLayer layer[MAX_STACK_LAYERS];
  int32_t n_layers;
  int32_t n_row_tiles;        // R_pad / 256
  const uint8_t* row_valid;   // [R_pad]
  const uint8_t* blk_valid;   // [R_pad/32]
  float* partial;             // [R_pad/32][2][C_last]  pooled partial sums of the last layer
  uint32_t* overflow_flag;
  uint32_t* done;             // [n_layers][n_row_tiles] completion counters, zero on entry
  int32_t debug;              // timing experiments only (results are WRONG): 1 = no dependency waits, 2 = no signalling
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Spin until done[row tile] of the previous layer has `need` arrivals (bounded like the mbarrier waits).
__device__ __forceinline__ void wait_rows_done(const uint32_t* flag, uint32_t need) {
  if (ld_acquire_gpu(flag) >= need) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (ld_acquire_gpu(flag) < need) {
    if ((++spins & 0xffu) == 0) {
      const uint64_t now = ptx::globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > XV_MBAR_TIMEOUT_NS) __trap();
    }
  }
}

template <bool LEAKY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
tdnn_stack_kernel(const __grid_constant__ StackArgs args) {
  constexpr int ATOMS = 2;
  constexpr int STAGE_K = ATOMS * BLOCK_K;
  constexpr int WGT_STAGE_BYTES = ATOMS * WGT_ATOM_BYTES;
  constexpr int BAR_DRAINED = 4 * MAX_STAGES + 4;    // one more barrier than tdnn_pair_kernel (fits: see static_assert)
  static_assert(OFF_BARS + (NUM_BARS + 1) * 8 <= OFF_TMEM_PTR, "barrier area");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t sAct = smem_base + OFF_RING;
  const uint32_t bar0 = smem_base + OFF_BARS;
  auto act_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto act_empty = [&](uint32_t s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto wgt_full = [&](uint32_t s) { return bar0 + 8u * (2 * MAX_STAGES + s); };
  auto wgt_empty = [&](uint32_t s) { return bar0 + 8u * (3 * MAX_STAGES + s); };
  auto t_full = [&](uint32_t s) { return bar0 + 8u * (4 * MAX_STAGES + s); };
  auto t_empty = [&](uint32_t s) { return bar0 + 8u * (4 * MAX_STAGES + 2 + s); };
  const uint32_t drained = bar0 + 8u * BAR_DRAINED;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int li = 0; li < args.n_layers; ++li) {
      ptx::prefetch_tmap(&args.layer[li].tmap_act);
      ptx::prefetch_tmap(&args.layer[li].tmap_wgt);
      if (args.layer[li].mode == 0) ptx::prefetch_tmap(&args.layer[li].tmap_out);
    }
    for (uint32_t s = 0; s < MAX_STAGES; ++s) {
      ptx::mbar_init(act_full(s), 1); ptx::mbar_init(act_empty(s), 1);
      ptx::mbar_init(wgt_full(s), 1); ptx::mbar_init(wgt_empty(s), 1);
    }
    for (uint32_t s = 0; s < 2; ++s) { ptx::mbar_init(t_full(s), 1); ptx::mbar_init(t_empty(s), 2 * NUM_EPI_WARPS); }
    ptx::mbar_init(drained, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), TMEM_COLS);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  __syncthreads();                            // (the cluster barrier already orders this; racecheck only models bar.sync)
  const uint32_t tmem_base = *tmem_ptr_smem;
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();

  if (warp == 0) {
    // ============================ TMA producer ===========================================
    uint32_t ae_ph = 0xffffffffu, we_ph = 0xffffffffu;       // parity to wait for on each empty barrier (fresh = free)
    const uint32_t act_full_leader = ptx::mapa_cluster(act_full(0), 0);
    const uint32_t wgt_full_leader = ptx::mapa_cluster(wgt_full(0), 0);
    for (int li = 0; li < args.n_layers; ++li) {
      const StackLayer& L = args.layer[li];
      const bool reuse = L.reuse != 0;
      const uint32_t act_atom_stride = reuse ? uint32_t(ACT_ATOM_BYTES) : uint32_t(ACT_BOX_ROWS_PLAIN * 128);
      const uint32_t act_stage_bytes = ATOMS * act_atom_stride;
      const uint32_t sWgt = sAct + uint32_t(L.n_act_stages) * act_stage_bytes;
      const uint32_t act_box_bytes = (reuse ? ACT_BOX_ROWS_REUSE : ACT_BOX_ROWS_PLAIN) * 128u;
      const uint32_t n_act = uint32_t(L.n_act_stages), n_wgt = uint32_t(L.n_wgt_stages);
      const int half_ctx = (L.taps - 1) >> 1, halo = half_ctx * L.dilation;
      const uint32_t need = li > 0 ? uint32_t(2 * NUM_EPI_WARPS * args.layer[li - 1].n_ch_tiles) : 0u;
      const uint32_t* prev_done = li > 0 ? args.done + size_t(li - 1) * args.n_row_tiles : nullptr;
      // the rings are re-carved for this layer: every MMA of the previous layer must have retired
      if (li > 0) ptx::mbar_wait(drained, uint32_t(li - 1) & 1u);
      uint32_t sa = 0, sb = 0;
      TileCursor tc;
      tc.init(cluster_id, n_clusters, L.n_ch_tiles, 1, false, 0);
      const int n_items = args.n_row_tiles * L.n_ch_tiles;
      // Rows a tile reads: row tiles t-1 .. t+1 of the previous layer.  Their counters are read one tile AHEAD (the
      // acquire loads are in flight while the current tile's TMA loads are issued), so the ~1 us of an L2 round
      // trip is not on the producer's critical path; only a tile that really is early falls into the spin.
      auto load_flags = [&](int row, uint32_t (&f)[3]) {
#pragma unroll
        for (int d = 0; d < 3; ++d) f[d] = ld_acquire_gpu(prev_done + min(max(row - 1 + d, 0), args.n_row_tiles - 1));
      };
      uint32_t fcur[3] = {need, need, need}, fnext[3] = {need, need, need};
      if (li > 0 && tc.item < n_items && !(args.debug & 1)) load_flags(tc.row, fcur);
      for (; tc.item < n_items; tc.next()) {
        if (li > 0 && !(args.debug & 1)) {
          TileCursor nx = tc;
          nx.next();
          if (nx.item < n_items) load_flags(nx.row, fnext);
          if (min(min(fcur[0], fcur[1]), fcur[2]) < need)
            for (int t = max(tc.row - 1, 0); t <= min(tc.row + 1, args.n_row_tiles - 1); ++t) wait_rows_done(prev_done + t, need);
          fence_proxy_async_all();                        // the loads below go through the async proxy
          fcur[0] = fnext[0]; fcur[1] = fnext[1]; fcur[2] = fnext[2];
        }
        const int r0 = tc.row * TILE_ROWS + int(rank) * CTA_ROWS;
        const int c0 = tc.ch * TILE_CH + int(rank) * CTA_CH;
        for (int cc = 0; cc < L.c_chunks; ++cc) {
          for (int j = 0; j < L.taps; ++j) {
            if (!reuse || j == 0) {
              ptx::mbar_wait(act_empty(sa), (ae_ph >> sa) & 1u);
              ae_ph ^= 1u << sa;
              if (ptx::elect_one()) {
                if (leader) ptx::mbar_arrive_expect_tx(act_full(sa), 2u * ATOMS * act_box_bytes);
                const int row = reuse ? (r0 - halo) : (r0 + (j - half_ctx) * L.dilation);
#pragma unroll
                for (int h = 0; h < ATOMS; ++h)
                  ptx::tma_load_2d_2sm(sAct + sa * act_stage_bytes + h * act_atom_stride, &L.tmap_act, act_full_leader + 8u * sa,
                                       cc * STAGE_K + h * BLOCK_K, row);
              }
              __syncwarp();
              if (++sa == n_act) sa = 0;
            }
            ptx::mbar_wait(wgt_empty(sb), (we_ph >> sb) & 1u);
            we_ph ^= 1u << sb;
            if (ptx::elect_one()) {
              if (leader) ptx::mbar_arrive_expect_tx(wgt_full(sb), 2u * WGT_STAGE_BYTES);
#pragma unroll
              for (int h = 0; h < ATOMS; ++h)
                ptx::tma_load_2d_2sm(sWgt + sb * WGT_STAGE_BYTES + h * WGT_ATOM_BYTES, &L.tmap_wgt, wgt_full_leader + 8u * sb,
                                     j * L.c_in_pad + cc * STAGE_K + h * BLOCK_K, c0);
            }
            __syncwarp();
            if (++sb == n_wgt) sb = 0;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA) ================================
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(TILE_ROWS, TILE_CH);
      const uint64_t desc_hi = ptx::make_sw128_kmajor_desc(0);
      uint32_t af_ph = 0, wf_ph = 0, it = 0;                 // parity to wait for on each full barrier
      for (int li = 0; li < args.n_layers; ++li) {
        const StackLayer& L = args.layer[li];
        const bool reuse = L.reuse != 0, swapped = L.mode == 1;
        const uint32_t act_atom_stride = reuse ? uint32_t(ACT_ATOM_BYTES) : uint32_t(ACT_BOX_ROWS_PLAIN * 128);
        const uint32_t act_stage_bytes = ATOMS * act_atom_stride;
        const uint32_t sWgt = sAct + uint32_t(L.n_act_stages) * act_stage_bytes;
        const uint32_t n_act = uint32_t(L.n_act_stages), n_wgt = uint32_t(L.n_wgt_stages);
        uint32_t sa = 0, sb = 0;
        const int n_items = args.n_row_tiles * L.n_ch_tiles;
        for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
          const uint32_t acc = it & 1u;
          ptx::mbar_wait_cluster(t_empty(acc), ((it >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TILE_CH;
          uint32_t accumulate = 0;
          for (int cc = 0; cc < L.c_chunks; ++cc) {
            for (int j = 0; j < L.taps; ++j) {
              if (!reuse || j == 0) {
                ptx::mbar_wait(act_full(sa), (af_ph >> sa) & 1u);
                af_ph ^= 1u << sa;
              }
              ptx::mbar_wait(wgt_full(sb), (wf_ph >> sb) & 1u);
              wf_ph ^= 1u << sb;
              ptx::tc_fence_after();
              const uint32_t act_addr = sAct + sa * act_stage_bytes + (reuse ? uint32_t(j * L.dilation) * 128u : 0u);
              const uint32_t wgt_addr = sWgt + sb * WGT_STAGE_BYTES;
              const uint64_t d_act = desc_hi | uint64_t((act_addr >> 4) & 0x3fffu);
              const uint64_t d_wgt = desc_hi | uint64_t((wgt_addr >> 4) & 0x3fffu);
              const bool last_tap = !reuse || j == L.taps - 1;
              if (ptx::elect_one()) {
#pragma unroll
                for (int h = 0; h < ATOMS; ++h) {
#pragma unroll
                  for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    const uint64_t da = d_act + uint64_t(h * (act_atom_stride >> 4) + 2 * k);
                    const uint64_t dw = d_wgt + uint64_t(h * (WGT_ATOM_BYTES >> 4) + 2 * k);
                    if (swapped) ptx::umma_f16_2sm(d_tmem, dw, da, idesc, accumulate | uint32_t(h | k));
                    else ptx::umma_f16_2sm(d_tmem, da, dw, idesc, accumulate | uint32_t(h | k));
                  }
                }
                ptx::umma_commit_2sm(wgt_empty(sb));
                if (last_tap) ptx::umma_commit_2sm(act_empty(sa));
              }
              __syncwarp();
              accumulate = 1;
              if (++sb == n_wgt) sb = 0;
              if (last_tap) { if (++sa == n_act) sa = 0; }
            }
          }
          if (ptx::elect_one()) ptx::umma_commit_2sm(t_full(acc));
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit_2sm(drained);   // all MMAs of this layer retired -> rings reusable
        __syncwarp();
      }
    }
  } else {
    // ============================ epilogue (8 warps per CTA) =============================
    const int e = warp - 2;
    const int q = warp & 3;
    const int colh = e >> 2;
    const int te = threadIdx.x - 64;
    const uint32_t t_empty_leader = ptx::mapa_cluster(t_empty(0), 0);
    const uint32_t sC = smem_base + OFF_C + uint32_t(e) * C_BUF_BYTES;
    const uint32_t swz = (uint32_t(lane) >> 1) & 3u;
    uint32_t hmax = 0;
    uint32_t it = 0;
    for (int li = 0; li < args.n_layers; ++li) {
      const StackLayer& L = args.layer[li];
      const int n_items = args.n_row_tiles * L.n_ch_tiles;
      TileCursor cur0;
      cur0.init(cluster_id, n_clusters, L.n_ch_tiles, 1, false, 0);
      if (L.mode == 0) {
        uint32_t* my_done = args.done + size_t(li) * args.n_row_tiles;
        auto fetch_params = [&](const TileCursor& tc, float& b, float& sc, float& sh, float& al, uint32_t& valid) {
          const int ch = tc.ch * TILE_CH + te;
          b = __ldg(L.bias + ch); sc = __ldg(L.scale + ch); sh = __ldg(L.shift + ch);
          if (LEAKY) al = __ldg(L.alpha + ch);
          valid = args.row_valid[tc.row * TILE_ROWS + int(rank) * CTA_ROWS + q * 32 + lane];
        };
        auto store_params = [&](uint32_t acc, float b, float sc, float sh, float al) {
          const uint32_t s_par = smem_base + OFF_PARAMS + acc * (PAR_ARRAYS * TILE_CH * 4);
          ptx::sts_f(s_par + uint32_t(te) * 4u, b);
          ptx::sts_f(s_par + uint32_t(TILE_CH + te) * 4u, sc);
          ptx::sts_f(s_par + uint32_t(2 * TILE_CH + te) * 4u, sh);
          if (LEAKY) ptx::sts_f(s_par + uint32_t(3 * TILE_CH + te) * 4u, al);
        };
        float nb = 0.f, nsc = 0.f, nsh = 0.f, nal = 0.f;
        uint32_t nvalid = 0;
        if (cur0.item < n_items) {
          fetch_params(cur0, nb, nsc, nsh, nal, nvalid);
          store_params(it & 1u, nb, nsc, nsh, nal);
        }
        int prev_row = -1;                                   // row tile whose stores are issued but not yet signalled
        TileCursor nx = cur0;
        for (TileCursor tc = cur0; tc.item < n_items; tc.next(), ++it) {
          const uint32_t acc = it & 1u;
          const int r_cta = tc.row * TILE_ROWS + int(rank) * CTA_ROWS;
          const int ch0 = tc.ch * TILE_CH;
          const uint32_t s_par = smem_base + OFF_PARAMS + acc * (PAR_ARRAYS * TILE_CH * 4);
          const bool valid = nvalid != 0;
          ptx::named_bar_sync(1, NUM_EPI_THREADS);
          nx.next();
          const bool has_next = nx.item < n_items;
          if (has_next) fetch_params(nx, nb, nsc, nsh, nal, nvalid);
          ptx::mbar_wait(t_full(acc), (it >> 1) & 1u);
          ptx::tc_fence_after();
          const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * TILE_CH + uint32_t(colh) * 128u;
          uint32_t v[2][32];
          ptx::tmem_ld_32x32(t_row, v[0]);
#pragma unroll
          for (int chunk = 0; chunk < 4; ++chunk) {
            ptx::tmem_ld_wait_dep(v[chunk & 1]);
            if (chunk < 3) {
              ptx::tmem_ld_32x32(t_row + (chunk + 1) * C_CHUNK, v[(chunk + 1) & 1]);
            } else {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_cluster(t_empty_leader + 8u * acc);
            }
            uint32_t p[16];
            epi_store_math<LEAKY>(v[chunk & 1], s_par, colh * 128 + chunk * C_CHUNK, valid, p, hmax);
            if (lane == 0) ptx::tma_store_wait_read<0>();
            __syncwarp();
            const uint32_t dst = sC + uint32_t(lane) * (C_CHUNK * 2);
#pragma unroll
            for (uint32_t c16 = 0; c16 < 4; ++c16) {
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((c16 ^ swz) << 4)),
                           "r"(p[c16 * 4 + 0]), "r"(p[c16 * 4 + 1]), "r"(p[c16 * 4 + 2]), "r"(p[c16 * 4 + 3])
                           : "memory");
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&L.tmap_out, sC, ch0 + colh * 128 + chunk * C_CHUNK, r_cta + q * 32);
              ptx::tma_store_commit();
            }
          }
          if (has_next) store_params(acc ^ 1u, nb, nsc, nsh, nal);
          // signal the PREVIOUS tile: its four stores are older than this tile's four, so they have completed once at
          // most four bulk groups are still pending
          if (lane == 0) {
            if (prev_row >= 0 && !(args.debug & 2)) {
              ptx::tma_store_wait_all<4>();                  // issued a whole tile ago: complete by now, no stall
              red_release_gpu_add(my_done + prev_row, 1u);
            }
          }
          prev_row = tc.row;
        }
        if (lane == 0 && prev_row >= 0) {
          ptx::tma_store_wait_all<0>();
          red_release_gpu_add(my_done + prev_row, 1u);
        }
      } else {
        for (TileCursor tc = cur0; tc.item < n_items; tc.next(), ++it) {
          const uint32_t acc = it & 1u;
          const int r_tile = tc.row * TILE_ROWS;
          const int ch = tc.ch * TILE_CH + int(rank) * CTA_CH + q * 32 + lane;
          const float b = __ldg(L.bias + ch), sc = __ldg(L.scale + ch), sh = __ldg(L.shift + ch);
          const float al = LEAKY ? __ldg(L.alpha + ch) : 0.f;
          const int blk0 = (r_tile + colh * 128) / POOL_BLOCK;
          const uint32_t nv4 = *reinterpret_cast<const uint32_t*>(args.blk_valid + blk0);
          ptx::mbar_wait(t_full(acc), (it >> 1) & 1u);
          ptx::tc_fence_after();
          const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * TILE_CH + uint32_t(colh) * 128u;
          uint32_t v[2][32];
          ptx::tmem_ld_32x32(t_row, v[0]);
#pragma unroll
          for (int chunk = 0; chunk < 4; ++chunk) {
            ptx::tmem_ld_wait_dep(v[chunk & 1]);
            if (chunk < 3) {
              ptx::tmem_ld_32x32(t_row + (chunk + 1) * C_CHUNK, v[(chunk + 1) & 1]);
            } else {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_cluster(t_empty_leader + 8u * acc);
            }
            const int nv = int((nv4 >> (8 * chunk)) & 0xffu);
            if (nv > 0) {
              float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
              if (nv >= POOL_BLOCK) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float y = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][i]), b, sc, sh, al);
                  s1[i & 3] += y;
                  s2[i & 3] = fmaf(y, y, s2[i & 3]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  float y = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][i]), b, sc, sh, al);
                  y = (i < nv) ? y : 0.f;
                  s1[i & 3] += y;
                  s2[i & 3] = fmaf(y, y, s2[i & 3]);
                }
              }
              float* dst = args.partial + size_t(blk0 + chunk) * 2 * L.c_out + ch;
              dst[0] = (s1[0] + s1[1]) + (s1[2] + s1[3]);
              dst[L.c_out] = (s2[0] + s2[1]) + (s2[2] + s2[3]);
            }
          }
        }
      }
    }
    if ((hmax & 0x7fffu) >= 0x7c00u || ((hmax >> 16) & 0x7fffu) >= 0x7c00u) atomicOr(args.overflow_flag, 1u);
  }
  __syncwarp();

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace tdnn2
