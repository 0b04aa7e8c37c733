// tdnn_tail.cuh -- the last two frame layers of the statistics-pooling topologies as ONE kernel (sm_100a).
//
// Both are context-free (k = 1) contractions: layer n-2 [rows, C_in] -> [rows, 512] (conv -> +b -> relu -> BatchNorm eval,
// local/tf/models.py:476-480) and layer n-1 [rows, 512] -> [rows, C_out] whose output is only pooled (tf.nn.moments over
// time, models.py:485).  Run as two launches of tdnn_pair_kernel they are bound by operand delivery, not by the tensor
// pipe: a 256 x 256 tile with K = 512 moves 256 KB per SM for 4096 tensor cycles, 47-62 B/clk/SM against the chip's
// ~43 B/clk/SM of L2 throughput (measured: 0.58 and 0.78 of the tensor peak), and the 512-wide intermediate makes a
// 105 MB round trip through L2 / HBM.  Here a CTA pair keeps the intermediate of its 256-row tile ON CHIP:
//
//   per row tile (8 accumulation jobs over two TMEM accumulators of 256 columns, alternating):
//     jobs 0,1   D[rows, 256 ch] = X[rows, C_in] . W3[nt]^T      (A = X, B = W3: store orientation, TMEM lane = row)
//                epilogue: relu(acc + b) * scale + shift, gap rows -> 0, fp16 -> written by the epilogue warps straight
//                into shared memory in the K-major SWIZZLE_128B layout of a UMMA B operand (8 atoms of 64 channels x
//                128 rows x 128 B per CTA = 128 KB): the next layer's activations never leave the SM
//     jobs 2..7  D[256 ch, rows] = W4[ct] . MID^T                  (A = W4 streamed, B = MID resident: pooled orientation)
//                epilogue: as tdnn_pair_kernel mode 1 (lane = channel, the 32 registers of a tcgen05.ld are 32 frames):
//                partial[block][{sum, sumsq}][channel]
//   operand bytes per SM per row tile: 192 KB (X: once into the intermediate's dead buffer for the first half of layer n-2 and
//   atoms 4..7 of the second, atoms 0..3 once more through the ring) + 256 KB (W3) + 768 KB (W4) = 1.19 MB for 32 768 tensor
//   cycles = 37 B/clk/SM -- under the L2 cap -- instead of 2 MB; nothing is written but the pooled partials.
//
// The arithmetic is that of the two separate launches, in the same order (K ascending in 64-wide atoms, same epilogue
// expressions), so results are BIT-IDENTICAL to them (tests/test_gpu_parity.py::test_fused_tail_is_bit_identical).
//
// Ring: 6 slots of 16 KB (one 64-wide K atom of one operand: [128 rows x 128 B]); layer n-2 uses two slots per K step (X, W3),
// layer n-1 one (W4).  Warp roles as in tdnn_pair_kernel: warp 0 TMA producer (both CTAs), warp 1 TMEM allocator + MMA issuer
// (leader), warps 2..9 epilogue.
#pragma once
#include "tdnn_pair.cuh"

namespace tdnn2 {

constexpr int FT_SLOT_BYTES = 16384;
constexpr int FT_SLOTS = 5;
constexpr int FT_MID_CH = 512;                                  // width of the intermediate (two 256-column accumulators)
constexpr int FT_MID_ATOMS = FT_MID_CH / BLOCK_K;               // 8
constexpr int FT_OFF_MID = 0;
constexpr int FT_OFF_RING = FT_MID_ATOMS * FT_SLOT_BYTES;       // 131072
constexpr int FT_OFF_PAR = FT_OFF_RING + FT_SLOTS * FT_SLOT_BYTES;    // 212992: bias | scale | shift | slope of the 512 intermediate channels
constexpr int FT_OFF_BARS = FT_OFF_PAR + 4 * FT_MID_CH * 4;           // 221184
constexpr int FT_NUM_BARS = 2 * FT_SLOTS + 6 + 2 * FT_MID_ATOMS; // full, empty | t_full[2], t_empty[2], mid_ready[2] | x_full[8], mid_free[8]
constexpr int FT_OFF_TMEM_PTR = FT_OFF_BARS + (FT_NUM_BARS + 2) * 8;
constexpr int FT_SMEM_BYTES = FT_OFF_TMEM_PTR + 16 + 1024;      // + slack for 1024-byte alignment
static_assert(FT_SMEM_BYTES <= 232448, "exceeds 227 KB of dynamic shared memory");

struct FusedTailArgs {
  int32_t n_row_tiles;      // R_pad / 256
  int32_t k_atoms_in;       // C_in / 64 of layer n-2
  int32_t n_ch_tiles;       // C_out / 256 of layer n-1
  int32_t c_out;
  // layer n-2 (512 channels): conv bias, folded BatchNorm, negative slope (staged in shared memory once per CTA: the
  // store-orientation epilogue needs all of a chunk's 32 channels in every thread; global or constant-bank reads made
  // that epilogue 2-4x slower); acc_scale as in PairArgs
  const float* bias3; const float* scale3; const float* shift3; const float* alpha3;
  float acc_scale3;
  // layer n-1
  const float* bias4; const float* scale4; const float* shift4; const float* alpha4;
  float acc_scale4;
  const uint8_t* row_valid; // [R_pad]
  const uint8_t* blk_valid; // [R_pad / 32]
  float* partial;           // [R_pad / 32][2][C_out]
  uint32_t* overflow_flag;
  uint32_t overflow_bit3;   // OR-ed in when a stored value of the intermediate overflowed fp16
  long long* trace;         // diagnostics: [cluster][rank][TRACE_TILES jobs][8] SM clock stamps (slots as in tdnn_pair.cuh), or null
};
__device__ __forceinline__ void ft_stamp(const FusedTailArgs& a, int cluster, uint32_t rank, uint32_t job, int slot) {
  if (a.trace != nullptr && job < TRACE_TILES) {
    a.trace[((size_t(cluster) * 2 + rank) * TRACE_TILES + job) * 8 + slot] = clock64();
    // slot 7 <- %globaltimer (ns) next to the "MMA start" stamp: calibrates the SM clock DURING the kernel (measured: 1 557 MHz
    // while nvidia-smi reports 1 965 -- the tensor kernels of a step run clocked down)
    if (slot == 2) a.trace[((size_t(cluster) * 2 + rank) * TRACE_TILES + job) * 8 + 7] = (long long)ptx::globaltimer_ns();
  }
}

// Order of the layer n-1 half-jobs (channel tile, K half) of a row tile: tiles 0 and 1 interleaved by halves -- (0, lo) (1, lo)
// (0, hi) (1, hi) -- then (2, lo) (2, hi) (3, lo) ...  Each accumulator still sums atoms 0..7 in order (same bits); but the
// second tile can start on atoms 0..3 of the intermediate while the epilogue of the second layer n-2 half is still writing
// atoms 4..7, instead of the first tile waiting for them alone.  (Host: at least two channel tiles.)
__device__ __forceinline__ void ft_half_job(int hj, int& ct, int& half) {
  if (hj < 4) { ct = hj & 1; half = hj >> 1; }
  else { ct = hj >> 1; half = hj & 1; }
}

__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <bool LEAKY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
tdnn_tail_fused_kernel(const __grid_constant__ CUtensorMap tmap_x,    // [R_pad, C_in] fp16 input of layer n-2
                       const __grid_constant__ CUtensorMap tmap_w3,   // [512, C_in] fp16 K-major
                       const __grid_constant__ CUtensorMap tmap_w4,   // [C_out, 512] fp16 K-major
                       const FusedTailArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t sMid = smem_base + FT_OFF_MID;
  const uint32_t sRing = smem_base + FT_OFF_RING;
  const uint32_t bar0 = smem_base + FT_OFF_BARS;
  auto full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty = [&](uint32_t s) { return bar0 + 8u * (FT_SLOTS + s); };
  auto t_full = [&](uint32_t s) { return bar0 + 8u * (2 * FT_SLOTS + s); };
  auto t_empty = [&](uint32_t s) { return bar0 + 8u * (2 * FT_SLOTS + 2 + s); };
  auto mid_ready = [&](uint32_t s) { return bar0 + 8u * (2 * FT_SLOTS + 4 + s); };
  // The INPUT rows of the next row tile wait in the intermediate's buffer while it is dead: atom a of it is free once the last
  // layer n-1 job of a tile has multiplied it (mid_free[a]) and holds input atom a from then (x_full[a]) until the epilogue of
  // the layer n-2 half that owns it writes the intermediate there.  The first half of layer n-2 then streams only W3 through
  // the ring (one slot per K step, like layer n-1) and the second half re-reads only input atoms 0..3: 448 KB instead of 512 KB
  // per SM and row tile in the phase that is bound by the 40 B/clk an SM ingests, and the input is requested ~2 500 cycles
  // before the phase starts instead of as ring slots free up.
  auto x_full = [&](uint32_t a) { return bar0 + 8u * (2 * FT_SLOTS + 6 + a); };
  auto mid_free = [&](uint32_t a) { return bar0 + 8u * (2 * FT_SLOTS + 6 + FT_MID_ATOMS + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + FT_OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w3);
    ptx::prefetch_tmap(&tmap_w4);
    for (uint32_t s = 0; s < FT_SLOTS; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    for (uint32_t s = 0; s < 2; ++s) {
      ptx::mbar_init(t_full(s), 1);
      ptx::mbar_init(t_empty(s), 2 * NUM_EPI_WARPS);
      ptx::mbar_init(mid_ready(s), 2 * NUM_EPI_WARPS);
    }
    for (uint32_t a = 0; a < FT_MID_ATOMS; ++a) { ptx::mbar_init(x_full(a), 1); ptx::mbar_init(mid_free(a), 1); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), TMEM_COLS);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  __syncthreads();
  const uint32_t tmem_base = *tmem_ptr_smem;
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();

  const int KA = args.k_atoms_in;
  const int NCT = args.n_ch_tiles;

  if (warp == 0) {
    // ============================ TMA producer ============================
    uint32_t s = 0, ph = 0, pg = 0;
    const uint32_t full_leader = ptx::mapa_cluster(full(0), 0);
    auto load = [&](const CUtensorMap* map, int col, int row) {
      ptx::mbar_wait(empty(s), ph ^ 1u);
      if (ptx::elect_one()) {
        if (leader) ptx::mbar_arrive_expect_tx(full(s), 2u * FT_SLOT_BYTES);            // both CTAs' boxes
        ptx::tma_load_2d_2sm(sRing + s * FT_SLOT_BYTES, map, full_leader + 8u * s, col, row);
      }
      __syncwarp();
      if (++s == FT_SLOTS) { s = 0; ph ^= 1u; }
    };
    const uint32_t x_full_leader = ptx::mapa_cluster(x_full(0), 0);
    uint32_t n_res = 0;                                          // row tiles whose input went into the intermediate's buffer so far
    // Everything the FIRST half of layer n-2 needs for one row tile: input atom a into the intermediate's atom a as soon as the
    // previous tile's last layer n-1 job has multiplied it, W3 atom a into the ring slot that job's W4 atom a just left -- both
    // are requested while that job still runs, in the order it frees them.
    auto first_half = [&](int row) {
      for (int a = 0; a < KA; ++a) {                              // (host: KA <= FT_MID_ATOMS)
        ptx::mbar_wait(mid_free(uint32_t(a)), (n_res & 1u) ^ 1u);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(x_full(uint32_t(a)), 2u * FT_SLOT_BYTES);
          ptx::tma_load_2d_2sm(sMid + uint32_t(a) * FT_SLOT_BYTES, &tmap_x, x_full_leader + 8u * uint32_t(a), a * BLOCK_K, row);
        }
        __syncwarp();
        load(&tmap_w3, a * BLOCK_K, int(rank) * CTA_CH);
      }
      ++n_res;
    };
    if (cluster_id < args.n_row_tiles) first_half(cluster_id * TILE_ROWS + int(rank) * CTA_ROWS);
    for (int tile = cluster_id; tile < args.n_row_tiles; tile += n_clusters) {
      const int r0 = tile * TILE_ROWS + int(rank) * CTA_ROWS;
      ++pg;                                                       // (the first half's loads were issued with the previous tile)
      for (int ka = 0; ka < KA; ++ka) {                           // second half: input atoms 0..3 once more (their place holds the
        if (ka < FT_MID_ATOMS / 2) load(&tmap_x, ka * BLOCK_K, r0);   // intermediate by now), W3
        if (ka == 0 && lane == 0) ft_stamp(args, cluster_id, rank, pg, 0);
        load(&tmap_w3, ka * BLOCK_K, TILE_CH + int(rank) * CTA_CH);
      }
      if (lane == 0) ft_stamp(args, cluster_id, rank, pg, 1);
      ++pg;
      for (int hj = 0; hj < 2 * NCT; ++hj) {                      // W4 atoms in the order the issuer multiplies them (ft_half_job)
        int ct, half;
        ft_half_job(hj, ct, half);
        if (hj == 4 && tile + n_clusters < args.n_row_tiles && ptx::elect_one()) {
          // the next row tile's input rows -> L2 now
          for (int ka = 0; ka < KA; ++ka) ptx::tma_prefetch_l2_2d(&tmap_x, ka * BLOCK_K, r0 + n_clusters * TILE_ROWS);
        }
        __syncwarp();
        for (int ka = half * (FT_MID_ATOMS / 2); ka < (half + 1) * (FT_MID_ATOMS / 2); ++ka) {
          load(&tmap_w4, ka * BLOCK_K, ct * TILE_CH + int(rank) * CTA_CH);
          if (ka == 0 && lane == 0) ft_stamp(args, cluster_id, rank, pg + uint32_t(ct), 0);
        }
        if (half == 1 && lane == 0) ft_stamp(args, cluster_id, rank, pg + uint32_t(ct), 1);
      }
      pg += uint32_t(NCT);
      if (tile + n_clusters < args.n_row_tiles) first_half(r0 + n_clusters * TILE_ROWS);
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader) ============================
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(TILE_ROWS, TILE_CH);
      const uint64_t desc_hi = ptx::make_sw128_kmajor_desc(0);
      auto desc = [&](uint32_t addr) { return desc_hi | uint64_t((addr >> 4) & 0x3fffu); };
      uint32_t s = 0, ph = 0, g = 0, ti = 0;
      for (int tile = cluster_id; tile < args.n_row_tiles; tile += n_clusters, ++ti) {
        // ---- layer n-2: two 256-channel halves, store orientation ----
        for (int nt = 0; nt < 2; ++nt, ++g) {
          const uint32_t acc = g & 1u;
          ptx::mbar_wait_cluster(t_empty(acc), ((g >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TILE_CH;
          if (lane == 0) ft_stamp(args, cluster_id, rank, g, 2);
          for (int ka = 0; ka < KA; ++ka) {
            // A operand: the resident input atom (first half: all of them, loaded once per row tile; second half: atoms 4..7,
            // which the first half's epilogue does not touch), else an input atom streamed through the ring
            const bool ring_x = nt == 1 && ka < FT_MID_ATOMS / 2;
            uint32_t a_addr = sMid + uint32_t(ka) * FT_SLOT_BYTES;
            uint32_t sx = 0;
            if (ring_x) {
              sx = s;
              ptx::mbar_wait(full(s), ph);
              a_addr = sRing + s * FT_SLOT_BYTES;
              if (++s == FT_SLOTS) { s = 0; ph ^= 1u; }
            } else if (nt == 0) {
              ptx::mbar_wait(x_full(uint32_t(ka)), ti & 1u);
            }
            const uint32_t sw = s;
            ptx::mbar_wait(full(s), ph);
            if (++s == FT_SLOTS) { s = 0; ph ^= 1u; }
            ptx::tc_fence_after();
            const uint64_t da = desc(a_addr), dw = desc(sRing + sw * FT_SLOT_BYTES);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                ptx::umma_f16_2sm(d_tmem, da + uint64_t(2 * k), dw + uint64_t(2 * k), idesc, uint32_t(ka | k));
              if (ring_x) ptx::umma_commit_2sm(empty(sx));
              ptx::umma_commit_2sm(empty(sw));
            }
            __syncwarp();
          }
          if (ptx::elect_one()) ptx::umma_commit_2sm(t_full(acc));
          __syncwarp();
          if (lane == 0) ft_stamp(args, cluster_id, rank, g, 3);
        }
        // ---- layer n-1: pooled orientation, B = the resident intermediate.  Half-jobs (channel tile, K half) in ft_half_job's
        //      order: the first two channel tiles start on atoms 0..3 while the second layer n-2 epilogue still writes 4..7.
        const uint32_t g0 = g;
        for (int hj = 0; hj < 2 * NCT; ++hj) {
          int ct, half;
          ft_half_job(hj, ct, half);
          const uint32_t gj = g0 + uint32_t(ct);
          const uint32_t acc = gj & 1u;
          if (half == 0) {
            ptx::mbar_wait_cluster(t_empty(acc), ((gj >> 1) & 1u) ^ 1u);
            ptx::tc_fence_after();
            if (lane == 0) ft_stamp(args, cluster_id, rank, gj, 2);
          }
          if (hj == 0 || hj == 2) {                                   // the half of the intermediate read from here on is written
            ptx::mbar_wait_cluster(mid_ready(hj == 0 ? 0u : 1u), ti & 1u);
            ptx::tc_fence_after();
          }
          const uint32_t d_tmem = tmem_base + acc * TILE_CH;
          for (int ka = half * (FT_MID_ATOMS / 2); ka < (half + 1) * (FT_MID_ATOMS / 2); ++ka) {
            ptx::mbar_wait(full(s), ph);
            ptx::tc_fence_after();
            const uint64_t dw = desc(sRing + s * FT_SLOT_BYTES), dm = desc(sMid + uint32_t(ka) * FT_SLOT_BYTES);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                ptx::umma_f16_2sm(d_tmem, dw + uint64_t(2 * k), dm + uint64_t(2 * k), idesc, uint32_t(ka | k));
              ptx::umma_commit_2sm(empty(s));
              if (ct == NCT - 1) ptx::umma_commit_2sm(mid_free(uint32_t(ka)));   // this atom of the intermediate has had its last reader
            }
            __syncwarp();
            if (++s == FT_SLOTS) { s = 0; ph ^= 1u; }
          }
          if (half == 1) {
            if (ptx::elect_one()) ptx::umma_commit_2sm(t_full(acc));
            __syncwarp();
            if (lane == 0) ft_stamp(args, cluster_id, rank, gj, 3);
          }
        }
        g += uint32_t(NCT);
      }
    }
  } else {
    // ============================ epilogue (8 warps per CTA) ============================
    const int e = warp - 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int colh = e >> 2;                         // which 128-column half of the accumulator
    const uint32_t t_empty_leader = ptx::mapa_cluster(t_empty(0), 0);
    const uint32_t mid_ready_leader = ptx::mapa_cluster(mid_ready(0), 0);
    const uint32_t sPar = smem_base + FT_OFF_PAR;
    for (int c = threadIdx.x - 64; c < FT_MID_CH; c += NUM_EPI_THREADS) {          // once: the parameters do not change with the tile
      ptx::sts_f(sPar + uint32_t(c) * 4u, __ldg(args.bias3 + c));
      ptx::sts_f(sPar + uint32_t(FT_MID_CH + c) * 4u, __ldg(args.scale3 + c));
      ptx::sts_f(sPar + uint32_t(2 * FT_MID_CH + c) * 4u, __ldg(args.shift3 + c));
      ptx::sts_f(sPar + uint32_t(3 * FT_MID_CH + c) * 4u, LEAKY ? __ldg(args.alpha3 + c) : 0.f);
    }
    ptx::named_bar_sync(1, NUM_EPI_THREADS);
    const float as3 = args.acc_scale3 != 0.f ? args.acc_scale3 : 1.f;
    const float as4 = args.acc_scale4 != 0.f ? args.acc_scale4 : 1.f;
    uint32_t hmax = 0, g = 0;
    for (int tile = cluster_id; tile < args.n_row_tiles; tile += n_clusters) {
      const int r_cta = tile * TILE_ROWS + int(rank) * CTA_ROWS;
      const int r_loc = q * 32 + lane;                                  // my row within the CTA's 128 (store orientation)
      const bool valid = args.row_valid[r_cta + r_loc] != 0;
      // ---- layer n-2: activation -> fp16 -> the resident B operand of layer n-1 ----
      for (int nt = 0; nt < 2; ++nt, ++g) {
        const uint32_t acc = g & 1u;
        ptx::mbar_wait(t_full(acc), (g >> 1) & 1u);
        ptx::tc_fence_after();
        if (e == 0 && lane == 0) ft_stamp(args, cluster_id, rank, g, 4);
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
            if (e == 0 && lane == 0) ft_stamp(args, cluster_id, rank, g, 5);
          }
          const int ch = nt * TILE_CH + colh * 128 + chunk * C_CHUNK;   // first of this chunk's 32 intermediate channels
          uint32_t p[16];
#pragma unroll
          for (int gq = 0; gq < 8; ++gq) {
            const uint32_t c4 = uint32_t(ch + gq * 4) * 4u;
            const float4 b4 = ptx::lds_f4(sPar + c4);
            const float4 s4 = ptx::lds_f4(sPar + FT_MID_CH * 4u + c4);
            const float4 h4 = ptx::lds_f4(sPar + 2u * FT_MID_CH * 4u + c4);
            float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (LEAKY) a4 = ptx::lds_f4(sPar + 3u * FT_MID_CH * 4u + c4);
            const float y0 = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][gq * 4 + 0]), b4.x, s4.x, h4.x, a4.x, as3);
            const float y1 = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][gq * 4 + 1]), b4.y, s4.y, h4.y, a4.y, as3);
            const float y2 = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][gq * 4 + 2]), b4.z, s4.z, h4.z, a4.z, as3);
            const float y3 = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][gq * 4 + 3]), b4.w, s4.w, h4.w, a4.w, as3);
            const uint32_t p0 = ptx::pack_half2(y0, y1), p1 = ptx::pack_half2(y2, y3);
            hmax = ptx::habs2_max(ptx::habs2_max(hmax, p0), p1);
            p[gq * 2 + 0] = valid ? p0 : 0u;                            // gap rows stay exact zeros
            p[gq * 2 + 1] = valid ? p1 : 0u;
          }
          // K-major SWIZZLE_128B: atom = 64 channels, row = 128 bytes, 16-byte unit u of row r stored at u ^ (r & 7)
          const uint32_t atom = uint32_t(nt * 4 + colh * 2 + (chunk >> 1));
          const uint32_t row_addr = sMid + atom * FT_SLOT_BYTES + uint32_t(r_loc) * 128u;
          const uint32_t ub = uint32_t(chunk & 1) * 4u, sw = uint32_t(r_loc) & 7u;
#pragma unroll
          for (uint32_t c16 = 0; c16 < 4; ++c16) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + (((ub + c16) ^ sw) << 4)),
                         "r"(p[c16 * 4 + 0]), "r"(p[c16 * 4 + 1]), "r"(p[c16 * 4 + 2]), "r"(p[c16 * 4 + 3])
                         : "memory");
          }
        }
        ptx::fence_proxy_async_smem();               // the tensor core reads these rows through the async proxy
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_release(mid_ready_leader + 8u * uint32_t(nt));
        if (e == 0 && lane == 0) ft_stamp(args, cluster_id, rank, g, 6);
      }
      // ---- layer n-1: pooled partial sums (lane = channel, registers = 32 consecutive frames) ----
      for (int ct = 0; ct < NCT; ++ct, ++g) {
        const uint32_t acc = g & 1u;
        const int r_tile = tile * TILE_ROWS;
        const int ch = ct * TILE_CH + int(rank) * CTA_CH + q * 32 + lane;
        const float b = __ldg(args.bias4 + ch), sc = __ldg(args.scale4 + ch), sh = __ldg(args.shift4 + ch);
        const float al = LEAKY ? __ldg(args.alpha4 + ch) : 0.f;
        const int blk0 = (r_tile + colh * 128) / POOL_BLOCK;
        const uint32_t nv4 = *reinterpret_cast<const uint32_t*>(args.blk_valid + blk0);   // 4 blocks, 1 byte each
        ptx::mbar_wait(t_full(acc), (g >> 1) & 1u);
        ptx::tc_fence_after();
        if (e == 0 && lane == 0) ft_stamp(args, cluster_id, rank, g, 4);
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
          const int nv = int((nv4 >> (8 * chunk)) & 0xffu);          // warp-uniform
          if (nv > 0) {
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
            if (nv >= POOL_BLOCK) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float y = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][i]), b, sc, sh, al, as4);
                s1[i & 3] += y;
                s2[i & 3] = fmaf(y, y, s2[i & 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float y = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][i]), b, sc, sh, al, as4);
                y = (i < nv) ? y : 0.f;                              // rows past the segment end do not count
                s1[i & 3] += y;
                s2[i & 3] = fmaf(y, y, s2[i & 3]);
              }
            }
            float* dst = args.partial + size_t(blk0 + chunk) * 2 * args.c_out + ch;
            dst[0] = (s1[0] + s1[1]) + (s1[2] + s1[3]);
            dst[args.c_out] = (s2[0] + s2[1]) + (s2[2] + s2[3]);
          }
        }
        if (e == 0 && lane == 0) ft_stamp(args, cluster_id, rank, g, 6);
      }
    }
    if ((hmax & 0x7fffu) >= 0x7c00u || ((hmax >> 16) & 0x7fffu) >= 0x7c00u)
      atomicOr(args.overflow_flag, args.overflow_bit3 ? args.overflow_bit3 : 1u);
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
