// wgrad_pair.cuh -- weight gradient of a frame-level TDNN layer on the tensor cores (sm_100a).
//
// Training-step counterpart of tdnn_pair.cuh: for the conv of local/tf/models.py:476 / :579 the gradient
// that tf.train.AdamOptimizer.minimize (models.py:516-519) needs is
//     dW[j, ci, co] = sum over packed rows r of  x[r + (j - (k-1)/2) * d, ci] * dz[r, co]
// a GEMM whose CONTRACTION runs over the rows.  Both operands live in HBM row-major ([R_pad, C] fp16, the layout
// the forward kernels write), so for the MMA they are "MN-major": the TMA box [64 channels x 128 rows] lands in
// shared memory as 128-byte rows (one per packed row, SWIZZLE_128B) and the UMMA descriptors read it transposed
// (instruction-descriptor bits 15/16 = MN-major; LBO = distance between the two 64-channel boxes, SBO = 1024 B
// between 8-row groups; one UMMA_K = 16 rows = 2048 B).  No transposed copy of the activations is ever made, and
// a temporal tap is just a row offset of the x box (rows outside [0, R_pad) are zero-filled by TMA; gap rows are
// exact zeros in both operands, which is the SAME padding of the forward pass).
//
// One work item = (K-split, tap, 256 x 256 tile of [C_in, C_out]) computed by a CTA pair (tcgen05 cta_group::2,
// UMMA 256x256x16): CTA r stages input channels [r*128, r*128+128) of x and output channels [r*128, ...) of dz.
// The raw fp32 accumulators of an item go to partial[split][tap][ci][co]; wgrad_reduce_kernel adds the splits in
// a fixed order (bit-reproducible) and removes the loss scale.
#pragma once
#include "ptx.cuh"
#include "tdnn_pair.cuh"

namespace wgrad {

constexpr int TILE = 256;                 // per pair, both dimensions
constexpr int CTA_CH = 128;
constexpr int BOX_CH = 64;                // fp16 channels per TMA box = one 128-byte swizzle row
constexpr int STAGE_ROWS = 128;           // contraction rows per pipeline stage (8 UMMAs)
constexpr int BOX_BYTES = BOX_CH * STAGE_ROWS * 2;        // 16384
constexpr int STAGE_BYTES = 4 * BOX_BYTES;                // x: 2 boxes, dz: 2 boxes
constexpr int N_STAGES = 3;
constexpr int NUM_THREADS = 320;
constexpr int NUM_EPI_WARPS = 8;
constexpr int OFF_BARS = N_STAGES * STAGE_BYTES;          // 196608
constexpr int OFF_TMEM_PTR = OFF_BARS + (2 * N_STAGES + 4) * 8 + 32;
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of dynamic shared memory");

struct WgradArgs {
  int32_t n_chunks;        // R_pad / 128
  int32_t chunks_per_split;
  int32_t k_splits;        // (k_splits - 1) * chunks_per_split < n_chunks
  int32_t taps;
  int32_t dilation;
  int32_t n_mt;            // ceil(C_in / 256)
  int32_t n_nt;            // C_out / 256
  int32_t c_in;            // rows of one tap's gradient that exist (stores beyond are skipped)
  int32_t c_out;
  int32_t lbo_bytes;       // descriptor fields (defaults 16384 / 1024; overridable for diagnostics)
  int32_t sbo_bytes;
  float* partial;          // [k_splits][taps][c_in][c_out]
};

__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmap_x,    // layer input  [R_pad, C_in]  fp16, box [64 x 128 rows]
                  const __grid_constant__ CUtensorMap tmap_dz,   // dL/dz        [R_pad, C_out] fp16, box [64 x 128 rows]
                  const WgradArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar0 = smem_base + OFF_BARS;
  auto full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty = [&](uint32_t s) { return bar0 + 8u * (N_STAGES + s); };
  auto t_full = [&](uint32_t s) { return bar0 + 8u * (2 * N_STAGES + s); };
  auto t_empty = [&](uint32_t s) { return bar0 + 8u * (2 * N_STAGES + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_dz);
    for (uint32_t s = 0; s < N_STAGES; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    for (uint32_t s = 0; s < 2; ++s) { ptx::mbar_init(t_full(s), 1); ptx::mbar_init(t_empty(s), 2 * NUM_EPI_WARPS); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  __syncthreads();
  const uint32_t tmem_base = *tmem_ptr_smem;
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();

  // item = ((split * taps + tap) * n_mt + mt) * n_nt + nt
  const int n_items = args.k_splits * args.taps * args.n_mt * args.n_nt;
  const int half_ctx = (args.taps - 1) >> 1;

  if (warp == 0) {
    // ============================ TMA producer =======================================
    uint32_t st = 0, ph = 0;
    const uint32_t full_leader = ptx::mapa_cluster(full(0), 0);
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int nt = item % args.n_nt;
      const int mt = (item / args.n_nt) % args.n_mt;
      const int tap = (item / (args.n_nt * args.n_mt)) % args.taps;
      const int split = item / (args.n_nt * args.n_mt * args.taps);
      const int c_begin = split * args.chunks_per_split;
      const int c_end = min(args.n_chunks, c_begin + args.chunks_per_split);
      const int ci0 = mt * TILE + int(rank) * CTA_CH;
      const int co0 = nt * TILE + int(rank) * CTA_CH;
      const int row_off = (tap - half_ctx) * args.dilation;
      for (int c = c_begin; c < c_end; ++c) {
        ptx::mbar_wait(empty(st), ph ^ 1u);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(full(st), 2u * STAGE_BYTES);
          const uint32_t dst = smem_base + st * STAGE_BYTES;
          const int r0 = c * STAGE_ROWS;
          ptx::tma_load_2d_2sm(dst, &tmap_x, full_leader + 8u * st, ci0, r0 + row_off);
          ptx::tma_load_2d_2sm(dst + BOX_BYTES, &tmap_x, full_leader + 8u * st, ci0 + BOX_CH, r0 + row_off);
          ptx::tma_load_2d_2sm(dst + 2 * BOX_BYTES, &tmap_dz, full_leader + 8u * st, co0, r0);
          ptx::tma_load_2d_2sm(dst + 3 * BOX_BYTES, &tmap_dz, full_leader + 8u * st, co0 + BOX_CH, r0);
        }
        __syncwarp();
        if (++st == N_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA) =============================
    if (leader) {
      // fp16 x fp16 -> fp32, both operands MN-major (bits 15, 16)
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(TILE, TILE) | (1u << 15) | (1u << 16);
      const uint64_t desc_hi = make_sw128_mnmajor_desc(uint32_t(args.lbo_bytes), uint32_t(args.sbo_bytes));
      uint32_t st = 0, ph = 0, it = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
        const int split = item / (args.n_nt * args.n_mt * args.taps);
        const int c_begin = split * args.chunks_per_split;
        const int c_end = min(args.n_chunks, c_begin + args.chunks_per_split);
        const uint32_t acc = it & 1u;
        ptx::mbar_wait_cluster(t_empty(acc), ((it >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TILE;
        uint32_t accumulate = 0;
        for (int c = c_begin; c < c_end; ++c) {
          ptx::mbar_wait(full(st), ph);
          ptx::tc_fence_after();
          const uint32_t a_addr = smem_base + st * STAGE_BYTES;
          const uint64_t d_a = desc_hi | uint64_t((a_addr >> 4) & 0x3fffu);
          const uint64_t d_b = desc_hi | uint64_t(((a_addr + 2 * BOX_BYTES) >> 4) & 0x3fffu);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < STAGE_ROWS / 16; ++k)              // 16 rows = 2048 bytes: +128 in the address field
              ptx::umma_f16_2sm(d_tmem, d_a + uint64_t(128 * k), d_b + uint64_t(128 * k), idesc, accumulate | uint32_t(k));
            ptx::umma_commit_2sm(empty(st));
          }
          __syncwarp();
          accumulate = 1;
          if (++st == N_STAGES) { st = 0; ph ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit_2sm(t_full(acc));
        __syncwarp();
      }
    }
  } else {
    // ============================ epilogue: raw fp32 accumulators -> partial ==========
    const int q = warp & 3;
    const int colh = (warp - 2) >> 2;
    const uint32_t t_empty_leader = ptx::mapa_cluster(t_empty(0), 0);
    uint32_t it = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
      const int nt = item % args.n_nt;
      const int mt = (item / args.n_nt) % args.n_mt;
      const int st_idx = item / (args.n_nt * args.n_mt);          // split * taps + tap
      const uint32_t acc = it & 1u;
      const int ci = mt * TILE + int(rank) * CTA_CH + q * 32 + lane;
      const int co0 = nt * TILE + colh * 128;
      float* dst = args.partial + (size_t(st_idx) * args.c_in + ci) * args.c_out + co0;
      ptx::mbar_wait(t_full(acc), (it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * TILE + uint32_t(colh) * 128u;
      uint32_t v[2][32];
      ptx::tmem_ld_32x32(t_row, v[0]);
#pragma unroll
      for (int chunk = 0; chunk < 4; ++chunk) {
        ptx::tmem_ld_wait_dep(v[chunk & 1]);
        if (chunk < 3) {
          ptx::tmem_ld_32x32(t_row + (chunk + 1) * 32, v[(chunk + 1) & 1]);
        } else {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(t_empty_leader + 8u * acc);
        }
        if (ci < args.c_in) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(dst + chunk * 32 + g * 4) =
                make_uint4(v[chunk & 1][g * 4], v[chunk & 1][g * 4 + 1], v[chunk & 1][g * 4 + 2], v[chunk & 1][g * 4 + 3]);
        }
      }
    }
  }
  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Tap-reuse variant for layers with temporal context (taps > 1): the x slab [128 + extra rows] of a K chunk is loaded ONCE
// and the taps of a group (<= 3, row span (g-1)*d <= 8) are addressed inside it by advancing the descriptor start by whole
// 128-byte rows -- the hardware applies the 128-byte swizzle on absolute shared-memory address bits, so a start that is
// not a multiple of 8 rows is legal, exactly as in the forward kernel's tap reuse -- each tap accumulating into its own
// 128 TMEM columns (UMMA 256 x 128 x 16).  Operand bytes per FLOP from L2 drop by ~1.9x against one work item per tap, but
// the narrower UMMA reads the A operand twice as often per FLOP from shared memory: measured on B200 it is NOT faster
// (option "wgrad_reuse", default 0; bit-for-bit the same tight parity tests pass with it on).
// One work item = (K-split, tap group, 256 input channels, 128 output channels).
namespace reuse {
constexpr int TILE_N = 128;                // output channels per pair and per tap
constexpr int MAX_GROUP = 3;               // taps per item: 3 x 128 TMEM columns
constexpr int X_BOX_ROWS = STAGE_ROWS + 8; // 136: rows of the x slab (supports (g-1)*d <= 8)
constexpr int X_BOX_BYTES = BOX_CH * X_BOX_ROWS * 2;      // 17408 (a multiple of 1024)
constexpr int Z_BOX_BYTES = BOX_CH * STAGE_ROWS * 2;      // 16384: this CTA's 64 output channels
constexpr int STAGE_BYTES = 2 * X_BOX_BYTES + Z_BOX_BYTES; // 51200
constexpr int N_STAGES = 4;
constexpr int OFF_BARS = N_STAGES * STAGE_BYTES;           // 204800
constexpr int OFF_TMEM_PTR = OFF_BARS + (2 * N_STAGES + 2) * 8 + 32;
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of dynamic shared memory");
}  // namespace reuse

struct WgradReuseArgs {
  int32_t n_chunks, chunks_per_split, k_splits;
  int32_t taps, dilation;
  int32_t n_groups;        // ceil(taps / group)
  int32_t group;           // taps per item (<= 3)
  int32_t n_mt;            // ceil(C_in / 256)
  int32_t n_nt;            // C_out / 128
  int32_t c_in, c_out;
  float* partial;          // [k_splits][taps][c_in][c_out]
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
wgrad_reuse_kernel(const __grid_constant__ CUtensorMap tmap_x,    // [R_pad, C_in]  fp16, box [64 x 136 rows]
                   const __grid_constant__ CUtensorMap tmap_dz,   // [R_pad, C_out] fp16, box [64 x 128 rows]
                   const WgradReuseArgs args) {
  constexpr int TILE_N = reuse::TILE_N, X_BOX_BYTES = reuse::X_BOX_BYTES, Z_BOX_BYTES = reuse::Z_BOX_BYTES;
  constexpr int STAGE_BYTES = reuse::STAGE_BYTES, N_STAGES = reuse::N_STAGES, OFF_BARS = reuse::OFF_BARS;
  constexpr int OFF_TMEM_PTR = reuse::OFF_TMEM_PTR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bar0 = smem_base + OFF_BARS;
  auto full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty = [&](uint32_t s) { return bar0 + 8u * (N_STAGES + s); };
  const uint32_t t_full = bar0 + 8u * (2 * N_STAGES), t_empty = bar0 + 8u * (2 * N_STAGES + 1);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_dz);
    for (uint32_t s = 0; s < N_STAGES; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    ptx::mbar_init(t_full, 1);
    ptx::mbar_init(t_empty, 2 * NUM_EPI_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  __syncthreads();
  const uint32_t tmem_base = *tmem_ptr_smem;
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();

  // item = ((split * n_groups + grp) * n_mt + mt) * n_nt + nt
  const int n_items = args.k_splits * args.n_groups * args.n_mt * args.n_nt;
  const int half_ctx = (args.taps - 1) >> 1;

  if (warp == 0) {
    // ============================ TMA producer =======================================
    uint32_t st = 0, ph = 0;
    const uint32_t full_leader = ptx::mapa_cluster(full(0), 0);
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      const int nt = item % args.n_nt;
      const int mt = (item / args.n_nt) % args.n_mt;
      const int grp = (item / (args.n_nt * args.n_mt)) % args.n_groups;
      const int split = item / (args.n_nt * args.n_mt * args.n_groups);
      const int c_begin = split * args.chunks_per_split;
      const int c_end = min(args.n_chunks, c_begin + args.chunks_per_split);
      const int ci0 = mt * TILE + int(rank) * CTA_CH;
      const int co0 = nt * TILE_N + int(rank) * BOX_CH;
      const int row_off = (grp * args.group - half_ctx) * args.dilation;     // row offset of the group's first tap
      for (int c = c_begin; c < c_end; ++c) {
        ptx::mbar_wait(empty(st), ph ^ 1u);
        if (ptx::elect_one()) {
          if (leader) ptx::mbar_arrive_expect_tx(full(st), 2u * STAGE_BYTES);
          const uint32_t dst = smem_base + st * STAGE_BYTES;
          const int r0 = c * STAGE_ROWS;
          ptx::tma_load_2d_2sm(dst, &tmap_x, full_leader + 8u * st, ci0, r0 + row_off);
          ptx::tma_load_2d_2sm(dst + X_BOX_BYTES, &tmap_x, full_leader + 8u * st, ci0 + BOX_CH, r0 + row_off);
          ptx::tma_load_2d_2sm(dst + 2 * X_BOX_BYTES, &tmap_dz, full_leader + 8u * st, co0, r0);
        }
        __syncwarp();
        if (++st == N_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA) =============================
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(TILE, TILE_N) | (1u << 15) | (1u << 16);   // both operands MN-major
      const uint64_t desc_x = make_sw128_mnmajor_desc(X_BOX_BYTES, 1024);     // LBO: next 64 input channels = next x box
      const uint64_t desc_z = make_sw128_mnmajor_desc(Z_BOX_BYTES, 1024);     // (N = 64 per CTA: one box, LBO unused)
      uint32_t st = 0, ph = 0, it = 0;
      for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
        const int grp = (item / (args.n_nt * args.n_mt)) % args.n_groups;
        const int split = item / (args.n_nt * args.n_mt * args.n_groups);
        const int c_begin = split * args.chunks_per_split;
        const int c_end = min(args.n_chunks, c_begin + args.chunks_per_split);
        const int g_taps = min(args.group, args.taps - grp * args.group);
        ptx::mbar_wait_cluster(t_empty, (it & 1u) ^ 1u);              // the previous item's accumulators are drained
        ptx::tc_fence_after();
        uint32_t accumulate = 0;
        for (int c = c_begin; c < c_end; ++c) {
          ptx::mbar_wait(full(st), ph);
          ptx::tc_fence_after();
          const uint32_t x_addr = smem_base + st * STAGE_BYTES;
          const uint64_t d_z = desc_z | uint64_t(((x_addr + 2 * X_BOX_BYTES) >> 4) & 0x3fffu);
          if (ptx::elect_one()) {
            for (int g = 0; g < g_taps; ++g) {
              const uint64_t d_x = desc_x | uint64_t(((x_addr + uint32_t(g * args.dilation) * 128u) >> 4) & 0x3fffu);
              const uint32_t d_tmem = tmem_base + uint32_t(g) * TILE_N;
#pragma unroll
              for (int k = 0; k < STAGE_ROWS / 16; ++k)
                ptx::umma_f16_2sm(d_tmem, d_x + uint64_t(128 * k), d_z + uint64_t(128 * k), idesc, accumulate | uint32_t(k));
            }
            ptx::umma_commit_2sm(empty(st));
          }
          __syncwarp();
          accumulate = 1;
          if (++st == N_STAGES) { st = 0; ph ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit_2sm(t_full);
        __syncwarp();
      }
    }
  } else {
    // ============================ epilogue: raw fp32 accumulators -> partial ==========
    const int q = warp & 3;
    const int colh = (warp - 2) >> 2;                 // which 64 of the 128 output channels
    const uint32_t t_empty_leader = ptx::mapa_cluster(t_empty, 0);
    uint32_t it = 0;
    for (int item = cluster_id; item < n_items; item += n_clusters, ++it) {
      const int nt = item % args.n_nt;
      const int mt = (item / args.n_nt) % args.n_mt;
      const int grp = (item / (args.n_nt * args.n_mt)) % args.n_groups;
      const int split = item / (args.n_nt * args.n_mt * args.n_groups);
      const int g_taps = min(args.group, args.taps - grp * args.group);
      const int ci = mt * TILE + int(rank) * CTA_CH + q * 32 + lane;
      ptx::mbar_wait(t_full, it & 1u);
      ptx::tc_fence_after();
      for (int g = 0; g < g_taps; ++g) {
        const int tap = grp * args.group + g;
        float* dst = args.partial + ((size_t(split) * args.taps + tap) * args.c_in + ci) * args.c_out + nt * TILE_N + colh * 64;
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(g) * TILE_N + uint32_t(colh) * 64u;
        uint32_t v[2][32];
        ptx::tmem_ld_32x32(t_row, v[0]);
        ptx::tmem_ld_32x32(t_row + 32, v[1]);
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {
          ptx::tmem_ld_wait_dep(v[chunk]);
          if (ci < args.c_in) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              *reinterpret_cast<uint4*>(dst + chunk * 32 + e * 4) =
                  make_uint4(v[chunk][e * 4], v[chunk][e * 4 + 1], v[chunk][e * 4 + 2], v[chunk][e * 4 + 3]);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(t_empty_leader);
    }
  }
  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, 512);
  }
}

// grad[i] = scale * sum_s partial[s][i]   (fixed order; i over taps * c_in * c_out)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, int64_t n4, int32_t splits, float scale) {
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4* p = reinterpret_cast<const float4*>(partial) + i;
  float4 s = p[0];
  for (int k = 1; k < splits; ++k) {
    const float4 v = p[int64_t(k) * n4];
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
  reinterpret_cast<float4*>(grad)[i] = s;
}

}  // namespace wgrad
