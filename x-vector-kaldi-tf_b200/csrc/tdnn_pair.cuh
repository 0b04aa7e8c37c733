// tdnn_pair.cuh -- the fused frame-level TDNN layer kernel, CTA-pair edition (sm_100a).
//
// Replaces, for a whole batch of segments at once, the five TensorFlow ops the reference runs per
// layer (local/tf/models.py:476-480: conv1d/convolution SAME -> bias_add -> relu ->
// batch_norm_wrapper(eval)), and for the LAST frame layer also the first half of statistics
// pooling (tf.nn.moments over time, models.py:485) -- so the [frames, 1536] activation of the last
// layer never touches HBM.
//
// Data layout ("packed rows"): all segments are stacked along the row axis of one [R_pad, C]
// fp16 matrix.  A segment starts at a row that is a multiple of 32 and is followed by >= gap
// all-zero rows (gap >= largest half-context (k-1)/2*d), so a tap reaching outside its segment
// reads zeros -- TF's SAME padding -- and every aligned 32-row block holds rows of ONE segment.
//
// One tile = 256 activation rows x 256 output channels, computed by a CTA PAIR (cluster of 2,
// tcgen05 cta_group::2, UMMA 256x256x16): CTA r stages activation rows [r*128, r*128+128) and
// weight rows (channels) [r*128, r*128+128) of the tile -- half the operand bytes per SM of a
// single-CTA 128x256 tile -- the leader issues the MMAs, completion is multicast to both CTAs.
//   mode 0 (store):  A = activations (M = rows), B = weights.  TMEM lane = row; the epilogue
//                    applies relu(acc+b)*scale+shift, zeroes gap rows, converts to fp16 and
//                    TMA-stores [32 rows x 32 ch] boxes per warp.
//   mode 1 (pool):   operands swapped: A = weights (M = channels), B = activations.  TMEM lane =
//                    channel and the 32 registers of one tcgen05.ld are 32 consecutive FRAMES of
//                    that channel, so the per-block sum / sum of squares is a private register
//                    reduction (no shuffles, fixed order -> bit-reproducible).  Output:
//                    partial[block][{sum,sumsq}][channel] fp32, one aligned 32-row block each.
//   mode 2 (GEMM):   store orientation, no activation function: the raw fp32 accumulators of one K-split go to
//                    out_f32[split][row][channel] (embed_layer-0 as a split-fp16 tensor-core GEMM, see xvec_api.cu).
//   mode 3 (score):  store orientation; self-attention scores of the attention-pooling topology (models.py:1043-1044):
//                    per row, sum over the tile's channels of v_c * tanh(acc + b_c) -> partial[row][2*ch_tile + half]
//                    (a private register reduction: lane = row, registers = channels); nothing else is stored.
// Temporal taps: the activation slab [136 rows x 128 ch] is loaded ONCE per channel chunk and tap
// j is addressed by advancing the UMMA descriptor start by j*d rows (reuse = 1); for half
// contexts > 4 rows one box per tap is loaded instead (reuse = 0).
//
// Warp roles (320 threads): warp 0 = TMA producer (both CTAs), warp 1 = TMEM allocator (both) +
// MMA issuer (leader), warps 2..9 = epilogue (TMEM lane quarter = warp % 4, column half =
// (warp-2)/4).  Pipelines: activation ring, weight ring (full barriers in the leader, empty
// barriers per CTA), double-buffered 128x256 fp32 accumulators in TMEM.
#pragma once
#include "ptx.cuh"

namespace tdnn2 {

constexpr int TILE_ROWS = 256;                     // activation rows per tile (pair)
constexpr int TILE_CH = 256;                       // output channels per tile (pair)
constexpr int CTA_ROWS = 128;
constexpr int CTA_CH = 128;
constexpr int BLOCK_K = 64;                        // fp16 elements = one 128-byte swizzle row (one TMA box / UMMA atom)
constexpr int UMMA_K = 16;                         // K per pipeline stage = ATOMS x 64 (template parameter):
                                                   //   2 atoms (8 UMMAs per barrier round trip) when weights stream,
                                                   //   1 atom when the weights of the channel tile stay resident
constexpr int ACT_BOX_ROWS_PLAIN = CTA_ROWS;
constexpr int ACT_BOX_ROWS_REUSE = CTA_ROWS + 8;   // supports half contexts <= 4 rows
constexpr int MAX_REUSE_HALO = 4;
constexpr int ACT_ATOM_BYTES = ACT_BOX_ROWS_REUSE * 128;    // 17408
constexpr int WGT_ATOM_BYTES = CTA_CH * 128;                // 16384
constexpr int RING_BYTES = 200704;                 // activation + weight rings (carved at run time); the pooled
                                                   // mode has no output staging and may also use the 16 KB after it
constexpr int MAX_STAGES = 8;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = 64 + NUM_EPI_THREADS;  // 320
constexpr int C_CHUNK = 32;                        // output columns per epilogue step
constexpr int C_BUF_BYTES = 32 * C_CHUNK * 2;      // 2048: one warp's [32 rows x 32 ch] fp16 box
constexpr int TMEM_COLS = 512;                     // 2 accumulators x 256 fp32 columns
constexpr int POOL_BLOCK = 32;                     // rows per pooled partial block

constexpr int OFF_RING = 0;
constexpr int OFF_C = OFF_RING + RING_BYTES;
constexpr int OFF_PARAMS = OFF_C + NUM_EPI_WARPS * C_BUF_BYTES;         // + 16384 (one staging box per epilogue warp)
constexpr int PAR_ARRAYS = 4;                      // bias | scale | shift | negative slope
constexpr int OFF_BARS = OFF_PARAMS + 2 * PAR_ARRAYS * TILE_CH * 4;     // + 8192
constexpr int NUM_BARS = 4 * MAX_STAGES + 4;
constexpr int OFF_TMEM_PTR = OFF_BARS + (NUM_BARS + 2) * 8 + 32;        // its own 16-byte slot, away from the barriers
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;                    // + slack for 1024-byte alignment
static_assert(RING_BYTES % 1024 == 0 && OFF_C % 1024 == 0, "swizzled stages must be 1024-byte aligned");
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of dynamic shared memory");

struct PairArgs {
  int32_t n_row_tiles;      // R_pad / 256
  int32_t n_ch_tiles;       // C_out / 256
  int32_t c_chunks;         // C_in_pad / (ATOMS * 64)
  int32_t taps;
  int32_t dilation;
  int32_t c_in_pad;         // column stride between taps in the packed weight matrix
  int32_t reuse;            // 0 / 1 (see header comment)
  int32_t n_act_stages;     // n_act * ATOMS * (reuse ? 17408 : 16384) + n_wgt * ATOMS * 16384 <= RING_BYTES (+ 16384 in mode 1)
  int32_t n_wgt_stages;
  int32_t mode;             // 0 store, 1 pool, 2 raw fp32 split-K partials (embedding GEMM)
  int32_t k_splits;         // mode 2: the K range [0, k_splits * c_chunks * 128) is cut into k_splits work items
  int32_t n_rows;           // mode 2: rows that exist (stores beyond are skipped)
  int32_t prefetch;         // 1: warm L2 with the activation boxes of this cluster's next work item
  int32_t wgt_resident;     // 1: n_wgt_stages == taps * c_chunks; every cluster owns ONE channel tile, loads its
                            //    weights once and streams activations only (halves the bytes an SM must ingest)
  int32_t c_out;
  const float* bias;        // [C_out]  conv bias b
  const float* scale;       // [C_out]  gamma * rsqrt(var + eps)
  const float* shift;       // [C_out]  beta - mean * scale
  const float* alpha;       // [C_out]  LEAKY kernels: negative slope (0.2 everywhere = leaky_relu models.py:912; PReLU's
                            //          learned per-channel slope tf_block.py:38-47)
  const uint8_t* row_valid; // [R_pad]     mode 0: 1 = row belongs to a segment, 0 = gap / tail
  const uint8_t* blk_valid; // [R_pad/32]  mode 1: valid rows in the block (they are its first rows)
  float* partial;           // [R_pad/32][2][C_out]  mode 1;  [R_pad][2 * n_ch_tiles]  mode 3 (score partials)
  float* out_f32;           // [k_splits][n_rows][C_out]  mode 2: raw accumulators
  uint32_t* overflow_flag;  // overflow_bit is OR-ed in if an fp16 output overflowed to inf
  uint32_t overflow_bit;    // 0 = 1u
  int32_t split_c;          // > 0: split-precision operands.  The input rows hold [hi (split_c) | ... | lo at split_lo_off]: two fp16
  int32_t split_lo_off;     //   terms per value (x = hi + lo to ~2^-22), the weights [w_hi | w_lo | w_hi] per tap, and the K loop walks
                            //   the VIRTUAL columns [hi | hi | lo] (3 * split_c) = hi*w_hi + hi*w_lo + lo*w_hi in one accumulator
  float acc_scale;          // 0 = 1: the pre-activation is acc * acc_scale + bias (acc_scale = 2^e when the INPUT rows are stored
                            // divided by 2^e: the fp16 range rescue of xvec_api.cu; exact, powers of two)
  long long* trace;         // diagnostics (tools/trace_tiles.py): [cluster][rank][TRACE_TILES][8] SM clock stamps, or null
};
constexpr int TRACE_TILES = 16;
// slots: 0 producer first load issued, 1 producer last load issued, 2 MMA start (accumulator free),
//        3 MMA all issued, 4 epilogue sees accumulator, 5 epilogue released accumulator, 6 epilogue done
__device__ __forceinline__ void trace_stamp(const PairArgs& a, int cluster, uint32_t rank, uint32_t it, int slot) {
  if (a.trace != nullptr && it < TRACE_TILES)
    a.trace[((size_t(cluster) * 2 + rank) * TRACE_TILES + it) * 8 + slot] = clock64();
}

// One tcgen05.ld chunk of the store epilogue: 32 channels of one row -> 16 packed half2.
// act(a): relu (models.py:479), or with LEAKY max(0,a) + alpha*min(0,a) (tf_block.py:47 / tf.nn.leaky_relu)
template <bool LEAKY>
__device__ __forceinline__ float act_bn(float acc, float b, float sc, float sh, float al, float as) {
  const float a = fmaf(acc, as, b);                 // as = 1: acc + b, bit for bit
  const float r = LEAKY ? fmaf(al, fminf(a, 0.f), fmaxf(a, 0.f)) : fmaxf(a, 0.f);
  return fmaf(r, sc, sh);                           // BatchNorm eval branch folded: r * inv + shift (tf_block.py:26)
}

// s_par: shared-memory address of this tile's [bias(256) | scale(256) | shift(256) | alpha(256)] floats.
template <bool LEAKY, bool SPLIT = false>
__device__ __forceinline__ void epi_store_math(const uint32_t (&v)[32], uint32_t s_par, int c, bool valid,
                                               uint32_t (&p)[16], uint32_t& hmax, float as, uint32_t* plo = nullptr) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 b4 = ptx::lds_f4(s_par + uint32_t(c + g * 4) * 4u);
    const float4 s4 = ptx::lds_f4(s_par + uint32_t(TILE_CH + c + g * 4) * 4u);
    const float4 h4 = ptx::lds_f4(s_par + uint32_t(2 * TILE_CH + c + g * 4) * 4u);
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (LEAKY) a4 = ptx::lds_f4(s_par + uint32_t(3 * TILE_CH + c + g * 4) * 4u);
    const float y0 = act_bn<LEAKY>(__uint_as_float(v[g * 4 + 0]), b4.x, s4.x, h4.x, a4.x, as);
    const float y1 = act_bn<LEAKY>(__uint_as_float(v[g * 4 + 1]), b4.y, s4.y, h4.y, a4.y, as);
    const float y2 = act_bn<LEAKY>(__uint_as_float(v[g * 4 + 2]), b4.z, s4.z, h4.z, a4.z, as);
    const float y3 = act_bn<LEAKY>(__uint_as_float(v[g * 4 + 3]), b4.w, s4.w, h4.w, a4.w, as);
    const uint32_t p0 = ptx::pack_half2(y0, y1), p1 = ptx::pack_half2(y2, y3);
    hmax = ptx::habs2_max(ptx::habs2_max(hmax, p0), p1);
    p[g * 2 + 0] = valid ? p0 : 0u;                  // gap rows stay exact zeros
    p[g * 2 + 1] = valid ? p1 : 0u;
    if (SPLIT) {                                     // second term: what the fp16 rounding of y dropped
      const float2 h0 = ptx::unpack_half2(p0), h1 = ptx::unpack_half2(p1);
      const uint32_t l0 = ptx::pack_half2(y0 - h0.x, y1 - h0.y), l1 = ptx::pack_half2(y2 - h1.x, y3 - h1.y);
      plo[g * 2 + 0] = valid ? l0 : 0u;
      plo[g * 2 + 1] = valid ? l1 : 0u;
    }
  }
}

// Walks one cluster's work items (item, item + step, ...) keeping (row tile, K split, channel tile) as mixed-radix
// digits, so that no integer division sits on any role's per-tile path (they cost ~1500 cycles per tile there).
struct TileCursor {
  int item, row, split, ch;
  int step, d_row, d_split, d_ch, n_ch, n_split;
  __device__ __forceinline__ void init(int first, int step_, int n_ch_, int n_split_, bool resident, int my_ch) {
    item = first; step = step_; n_ch = n_ch_; n_split = n_split_;
    if (resident) {                       // items are row tiles of a fixed channel tile
      row = first; split = 0; ch = my_ch; d_row = step_; d_split = 0; d_ch = 0;
    } else {                              // item = (row * n_split + split) * n_ch + ch
      ch = first % n_ch_; split = (first / n_ch_) % n_split_; row = first / (n_ch_ * n_split_);
      d_ch = step_ % n_ch_; d_split = (step_ / n_ch_) % n_split_; d_row = step_ / (n_ch_ * n_split_);
    }
  }
  __device__ __forceinline__ void next() {
    item += step; ch += d_ch; split += d_split; row += d_row;
    if (ch >= n_ch) { ch -= n_ch; ++split; }
    if (split >= n_split) { split -= n_split; ++row; }
  }
};

// STATS (mode 0 only, training): the epilogue also leaves the column sums of every [32 rows x 32 channels] box it stores --
// of the fp16 values as stored -- in partial[row_block][{sum, sum of squares}][channel]: the batch-norm moments of the training
// branch (tf_block.py:19) and the pooling sums of the last layer come out of the layer kernel instead of a second pass over HBM.
// SPLIT (mode 0 only): the output is stored as two fp16 terms, hi = rn(y) at column c and lo = rn(y - hi) at column C_out + c
// (the next layer's split-precision input); the caller's tmap_out then spans 2 * C_out columns.
template <int MODE, int ATOMS, bool LEAKY, bool STATS = false, bool SPLIT = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
tdnn_pair_kernel(const __grid_constant__ CUtensorMap tmap_act,   // activations in  [R_pad, C_in_pad] fp16
                 const __grid_constant__ CUtensorMap tmap_wgt,   // weights [C_out, taps*C_in_pad] fp16 (K-major)
                 const __grid_constant__ CUtensorMap tmap_out,   // activations out [R_pad, C_out] fp16 (mode 0)
                 const PairArgs args) {
  constexpr int STAGE_K = ATOMS * BLOCK_K;
  constexpr int WGT_STAGE_BYTES = ATOMS * WGT_ATOM_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t sAct = smem_base + OFF_RING;
  // activation atom stride: 136 rows when taps are addressed inside the slab, else 128 (both multiples of 1024 B)
  const uint32_t ACT_ATOM_STRIDE = args.reuse ? uint32_t(ACT_ATOM_BYTES) : uint32_t(ACT_BOX_ROWS_PLAIN * 128);
  const uint32_t ACT_STAGE_BYTES = ATOMS * ACT_ATOM_STRIDE;
  const uint32_t sWgt = sAct + uint32_t(args.n_act_stages) * ACT_STAGE_BYTES;
  const uint32_t bar0 = smem_base + OFF_BARS;
  auto act_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto act_empty = [&](uint32_t s) { return bar0 + 8u * (MAX_STAGES + s); };
  auto wgt_full = [&](uint32_t s) { return bar0 + 8u * (2 * MAX_STAGES + s); };
  auto wgt_empty = [&](uint32_t s) { return bar0 + 8u * (3 * MAX_STAGES + s); };
  auto t_full = [&](uint32_t s) { return bar0 + 8u * (4 * MAX_STAGES + s); };
  auto t_empty = [&](uint32_t s) { return bar0 + 8u * (4 * MAX_STAGES + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;          // warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_act);
    ptx::prefetch_tmap(&tmap_wgt);
    if (MODE == 0) ptx::prefetch_tmap(&tmap_out);
    for (uint32_t s = 0; s < MAX_STAGES; ++s) {
      ptx::mbar_init(act_full(s), 1); ptx::mbar_init(act_empty(s), 1);
      ptx::mbar_init(wgt_full(s), 1); ptx::mbar_init(wgt_empty(s), 1);
    }
    for (uint32_t s = 0; s < 2; ++s) { ptx::mbar_init(t_full(s), 1); ptx::mbar_init(t_empty(s), 2 * NUM_EPI_WARPS); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_2sm(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), TMEM_COLS);
    ptx::tmem_relinquish_2sm();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();                    // barrier inits + TMEM allocation visible to both CTAs
  ptx::tc_fence_after();
  __syncthreads();                            // (the cluster barrier already orders this; racecheck only models bar.sync)
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Programmatic dependent launch: everything above (barriers, TMEM, tensor-map prefetch) overlapped the tail of
  // the previous kernel; its output is read -- and ours written -- only after this point.
  cudaTriggerProgrammaticLaunchCompletion();
  cudaGridDependencySynchronize();

  // Work items of this cluster.  Streaming weights: item = tile index (channel tile fastest, so that
  // neighbouring clusters share activation rows in L2).  Resident weights: the cluster keeps channel tile
  // cluster_id % n_ch_tiles and its items are row tiles.
  const bool resident = args.wgt_resident != 0;
  const int my_ch_tile = cluster_id % args.n_ch_tiles;
  const int k_splits = MODE == 2 ? args.k_splits : 1;
  const int n_items = resident ? args.n_row_tiles : args.n_row_tiles * k_splits * args.n_ch_tiles;
  const int item_first = resident ? cluster_id / args.n_ch_tiles : cluster_id;
  const int item_step = resident ? n_clusters / args.n_ch_tiles : n_clusters;     // host: n_clusters % n_ch_tiles == 0
  TileCursor cur0;
  cur0.init(item_first, item_step, args.n_ch_tiles, k_splits, resident, my_ch_tile);
  const int half_ctx = (args.taps - 1) >> 1;
  const int halo = half_ctx * args.dilation;
  const bool reuse = args.reuse != 0;
  const uint32_t act_box_bytes = (reuse ? ACT_BOX_ROWS_REUSE : ACT_BOX_ROWS_PLAIN) * 128u;
  const uint32_t n_act = uint32_t(args.n_act_stages), n_wgt = uint32_t(args.n_wgt_stages);

  if (warp == 0) {
    // ============================ TMA producer (warp 0 of each CTA; one elected lane issues) ====
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;         // stage index / phase of each ring
    const uint32_t act_full_leader = ptx::mapa_cluster(act_full(0), 0);   // transaction bytes are counted there
    const uint32_t wgt_full_leader = ptx::mapa_cluster(wgt_full(0), 0);
    uint32_t pit = 0;
    for (TileCursor tc = cur0; tc.item < n_items; tc.next(), ++pit) {
      const int r0 = tc.row * TILE_ROWS + int(rank) * CTA_ROWS;
      const int c0 = tc.ch * TILE_CH + int(rank) * CTA_CH;
      const bool load_wgt = !resident || pit == 0;
      const int k_base = tc.split * args.c_chunks;                     // first K chunk of this item (mode 2 split-K)
      for (int cc = 0; cc < args.c_chunks; ++cc) {
        for (int j = 0; j < args.taps; ++j) {
          if (!reuse || j == 0) {
            ptx::mbar_wait(act_empty(sa), pa ^ 1u);
            if (ptx::elect_one()) {
              if (cc == 0 && j == 0) trace_stamp(args, cluster_id, rank, pit, 0);
              if (cc == args.c_chunks - 1) trace_stamp(args, cluster_id, rank, pit, 1);
              if (leader) ptx::mbar_arrive_expect_tx(act_full(sa), 2u * ATOMS * act_box_bytes);   // both CTAs' boxes
              const int row = reuse ? (r0 - halo) : (r0 + (j - half_ctx) * args.dilation);
#pragma unroll
              for (int h = 0; h < ATOMS; ++h) {
                int col = (k_base + cc) * STAGE_K + h * BLOCK_K;
                if (args.split_c > 0)                  // virtual [hi | hi | lo] -> stored [hi | .. | lo]
                  col = col < args.split_c ? col : (col < 2 * args.split_c ? col - args.split_c : col - 2 * args.split_c + args.split_lo_off);
                ptx::tma_load_2d_2sm(sAct + sa * ACT_STAGE_BYTES + h * ACT_ATOM_STRIDE, &tmap_act, act_full_leader + 8u * sa, col, row);
              }
              // the same boxes of this cluster's NEXT item -> L2 now: with resident weights the ring holds too few
              // bytes to cover an HBM round trip (~2000 cycles), an L2 hit (~700) it does cover
              if (args.prefetch && tc.item + tc.step < n_items) {
                TileCursor nx = tc;
                nx.next();
                const int row_next = row + (nx.row - tc.row) * TILE_ROWS;
#pragma unroll
                for (int h = 0; h < ATOMS; ++h) ptx::tma_prefetch_l2_2d(&tmap_act, cc * STAGE_K + h * BLOCK_K, row_next);
              }
            }
            __syncwarp();
            if (++sa == n_act) { sa = 0; pa ^= 1u; }
          }
          if (load_wgt) {
            ptx::mbar_wait(wgt_empty(sb), pb ^ 1u);
            if (ptx::elect_one()) {
              if (leader) ptx::mbar_arrive_expect_tx(wgt_full(sb), 2u * WGT_STAGE_BYTES);
#pragma unroll
              for (int h = 0; h < ATOMS; ++h)
                ptx::tma_load_2d_2sm(sWgt + sb * WGT_STAGE_BYTES + h * WGT_ATOM_BYTES, &tmap_wgt, wgt_full_leader + 8u * sb,
                                     j * args.c_in_pad + (k_base + cc) * STAGE_K + h * BLOCK_K, c0);
            }
            __syncwarp();
          }
          if (++sb == n_wgt) { sb = 0; pb ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (warp 1 of the leader CTA; one elected lane issues) =
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(TILE_ROWS, TILE_CH);   // 256 x 256 either orientation
      // descriptor without the start address: SBO = 1024 B, version 1, SWIZZLE_128B; the base-offset bits
      // stay 0: the hardware swizzles on absolute smem address bits, which is what lets a tap start at
      // any 128-byte row of the activation slab
      const uint64_t desc_hi = ptx::make_sw128_kmajor_desc(0);
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
      if (ATOMS == 1 && resident) {
        // Weight-stationary schedule (host guarantees taps == 1, an even number of 64-wide chunks, weight stage
        // index == chunk index).  Two activation stages per trip: 8 UMMAs amortise the fixed cost of a trip
        // (barrier round trip, election, descriptor arithmetic ~ 600 cycles), which 4 UMMAs (512 cycles) do not.
        for (int item = item_first; item < n_items; item += item_step, ++it) {
          const uint32_t acc = it & 1u;
          ptx::mbar_wait_cluster(t_empty(acc), ((it >> 1) & 1u) ^ 1u);
          ptx::tc_fence_after();
          if (lane == 0) trace_stamp(args, cluster_id, rank, it, 2);
          const uint32_t d_tmem = tmem_base + acc * TILE_CH;
          for (int cc = 0; cc < args.c_chunks; cc += 2) {
            const bool wrap = sa + 1 == n_act;
            const uint32_t sa2 = wrap ? 0u : sa + 1u, pa2 = wrap ? pa ^ 1u : pa;
            ptx::mbar_wait(act_full(sa), pa);
            ptx::mbar_wait(act_full(sa2), pa2);
            if (it == 0) { ptx::mbar_wait(wgt_full(cc), 0); ptx::mbar_wait(wgt_full(cc + 1), 0); }
            ptx::tc_fence_after();
            const uint64_t d_a0 = desc_hi | uint64_t(((sAct + sa * ACT_STAGE_BYTES) >> 4) & 0x3fffu);
            const uint64_t d_a1 = desc_hi | uint64_t(((sAct + sa2 * ACT_STAGE_BYTES) >> 4) & 0x3fffu);
            const uint64_t d_w0 = desc_hi | uint64_t(((sWgt + uint32_t(cc) * WGT_STAGE_BYTES) >> 4) & 0x3fffu);
            const uint64_t d_w1 = d_w0 + uint64_t(WGT_STAGE_BYTES >> 4);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                if (MODE == 1) ptx::umma_f16_2sm(d_tmem, d_w0 + 2u * k, d_a0 + 2u * k, idesc, uint32_t(cc | k));
                else ptx::umma_f16_2sm(d_tmem, d_a0 + 2u * k, d_w0 + 2u * k, idesc, uint32_t(cc | k));
              }
              ptx::umma_commit_2sm(act_empty(sa));
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                if (MODE == 1) ptx::umma_f16_2sm(d_tmem, d_w1 + 2u * k, d_a1 + 2u * k, idesc, 1u);
                else ptx::umma_f16_2sm(d_tmem, d_a1 + 2u * k, d_w1 + 2u * k, idesc, 1u);
              }
              ptx::umma_commit_2sm(act_empty(sa2));
            }
            __syncwarp();
            sa = sa2; pa = pa2;
            if (++sa == n_act) { sa = 0; pa ^= 1u; }
          }
          if (ptx::elect_one()) ptx::umma_commit_2sm(t_full(acc));
          __syncwarp();
          if (lane == 0) trace_stamp(args, cluster_id, rank, it, 3);
        }
      } else
      for (int item = item_first; item < n_items; item += item_step, ++it) {
        const uint32_t acc = it & 1u;
        const bool wait_wgt = !resident || it == 0;                    // resident weights land once, in phase 0
        ptx::mbar_wait_cluster(t_empty(acc), ((it >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained it
        ptx::tc_fence_after();
        if (lane == 0) trace_stamp(args, cluster_id, rank, it, 2);
        const uint32_t d_tmem = tmem_base + acc * TILE_CH;
        uint32_t accumulate = 0;
        for (int cc = 0; cc < args.c_chunks; ++cc) {
          for (int j = 0; j < args.taps; ++j) {
            if (!reuse || j == 0) ptx::mbar_wait(act_full(sa), pa);
            if (wait_wgt) ptx::mbar_wait(wgt_full(sb), pb);
            ptx::tc_fence_after();
            const uint32_t act_addr = sAct + sa * ACT_STAGE_BYTES + (reuse ? uint32_t(j * args.dilation) * 128u : 0u);
            const uint32_t wgt_addr = sWgt + sb * WGT_STAGE_BYTES;
            const uint64_t d_act = desc_hi | uint64_t((act_addr >> 4) & 0x3fffu);
            const uint64_t d_wgt = desc_hi | uint64_t((wgt_addr >> 4) & 0x3fffu);
            if (ptx::elect_one()) {
#pragma unroll
              for (int h = 0; h < ATOMS; ++h) {
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {         // +32 bytes of K per UMMA: +2 in the address field
                  const uint64_t da = d_act + uint64_t(h * (ACT_ATOM_STRIDE >> 4) + 2 * k);
                  const uint64_t dw = d_wgt + uint64_t(h * (WGT_ATOM_BYTES >> 4) + 2 * k);
                  if (MODE == 1) ptx::umma_f16_2sm(d_tmem, dw, da, idesc, accumulate | uint32_t(h | k));
                  else ptx::umma_f16_2sm(d_tmem, da, dw, idesc, accumulate | uint32_t(h | k));
                }
              }
              if (!resident) ptx::umma_commit_2sm(wgt_empty(sb));      // weight slot free in both CTAs
              if (!reuse || j == args.taps - 1) ptx::umma_commit_2sm(act_empty(sa));
            }
            __syncwarp();
            accumulate = 1;
            if (++sb == n_wgt) { sb = 0; pb ^= 1u; }
            if (!reuse || j == args.taps - 1) {
              if (++sa == n_act) { sa = 0; pa ^= 1u; }
            }
          }
        }
        if (ptx::elect_one()) ptx::umma_commit_2sm(t_full(acc));       // accumulator ready in both CTAs
        __syncwarp();
        if (lane == 0) trace_stamp(args, cluster_id, rank, it, 3);
      }
    }
  } else {
    // ============================ epilogue (8 warps per CTA) ===========================
    const int e = warp - 2;                          // 0..7
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int colh = e >> 2;                         // which 128-column half of the accumulator
    const int te = threadIdx.x - 64;                 // 0..255
    const uint32_t t_empty_leader = ptx::mapa_cluster(t_empty(0), 0);   // + 8 * acc
    const float acc_scale = args.acc_scale != 0.f ? args.acc_scale : 1.f;
    uint32_t it = 0;
    if (MODE == 2) {
      for (TileCursor tc = cur0; tc.item < n_items; tc.next(), ++it) {
        const uint32_t acc = it & 1u;
        const int row = tc.row * TILE_ROWS + int(rank) * CTA_ROWS + q * 32 + lane;
        const int ch0 = tc.ch * TILE_CH + colh * 128;
        float* dst = args.out_f32 + (size_t(tc.split) * args.n_rows + row) * args.c_out + ch0;
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
          if (row < args.n_rows) {
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<uint4*>(dst + chunk * C_CHUNK + g * 4) =
                  make_uint4(v[chunk & 1][g * 4], v[chunk & 1][g * 4 + 1], v[chunk & 1][g * 4 + 2], v[chunk & 1][g * 4 + 3]);
          }
        }
      }
    } else if (MODE == 3) {
      // bias -> b, scale -> v of "attention/" (models.py:1038-1039), staged per tile in shared memory
      const int n_part = 2 * args.n_ch_tiles;
      for (TileCursor tc = cur0; tc.item < n_items; tc.next(), ++it) {
        const uint32_t acc = it & 1u;
        const int row = tc.row * TILE_ROWS + int(rank) * CTA_ROWS + q * 32 + lane;
        const uint32_t s_par = smem_base + OFF_PARAMS + acc * (PAR_ARRAYS * TILE_CH * 4);
        {
          const int ch = tc.ch * TILE_CH + te;
          ptx::sts_f(s_par + uint32_t(te) * 4u, __ldg(args.bias + ch));
          ptx::sts_f(s_par + uint32_t(TILE_CH + te) * 4u, __ldg(args.scale + ch));
        }
        ptx::named_bar_sync(1, NUM_EPI_THREADS);       // parameters of this tile visible (double-buffered by accumulator)
        ptx::mbar_wait(t_full(acc), (it >> 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * TILE_CH + uint32_t(colh) * 128u;
        uint32_t v[2][32];
        ptx::tmem_ld_32x32(t_row, v[0]);
        float sum[4] = {0.f, 0.f, 0.f, 0.f};
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
          const int c = colh * 128 + chunk * C_CHUNK;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b4 = ptx::lds_f4(s_par + uint32_t(c + g * 4) * 4u);
            const float4 v4 = ptx::lds_f4(s_par + uint32_t(TILE_CH + c + g * 4) * 4u);
            const float bb[4] = {b4.x, b4.y, b4.z, b4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float x = fmaf(__uint_as_float(v[chunk & 1][g * 4 + k]), acc_scale, bb[k]);
              const float th = 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f);      // tanh(x); saturates cleanly at +-1
              sum[k] = fmaf(vv[k], th, sum[k]);
            }
          }
        }
        args.partial[size_t(row) * n_part + tc.ch * 2 + colh] = (sum[0] + sum[1]) + (sum[2] + sum[3]);
      }
    } else if (MODE == 0) {
      const uint32_t sC = smem_base + OFF_C + uint32_t(e) * C_BUF_BYTES;
      const uint32_t swz = (uint32_t(lane) >> 1) & 3u;   // SWIZZLE_64B phase of this row in the staging box
      uint32_t hmax = 0;
      // Per-tile parameters (bias | scale | shift of the tile's 256 channels) live in shared memory, double-buffered
      // by accumulator; the NEXT tile's are fetched into registers at the top of a tile and stored at its end, so
      // no global-load latency sits between two tiles.
      auto fetch_params = [&](const TileCursor& tc, float& b, float& sc, float& sh, float& al, uint32_t& valid) {   // loads only
        const int ch = tc.ch * TILE_CH + te;
        b = __ldg(args.bias + ch); sc = __ldg(args.scale + ch); sh = __ldg(args.shift + ch);
        if (LEAKY) al = __ldg(args.alpha + ch);
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
      if (item_first < n_items) {
        fetch_params(cur0, nb, nsc, nsh, nal, nvalid);
        store_params(0, nb, nsc, nsh, nal);
      }
      TileCursor nx = cur0;
      for (TileCursor tc = cur0; tc.item < n_items; tc.next(), ++it) {
        const uint32_t acc = it & 1u;
        const int r_cta = tc.row * TILE_ROWS + int(rank) * CTA_ROWS;
        const int ch0 = tc.ch * TILE_CH;
        const uint32_t s_par = smem_base + OFF_PARAMS + acc * (PAR_ARRAYS * TILE_CH * 4);
        const bool valid = nvalid != 0;
        ptx::named_bar_sync(1, NUM_EPI_THREADS);       // this tile's parameters visible; the other buffer is free
        if (e == 0 && lane == 0) trace_stamp(args, cluster_id, rank, it, 5);
        nx.next();                                     // nx is always one item ahead of tc
        const bool has_next = nx.item < n_items;
        if (has_next) fetch_params(nx, nb, nsc, nsh, nal, nvalid);
        ptx::mbar_wait(t_full(acc), (it >> 1) & 1u);
        ptx::tc_fence_after();
        if (e == 0 && lane == 0) trace_stamp(args, cluster_id, rank, it, 4);
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * TILE_CH + uint32_t(colh) * 128u;
        uint32_t v[2][32];
        ptx::tmem_ld_32x32(t_row, v[0]);
#pragma unroll
        for (int chunk = 0; chunk < 4; ++chunk) {
          ptx::tmem_ld_wait_dep(v[chunk & 1]);
          if (chunk < 3) {
            ptx::tmem_ld_32x32(t_row + (chunk + 1) * C_CHUNK, v[(chunk + 1) & 1]);
          } else {                                   // all of this warp's loads done: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(t_empty_leader + 8u * acc);
          }
          uint32_t p[16], plo[SPLIT ? 16 : 1];
          epi_store_math<LEAKY, SPLIT>(v[chunk & 1], s_par, colh * 128 + chunk * C_CHUNK, valid, p, hmax, acc_scale, plo);
          const uint32_t buf = sC;                        // one box per warp: the previous chunk's store has had the
          if (lane == 0) ptx::tma_store_wait_read<0>();   // whole tcgen05.ld + math of this chunk to read it
          __syncwarp();
          const uint32_t dst = buf + uint32_t(lane) * (C_CHUNK * 2);
#pragma unroll
          for (uint32_t c16 = 0; c16 < 4; ++c16) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((c16 ^ swz) << 4)),
                         "r"(p[c16 * 4 + 0]), "r"(p[c16 * 4 + 1]), "r"(p[c16 * 4 + 2]), "r"(p[c16 * 4 + 3])
                         : "memory");
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_out, buf, ch0 + colh * 128 + chunk * C_CHUNK, r_cta + q * 32);
            ptx::tma_store_commit();
          }
          if (SPLIT) {                                    // the lo terms: same box, once the hi store has read it
            if (lane == 0) ptx::tma_store_wait_read<0>();
            __syncwarp();
#pragma unroll
            for (uint32_t c16 = 0; c16 < 4; ++c16) {
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((c16 ^ swz) << 4)),
                           "r"(plo[c16 * 4 + 0]), "r"(plo[c16 * 4 + 1]), "r"(plo[c16 * 4 + 2]), "r"(plo[c16 * 4 + 3])
                           : "memory");
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&tmap_out, buf, args.c_out + ch0 + colh * 128 + chunk * C_CHUNK, r_cta + q * 32);
              ptx::tma_store_commit();
            }
          }
          if (STATS) {
            // lane = channel: read the staged box column-wise (row r is one 64-byte line whose 16-byte pieces are XOR-swizzled
            // by (r >> 1) & 3, so the 32 lanes of a step read one whole line: conflict free), fixed order over the rows
            float c1 = 0.f, c2 = 0.f;
            const uint32_t piece = uint32_t(lane) >> 3, within = (uint32_t(lane) & 7u) * 2u;
#pragma unroll
            for (uint32_t r = 0; r < 32; ++r) {
              const uint32_t h = ptx::lds_u16(buf + r * (C_CHUNK * 2) + ((piece ^ ((r >> 1) & 3u)) << 4) + within);
              const float f = __half2float(__ushort_as_half(static_cast<unsigned short>(h)));
              c1 += f;
              c2 = fmaf(f, f, c2);
            }
            float* dstp = args.partial + size_t((r_cta + q * 32) >> 5) * 2 * args.c_out + ch0 + colh * 128 + chunk * C_CHUNK + lane;
            dstp[0] = c1;
            dstp[args.c_out] = c2;
          }
        }
        if (has_next) store_params(acc ^ 1u, nb, nsc, nsh, nal);
        if (e == 0 && lane == 0) trace_stamp(args, cluster_id, rank, it, 6);
        if (lane == 0 && args.trace != nullptr && it < TRACE_TILES)      // slot 7: when the slowest epilogue warp finished
          atomicMax(reinterpret_cast<unsigned long long*>(args.trace) + ((size_t(cluster_id) * 2 + rank) * TRACE_TILES + it) * 8 + 7,
                    static_cast<unsigned long long>(clock64()));
      }
      if (lane == 0) ptx::tma_store_wait_all<0>();
      if ((hmax & 0x7fffu) >= 0x7c00u || ((hmax >> 16) & 0x7fffu) >= 0x7c00u)
        atomicOr(args.overflow_flag, args.overflow_bit ? args.overflow_bit : 1u);
    } else {
      for (TileCursor tc = cur0; tc.item < n_items; tc.next(), ++it) {
        const uint32_t acc = it & 1u;
        const int r_tile = tc.row * TILE_ROWS;
        const int ch = tc.ch * TILE_CH + int(rank) * CTA_CH + q * 32 + lane;
        const float b = __ldg(args.bias + ch), sc = __ldg(args.scale + ch), sh = __ldg(args.shift + ch);
        const float al = LEAKY ? __ldg(args.alpha + ch) : 0.f;
        const int blk0 = (r_tile + colh * 128) / POOL_BLOCK;
        const uint32_t nv4 = *reinterpret_cast<const uint32_t*>(args.blk_valid + blk0);   // 4 blocks, 1 byte each
        ptx::mbar_wait(t_full(acc), (it >> 1) & 1u);
        ptx::tc_fence_after();
        if (e == 0 && lane == 0) trace_stamp(args, cluster_id, rank, it, 4);
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
            if (e == 0 && lane == 0) trace_stamp(args, cluster_id, rank, it, 5);
          }
          const int nv = int((nv4 >> (8 * chunk)) & 0xffu);          // warp-uniform
          if (nv > 0) {
            float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
            if (nv >= POOL_BLOCK) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float y = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][i]), b, sc, sh, al, acc_scale);
                s1[i & 3] += y;
                s2[i & 3] = fmaf(y, y, s2[i & 3]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float y = act_bn<LEAKY>(__uint_as_float(v[chunk & 1][i]), b, sc, sh, al, acc_scale);
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
        if (e == 0 && lane == 0) trace_stamp(args, cluster_id, rank, it, 6);
      }
    }
  }
  __syncwarp();

  ptx::tc_fence_before();
  ptx::cluster_sync_all();                    // the peer may still be reading our smem / signalling our barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

}  // namespace tdnn2
