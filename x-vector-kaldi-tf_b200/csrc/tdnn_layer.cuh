// tdnn_layer.cuh -- ONE fused kernel per frame-level TDNN layer (sm_100a).
//
// Replaces, for a whole batch of segments at once, the five TensorFlow ops the reference runs
// per layer (local/tf/models.py:476-480):
//     conv = tf.nn.conv1d / tf.nn.convolution(h, w, SAME[, dilation])     (implicit GEMM, tcgen05)
//     h    = tf.nn.bias_add(conv, b); h = tf.nn.relu(h)                   (epilogue)
//     h    = batch_norm_wrapper(h, is_training=False)                     (epilogue: *inv + shift)
//
// Data layout ("packed rows"): all segments of the batch are stacked along the row axis of one
// [R_pad, C] fp16 matrix with `gap` all-zero rows before, between and after segments.
// gap >= the largest half-context (k-1)/2*d of any layer, so a tap that reaches outside its
// segment reads a zero row -- exactly TF's SAME zero padding -- and an M-tile of 128 rows may
// span several segments: ragged batches need no per-utterance padding.  Every layer writes its
// gap rows as exact zeros (row_valid mask), because they are the next layer's padding.
//
// GEMM view per tile: D[128 rows, 256 out-channels] = sum over (channel chunk cc, tap j) of
//     A_j[128, 64] (rows m0 + (j-half)*d .., channels cc*64 ..)  x  W[256, 64]^T
// A and W tiles are K-major, 128-byte rows, SWIZZLE_128B as written by TMA.  Two ways to get A_j:
//   reuse = 0 : one TMA box [64 ch x 128 rows] per (cc, j) at row m0 + (j-half)*d
//   reuse = 1 : one TMA box [64 ch x 136 rows] per cc at row m0 - halo; tap j is addressed inside
//               the slab by advancing the UMMA descriptor start by j*d rows (128 B each)
//               -> activation traffic from L2 drops by the number of taps.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quarter = warp % 4).  Pipelines: A ring, B ring
// (full/empty mbarriers), double-buffered 128x256 fp32 accumulator in TMEM (full/empty),
// double-buffered 8 KB output staging for TMA stores.  Persistent grid, static tile schedule.
#pragma once
#include "ptx.cuh"

namespace tdnn {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;                        // fp16 elements = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_BOX_ROWS_PLAIN = BLOCK_M;
constexpr int A_BOX_ROWS_REUSE = BLOCK_M + 8;      // supports halo <= 4
constexpr int MAX_REUSE_HALO = 4;
constexpr int A_STAGE_BYTES = A_BOX_ROWS_REUSE * 128;   // 17408
constexpr int B_STAGE_BYTES = BLOCK_N * 128;            // 32768
constexpr int NUM_A_STAGES = 4;
constexpr int NUM_B_STAGES = 4;
constexpr int C_CHUNK = 32;                              // output columns per epilogue step
constexpr int C_STAGE_BYTES = BLOCK_M * C_CHUNK * 2;     // 8192
constexpr int NUM_C_STAGES = 2;
constexpr int TMEM_COLS = 512;                           // 2 accumulators x 256 fp32 columns
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_THREADS = 128;

constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + NUM_A_STAGES * A_STAGE_BYTES;
constexpr int OFF_C = OFF_B + NUM_B_STAGES * B_STAGE_BYTES;
constexpr int OFF_PARAMS = OFF_C + NUM_C_STAGES * C_STAGE_BYTES;
constexpr int OFF_BARS = OFF_PARAMS + 3 * BLOCK_N * 4;
constexpr int NUM_BARS = 2 * NUM_A_STAGES + 2 * NUM_B_STAGES + 4;
constexpr int OFF_TMEM_PTR = OFF_BARS + NUM_BARS * 8;
constexpr int SMEM_BYTES = OFF_TMEM_PTR + 16 + 1024;     // + slack for 1024-byte alignment
static_assert(OFF_B % 1024 == 0 && OFF_C % 1024 == 0, "swizzled stages must be 1024-byte aligned");
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of dynamic shared memory");

struct LayerArgs {
  int32_t n_m_tiles;        // R_pad / 128
  int32_t n_n_tiles;        // C_out / 256
  int32_t c_chunks;         // C_in_pad / 64
  int32_t taps;
  int32_t dilation;
  int32_t c_in_pad;         // column stride between taps in the packed weight matrix
  int32_t reuse;            // 0 / 1 (see header comment)
  int32_t desc_base_offset; // 1: set descriptor base-offset bits from the start address; 0: leave 0
  const float* bias;        // [C_out]  conv bias b
  const float* scale;       // [C_out]  gamma * rsqrt(var + eps)
  const float* shift;       // [C_out]  beta - mean * scale
  const uint8_t* row_valid; // [R_pad]  1 = row belongs to a segment, 0 = gap / tail
  uint32_t* overflow_flag;  // set to 1 if |activation| > 65504 (fp16 max)
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
tdnn_layer_kernel(const __grid_constant__ CUtensorMap tmap_a,     // activations in  [R_pad, C_in_pad] fp16
                  const __grid_constant__ CUtensorMap tmap_w,     // weights [C_out, taps*C_in_pad] fp16 (K-major)
                  const __grid_constant__ CUtensorMap tmap_c,     // activations out [R_pad, C_out] fp16
                  const LayerArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t sA = smem_base + OFF_A, sB = smem_base + OFF_B, sC = smem_base + OFF_C;
  float* s_bias = reinterpret_cast<float*>(smem + OFF_PARAMS);
  float* s_scale = s_bias + BLOCK_N;
  float* s_shift = s_scale + BLOCK_N;
  const uint32_t bar0 = smem_base + OFF_BARS;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (NUM_A_STAGES + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * NUM_A_STAGES + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * NUM_A_STAGES + NUM_B_STAGES + s); };
  auto t_full = [&](int s) { return bar0 + 8u * (2 * NUM_A_STAGES + 2 * NUM_B_STAGES + s); };
  auto t_empty = [&](int s) { return bar0 + 8u * (2 * NUM_A_STAGES + 2 * NUM_B_STAGES + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;          // warp-uniform
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_c);
    for (int s = 0; s < NUM_A_STAGES; ++s) { ptx::mbar_init(a_full(s), 1); ptx::mbar_init(a_empty(s), 1); }
    for (int s = 0; s < NUM_B_STAGES; ++s) { ptx::mbar_init(b_full(s), 1); ptx::mbar_init(b_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(t_full(s), 1); ptx::mbar_init(t_empty(s), NUM_EPI_THREADS); }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total_tiles = args.n_m_tiles * args.n_n_tiles;
  const int half = (args.taps - 1) >> 1;
  const int halo = half * args.dilation;
  const bool reuse = args.reuse != 0;
  const uint32_t a_box_bytes = (reuse ? A_BOX_ROWS_REUSE : A_BOX_ROWS_PLAIN) * 128u;

  if (warp == 0) {
    // ============================ TMA producer (one thread) ============================
    if (lane == 0) {
      uint32_t ia = 0, ib = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / args.n_n_tiles) * BLOCK_M;
        const int n0 = (tile % args.n_n_tiles) * BLOCK_N;
        for (int cc = 0; cc < args.c_chunks; ++cc) {
          for (int j = 0; j < args.taps; ++j) {
            if (!reuse || j == 0) {
              const uint32_t s = ia % NUM_A_STAGES, ph = (ia / NUM_A_STAGES) & 1u;
              ptx::mbar_wait(a_empty(s), ph ^ 1u);
              ptx::mbar_arrive_expect_tx(a_full(s), a_box_bytes);
              const int row = reuse ? (m0 - halo) : (m0 + (j - half) * args.dilation);
              ptx::tma_load_2d(sA + s * A_STAGE_BYTES, &tmap_a, a_full(s), cc * BLOCK_K, row);
              ++ia;
            }
            const uint32_t s = ib % NUM_B_STAGES, ph = (ib / NUM_B_STAGES) & 1u;
            ptx::mbar_wait(b_empty(s), ph ^ 1u);
            ptx::mbar_arrive_expect_tx(b_full(s), B_STAGE_BYTES);
            ptx::tma_load_2d(sB + s * B_STAGE_BYTES, &tmap_w, b_full(s), j * args.c_in_pad + cc * BLOCK_K, n0);
            ++ib;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer (one thread) ==============================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(BLOCK_M, BLOCK_N);
      const uint64_t desc_mask = args.desc_base_offset ? ~0ull : ~(0x7ull << 49);
      uint32_t ia = 0, ib = 0, it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1u;
        ptx::mbar_wait(t_empty(acc), ((it >> 1) & 1u) ^ 1u);      // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;
        uint32_t sa = 0;
        for (int cc = 0; cc < args.c_chunks; ++cc) {
          for (int j = 0; j < args.taps; ++j) {
            if (!reuse || j == 0) {
              sa = ia % NUM_A_STAGES;
              ptx::mbar_wait(a_full(sa), (ia / NUM_A_STAGES) & 1u);
            }
            const uint32_t sb = ib % NUM_B_STAGES;
            ptx::mbar_wait(b_full(sb), (ib / NUM_B_STAGES) & 1u);
            ptx::tc_fence_after();
            const uint32_t a_addr = sA + sa * A_STAGE_BYTES + (reuse ? uint32_t(j * args.dilation) * 128u : 0u);
            const uint32_t b_addr = sB + sb * B_STAGE_BYTES;
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t da = ptx::make_sw128_kmajor_desc(a_addr + k * (UMMA_K * 2)) & desc_mask;
              const uint64_t db = ptx::make_sw128_kmajor_desc(b_addr + k * (UMMA_K * 2));
              ptx::umma_f16(d_tmem, da, db, idesc, accumulate);
              accumulate = 1;
            }
            ptx::umma_commit(b_empty(sb));                         // weights slot free once these MMAs retire
            ++ib;
            if (!reuse || j == args.taps - 1) {
              ptx::umma_commit(a_empty(sa));
              ++ia;
            }
          }
        }
        ptx::umma_commit(t_full(acc));                             // accumulator ready for the epilogue
      }
    }
  } else {
    // ============================ epilogue (4 warps, 128 threads) ======================
    const int te = threadIdx.x - 64;                 // 0..127
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                   // row of the tile owned by this thread
    const uint32_t swz = (uint32_t(row) >> 1) & 3u;  // SWIZZLE_64B phase of this row in the staging tile
    float amax = 0.f;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1u;
      const int m0 = (tile / args.n_n_tiles) * BLOCK_M;
      const int n0 = (tile % args.n_n_tiles) * BLOCK_N;
      ptx::named_bar_sync(1, NUM_EPI_THREADS);       // previous tile's parameter reads are done
      for (int i = te; i < BLOCK_N; i += NUM_EPI_THREADS) {
        s_bias[i] = __ldg(args.bias + n0 + i);
        s_scale[i] = __ldg(args.scale + n0 + i);
        s_shift[i] = __ldg(args.shift + n0 + i);
      }
      const bool valid = args.row_valid[m0 + row] != 0;
      ptx::named_bar_sync(1, NUM_EPI_THREADS);
      ptx::mbar_wait(t_full(acc), (it >> 1) & 1u);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / C_CHUNK; ++chunk) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(t_row + chunk * C_CHUNK, v);
        ptx::tmem_ld_wait();
        if (chunk == BLOCK_N / C_CHUNK - 1) {        // accumulator fully read: hand it back to the MMA warp
          ptx::tc_fence_before();
          ptx::mbar_arrive(t_empty(acc));
        }
        uint32_t p[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const int c = chunk * C_CHUNK + g * 4;
          const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c);
          const float4 s4 = *reinterpret_cast<const float4*>(s_scale + c);
          const float4 h4 = *reinterpret_cast<const float4*>(s_shift + c);
          // relu(acc + b) * inv + shift      (models.py:477-480, tf_block.py:26)
          const float y0 = fmaf(fmaxf(__uint_as_float(v[g * 4 + 0]) + b4.x, 0.f), s4.x, h4.x);
          const float y1 = fmaf(fmaxf(__uint_as_float(v[g * 4 + 1]) + b4.y, 0.f), s4.y, h4.y);
          const float y2 = fmaf(fmaxf(__uint_as_float(v[g * 4 + 2]) + b4.z, 0.f), s4.z, h4.z);
          const float y3 = fmaf(fmaxf(__uint_as_float(v[g * 4 + 3]) + b4.w, 0.f), s4.w, h4.w);
          amax = fmaxf(amax, fmaxf(fmaxf(fabsf(y0), fabsf(y1)), fmaxf(fabsf(y2), fabsf(y3))));
          p[g * 2 + 0] = valid ? ptx::pack_half2(y0, y1) : 0u;    // gap rows stay exact zeros
          p[g * 2 + 1] = valid ? ptx::pack_half2(y2, y3) : 0u;
        }
        const uint32_t buf = chunk & 1u;
        if (te == 0) ptx::tma_store_wait_read<1>();  // the store that last used this buffer has read it
        ptx::named_bar_sync(1, NUM_EPI_THREADS);
        const uint32_t dst = sC + buf * C_STAGE_BYTES + uint32_t(row) * (C_CHUNK * 2);
#pragma unroll
        for (uint32_t c16 = 0; c16 < 4; ++c16) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((c16 ^ swz) << 4)),
                       "r"(p[c16 * 4 + 0]), "r"(p[c16 * 4 + 1]), "r"(p[c16 * 4 + 2]), "r"(p[c16 * 4 + 3])
                       : "memory");
        }
        ptx::fence_proxy_async_smem();
        ptx::named_bar_sync(1, NUM_EPI_THREADS);
        if (te == 0) {
          ptx::tma_store_2d(&tmap_c, sC + buf * C_STAGE_BYTES, n0 + chunk * C_CHUNK, m0);
          ptx::tma_store_commit();
        }
      }
    }
    if (te == 0) ptx::tma_store_wait_all<0>();
    if (amax > 65504.f) atomicOr(args.overflow_flag, 1u);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tdnn
