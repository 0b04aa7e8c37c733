// tdnn_first.cuh -- the FIRST frame layer with its input splice done inside the kernel (sm_100a).
//
// Replaces two launches -- pack_im2col_kernel (fp32 / fp16 MFCC rows -> spliced fp16 [R_pad, 128] matrix in HBM) and
// tdnn_pair_kernel<0> on that matrix -- for the layer local/tf/models.py:476-480 builds first: conv1d SAME over 5 taps of the
// 23 cepstra -> bias_add -> relu -> batch_norm_wrapper(eval).  That pair of launches was the least efficient part of an
// extraction step: the splice made a 27 MB round trip through HBM for 9 MB of features, and the layer itself (K = 128,
// 8 UMMAs per tile) waited on its store epilogue, not on the tensor pipe (3 300 of 4 200 cycles per tile with 8 epilogue warps).
//
// Here, per CTA pair and 256-row tile:
//   builder warps (4 per CTA)  copy the feature rows of the CTA's four aligned 32-row blocks (+ the layer's context) with
//                              16-byte cp.async one tile ahead, and write the spliced rows (taps outside the segment
//                              are zeros: TF's SAME padding) STRAIGHT INTO SHARED MEMORY in the
//                              K-major SWIZZLE_128B layout of a UMMA A operand ([128 rows x 64 k] atoms; 16-byte unit u of
//                              row r at u ^ (r & 7)), double-buffered; they also emit row_valid / blk_valid (what every
//                              later epilogue reads) and zero the counters of embed_fc_kernel
//   MMA warp (leader CTA)      D[256 rows, 256 ch] = A . W0[ct]^T for both channel tiles (weights resident in shared
//                              memory, loaded once per CTA by TMA), one TMEM accumulator per channel tile
//   epilogue warps (16/CTA)    relu(acc + b) * scale + shift, gap rows -> 0, fp16, per-warp swizzled staging box,
//                              TMA store of [32 rows x 32 ch] boxes
// The spliced matrix never exists in HBM, and twice the epilogue warps drain the accumulators (the kernel is bound by the
// 1 KB per row it must write).  Arithmetic: the values, the K order and the epilogue expressions are those of the two
// launches it replaces, so the stored rows are BIT-IDENTICAL (tests/test_gpu_parity.py::test_fused_first_layer_*).
//
// Used when the spliced width is 128 (5 x 23 -> 115 -> 128) with dilation 1, the layer is 256 or 512 wide, plain fp16 operands,
// the caller's feature matrix starts on a 16-byte boundary and the staged span of a block fits 4 KB (else: the two launches).
// Option "fuse_first" = 0 gives the two launches.
#pragma once
#include <type_traits>

#include "tdnn_tail.cuh"

namespace tdnn2 {

constexpr int F1_K = 128;                                  // spliced input width (two 64-wide atoms)
constexpr int F1_BUILD_WARPS = 4;
constexpr int F1_BUILD_THREADS = 32 * F1_BUILD_WARPS;      // 128 = one thread per row of the CTA's half tile
constexpr int F1_EPI_WARPS = 16;
constexpr int F1_EPI_THREADS = 32 * F1_EPI_WARPS;
constexpr int F1_FIRST_BUILD_WARP = 2;
constexpr int F1_FIRST_EPI_WARP = F1_FIRST_BUILD_WARP + F1_BUILD_WARPS;       // 6
constexpr int F1_THREADS = 32 * (F1_FIRST_EPI_WARP + F1_EPI_WARPS);          // 704
constexpr int F1_CHUNKS = 8;                               // 16-byte pieces of a block's staged span per builder lane
constexpr int F1_STAGE_BYTES = F1_CHUNKS * 32 * 16;        // 4096: (32 + 2 halo) * D * 4 + 30 must fit (D = 23, halo 2: 3 342)
constexpr int F1_MAX_CT = 2;                               // channel tiles (c_out <= 512): one TMEM accumulator each

constexpr int F1_OFF_W = 0;                                                  // [ct][atom][128 ch x 128 B]
constexpr int F1_OFF_A = F1_OFF_W + F1_MAX_CT * 2 * WGT_ATOM_BYTES;          // 65536: [buf][atom][128 rows x 128 B]
constexpr int F1_OFF_C = F1_OFF_A + 2 * 2 * WGT_ATOM_BYTES;                  // 131072: [warp][32 x 64 B]
constexpr int F1_OFF_FEAT = F1_OFF_C + F1_EPI_WARPS * C_BUF_BYTES;           // 163840: [4 builder warps][2 buffers][F1_STAGE_BYTES]
constexpr int F1_OFF_PAR = F1_OFF_FEAT + 4 * 2 * F1_STAGE_BYTES;                // [ct][bias | scale | shift | slope][256]
constexpr int F1_OFF_BARS = F1_OFF_PAR + F1_MAX_CT * PAR_ARRAYS * TILE_CH * 4;
constexpr int F1_NUM_BARS = 1 + 2 + 2 + 2 + 2;             // w_full | a_full[2] | a_empty[2] | t_full[2] | t_empty[2]
constexpr int F1_OFF_TMEM_PTR = F1_OFF_BARS + (F1_NUM_BARS + 1) * 8;
constexpr int F1_SMEM_BYTES = F1_OFF_TMEM_PTR + 16 + 1024;
static_assert(F1_OFF_A % 1024 == 0 && F1_OFF_C % 1024 == 0, "swizzled regions must be 1024-byte aligned");
static_assert(F1_SMEM_BYTES <= 232448, "exceeds 227 KB of dynamic shared memory");

struct FirstArgs {
  int32_t n_row_tiles;        // R_pad / 256
  int32_t n_ch_tiles;         // C_out / 256 (1 or 2)
  int32_t c_out;
  int32_t feat_dim, halo;     // D; half context of the layer in frames ((taps - 1) / 2 * dilation)
  const void* feats;          // [total_frames, D] fp32, or float16 values when feats_f16; 16-byte aligned
  int32_t feats_f16;
  int64_t n_feat_bytes;       // total_frames * D * (4 or 2)
  int32_t taps;
  const int4* blk_info;       // [R_pad / 32] {first feature row, first frame in segment, segment length, valid rows}
  const float* bias; const float* scale; const float* shift; const float* alpha;
  uint8_t* row_valid;         // [R_pad]      written here
  uint8_t* blk_valid;         // [R_pad / 32] written here
  uint32_t* counters;         // zeroed here (embed_fc_kernel)
  int32_t n_counters;
  uint32_t* overflow_flag;
  uint32_t overflow_bit;
  long long* trace;           // diagnostics (tools/trace_tiles.py 0): [cluster][rank][TRACE_TILES][8] SM clock stamps, or null:
                              // 0 builder has staged the tile, 1 builder published A, 2 MMA sees A, 3 MMAs issued,
                              // 4 epilogue sees accumulator 0, 5 epilogue released the last accumulator, 6 epilogue done, 7 A buffer free
};

__device__ __forceinline__ void f1_stamp(const FirstArgs& a, int cluster, uint32_t rank, uint32_t it, int slot) {
  if (a.trace != nullptr && it < TRACE_TILES)
    a.trace[((size_t(cluster) * 2 + rank) * TRACE_TILES + it) * 8 + slot] = clock64();
}

template <bool LEAKY>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F1_THREADS, 1)
tdnn_first_kernel(const __grid_constant__ CUtensorMap tmap_w,     // [C_out, 128] fp16 K-major
                  const __grid_constant__ CUtensorMap tmap_out,   // [R_pad, C_out] fp16
                  const __grid_constant__ FirstArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t sW = smem_base + F1_OFF_W;
  const uint32_t sA = smem_base + F1_OFF_A;
  const uint32_t bar0 = smem_base + F1_OFF_BARS;
  const uint32_t w_full = bar0;
  auto a_full = [&](uint32_t s) { return bar0 + 8u * (1 + s); };
  auto a_empty = [&](uint32_t s) { return bar0 + 8u * (3 + s); };
  auto t_full = [&](uint32_t s) { return bar0 + 8u * (5 + s); };
  auto t_empty = [&](uint32_t s) { return bar0 + 8u * (7 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + F1_OFF_TMEM_PTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int NCT = args.n_ch_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_out);
    ptx::mbar_init(w_full, 1);
    for (uint32_t s = 0; s < 2; ++s) {
      ptx::mbar_init(a_full(s), 2 * F1_BUILD_WARPS);
      ptx::mbar_init(a_empty(s), 1);
      ptx::mbar_init(t_full(s), 1);
      ptx::mbar_init(t_empty(s), 2 * F1_EPI_WARPS);
    }
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

  if (warp == 0) {
    // ============================ weights: once per CTA (its 128 channels of every channel tile) ============================
    if (ptx::elect_one()) {
      const uint32_t w_full_leader = ptx::mapa_cluster(w_full, 0);
      if (leader) ptx::mbar_arrive_expect_tx(w_full, 2u * uint32_t(NCT) * 2u * WGT_ATOM_BYTES);      // both CTAs' boxes
      for (int ct = 0; ct < NCT; ++ct)
        for (int h = 0; h < 2; ++h)
          ptx::tma_load_2d_2sm(sW + uint32_t(ct * 2 + h) * WGT_ATOM_BYTES, &tmap_w, w_full_leader, h * BLOCK_K,
                               ct * TILE_CH + int(rank) * CTA_CH);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================ MMA issuer (leader) ============================
    if (leader) {
      constexpr uint32_t idesc = ptx::make_idesc_f16_f32(TILE_ROWS, TILE_CH);
      const uint64_t desc_hi = ptx::make_sw128_kmajor_desc(0);
      auto desc = [&](uint32_t addr) { return desc_hi | uint64_t((addr >> 4) & 0x3fffu); };
      ptx::mbar_wait(w_full, 0);
      uint32_t ti = 0;
      for (int tile = cluster_id; tile < args.n_row_tiles; tile += n_clusters, ++ti) {
        const uint32_t buf = ti & 1u;
        ptx::mbar_wait_cluster(a_full(buf), (ti >> 1) & 1u);              // both CTAs' builders wrote their rows
        ptx::tc_fence_after();
        if (lane == 0) f1_stamp(args, cluster_id, rank, ti, 2);
        for (int ct = 0; ct < NCT; ++ct) {
          ptx::mbar_wait_cluster(t_empty(uint32_t(ct)), (ti & 1u) ^ 1u);  // both CTAs' epilogues drained this accumulator
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + uint32_t(ct) * TILE_CH;
          const uint64_t da = desc(sA + buf * (2 * WGT_ATOM_BYTES)), dw = desc(sW + uint32_t(ct) * (2 * WGT_ATOM_BYTES));
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < F1_K / UMMA_K; ++k) {
              const uint64_t off = uint64_t((k >> 2) * (WGT_ATOM_BYTES >> 4) + 2 * (k & 3));
              ptx::umma_f16_2sm(d_tmem, da + off, dw + off, idesc, uint32_t(k));
            }
            ptx::umma_commit_2sm(t_full(uint32_t(ct)));
            if (ct == NCT - 1) ptx::umma_commit_2sm(a_empty(buf));        // the A buffer is free in both CTAs
          }
          __syncwarp();
        }
        if (lane == 0) f1_stamp(args, cluster_id, rank, ti, 3);
      }
    }
  } else if (warp < F1_FIRST_EPI_WARP) {
    // ============================ builders: feature rows -> spliced A operand in shared memory ============================
    // Warp j owns block j of the CTA's four aligned 32-row blocks from the global loads to the published rows: no barrier
    // between builder warps.  The block's feature rows (+ the layer's context) are ONE contiguous span of the caller's matrix:
    // it is copied as it lies -- 16-byte cp.async pieces from the 16-byte aligned address below its first byte, no registers,
    // one tile ahead, double-buffered -- and the rows are spliced out of that copy.  Frames outside the segment (SAME
    // padding, models.py:476) are masked per tap while splicing; pieces outside the caller's matrix are zero-filled.
    const int j = warp - F1_FIRST_BUILD_WARP;                   // my block
    const int b = j * 32 + lane;                                // my row of the CTA's half tile
    const int D = args.feat_dim, halo = args.halo, taps = args.taps;
    const bool f16 = args.feats_f16 != 0;
    const int esz = f16 ? 2 : 4;
    const int span_bytes = (32 + 2 * halo) * D * esz;
    const uint32_t sStage = smem_base + F1_OFF_FEAT + uint32_t(j) * (2 * F1_STAGE_BYTES);
    const uint32_t a_full_leader = ptx::mapa_cluster(a_full(0), 0);
    const char* feats_b = reinterpret_cast<const char*>(args.feats);
    for (int i = blockIdx.x * F1_BUILD_THREADS + b; i < args.n_counters; i += gridDim.x * F1_BUILD_THREADS) args.counters[i] = 0u;
    auto blk_of = [&](int tile) { return ((tile * TILE_ROWS + int(rank) * CTA_ROWS) >> 5) + j; };
    // Block descriptors: lane l holds the one of this warp's block in the (32 g + l)-th tile of the cluster, reloaded every
    // 32 tiles; a tile's descriptor is then four shuffles.
    int4 bi_lane = make_int4(0, 0, 0, 0);
    auto load_group = [&](uint32_t t_first) {                    // t_first: multiple of 32
      const int tile = cluster_id + int(t_first + uint32_t(lane)) * n_clusters;
      bi_lane = tile < args.n_row_tiles ? __ldg(args.blk_info + blk_of(tile)) : make_int4(0, 0, 0, 0);
    };
    auto desc_of = [&](uint32_t t) {
      const int src = int(t & 31u);
      return make_int4(__shfl_sync(0xffffffffu, bi_lane.x, src), __shfl_sync(0xffffffffu, bi_lane.y, src),
                       __shfl_sync(0xffffffffu, bi_lane.z, src), __shfl_sync(0xffffffffu, bi_lane.w, src));
    };
    // async copy of a block's span into a staging buffer; returns the byte offset of the span's first value in it
    auto issue = [&](const int4 bi, uint32_t sbuf) {
      const int64_t first = (int64_t(bi.x) - halo) * D * esz;   // byte offset of the span in the caller's matrix (may be < 0)
      const int64_t a0 = first & ~int64_t(15);                  // floor to 16 bytes (two's complement: also below zero)
      const int shift = int(first - a0);
      if (bi.w > 0) {                                           // (a block without valid rows is never read)
        const int n_chunks = (shift + span_bytes + 15) >> 4;
#pragma unroll
        for (int u = 0; u < F1_CHUNKS; ++u) {
          const int c = lane + u * 32;
          if (c < n_chunks) {
            const int64_t g = a0 + int64_t(c) * 16;
            const int64_t left = args.n_feat_bytes - g;
            const uint32_t sz = (g < 0 || left <= 0) ? 0u : (left < 16 ? uint32_t(left) : 16u);    // bytes that exist; the rest: zeros
            const char* src = feats_b + (sz ? g : int64_t(0));
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(sbuf + uint32_t(c) * 16u), "l"(src), "r"(sz) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      return shift;
    };
    load_group(0);
    int4 cur = desc_of(0);
    int cur_shift = issue(cur, sStage);
    uint32_t ti = 0;
    for (int tile = cluster_id; tile < args.n_row_tiles; tile += n_clusters, ++ti) {
      const uint32_t buf = ti & 1u;
      const int r_cta = tile * TILE_ROWS + int(rank) * CTA_ROWS;
      const uint32_t sbuf = sStage + buf * F1_STAGE_BYTES;
      // (1) this tile's span has landed (issued one tile ago); the next tile's goes into the other buffer, which the build of
      //     the previous tile has finished reading (the __syncwarp that ends it)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      if (b == 0) f1_stamp(args, cluster_id, rank, ti, 0);
      int4 nxt = make_int4(0, 0, 0, 0);
      int nxt_shift = 0;
      if (tile + n_clusters < args.n_row_tiles) {
        if (((ti + 1u) & 31u) == 0u) load_group(ti + 1u);
        nxt = desc_of(ti + 1u);
        nxt_shift = issue(nxt, sStage + (buf ^ 1u) * F1_STAGE_BYTES);
      }
      const int t_row = cur.y + lane, len = cur.z, nv = cur.w;  // my row's frame in its segment
      const bool valid = lane < nv;
      args.row_valid[r_cta + b] = valid ? 1 : 0;
      if (lane == 0) args.blk_valid[(r_cta >> 5) + j] = uint8_t(nv);
      // (2) spliced row -> both atoms of A[buf], once the MMAs of two tiles ago have read it.  The layer's dilation is 1, so
      //     the taps * D values of a spliced row are one contiguous run of the staged frames, starting at the row's own
      //     frame minus the half context = staged value lane * D.  Tap `tap` is frame t_row - halo + tap: zeros unless it
      //     lies in [0, len).
      ptx::mbar_wait(a_empty(buf), ((ti >> 1) & 1u) ^ 1u);
      if (b == 0) f1_stamp(args, cluster_id, rank, ti, 7);
      const int kv = taps * D;                                                  // real columns of a spliced row (the rest: padding)
      const int k_lo = valid ? max(0, halo - t_row) * D : 0;                    // columns [k_lo, k_hi) of my row exist
      const int k_hi = valid ? min(taps, len - t_row + halo) * D : 0;
      const bool whole = __all_sync(0xffffffffu, valid && k_lo == 0 && k_hi == kv);   // an interior block: nothing to mask
      const uint32_t row_src = sbuf + uint32_t(cur_shift) + uint32_t(lane * D * esz);
      const uint32_t row_a = sA + buf * (2 * WGT_ATOM_BYTES) + uint32_t(b) * 128u;
      const uint32_t sw = uint32_t(b) & 7u;
      auto put = [&](int u, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_a + uint32_t(u >> 3) * WGT_ATOM_BYTES + (((uint32_t(u) & 7u) ^ sw) << 4)),
                     "r"(p0), "r"(p1), "r"(p2), "r"(p3)
                     : "memory");
      };
      // F16: the staged values are float16 (the product's feed) -- two of them ARE a packed pair; else fp32, rounded here
      auto splice = [&](auto f16_tag) {
        constexpr bool F16 = decltype(f16_tag)::value;
        auto ld = [&](int k) {             // one value of my row (volatile: the address does not change from tile to tile, the span does)
          uint32_t v;
          if (F16) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(row_src + uint32_t(k) * 2u) : "memory");
          else asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(row_src + uint32_t(k) * 4u) : "memory");
          return v;
        };
        auto pack = [&](uint32_t x0, uint32_t x1) {
          return F16 ? (x0 | (x1 << 16)) : ptx::pack_half2(__uint_as_float(x0), __uint_as_float(x1));
        };
        if (whole) {
#pragma unroll
          for (int u = 0; u < F1_K / 8; ++u) {
            if (u * 8 + 8 <= kv) {                               // (warp-uniform) every value of this 16-byte piece exists
              uint32_t x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = ld(u * 8 + e);
              put(u, pack(x[0], x[1]), pack(x[2], x[3]), pack(x[4], x[5]), pack(x[6], x[7]));
            } else if (u * 8 >= kv) {
              put(u, 0u, 0u, 0u, 0u);
            } else {
              uint32_t x[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) x[e] = (u * 8 + e < kv) ? ld(u * 8 + e) : 0u;
              put(u, pack(x[0], x[1]), pack(x[2], x[3]), pack(x[4], x[5]), pack(x[6], x[7]));
            }
          }
        } else {                                                 // first / last blocks of a segment: per-value bounds
#pragma unroll
          for (int u = 0; u < F1_K / 8; ++u) {
            uint32_t x[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) x[e] = (u * 8 + e >= k_lo && u * 8 + e < k_hi) ? ld(u * 8 + e) : 0u;
            put(u, pack(x[0], x[1]), pack(x[2], x[3]), pack(x[4], x[5]), pack(x[6], x[7]));
          }
        }
      };
      if (nv == 0) {                                             // a gap block: zero rows
#pragma unroll
        for (int u = 0; u < F1_K / 8; ++u) put(u, 0u, 0u, 0u, 0u);
      } else if (f16) {
        splice(std::true_type{});
      } else {
        splice(std::false_type{});
      }
      ptx::fence_proxy_async_smem();                 // the tensor core reads these rows through the async proxy
      __syncwarp();                                  // ... and every lane has read the staged span: the copy after next may overwrite it
      if (lane == 0) mbar_arrive_cluster_release(a_full_leader + 8u * buf);
      if (b == 0) f1_stamp(args, cluster_id, rank, ti, 1);
      cur = nxt;
      cur_shift = nxt_shift;
    }
  } else {
    // ============================ epilogue (16 warps per CTA) ============================
    const int e = warp - F1_FIRST_EPI_WARP;          // 0..15
    const int q = warp & 3;                          // TMEM lane quarter this warp may read
    const int colq = e >> 2;                         // which 64-column quarter of an accumulator
    const uint32_t t_empty_leader = ptx::mapa_cluster(t_empty(0), 0);
    const uint32_t sPar = smem_base + F1_OFF_PAR;
    const uint32_t sC = smem_base + F1_OFF_C + uint32_t(e) * C_BUF_BYTES;
    const uint32_t swz = (uint32_t(lane) >> 1) & 3u; // SWIZZLE_64B phase of this row in the staging box
    for (int c = threadIdx.x - 32 * F1_FIRST_EPI_WARP; c < NCT * TILE_CH; c += F1_EPI_THREADS) {   // once: same for every tile
      const uint32_t dst = sPar + uint32_t(c >> 8) * (PAR_ARRAYS * TILE_CH * 4) + uint32_t(c & 255) * 4u;
      ptx::sts_f(dst, __ldg(args.bias + c));
      ptx::sts_f(dst + TILE_CH * 4u, __ldg(args.scale + c));
      ptx::sts_f(dst + 2u * TILE_CH * 4u, __ldg(args.shift + c));
      ptx::sts_f(dst + 3u * TILE_CH * 4u, LEAKY ? __ldg(args.alpha + c) : 0.f);
    }
    ptx::named_bar_sync(1, F1_EPI_THREADS);
    uint32_t hmax = 0, ti = 0;
    auto nv_of = [&](int tile) { return __ldg(&args.blk_info[((tile * TILE_ROWS + int(rank) * CTA_ROWS) >> 5) + q].w); };
    int nv_next = cluster_id < args.n_row_tiles ? nv_of(cluster_id) : 0;         // my row = lane of the CTA's block q
    for (int tile = cluster_id; tile < args.n_row_tiles; tile += n_clusters, ++ti) {
      const int r_cta = tile * TILE_ROWS + int(rank) * CTA_ROWS;
      const bool valid = lane < nv_next;
      int nv_load = 0;
      if (tile + n_clusters < args.n_row_tiles) nv_load = nv_of(tile + n_clusters);   // consumed at the END of this tile
      for (int ct = 0; ct < NCT; ++ct) {
        ptx::mbar_wait(t_full(uint32_t(ct)), ti & 1u);
        ptx::tc_fence_after();
        if (e == 0 && lane == 0 && ct == 0) f1_stamp(args, cluster_id, rank, ti, 4);
        const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(ct) * TILE_CH + uint32_t(colq) * 64u;
        const uint32_t s_par = sPar + uint32_t(ct) * (PAR_ARRAYS * TILE_CH * 4);
#pragma unroll
        for (int chunk = 0; chunk < 2; ++chunk) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_row + chunk * C_CHUNK, v);
          ptx::tmem_ld_wait_dep(v);
          if (chunk == 1) {                            // all of this warp's loads done: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(t_empty_leader + 8u * uint32_t(ct));
            if (e == 0 && lane == 0 && ct == NCT - 1) f1_stamp(args, cluster_id, rank, ti, 5);
          }
          uint32_t p[16];
          epi_store_math<LEAKY>(v, s_par, colq * 64 + chunk * C_CHUNK, valid, p, hmax, 1.f);
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
            ptx::tma_store_2d(&tmap_out, buf, ct * TILE_CH + colq * 64 + chunk * C_CHUNK, r_cta + q * 32);
            ptx::tma_store_commit();
          }
        }
      }
      if (e == 0 && lane == 0) f1_stamp(args, cluster_id, rank, ti, 6);
      asm volatile("mov.b32 %0, %1;" : "=r"(nv_next) : "r"(nv_load));      // (volatile: keeps the wait for the load down here)
    }
    if (lane == 0) ptx::tma_store_wait_all<0>();
    if ((hmax & 0x7fffu) >= 0x7c00u || ((hmax >> 16) & 0x7fffu) >= 0x7c00u)
      atomicOr(args.overflow_flag, args.overflow_bit ? args.overflow_bit : 1u);
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
