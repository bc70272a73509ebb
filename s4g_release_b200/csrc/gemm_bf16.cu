// Persistent tcgen05 GEMM for the TRAINING path:   C[P][N] (bf16) = A[P][K] (bf16) · B[N][K]^T (bf16),  fp32 accumulation.
//
// The training step cannot use the fused inference chains (csrc/mlp_chain.cu): train-mode BatchNorm needs the statistics
// of a layer's whole output before the next layer can start, and the backward pass needs every layer's input and
// pre-activation again.  So a shared-MLP layer (reference nn_utils/conv.py:30-36,70-76 in training mode) is, on this
// path, one GEMM over channel-last bf16 rows followed by small fused element-wise kernels (csrc/train_ops.cu); this
// kernel is that GEMM — forward (Y = X W^T) and input gradient (dX = dY W, with W^T handed in as B).
//
//   * one persistent CTA per SM walks 128 x 128 — or, for the K >= 256 layers, 128 x 256 — output tiles, n-tile fastest
//     (the CTAs that run side by side share the A rows through L2);
//   * warp 0 and the LAST warp: TMA producers — per K-slab of 64 bf16 (= one 128-byte swizzle row) one A box {64, 128}
//     and one B box {64, 128 or 256} into a 3-5-stage ring (producer 0 issues the A boxes, producer 1 the B boxes: one
//     thread issues a bulk copy every ~0.26 us whatever its size); rows / channels out of range are zero-filled by the
//     copy engine, so P, N, K need no padding;
//   * warp 1: tcgen05.mma kind::f16 (M = 128, N = 128 or 256, K = 16), 4 per slab, into one of TWO TMEM accumulators;
//     tcgen05.commit frees the stage / publishes the accumulator.  The issuing thread costs ~130-190 cycles per MMA
//     whatever N (profiles/r01/mma_loop.txt), hence N = 256 where there are many K steps per tile;
//   * warps 2-5 (and 6-9): epilogue — drain the other accumulator (tcgen05.ld, bf16 pack) into a 128-byte-swizzled staging
//     tile in shared memory and hand it to the copy engine (cp.async.bulk.tensor store, two {64, 128} boxes; rows >= P and
//     columns >= N are clipped by the tensor map) while the tensor pipe fills the next accumulator.  (First version:
//     every thread stored its own row with 16-byte st.global — 32 half-filled sectors per warp instruction; the store
//     path, not HBM, bounded the memory-bound layers: 25.8 ms of GEMM per training step against ~14 ms of traffic.)
//   * the epilogue comes in one or TWO groups of 4 warps, a staging tile each: with 128-column tiles the groups alternate
//     tiles (tile parity = accumulator = group) — with one group the K <= 128 layers ran at the epilogue's pace (ncu,
//     profiles/r02/ncu_training.txt: 128 -> 128 and 128 -> 256 on 10.5 M rows both at ~3 000 cycles per tile, 75 % resp.
//     58 % of the DRAM peak although the tile's traffic differs by 1.35 x); with 256-column tiles they split every tile.
//   What bounds a layer (profiles/r02/gemm_layers_v6.txt): HBM for the 10.5 M-row layers (0.9-1.0 of the copy bandwidth);
//   for K, N >= 512 latency x bytes in flight (ncu, profiles/r02/ncu_gemm_tile256.txt: tensor pipe 47 %, L2 35 %, DRAM
//   33 % of their peaks, stalls on the operand loads; the 144 KB ring turns over once per ~2 us and the shared memory is
//   full) — sharing the weight slab in a CTA pair / cluster multicast is the next step there.
//
// BWD (input-gradient GEMM of block l, dX = dY W): the output tile IS the upstream gradient of block l-1, so the epilogue
// also does what the first pass of that block's BatchNorm backward would do: the producer warp fetches the matching
// 128 x 128 tile of block l-1's pre-activations y into shared memory (TMA, same swizzle as the staging tile); after the
// accumulator is staged, the epilogue threads walk the tile COLUMN-wise (2 columns x 64 rows per thread, conflict-free),
// mask the staged gradient with ReLU' / dropout of block l-1 in place, and accumulate sum g and sum g*y per column;
// the MASKED tile goes out through the TMA store — s4g_train_bn_bwd_reduce never runs for a block whose gradient comes
// out of a GEMM.  One epilogue group, two y tiles (so that the producer can run two tiles ahead).  Measured: the column
// walk is ~2 800 instructions per epilogue warp and tile at 0.36 IPC (one warp per scheduler) — 7 700 cycles per tile
// against 3 700 of HBM time, slower than the plain GEMM + the staged separate pass; kept as an opt-in
// (train_engine.FUSED_BWD_REDUCE).  A first version (profiles/r02/gemm_layers_v1.txt) was slower still: every thread read its own ROW of y from global memory and the column sums were a register
// transpose-reduce over the warp (31 shuffles per 32 columns and quantity) — ~2 300 instructions per thread and tile,
// 2.6 ms for the 128 -> 128 layer on 10.5 M rows against 0.87 ms for the plain GEMM + 0.9 ms for the separate pass.
//
// WS (weight-stationary, 128-column tiles, K <= 384: the slice must leave room for >= 5 A stages): a CTA keeps ONE n-tile
// for all its m-tiles and loads that [128][K] slice of B into shared memory once; the ring then only carries A slabs (16 KB each, up to 6 in flight).  Without it every 128 x 128
// tile re-reads its B slice from L2 — ncu on the 264 -> 256 layer of the second set-abstraction level (2.1 M rows):
// 30 % of the DRAM peak, tensor pipe 22 %, stalls `long_scoreboard`, 6.3 TB/s of L2 -> SM traffic, i.e. L2-bound
// (profiles/r02/ncu_training.txt).  Measured per launch (profiles/r02/train_kernels_v5.txt): 3 -> 128 layer on 10.5 M rows
// 2.24 -> 0.80 ms, 264 -> 256 on 2.1 M rows 0.85 -> 0.46 ms; K = 512 layers got 20-30 % slower with only 3 A stages and stay
// on the streaming schedule.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace s4g {
namespace gemm {

constexpr int kMaxThreads = 64 + 2 * 128 + 32;    // producer warp, MMA warp, 1 or 2 epilogue groups of 4 warps, second producer warp
constexpr int kStages = 5;                         // streaming mode: at most this many stages of A + B slabs (32 KB)
constexpr int kMaxStagesA = 6;                     // weight-stationary mode: stages of A slabs (16 KB)
constexpr int kMaxSlabsWS = 8;                     // K <= 512
constexpr int kTile = 128;
constexpr int kSlab = 64;                          // K elements per stage: 64 bf16 = 128 B
constexpr int kOperandBytes = kTile * kSlab * 2;   // 16 KB
constexpr int kStagingBytes = kTile * kTile * 2;    // 32 KB: the bf16 output tile as two {64 columns, 128 rows} swizzled halves

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();  // a protocol bug traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
      "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// one box of the output tile: shared memory (swizzled like the tensor map) -> global, bulk async group
__device__ __forceinline__ void tma_store_2d(const void* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)),
               "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void epi_barrier(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }  // one group
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// STATS: the epilogue also accumulates, per output column, the sum and the sum of squares of the bf16-ROUNDED results
// (what is stored is what BatchNorm normalises): once the tile is staged in shared memory for the store, the epilogue
// threads re-read it COLUMN-wise (one 32-bit word = two columns per thread, 64 rows each), shared-memory atomics collect
// the sums per CTA over all its tiles, one fp64 global atomic per column and CTA at the end.  Rows >= P and columns >= N
// are zero-filled operands: they add nothing.  (A register-level shuffle butterfly was measured first: ~5 800 cycles per
// tile for the 4 epilogue warps against ~2 800 cycles of HBM time.)
// what the BWD epilogue needs of block l-1 (the block whose output this GEMM's result is the gradient of)
struct BwdEpilogue {
  const __nv_bfloat16* y;   // its pre-BatchNorm rows [P][N], leading dimension ldy
  long long ldy;
  const float* scale;       // its folded BatchNorm scale / shift [N]
  const float* shift;
  int relu;
  unsigned seed, thresh;    // its dropout (thresh 0 = none): same counter hash as csrc/train_ops.cu
  float keep_scale;
};
__device__ __forceinline__ bool keep_elem(unsigned seed, long long row, int ch, int C, unsigned thresh) {
  unsigned long long idx = (unsigned long long)row * (unsigned)C + (unsigned)ch;
  unsigned h = (unsigned)idx ^ (unsigned)(idx >> 32) * 0x9E3779B9u ^ seed;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h >= thresh;
}

constexpr int kPlain = 0, kStats = 1, kBwd = 2;

template <int MODE, bool WS>
__global__ void __launch_bounds__(kMaxThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_y, int P, int N, int K,
                 int n_stages, int groups, int bn, double* __restrict__ stats, const BwdEpilogue bw) {
  constexpr bool STATS = MODE != kPlain;  // per-column sums collected in shared memory, fp64 global atomics at the end
  const int kThreads = (int)blockDim.x;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kMaxStagesA], empty[kMaxStagesA], acc_full[2], acc_empty[2], w_full, y_full[2], y_empty[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // bn = output columns per tile: 128, or 256 (streaming schedule only): ONE N = 256 MMA per K step instead of two N = 128
  // ones — the MMA-issuing thread, not the tensor pipe or the loads, paces the K >= 256 layers (~190 cycles per MMA
  // whatever its N: profiles/r02/gemm_layers_v3.txt, gemm_layers_v4_two_producers.txt; profiles/r01/mma_loop.txt).  The two
  // epilogue groups then split every tile by columns instead of alternating tiles.
  const bool split = bn > kTile;
  const int tiles_n = (N + bn - 1) / bn;
  const int ncols = tiles_n * bn;  // width of the per-column tables
  const int tiles_m = (P + kTile - 1) / kTile;
  const int n_slabs = (K + kSlab - 1) / kSlab;
  // shared memory: [WS: B slice, n_slabs x 16 KB][ring][staging: groups x 32 KB][BWD: 2 y tiles of 32 KB]
  //                [STATS / BWD: 2 x tiles_n x 128 floats of column sums]
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms are 1024-byte aligned
  uint8_t* wreg = base;
  uint8_t* ring = base + (WS ? (size_t)n_slabs * kOperandBytes : 0);
  const int stage_bytes = WS ? kOperandBytes : kOperandBytes + (bn / kTile) * kOperandBytes;
  uint8_t* staging_all = ring + (size_t)n_stages * stage_bytes;
  uint8_t* ytile = staging_all + (size_t)groups * kStagingBytes;
  float* s_stat = reinterpret_cast<float*>(ytile + (MODE == kBwd ? 2 * kStagingBytes : 0));
  // this CTA's tiles: streaming = every gridDim-th tile, n fastest; WS = one n-tile, every (gridDim / tiles_n)-th m-tile
  const int per_n = WS ? (int)gridDim.x / tiles_n : 0;
  const int my_n = WS ? (int)blockIdx.x % tiles_n : 0;
  const int my_m0 = WS ? (int)blockIdx.x / tiles_n : 0;
  const int n_my = WS ? (((int)blockIdx.x < per_n * tiles_n && my_m0 < tiles_m) ? (tiles_m - my_m0 + per_n - 1) / per_n : 0)
                      : (((int)blockIdx.x < tiles_m * tiles_n) ? (tiles_m * tiles_n - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0);
  auto tile_of = [&](int i, int& row0, int& col0) {
    if (WS) { row0 = (my_m0 + i * per_n) * kTile; col0 = my_n * kTile; }
    else { const int t = (int)blockIdx.x + i * (int)gridDim.x; row0 = (t / tiles_n) * kTile; col0 = (t % tiles_n) * bn; }
  };

  if constexpr (STATS) {
    for (int i = threadIdx.x; i < 2 * ncols; i += kThreads) s_stat[i] = 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStagesA; ++s) { mbar_init(&full[s], WS ? 1 : 2); mbar_init(&empty[s], 1); }  // streaming: A and B producers
    mbar_init(&w_full, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&y_full[a], 1); mbar_init(&y_empty[a], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], split ? 8 : 4); }  // epilogue warps per tile
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(2u * (unsigned)bn) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  const int warp_p1 = 2 + 4 * groups;  // the second producer warp (the last one)
  if (warp == 0 || warp == warp_p1) {
    // Two producer lanes in two warps: ONE thread issues a bulk copy every ~0.26 us whatever its size
    // (profiles/r01/tma_bw.txt); with 256-column tiles a K slab is an A box of 16 KB and a B box of 32 KB, and one lane
    // issuing both paced the K = 512 layers at ~1 360 cycles per slab (gemm_layers_v5_tile256.txt).  Streaming: producer 0
    // loads the A boxes, producer 1 the B boxes (each arms the stage's barrier with its own bytes).  Weight-stationary:
    // they alternate slabs; producer 1 also loads the resident B slice and, in BWD mode, the y tiles.
    const int who = warp == 0 ? 0 : 1;
    if (elect_one() && n_my > 0) {
      if (WS && who == 1) {  // the CTA's B slice, once
        mbar_expect_tx(&w_full, (unsigned)n_slabs * kOperandBytes);
        for (int k = 0; k < n_slabs; ++k) tma_load_2d(wreg + (size_t)k * kOperandBytes, &map_b, k * kSlab, my_n * kTile, &w_full);
      }
      unsigned g = 0;  // running slab index over all of this CTA's tiles
      for (int i = 0; i < n_my; ++i) {
        int row0, col0;
        tile_of(i, row0, col0);
        for (int k = 0; k < n_slabs; ++k, ++g) {
          if (WS && (int)(g & 1u) != who) continue;
          const unsigned s = g % (unsigned)n_stages, use = g / (unsigned)n_stages;
          if (use > 0) mbar_wait(&empty[s], (use - 1) & 1u);
          // (zero-filled out-of-range elements count as transferred bytes)
          if (WS || who == 0) {
            mbar_expect_tx(&full[s], (unsigned)kOperandBytes);
            tma_load_2d(ring + (size_t)s * stage_bytes, &map_a, k * kSlab, row0, &full[s]);
          } else {
            mbar_expect_tx(&full[s], (unsigned)(stage_bytes - kOperandBytes));
            tma_load_2d(ring + (size_t)s * stage_bytes + kOperandBytes, &map_b, k * kSlab, col0, &full[s]);
          }
        }
        if constexpr (MODE == kBwd) {  // the previous block's pre-activation tile, for this tile's epilogue
          if (who == 1) {
            // (two buffers: with one, the producer could not run more than a tile ahead of the epilogue)
            const int yb = i & 1;
            if (i >= 2) mbar_wait(&y_empty[yb], (unsigned)((i >> 1) - 1) & 1u);
            const bool two = col0 + 64 < N;
            uint8_t* yt = ytile + (size_t)yb * kStagingBytes;
            mbar_expect_tx(&y_full[yb], two ? (unsigned)kStagingBytes : (unsigned)kStagingBytes / 2);
            tma_load_2d(yt, &map_y, col0, row0, &y_full[yb]);
            if (two) tma_load_2d(yt + kStagingBytes / 2, &map_y, col0 + 64, row0, &y_full[yb]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // instruction descriptor, kind::f16: D = f32 (bit 4), A = B = bf16 (1 at bits 7, 10), both K-major, N at 17, M at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
    // shared-memory descriptor of a 128-byte-swizzled K-major tile: SBO = 1024 B (8 rows), version 1, layout 2; LBO unused
    const uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    unsigned g = 0;
    if (WS && n_my > 0) mbar_wait(&w_full, 0u);
    for (int i = 0; i < n_my; ++i) {
      const unsigned a = (unsigned)i & 1u, ause = (unsigned)i >> 1;
      if (ause > 0) mbar_wait(&acc_empty[a], (ause - 1) & 1u);  // the epilogue has drained this accumulator's previous tile
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_addr = tmem + a * (uint32_t)bn;
      for (int k = 0; k < n_slabs; ++k, ++g) {
        const unsigned s = g % (unsigned)n_stages;
        mbar_wait(&full[s], (g / (unsigned)n_stages) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t a16 = smem_u32(ring + (size_t)s * stage_bytes) >> 4;
          const uint32_t b16 = WS ? (smem_u32(wreg + (size_t)k * kOperandBytes) >> 4) : a16 + (kOperandBytes >> 4);
#pragma unroll
          for (int j = 0; j < kSlab / 16; ++j)  // K = 16 per MMA = 32 B inside the 128-byte row: +2 in 16-byte units
            umma_bf16(d_addr, desc_hi | (uint64_t)((a16 + 2u * j) | (1u << 16)), desc_hi | (uint64_t)((b16 + 2u * j) | (1u << 16)),
                      idesc, (k > 0 || j > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
          if (k == n_slabs - 1) umma_commit(&acc_full[a]);
        }
        __syncwarp();
      }
    }
  } else if (warp < warp_p1) {
    const int qd = warp & 3;  // the TMEM lane quadrant a warp may read is fixed by warp id % 4
    const int r = qd * 32 + lane;  // this thread's row of the tile (= TMEM lane)
    const int grp = (warp - 2) >> 2;  // epilogue group: tiles i = grp, grp + groups, ... — or, split, its 128 columns of every tile
    const int bar_id = 1 + grp;
    const bool issuer = (((warp - 2) & 3) == 0 && lane == 0);
    uint8_t* staging = staging_all + (size_t)grp * kStagingBytes;
    const int coff = split ? grp * kTile : 0;
    for (int i = split ? 0 : grp; i < n_my; i += split ? 1 : groups) {
      const unsigned a = (unsigned)i & 1u;
      int row0, col0;
      tile_of(i, row0, col0);
      col0 += coff;
      mbar_wait(&acc_full[a], ((unsigned)i >> 1) & 1u);
      if (col0 >= N) {  // (split, ragged N: this group's half of the tile does not exist)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[a]);
        continue;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // the group's previous tile's stores must have finished READING the staging tile before it is overwritten
      if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      epi_barrier(bar_id);
      const uint32_t t_addr = tmem + ((uint32_t)(qd * 32) << 16) + a * (uint32_t)bn + (uint32_t)coff;
#pragma unroll 1
      for (int cc = 0; cc < kTile; cc += 32) {
        float v[32];
        tmem_ld32(t_addr + (uint32_t)cc, v);
        // 32 columns = four 16-byte chunks of this row's 128-byte line in half cc / 64; SWIZZLE_128B: chunk ^= row % 8
        uint8_t* line = staging + (size_t)(cc >> 6) * (kStagingBytes / 2) + (size_t)r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (((cc & 63) >> 3) + q) ^ (r & 7);
          *reinterpret_cast<uint4*>(line + chunk * 16) =
              make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                         pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[a]);  // the accumulator is in shared memory now
      // column walk: thread = (column half h, 64-row half rh, 32-bit word w = 2 columns); a warp reads the 32 words of one
      // 128-byte line per step: conflict-free in spite of the swizzle
      const int tid = ((warp - 2) & 3) * 32 + lane;
      const int h = tid >> 6, rh = (tid >> 5) & 1, w = tid & 31;
      const size_t half_off = (size_t)h * (kStagingBytes / 2);
      if constexpr (MODE == kBwd) {
        const int c0 = col0 + h * 64 + 2 * w;  // this thread's two columns (N is even)
        const bool cok = c0 < N;
        const float sc0 = cok ? __ldg(bw.scale + c0) : 0.f, sc1 = cok ? __ldg(bw.scale + c0 + 1) : 0.f;
        const float sh0 = cok ? __ldg(bw.shift + c0) : 0.f, sh1 = cok ? __ldg(bw.shift + c0 + 1) : 0.f;
        mbar_wait(&y_full[i & 1], ((unsigned)i >> 1) & 1u);
        epi_barrier(bar_id);  // every row of the staged tile is written
        const uint8_t* yt = ytile + (size_t)(i & 1) * kStagingBytes;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        // 8 rows per batch, the next batch's words loaded BEFORE this batch's write-backs: the compiler cannot prove that
        // the in-place stores do not alias later loads, so a plain row loop ran one shared-memory latency per row
        // (~7 800 cycles per tile measured)
        const uint32_t col_off = (uint32_t)half_off + (uint32_t)(w & 3) * 4u;
        auto word_off = [&](int rr) { return col_off + (uint32_t)rr * 128u + ((uint32_t)((w >> 2) ^ (rr & 7)) << 4); };
        uint32_t gw[8], yw[8], gn[8], yn[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t off = word_off(rh * 64 + u);
          gw[u] = *reinterpret_cast<const uint32_t*>(staging + off);
          yw[u] = *reinterpret_cast<const uint32_t*>(yt + off);
        }
#pragma unroll 1
        for (int r8 = rh * 64; r8 < rh * 64 + 64; r8 += 8) {
          if (r8 + 8 < rh * 64 + 64) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t off = word_off(r8 + 8 + u);
              gn[u] = *reinterpret_cast<const uint32_t*>(staging + off);
              yn[u] = *reinterpret_cast<const uint32_t*>(yt + off);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float y0 = __uint_as_float(yw[u] << 16), y1 = __uint_as_float(yw[u] & 0xffff0000u);
            float g0 = __uint_as_float(gw[u] << 16), g1 = __uint_as_float(gw[u] & 0xffff0000u);
            if (bw.relu) {
              if (!(fmaf(y0, sc0, sh0) > 0.f)) g0 = 0.f;
              if (!(fmaf(y1, sc1, sh1) > 0.f)) g1 = 0.f;
            }
            uint32_t out = (__float_as_uint(g0) >> 16) | (__float_as_uint(g1) & 0xffff0000u);  // exact: bf16 values or 0
            if (bw.thresh) {
              const long long grow = (long long)row0 + r8 + u;
              g0 = keep_elem(bw.seed, grow, c0, N, bw.thresh) ? g0 * bw.keep_scale : 0.f;
              g1 = keep_elem(bw.seed, grow, c0 + 1, N, bw.thresh) ? g1 * bw.keep_scale : 0.f;
              out = pack_bf16(g0, g1);  // the sums are those of the STORED gradient
              g0 = __uint_as_float(out << 16);
              g1 = __uint_as_float(out & 0xffff0000u);
            }
            *reinterpret_cast<uint32_t*>(staging + word_off(r8 + u)) = out;
            s0 += g0; s1 += g1;
            q0 = fmaf(g0, y0, q0); q1 = fmaf(g1, y1, q1);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) { gw[u] = gn[u]; yw[u] = yn[u]; }
        }
        float* dst = s_stat + col0 + h * 64 + 2 * w;
        atomicAdd(dst, s0);
        atomicAdd(dst + 1, s1);
        atomicAdd(dst + ncols, q0);
        atomicAdd(dst + ncols + 1, q1);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging writes -> visible to the copy engine
      epi_barrier(bar_id);
      if (issuer) {
        tma_store_2d(&map_c, staging, col0, row0);
        if (col0 + 64 < N) tma_store_2d(&map_c, staging + kStagingBytes / 2, col0 + 64, row0);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if constexpr (MODE == kBwd) mbar_arrive(&y_empty[i & 1]);  // (after the barrier: every thread is done with the y tile)
      }
      if constexpr (MODE == kStats) {
        // column sums of the STAGED (bf16) tile
        const uint8_t* base = staging + half_off;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
        for (int rr = rh * 64; rr < rh * 64 + 64; ++rr) {
          const uint32_t word = *reinterpret_cast<const uint32_t*>(base + (size_t)rr * 128 + ((((w >> 2) ^ (rr & 7))) << 4) + (w & 3) * 4);
          const float x0 = __uint_as_float(word << 16), x1 = __uint_as_float(word & 0xffff0000u);
          s0 += x0; s1 += x1;
          q0 = fmaf(x0, x0, q0); q1 = fmaf(x1, x1, q1);
        }
        float* dst = s_stat + col0 + h * 64 + 2 * w;
        atomicAdd(dst, s0);
        atomicAdd(dst + 1, s1);
        atomicAdd(dst + ncols, q0);
        atomicAdd(dst + ncols + 1, q1);
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(2u * (unsigned)bn) : "memory");
  if constexpr (STATS) {
    if (n_my > 0) {
      for (int i = threadIdx.x; i < 2 * ncols; i += kThreads) {
        const int half = i / ncols, col = i - half * ncols;
        if (col < N) atomicAdd(stats + (size_t)half * N + col, (double)s_stat[i]);
      }
    }
  }
}

static int encode_bf16_map(CUtensorMap* map, const void* base, long long width, long long ld, long long rows, int box_rows = kTile) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    S4G_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    S4G_CHECK_ARG(fn != nullptr && qres == cudaDriverEntryPointSuccess, "gemm_bf16: cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2u};
  const cuuint32_t box[2] = {(cuuint32_t)kSlab, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  S4G_CHECK_ARG(r == CUDA_SUCCESS, "gemm_bf16: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return S4G_OK;
}

}  // namespace gemm
}  // namespace s4g

static bool g_gemm_ws = true;
static int g_epi_groups = 0;
static int g_tile_n = 0;
// A/B switches for measurements.  weight_stationary: 0 = always stream B through the ring, 1 = weight-stationary where it
// applies (default).  epilogue_groups: 1 or 2 groups of 4 epilogue warps, 0 (default) = chosen per launch.  Both return
// the previous value.
extern "C" int s4g_gemm_bf16_set_weight_stationary(int on) {
  const int prev = g_gemm_ws ? 1 : 0;
  g_gemm_ws = on != 0;
  return prev;
}
extern "C" int s4g_gemm_bf16_set_epilogue_groups(int groups) {
  const int prev = g_epi_groups;
  if (groups >= 0 && groups <= 2) g_epi_groups = groups;
  return prev;
}

// A/B switch for measurements: output columns per tile, 128 or 256, 0 (default) = chosen per launch.  Returns the previous value.
extern "C" int s4g_gemm_bf16_set_tile_n(int bn) {
  const int prev = g_tile_n;
  if (bn == 0 || bn == 128 || bn == 256) g_tile_n = bn;
  return prev;
}

// How one launch is laid out: tile width, epilogue groups, schedule, ring depth, shared memory, grid.  Pure host arithmetic
// (exported as s4g_gemm_bf16_plan so that the CPU tests can check the shared-memory budget of every shape of the model).
struct GemmPlan {
  int bn, groups, ws, n_stages, grid, threads;
  long long smem;
};
constexpr int kMaxDynSmem = 226 * 1024;  // (the kernel also has ~170 bytes of static shared memory; the limit is 227 KB)

static int plan_gemm(long long P, int N, int K, int mode, int sms, GemmPlan* out) {
  using namespace s4g::gemm;
  const int n_slabs = (K + kSlab - 1) / kSlab;
  const long long tiles_m = (P + kTile - 1) / kTile;
  // 256-column tiles (streaming schedule, two epilogue groups splitting the tile): where the MMA-issuing thread sets the
  // pace, i.e. many K slabs per tile and at least 256 output columns.  Measured per shape (profiles/r02/gemm_layers_v6.txt):
  // 10-25 % faster for K >= 256 (512 -> 1024 on 0.5 M rows 0.72 -> 0.58 ms, 1024 -> 512 0.72 -> 0.54), except K = 264 (a fifth
  // slab of 8 columns costs a whole 48 KB stage here but 16 KB on the weight-stationary schedule: 0.58 vs 0.70 ms)
  int bn = kTile;
  if (mode != kBwd && N > kTile) {
    if (g_tile_n == 256) bn = 256;
    else if (g_tile_n == 0 && n_slabs >= 4 && (K % kSlab == 0 || n_slabs >= 8)) bn = 256;
  }
  const int tiles_n = (N + bn - 1) / bn;
  // epilogue groups: two pay where the epilogue, not HBM or the tensor pipe, sets the pace — short K (<= 2 slabs: measured
  // 3 -> 128 on 10.5 M rows 0.82 -> 0.55 ms, 128 -> 256 1.63 -> 1.28 ms); with more slabs per tile the second staging tile
  // only costs ring stages (264 -> 256 on 2.1 M rows 0.49 -> 0.64 ms: it loses the weight-stationary schedule).  The BWD
  // epilogue has one group (its y tiles take the second staging tile's place); 256-column tiles always have two.
  const int groups = bn == 256 ? 2 : mode == kBwd ? 1 : g_epi_groups ? g_epi_groups : (n_slabs <= 2 ? 2 : 1);
  // per-column sums in shared memory
  const size_t table_bytes = sizeof(float) * (size_t)tiles_n * bn * (mode == kPlain ? 0 : 2);
  const long long room = (long long)kMaxDynSmem - 1024 - (long long)(groups + (mode == kBwd ? 2 : 0)) * kStagingBytes -
                         (long long)table_bytes;
  const int stage_bytes = kOperandBytes + (bn / kTile) * kOperandBytes;
  S4G_CHECK_ARG(room >= 2LL * stage_bytes, "gemm_bf16: too many output columns for the fused statistics");
  // weight-stationary when the B slice fits beside >= 5 A stages and every CTA gets >= 2 m-tiles
  // (K = 512 with 3 A stages beside its 128 KB slice: measured 20-30 % SLOWER than streaming)
  int n_stages = (int)(room / stage_bytes);
  if (n_stages > kStages) n_stages = kStages;
  bool ws = false;
  if (bn == kTile && g_gemm_ws && n_slabs <= kMaxSlabsWS && tiles_n <= sms) {
    int stages_a = (int)(room / kOperandBytes) - n_slabs;
    if (stages_a > kMaxStagesA) stages_a = kMaxStagesA;
    ws = stages_a >= 5 && tiles_m >= 2LL * (sms / tiles_n);
    if (ws) n_stages = stages_a;
  }
  const size_t smem = 1024 + (size_t)(groups + (mode == kBwd ? 2 : 0)) * kStagingBytes + table_bytes +
                      (ws ? (size_t)(n_slabs + n_stages) * kOperandBytes : (size_t)n_stages * stage_bytes);
  const long long tiles = tiles_m * tiles_n;
  int grid = (int)(tiles < sms ? tiles : sms);
  if (ws) grid = (sms / tiles_n) * tiles_n;  // every n-tile gets the same number of CTAs
  *out = GemmPlan{bn, groups, ws ? 1 : 0, n_stages, grid, 64 + 128 * groups + 32, (long long)smem};
  return S4G_OK;
}

// out7 = {tile columns, epilogue groups, weight-stationary, ring stages, grid, threads, dynamic shared memory bytes} of the
// launch s4g_gemm_bf16 / _stats / _bwd (mode 0 / 1 / 2) would make for this shape on a GPU with `sms` SMs (0 = the current
// device).  No GPU work.
extern "C" int s4g_gemm_bf16_plan(long long P, int N, int K, int mode, int sms, long long* out7) {
  S4G_CHECK_ARG(out7 && P > 0 && N > 0 && K > 0 && mode >= 0 && mode <= 2, "gemm_bf16_plan: bad arguments");
  GemmPlan pl;
  const int rc = plan_gemm(P, N, K, mode, sms > 0 ? sms : s4g::num_sms(), &pl);
  if (rc != S4G_OK) return rc;
  out7[0] = pl.bn; out7[1] = pl.groups; out7[2] = pl.ws; out7[3] = pl.n_stages; out7[4] = pl.grid; out7[5] = pl.threads;
  out7[6] = pl.smem;
  return S4G_OK;
}

static int gemm_launch(const void* a, long long lda, const void* b, long long ldb, void* c, long long ldc, long long P, int N,
                       int K, int mode, double* stats, const s4g::gemm::BwdEpilogue& bw, void* stream) {
  using namespace s4g::gemm;
  S4G_CHECK_ARG(a && b && c, "gemm_bf16: null pointer");
  S4G_CHECK_ARG(P >= 0 && P < (1ll << 31) - kTile && N > 0 && K > 0, "gemm_bf16: bad shape");
  S4G_CHECK_ARG(lda >= K && ldb >= K && ldc >= N, "gemm_bf16: leading dimension smaller than the row");
  S4G_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0 && ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0 &&
                    ((uintptr_t)c & 15) == 0,
                "gemm_bf16: rows must be 16-byte aligned (leading dimensions multiples of 8 bf16)");
  if (P == 0) return S4G_OK;
  GemmPlan pl;
  int rc = plan_gemm(P, N, K, mode, s4g::num_sms(), &pl);
  if (rc != S4G_OK) return rc;
  const int bn = pl.bn, groups = pl.groups, n_stages = pl.n_stages, grid = pl.grid, threads = pl.threads;
  const bool ws = pl.ws != 0;
  const size_t smem = (size_t)pl.smem;
  CUtensorMap ma, mb;
  rc = encode_bf16_map(&ma, a, K, lda, P);
  if (rc != S4G_OK) return rc;
  rc = encode_bf16_map(&mb, b, K, ldb, N, bn);
  if (rc != S4G_OK) return rc;
  CUtensorMap mc;
  rc = encode_bf16_map(&mc, c, N, ldc, P);
  if (rc != S4G_OK) return rc;
  CUtensorMap my = mc;
  if (mode == kBwd) {
    rc = encode_bf16_map(&my, bw.y, N, bw.ldy, P);
    if (rc != S4G_OK) return rc;
  }
  static bool attr_set[64] = {};
  if (s4g::first_use_on_device(attr_set)) {
    S4G_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<kPlain, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    S4G_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<kPlain, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    S4G_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<kStats, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    S4G_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<kStats, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    S4G_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<kBwd, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    S4G_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<kBwd, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (stats) S4G_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * N, st));
#define S4G_GEMM_GO(MODE_, WS_) \
  gemm_bf16_kernel<MODE_, WS_><<<grid, threads, smem, st>>>(ma, mb, mc, my, (int)P, N, K, n_stages, groups, bn, stats, bw)
  if (mode == kBwd) { if (ws) S4G_GEMM_GO(kBwd, true); else S4G_GEMM_GO(kBwd, false); }
  else if (mode == kStats) { if (ws) S4G_GEMM_GO(kStats, true); else S4G_GEMM_GO(kStats, false); }
  else { if (ws) S4G_GEMM_GO(kPlain, true); else S4G_GEMM_GO(kPlain, false); }
#undef S4G_GEMM_GO
  S4G_LAUNCH_CHECK("gemm_bf16");
  return S4G_OK;
}

extern "C" int s4g_gemm_bf16(const void* a, long long lda, const void* b, long long ldb, void* c, long long ldc, long long P,
                             int N, int K, void* stream) {
  return gemm_launch(a, lda, b, ldb, c, ldc, P, N, K, s4g::gemm::kPlain, nullptr, s4g::gemm::BwdEpilogue{}, stream);
}

// the same with the per-column sum / sum of squares of the stored (bf16-rounded) result: stats2n[0..N) = sum_r c[r][n],
// stats2n[N..2N) = sum_r c[r][n]^2 (fp64, zeroed here) — the BatchNorm batch statistics without another pass over C
extern "C" int s4g_gemm_bf16_stats(const void* a, long long lda, const void* b, long long ldb, void* c, long long ldc,
                                   long long P, int N, int K, double* stats2n, void* stream) {
  S4G_CHECK_ARG(stats2n != nullptr, "gemm_bf16_stats: null statistics buffer");
  return gemm_launch(a, lda, b, ldb, c, ldc, P, N, K, s4g::gemm::kStats, stats2n, s4g::gemm::BwdEpilogue{}, stream);
}

// input-gradient GEMM whose result is the upstream gradient of the block that produced y_prev = the rows [P][N] BEFORE its
// BatchNorm: c = (a · b^T) * relu'(y_prev * scale + shift) * dropout mask  (stored masked, bf16), and
// sums2n[0..N) = sum_r c[r][n], sums2n[N..2N) = sum_r c[r][n] * y_prev[r][n]  (fp64, zeroed here) — what
// s4g_train_bn_bwd_reduce_bf16 would compute from c and y_prev in another pass.
extern "C" int s4g_gemm_bf16_bwd(const void* a, long long lda, const void* b, long long ldb, void* c, long long ldc, long long P,
                                 int N, int K, const void* y_prev, long long ldy, const float* scale, const float* shift,
                                 int relu, unsigned seed, float drop_p, double* sums2n, void* stream) {
  S4G_CHECK_ARG(y_prev && scale && shift && sums2n, "gemm_bf16_bwd: null pointer");
  S4G_CHECK_ARG(ldy >= N && ldy % 8 == 0 && N % 8 == 0 && ((uintptr_t)y_prev & 15) == 0 && drop_p >= 0.f && drop_p < 1.f,
                "gemm_bf16_bwd: y_prev rows must be 16-byte aligned, N a multiple of 8");
  s4g::gemm::BwdEpilogue bw;
  bw.y = reinterpret_cast<const __nv_bfloat16*>(y_prev);
  bw.ldy = ldy;
  bw.scale = scale;
  bw.shift = shift;
  bw.relu = relu;
  bw.seed = seed;
  bw.thresh = drop_p > 0.f ? (unsigned)((double)drop_p * 4294967296.0) : 0u;
  bw.keep_scale = 1.f / (1.f - drop_p);
  return gemm_launch(a, lda, b, ldb, c, ldc, P, N, K, s4g::gemm::kBwd, sums2n, bw, stream);
}
