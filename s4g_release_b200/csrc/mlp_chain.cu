// Fused shared-MLP chain on the 5th-generation tensor cores (tcgen05 + TMEM) for sm_100a.
//
// Replaces, for eval-mode inference, the reference's per-layer Conv{1,2}d -> BatchNorm -> ReLU kernels
// (nn_utils/conv.py:30-36,70-76; nn_utils/mlp.py:95-106) together with what surrounds them in
//   * PointNetSAModule.forward (pointnet2_utils/modules.py:208-244): group_points gather, centroid
//     subtraction, channel concat, 3 layers, max over the K neighbours;
//   * PointnetFPModule.forward (:498-507) and the four heads of PointNet2.forward
//     (models/PointNet2_tcls.py:126-148): plain chains, the last 1x1 conv with bias (+ sigmoid).
// BatchNorm (running statistics) is folded into the bf16 weights and an fp32 per-channel shift on the
// host, so a layer is  y = relu(W' x + b').
//
// One CTA owns a tile of 128 "positions" (rows) and walks the whole chain for it; intermediate
// activations never leave the SM:
//     act  (smem, bf16, K-major core-matrix layout [C/8][128][8])   -- A operand of the next layer
//     W    (global, pre-tiled by the host in the same layout, streamed by cp.async.bulk = TMA 1-D
//           copies into a ring of 32 KB stages, completion on mbarriers)
//     D    (TMEM, fp32, 128 lanes x <=512 columns)                   -- tcgen05.mma accumulators
// Per layer the control thread issues tcgen05.mma (M=128, N<=256, K=16, kind::f16 with bf16 inputs),
// tcgen05.commit signals completion, and the four worker warps run the epilogue straight out of TMEM
// (tcgen05.ld 32x32b): + shift, ReLU, bf16 pack, 16-byte conflict-free stores back into `act`.
// The last layer of a set-abstraction chain runs TRANSPOSED (A = weights, B = act), so TMEM lanes are
// output channels and the K=64 neighbours of a centroid are 64 adjacent columns: the max-pool is an
// in-register reduction inside the epilogue and only (centroid, channel) results reach HBM.
//
// Shared-memory operand layout (no swizzle, "interleaved" canonical K-major layout of UMMA):
//   element (row r, channel k) at byte  ((k/8) * ROWS + r) * 16 + (k%8)*2
//   -> core matrix = 8 rows x 16 B contiguous, SBO (8-row groups) = 128 B, LBO (K chunks) = ROWS*16 B.
#include <cuda_bf16.h>

#include "common.cuh"

namespace s4g {

constexpr int kTileRows = 128;
constexpr int kWorkerThreads = 256;                 // 8 warps: two per TMEM lane quadrant
constexpr int kControlWarp = kWorkerThreads / 32;
constexpr int kProducerWarp = kControlWarp + 1;
constexpr int kChainThreads = kWorkerThreads + 64;  // + MMA-issue warp + TMA-producer warp
constexpr int kStageBytes = 32768;                  // one weight chunk: <=256 rows x 64 channels bf16
constexpr int kMaxPhases = 16;
constexpr int kMaxLayers = 6;

enum InMode { IN_ROWS = 0, IN_GATHER = 1 };
enum Action { ACT_LOAD_A = 0, ACT_EPI_HIDDEN = 1, ACT_EPI_ROWS = 2, ACT_EPI_MAXPOOL = 3, ACT_EPI_LOGITS = 4 };

struct Phase {
  int layer;
  int k_begin, k_end;  // input-channel range accumulated in this phase (multiples of 16)
  int n_begin, n_end;  // output-channel range produced in this phase
  int n_chunk;         // rows per weight chunk (<=256; 128 when transposed)
  int k_chunk;         // channels per weight chunk: n_chunk * k_chunk * 2 B <= 32 KB (one ring stage)
  int transposed;      // 1: A = weights, B = act
  int first;           // 1: first accumulation phase of its (layer, n-range): overwrite TMEM
  int action;          // what the workers do once the phase's MMAs are complete
  int load_begin, load_end;  // ACT_LOAD_A: channel range of layer-0 input to stage next; also used by
                             // an epilogue phase that must restage part 0 for the next n-range
  int reload;          // 1: after the epilogue, restage [load_begin, load_end) of the layer-0 input
  int relu;
};

struct ChainParams {
  Phase phase[kMaxPhases];
  int n_phases;
  const __nv_bfloat16* weights;  // all chunks of one tile, in consumption order
  unsigned w_bytes;              // total bytes of `weights`
  const float* bias[kMaxLayers];
  int bias_off[kMaxLayers], bias_len[kMaxLayers], bias_total;  // shifts staged in smem (floats)
  int n_layers;
  int P;                         // rows (positions)
  int act_c;                     // capacity of the act buffer in channels
  int stages;
  int tmem_cols;
  // input
  int in_mode;
  const __nv_bfloat16* in_rows;  // IN_ROWS: [P][in_stride] channel-last
  int in_stride;
  int in_first_load_end;         // channels staged before phase 0
  const __nv_bfloat16* feat;     // IN_GATHER: [B*N][feat_c] channel-last features (may be null: feat_c = 0)
  int feat_c;
  const float* xyz;              // (B,3,N)
  const float* ctr;              // (B,3,M)
  const int* nbr;                // (B,M,K) int32
  int N, M, K;
  // output
  void* out;
  int out_c;                     // real output channels (row stride of ROWS / MAXPOOL outputs)
  int group;                     // MAXPOOL: rows per group (K neighbours), divides 128
  int n_points;                  // LOGITS: points per batch element (channel-first output (B, out_c, n_points))
  int sigmoid;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
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
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> cudaErrorLaunchFailure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const unsigned n = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, unsigned cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, unsigned cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; both operands K-major bf16, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc),
      "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor: SWIZZLE_NONE, K-major, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor: kind::f16, A = B = bf16, D = f32, both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// worker: stage channels [c_begin, c_end) of the layer-0 input of `tile` into act (local channel 0..).
// Thread (r, hh): row r of the tile, 16-byte pieces hh, hh+2, ... of that row (conflict-free smem
// writes: consecutive rows are consecutive 16-byte slots of one K chunk).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_input(const ChainParams& p, uint8_t* act, int tile, int c_begin, int c_end,
                                            int r, int hh) {
  const long long row = (long long)tile * kTileRows + r;
  const bool valid = row < p.P;
  uint8_t* dst = act + (size_t)r * 16;
  if (p.in_mode == IN_ROWS) {
    const __nv_bfloat16* src = p.in_rows + (valid ? row : 0) * (long long)p.in_stride + c_begin;
    const int pieces = (c_end - c_begin) >> 3;
    for (int c = hh; c < pieces; c += 2) cp_async16(dst + (size_t)c * (kTileRows * 16), src + c * 8, valid);
  } else {
    // row -> (b, m, k); gathered feature row first, then the 16-wide relative-xyz chunk
    const int per_b = p.M * p.K;
    const long long rr = valid ? row : 0;
    const int b = (int)(rr / per_b);
    const int m = (int)((rr - (long long)b * per_b) / p.K);
    const int j = __ldg(p.nbr + rr);
    const int fc = p.feat_c;
    if (c_begin < fc) {
      const int e = min(c_end, fc);
      const __nv_bfloat16* src = p.feat + ((long long)b * p.N + j) * fc + c_begin;
      const int pieces = (e - c_begin) >> 3;
      for (int c = hh; c < pieces; c += 2) cp_async16(dst + (size_t)c * (kTileRows * 16), src + c * 8, valid);
    }
    if (c_end > fc && hh == 0) {  // the xyz chunk [fc, fc+16) lies in this part
      const float* X = p.xyz + (long long)b * 3 * p.N;
      const float* C = p.ctr + (long long)b * 3 * p.M;
      float dx = 0.f, dy = 0.f, dz = 0.f;
      if (valid) {
        dx = __fsub_rn(__ldg(X + j), __ldg(C + m));
        dy = __fsub_rn(__ldg(X + p.N + j), __ldg(C + p.M + m));
        dz = __fsub_rn(__ldg(X + 2 * p.N + j), __ldg(C + 2 * p.M + m));
      }
      const int c0 = (fc - c_begin) >> 3;
      *reinterpret_cast<uint4*>(dst + (size_t)c0 * (kTileRows * 16)) =
          make_uint4(pack_bf16(dx, dy), pack_bf16(dz, 0.f), 0u, 0u);
      *reinterpret_cast<uint4*>(dst + (size_t)(c0 + 1) * (kTileRows * 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
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

// issue the (non-blocking) TMEM load of 32 / 16 accumulator columns; pair with tmem_wait()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* v) {
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
}
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// one 32-column slab of a row-oriented epilogue: + shift, ReLU, bf16 pack, 4 x 16-byte stores
template <bool TO_GLOBAL>
__device__ __forceinline__ void epi_store32(const float* v, const float* sb, int relu, uint8_t* dst, size_t piece_stride) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float4 b0 = *reinterpret_cast<const float4*>(sb + g * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(sb + g * 8 + 4);
    float o[8] = {v[g * 8 + 0] + b0.x, v[g * 8 + 1] + b0.y, v[g * 8 + 2] + b0.z, v[g * 8 + 3] + b0.w,
                  v[g * 8 + 4] + b1.x, v[g * 8 + 5] + b1.y, v[g * 8 + 6] + b1.z, v[g * 8 + 7] + b1.w};
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaxf(o[e], 0.f);
    }
    *reinterpret_cast<uint4*>(dst + g * piece_stride) =
        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel: warps 0..7 = workers (input staging + epilogues), warp 8 = control (TMA + MMA issue)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kChainThreads, 1) mlp_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // carve: [act | weight ring | bias | barriers]
  uint8_t* act = smem;
  uint8_t* ring = smem + (size_t)p.act_c * kTileRows * 2;
  float* s_bias = reinterpret_cast<float*>(ring + (size_t)p.stages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + p.bias_total);
  uint64_t* full = bars;                     // [stages] weights landed
  uint64_t* empty = bars + p.stages;         // [stages] MMAs that read the stage are complete
  uint64_t* mma_done = bars + 2 * p.stages;  // a phase's MMAs are complete
  uint64_t* act_ready = mma_done + 1;        // workers staged / rewrote act and drained TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 1);

  const int warp = threadIdx.x >> 5;
  const int n_tiles = (p.P + kTileRows - 1) / kTileRows;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(mma_done, 1);
    mbar_init(act_ready, kWorkerThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int l = 0; l < p.n_layers; ++l)
    for (int c = threadIdx.x; c < p.bias_len[l]; c += kChainThreads) s_bias[p.bias_off[l] + c] = __ldg(p.bias[l] + c);
  if (warp == kControlWarp) tmem_alloc(tmem_slot, (unsigned)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // ============== TMA producer warp: streams the weight chunks of every tile, in consumption order ==============
    int s = 0, wrap = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.weights);
      for (int q = 0; q < p.n_phases; ++q) {
        const Phase& ph = p.phase[q];
        const int n_n = (ph.n_end - ph.n_begin) / ph.n_chunk;
        for (int k = ph.k_begin; k < ph.k_end; k += ph.k_chunk) {
          const unsigned bytes = (unsigned)min(ph.k_chunk, ph.k_end - k) * (unsigned)ph.n_chunk * 2u;
          for (int i = 0; i < n_n; ++i) {
            if (wrap > 0) mbar_wait(&empty[s], (unsigned)((wrap - 1) & 1));
            if (elect_one()) {
              mbar_expect_tx(&full[s], bytes);
              bulk_g2s(ring + (size_t)s * kStageBytes, src, bytes, &full[s]);
            }
            __syncwarp();
            src += bytes;
            if (++s == p.stages) { s = 0; ++wrap; }
          }
        }
      }
    }
  } else if (warp == kControlWarp) {
    // ============== MMA warp (converged; one elected lane issues tcgen05.mma / commit) ==============
    int s = 0, wrap = 0;
    unsigned ready_count = 0;
    const uint32_t act16 = smem_u32(act) >> 4;
    const uint32_t ring16 = smem_u32(ring) >> 4;
    // descriptor high word: SBO = 128 B, version 1;  low word: (addr >> 4) | (LBO >> 4) << 16
    const uint64_t desc_hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint32_t a_lbo16 = (uint32_t)kTileRows;  // (128 rows * 16 B) >> 4
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int q = 0; q < p.n_phases; ++q) {
        const Phase& ph = p.phase[q];
        const uint32_t idesc = make_idesc(128, ph.transposed ? kTileRows : ph.n_chunk);
        const uint32_t w_lbo16 = (uint32_t)ph.n_chunk;  // (n_chunk rows * 16 B) >> 4
        const int n_n = (ph.n_end - ph.n_begin) / ph.n_chunk;
        const bool transposed = ph.transposed != 0;
        mbar_wait(act_ready, ready_count & 1);  // act staged / previous epilogue done with act and TMEM
        ++ready_count;
        for (int k = ph.k_begin; k < ph.k_end; k += ph.k_chunk) {
          const int steps = min(ph.k_chunk, ph.k_end - k) >> 4;
          const bool last_k = k + ph.k_chunk >= ph.k_end;
          const uint32_t a_lo0 = (act16 + (uint32_t)((k - ph.k_begin) >> 3) * a_lbo16) | (a_lbo16 << 16);
          const uint32_t acc0 = (ph.first && k == ph.k_begin) ? 0u : 1u;
          for (int i = 0; i < n_n; ++i) {
            mbar_wait(&full[s], (unsigned)(wrap & 1));
            tc_fence_after();
            if (elect_one()) {
              uint32_t a_lo = a_lo0;
              uint32_t w_lo = (ring16 + (uint32_t)s * (kStageBytes >> 4)) | (w_lbo16 << 16);
              const uint32_t d_addr = tmem_base + (uint32_t)(i * ph.n_chunk);
              uint32_t acc = acc0;
#pragma unroll 4
              for (int j = 0; j < steps; ++j) {
                if (transposed) umma_bf16(d_addr, desc_hi | w_lo, desc_hi | a_lo, idesc, acc);
                else umma_bf16(d_addr, desc_hi | a_lo, desc_hi | w_lo, idesc, acc);
                acc = 1u;
                a_lo += 2u * a_lbo16;  // next 16 channels = 2 K pieces
                w_lo += 2u * w_lbo16;
              }
              umma_commit(&empty[s]);  // stage reusable once these MMAs have read it
              if (last_k && i == n_n - 1) umma_commit(mma_done);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ++wrap; }
          }
        }
      }
    }
  } else {
    // ======================== worker warps: input staging + epilogues ========================
    const int r = threadIdx.x & (kTileRows - 1);  // row of the tile == TMEM lane (quadrant = warp & 3)
    const int hh = threadIdx.x >> 7;              // which half of the columns / pieces this warp takes
    const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    unsigned done_count = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long row = (long long)tile * kTileRows + r;
      stage_input(p, act, tile, p.phase[0].k_begin, p.in_first_load_end, r, hh);
      cp_async_wait_all();
      fence_proxy_async();
      mbar_arrive(act_ready);
      for (int q = 0; q < p.n_phases; ++q) {
        const Phase& ph = p.phase[q];
        mbar_wait(mma_done, done_count & 1);
        ++done_count;
        tc_fence_after();
        const float* sb = s_bias + p.bias_off[ph.layer];
        if (ph.action == ACT_LOAD_A) {
          stage_input(p, act, tile, ph.load_begin, ph.load_end, r, hh);
          cp_async_wait_all();
        } else if (ph.action == ACT_EPI_HIDDEN || ph.action == ACT_EPI_ROWS) {
          // 32-column slabs, alternating between the two warps of a lane quadrant; the TMEM load of the
          // next slab is in flight while the current one is processed
          const bool to_global = ph.action == ACT_EPI_ROWS;
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
          const int width = ph.n_end - ph.n_begin;
          if ((width & 31) == 0) {
            float va[32], vb[32];
            int c = ph.n_begin + 32 * hh;
            if (c < ph.n_end) tmem_ld32_issue(lane_base + (uint32_t)(c - ph.n_begin), va);
            for (; c < ph.n_end; c += 128) {
              tmem_wait();
              const int c2 = c + 64;
              if (c2 < ph.n_end) tmem_ld32_issue(lane_base + (uint32_t)(c2 - ph.n_begin), vb);
              if (to_global) {
                if (row < p.P) epi_store32<true>(va, sb + c, ph.relu, reinterpret_cast<uint8_t*>(out + row * (long long)p.out_c + c), 16);
              } else {
                epi_store32<false>(va, sb + c, ph.relu, act + ((size_t)(c >> 3) * kTileRows + r) * 16, kTileRows * 16);
              }
              if (c2 < ph.n_end) {
                tmem_wait();
                const int c3 = c2 + 64;
                if (c3 < ph.n_end) tmem_ld32_issue(lane_base + (uint32_t)(c3 - ph.n_begin), va);
                if (to_global) {
                  if (row < p.P) epi_store32<true>(vb, sb + c2, ph.relu, reinterpret_cast<uint8_t*>(out + row * (long long)p.out_c + c2), 16);
                } else {
                  epi_store32<false>(vb, sb + c2, ph.relu, act + ((size_t)(c2 >> 3) * kTileRows + r) * 16, kTileRows * 16);
                }
              }
            }
          } else {
            // narrow layers (width a multiple of 16 only): 16-column slabs, hidden activations only
            for (int c = ph.n_begin + 16 * hh; c < ph.n_end; c += 32) {
              float v[16];
              tmem_ld16(lane_base + (uint32_t)(c - ph.n_begin), v);
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  const float t = v[g * 8 + e] + sb[c + g * 8 + e];
                  o[e] = ph.relu ? fmaxf(t, 0.f) : t;
                }
                const uint4 pk = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]),
                                            pack_bf16(o[6], o[7]));
                if (to_global) {
                  if (row < p.P) *reinterpret_cast<uint4*>(out + row * (long long)p.out_c + c + g * 8) = pk;
                } else {
                  *reinterpret_cast<uint4*>(act + ((size_t)((c >> 3) + g) * kTileRows + r) * 16) = pk;
                }
              }
            }
          }
        } else if (ph.action == ACT_EPI_MAXPOOL) {
          // transposed: TMEM lane = output channel, columns = the tile's 128 positions; a group of
          // `group` adjacent columns is one centroid's neighbourhood.  max commutes with (+shift, ReLU).
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
          const int G = p.group;
          const long long group0 = (long long)tile * (kTileRows / G);
          const long long n_groups = p.P / G;
          // groups <= 64 wide: warp half hh owns columns [64 hh, 64 hh + 64); a 128-wide group: half 0 only
          const int col0 = (G <= 64) ? 64 * hh : 0;
          const int col1 = (G <= 64) ? col0 + 64 : (hh == 0 ? kTileRows : 0);
          for (int cb = ph.n_begin; cb < ph.n_end; cb += kTileRows) {
            const int ch = cb + r;
            const float b = sb[ch];
            float run = 0.f;
            for (int s0 = col0; s0 < col1; s0 += 32) {
              float v[32];
              tmem_ld32(lane_base + (uint32_t)(cb - ph.n_begin + s0), v);
              if (G >= 32) {
                float m = v[0];
#pragma unroll
                for (int e = 1; e < 32; ++e) m = fmaxf(m, v[e]);
                run = (s0 % G == 0) ? m : fmaxf(run, m);
                if ((s0 + 32) % G == 0) {
                  const long long gi = group0 + s0 / G;
                  if (gi < n_groups && ch < p.out_c) {
                    const float t = run + b;
                    out[gi * p.out_c + ch] = __float2bfloat16(ph.relu ? fmaxf(t, 0.f) : t);
                  }
                }
              } else if (G == 16) {
#pragma unroll
                for (int g = 0; g < 32; g += 16) {
                  float m = v[g];
#pragma unroll
                  for (int e = 1; e < 16; ++e) m = fmaxf(m, v[g + e]);
                  const long long gi = group0 + (s0 + g) / 16;
                  if (gi < n_groups && ch < p.out_c) {
                    const float t = m + b;
                    out[gi * p.out_c + ch] = __float2bfloat16(ph.relu ? fmaxf(t, 0.f) : t);
                  }
                }
              } else {  // G == 8 (or smaller powers of two folded into 8-wide pieces by the host check)
#pragma unroll
                for (int g = 0; g < 32; g += 8) {
                  float m = v[g];
#pragma unroll
                  for (int e = 1; e < 8; ++e) m = fmaxf(m, v[g + e]);
                  const long long gi = group0 + (s0 + g) / 8;
                  if (gi < n_groups && ch < p.out_c) {
                    const float t = m + b;
                    out[gi * p.out_c + ch] = __float2bfloat16(ph.relu ? fmaxf(t, 0.f) : t);
                  }
                }
              }
            }
          }
        } else if (hh == 0) {  // ACT_EPI_LOGITS: fp32, channel-first (B, out_c, n_points), bias, optional sigmoid
          float* out = reinterpret_cast<float*>(p.out);
          float v[16];
          tmem_ld16(lane_base, v);
          if (row < p.P) {
            const long long bb = row / p.n_points;
            const long long n = row - bb * p.n_points;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              if (c < p.out_c) {
                float t = v[c] + sb[c];
                if (ph.relu) t = fmaxf(t, 0.f);
                if (p.sigmoid) t = 1.f / (1.f + __expf(-t));
                out[(bb * p.out_c + c) * p.n_points + n] = t;
              }
            }
          }
        }
        if (ph.reload) {
          stage_input(p, act, tile, ph.load_begin, ph.load_end, r, hh);
          cp_async_wait_all();
        }
        tc_fence_before();
        if (q + 1 < p.n_phases) {
          fence_proxy_async();
          mbar_arrive(act_ready);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kControlWarp) tmem_dealloc(tmem_base, (unsigned)p.tmem_cols);
}

}  // namespace s4g

// ================================================================================================
// Host side: chain plan (phases, smem/TMEM budget), weight packing, launch
// ================================================================================================
#include <new>
#include <vector>

struct s4g_chain {
  s4g::ChainParams prm;
  int n_layers;
  int cin_pad[s4g::kMaxLayers], cout_pad[s4g::kMaxLayers];
  int in_mode, out_mode;
  size_t smem_bytes;
  int ctas_per_sm;
};

namespace s4g {

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

static int pick_n_chunk(int width) {
  for (int c = 256; c >= 16; c -= 16)
    if (width % c == 0) return c;
  return 16;
}

static int plan_chain(s4g_chain* ch, int n_layers, const int* cin, const int* cout, const int* relu, int in_mode,
                      int feat_c, int out_mode, int out_c, int group, int sigmoid) {
  ChainParams& p = ch->prm;
  memset(ch, 0, sizeof(*ch));
  S4G_CHECK_ARG(n_layers >= 1 && n_layers <= kMaxLayers, "mlp_chain: 1..%d layers supported", kMaxLayers);
  S4G_CHECK_ARG(in_mode == IN_ROWS || in_mode == IN_GATHER, "mlp_chain: bad in_mode");
  S4G_CHECK_ARG(out_mode >= ACT_EPI_ROWS && out_mode <= ACT_EPI_LOGITS, "mlp_chain: bad out_mode");
  ch->n_layers = n_layers;
  ch->in_mode = in_mode;
  ch->out_mode = out_mode;
  const int L = n_layers;
  for (int l = 0; l < L; ++l) {
    S4G_CHECK_ARG(cin[l] > 0 && cout[l] > 0, "mlp_chain: bad layer width");
    if (l > 0) S4G_CHECK_ARG(cin[l] == cout[l - 1], "mlp_chain: layer %d input width != previous output width", l);
    ch->cin_pad[l] = round_up(cin[l], 16);
    ch->cout_pad[l] = round_up(cout[l], 16);
  }
  if (in_mode == IN_GATHER) {
    S4G_CHECK_ARG(feat_c % 8 == 0 && cin[0] == feat_c + 3, "mlp_chain: gather input must be feat_c(%%8==0) + 3 xyz");
    ch->cin_pad[0] = feat_c + 16;
  } else {
    S4G_CHECK_ARG(cin[0] % 8 == 0, "mlp_chain: row input width must be a multiple of 8");
  }
  for (int l = 0; l + 1 < L; ++l) {
    S4G_CHECK_ARG(ch->cout_pad[l] == cout[l] && cout[l] <= 512, "mlp_chain: hidden width must be a multiple of 16, <= 512");
  }
  if (out_mode == ACT_EPI_MAXPOOL) {
    ch->cout_pad[L - 1] = round_up(cout[L - 1], 128);
    S4G_CHECK_ARG(group == 8 || group == 16 || group == 32 || group == 64 || group == 128,
                  "mlp_chain: max-pool group (neighbours per centroid) must be 8, 16, 32, 64 or 128");
  } else if (out_mode == ACT_EPI_LOGITS) {
    S4G_CHECK_ARG(cout[L - 1] <= 16, "mlp_chain: logits layer supports <= 16 outputs");
    ch->cout_pad[L - 1] = 16;
  } else {
    S4G_CHECK_ARG(cout[L - 1] % 32 == 0, "mlp_chain: row output width must be a multiple of 32");
  }
  // layer-0 input parts (K passes) and last-layer output parts (N passes)
  std::vector<std::pair<int, int>> kparts, nparts;
  const int kMaxAct = 544;
  if (ch->cin_pad[0] <= kMaxAct) kparts.push_back({0, ch->cin_pad[0]});
  else {
    S4G_CHECK_ARG(in_mode == IN_ROWS, "mlp_chain: gathered input wider than %d channels", kMaxAct);
    for (int k = 0; k < ch->cin_pad[0]; k += 512) kparts.push_back({k, std::min(k + 512, ch->cin_pad[0])});
  }
  for (int n = 0; n < ch->cout_pad[L - 1]; n += 512) nparts.push_back({n, std::min(n + 512, ch->cout_pad[L - 1])});
  S4G_CHECK_ARG(kparts.size() == 1 || nparts.size() == 1 || L == 1, "mlp_chain: unsupported shape");
  int np = 0, act_c = 16, tmem = 32;
  for (auto& kp : kparts) act_c = std::max(act_c, kp.second - kp.first);
  for (int l = 0; l < L; ++l) {
    const bool last = (l == L - 1);
    std::vector<std::pair<int, int>> ks, ns;
    if (l == 0) ks = kparts; else ks.push_back({0, ch->cin_pad[l]});
    if (last) ns = nparts; else ns.push_back({0, ch->cout_pad[l]});
    if (!last) act_c = std::max(act_c, ch->cout_pad[l]);
    for (size_t ni = 0; ni < ns.size(); ++ni) {
      for (size_t ki = 0; ki < ks.size(); ++ki) {
        S4G_CHECK_ARG(np < kMaxPhases, "mlp_chain: too many phases");
        Phase& ph = p.phase[np++];
        ph.layer = l;
        ph.k_begin = ks[ki].first; ph.k_end = ks[ki].second;
        ph.n_begin = ns[ni].first; ph.n_end = ns[ni].second;
        ph.transposed = (last && out_mode == ACT_EPI_MAXPOOL) ? 1 : 0;
        ph.n_chunk = ph.transposed ? 128 : pick_n_chunk(ph.n_end - ph.n_begin);
        ph.k_chunk = std::max(16, (kStageBytes / 2 / ph.n_chunk) / 16 * 16);  // ~32 KB per TMA request
        ph.first = (ki == 0);
        ph.relu = relu[l];
        ph.reload = 0;
        if (ki + 1 < ks.size()) {
          ph.action = ACT_LOAD_A;
          ph.load_begin = ks[ki + 1].first; ph.load_end = ks[ki + 1].second;
        } else {
          ph.action = last ? out_mode : ACT_EPI_HIDDEN;
          if (ks.size() > 1 && ni + 1 < ns.size()) {
            ph.reload = 1;
            ph.load_begin = ks[0].first; ph.load_end = ks[0].second;
          }
        }
        tmem = std::max(tmem, ph.n_end - ph.n_begin);
      }
    }
  }
  p.n_phases = np;
  p.act_c = act_c;
  p.n_layers = L;
  int boff = 0;
  for (int l = 0; l < L; ++l) {
    p.bias_off[l] = boff;
    p.bias_len[l] = ch->cout_pad[l];
    boff += ch->cout_pad[l];
  }
  p.bias_total = boff;
  int cols = 32;
  while (cols < tmem) cols *= 2;
  S4G_CHECK_ARG(cols <= 512, "mlp_chain: accumulator wider than TMEM");
  p.tmem_cols = cols;
  p.in_mode = in_mode;
  p.in_first_load_end = kparts[0].second;
  p.feat_c = feat_c;
  p.out_c = out_c;
  p.group = group > 0 ? group : 1;
  p.sigmoid = sigmoid;
  // weight bytes per tile, in consumption order
  size_t wb = 0;
  for (int q = 0; q < np; ++q) {
    const Phase& ph = p.phase[q];
    wb += (size_t)(ph.k_end - ph.k_begin) * (ph.n_end - ph.n_begin) * 2;
  }
  p.w_bytes = (unsigned)wb;
  // shared memory: act + ring + barriers
  const size_t act_bytes = (size_t)act_c * kTileRows * 2;
  const size_t tail = (size_t)p.bias_total * 4 + 256;  // staged shifts + barriers
  const size_t budget = 227 * 1024 - tail;
  int stages = (int)((budget - act_bytes) / kStageBytes);
  S4G_CHECK_ARG(stages >= 2, "mlp_chain: activation tile leaves no room for the weight ring");
  // two CTAs per SM hide each other's epilogue when TMEM (<=256 columns) and smem allow it
  ch->ctas_per_sm = 1;
  if (cols <= 256) {
    const size_t half = (227 * 1024) / 2 - 1024 - tail;
    if (act_bytes + 2 * kStageBytes <= half) {
      ch->ctas_per_sm = 2;
      stages = std::min(stages, (int)((half - act_bytes) / kStageBytes));
    }
  }
  stages = std::min(stages, 4);
  p.stages = stages;
  ch->smem_bytes = act_bytes + (size_t)stages * kStageBytes + tail;
  return S4G_OK;
}

}  // namespace s4g

extern "C" s4g_chain* s4g_chain_create(int n_layers, const int* cin, const int* cout, const int* relu, int in_mode,
                                       int feat_c, int out_mode, int out_c, int group, int sigmoid) {
  s4g_chain* ch = new (std::nothrow) s4g_chain;
  if (!ch) return nullptr;
  if (s4g::plan_chain(ch, n_layers, cin, cout, relu, in_mode, feat_c, out_mode, out_c, group, sigmoid) != S4G_OK) {
    delete ch;
    return nullptr;
  }
  return ch;
}

extern "C" void s4g_chain_destroy(s4g_chain* ch) { delete ch; }

extern "C" size_t s4g_chain_weight_bytes(const s4g_chain* ch) { return ch ? ch->prm.w_bytes : 0; }

extern "C" int s4g_chain_cout_pad(const s4g_chain* ch, int layer) {
  return (ch && layer >= 0 && layer < ch->n_layers) ? ch->cout_pad[layer] : -1;
}

extern "C" int s4g_chain_info(const s4g_chain* ch, int* n_phases, int* act_c, int* stages, int* tmem_cols,
                              int* smem_bytes, int* ctas_per_sm) {
  S4G_CHECK_ARG(ch != nullptr, "mlp_chain: null chain");
  if (n_phases) *n_phases = ch->prm.n_phases;
  if (act_c) *act_c = ch->prm.act_c;
  if (stages) *stages = ch->prm.stages;
  if (tmem_cols) *tmem_cols = ch->prm.tmem_cols;
  if (smem_bytes) *smem_bytes = (int)ch->smem_bytes;
  if (ctas_per_sm) *ctas_per_sm = ch->ctas_per_sm;
  return S4G_OK;
}

// Pack one layer's folded fp32 weights W[cout_real][cin_real] (row-major, host memory) into the bf16
// chunk stream the kernel consumes.  For a gathered first layer the reference's channel order
// [xyz(3) | features] (modules.py:48) becomes [features | xyz(3) | 0 x 13].
extern "C" int s4g_chain_pack_weights(const s4g_chain* ch, int layer, const float* w_host, int cout_real, int cin_real,
                                      void* packed_host) {
  S4G_CHECK_ARG(ch && w_host && packed_host, "mlp_chain: null pointer");
  S4G_CHECK_ARG(layer >= 0 && layer < ch->n_layers, "mlp_chain: bad layer");
  const s4g::ChainParams& p = ch->prm;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(packed_host);
  size_t off = 0;  // in elements
  const bool gather0 = (layer == 0 && ch->in_mode == s4g::IN_GATHER);
  const int fc = p.feat_c;
  for (int q = 0; q < p.n_phases; ++q) {
    const s4g::Phase& ph = p.phase[q];
    for (int k = ph.k_begin; k < ph.k_end; k += ph.k_chunk) {
      const int kw = std::min(ph.k_chunk, ph.k_end - k);
      for (int n = ph.n_begin; n < ph.n_end; n += ph.n_chunk) {
        if (ph.layer == layer) {
          for (int kk = 0; kk < kw; ++kk) {
            const int c = k + kk;  // channel in the kernel's order
            int src_c;
            if (gather0) src_c = (c < fc) ? c + 3 : (c < fc + 3 ? c - fc : -1);
            else src_c = (c < cin_real) ? c : -1;
            for (int rr = 0; rr < ph.n_chunk; ++rr) {
              const int o = n + rr;
              float v = 0.f;
              if (src_c >= 0 && src_c < cin_real && o < cout_real) v = w_host[(size_t)o * cin_real + src_c];
              dst[off + ((size_t)(kk >> 3) * ph.n_chunk + rr) * 8 + (kk & 7)] = __float2bfloat16(v);
            }
          }
        }
        off += (size_t)kw * ph.n_chunk;
      }
    }
  }
  return S4G_OK;
}

extern "C" int s4g_chain_set_params(s4g_chain* ch, const void* weights_dev, const float* const* bias_dev) {
  S4G_CHECK_ARG(ch && weights_dev && bias_dev, "mlp_chain: null pointer");
  S4G_CHECK_ARG(((uintptr_t)weights_dev & 15) == 0, "mlp_chain: weights must be 16-byte aligned");
  ch->prm.weights = reinterpret_cast<const __nv_bfloat16*>(weights_dev);
  for (int l = 0; l < ch->n_layers; ++l) ch->prm.bias[l] = bias_dev[l];
  return S4G_OK;
}

static int s4g_chain_launch(const s4g_chain* ch, s4g::ChainParams& p, cudaStream_t stream) {
  S4G_CHECK_ARG(p.weights != nullptr, "mlp_chain: s4g_chain_set_params was not called");
  if (p.P <= 0) return S4G_OK;
  static bool attr_set = false;
  if (!attr_set) {
    S4G_CUDA(cudaFuncSetAttribute(s4g::mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int tiles = (p.P + s4g::kTileRows - 1) / s4g::kTileRows;
  const int grid = std::min(tiles, s4g::num_sms() * ch->ctas_per_sm);
  s4g::mlp_chain_kernel<<<grid, s4g::kChainThreads, ch->smem_bytes, stream>>>(p);
  S4G_LAUNCH_CHECK("mlp_chain");
  return S4G_OK;
}

// rows in: in_rows [P][in_stride] bf16 channel-last.  out: ROWS -> bf16 [P][out_c];
// LOGITS -> fp32 (P / n_points, out_c, n_points).
extern "C" int s4g_chain_run_rows(const s4g_chain* ch, const void* in_rows, int in_stride, long long P, void* out,
                                  int n_points, void* stream) {
  S4G_CHECK_ARG(ch && in_rows && out, "mlp_chain: null pointer");
  S4G_CHECK_ARG(ch->in_mode == s4g::IN_ROWS, "mlp_chain: chain was planned for gathered input");
  S4G_CHECK_ARG(P >= 0 && P < (1ll << 31), "mlp_chain: bad row count");
  S4G_CHECK_ARG(in_stride >= ch->cin_pad[0] && in_stride % 8 == 0 && ((uintptr_t)in_rows & 15) == 0,
                "mlp_chain: input rows must be 16-byte aligned and at least cin wide");
  if (ch->out_mode == s4g::ACT_EPI_LOGITS) S4G_CHECK_ARG(n_points > 0 && P % n_points == 0, "mlp_chain: bad n_points");
  s4g::ChainParams p = ch->prm;
  p.P = (int)P;
  p.in_rows = reinterpret_cast<const __nv_bfloat16*>(in_rows);
  p.in_stride = in_stride;
  p.out = out;
  p.n_points = n_points > 0 ? n_points : 1;
  return s4g_chain_launch(ch, p, (cudaStream_t)stream);
}

// gathered in (set abstraction): feat [B*N][feat_c] bf16 (or NULL when feat_c == 0), xyz (B,3,N) fp32,
// ctr (B,3,M) fp32, nbr (B,M,K) int32.  out: MAXPOOL -> bf16 [B*M][out_c]; ROWS -> bf16 [B*M*K][out_c].
extern "C" int s4g_chain_run_gather(const s4g_chain* ch, const void* feat, const float* xyz, const float* ctr,
                                    const int* nbr, int B, int N, int M, int K, void* out, void* stream) {
  S4G_CHECK_ARG(ch && xyz && ctr && nbr && out, "mlp_chain: null pointer");
  S4G_CHECK_ARG(ch->in_mode == s4g::IN_GATHER, "mlp_chain: chain was planned for row input");
  S4G_CHECK_ARG(ch->prm.feat_c == 0 || (feat != nullptr && ((uintptr_t)feat & 15) == 0), "mlp_chain: bad feature table");
  S4G_CHECK_ARG((long long)B * M * K < (1ll << 31), "mlp_chain: too many rows");
  if (ch->out_mode == s4g::ACT_EPI_MAXPOOL) S4G_CHECK_ARG(K == ch->prm.group, "mlp_chain: K != planned max-pool group");
  s4g::ChainParams p = ch->prm;
  p.P = B * M * K;
  p.feat = reinterpret_cast<const __nv_bfloat16*>(feat);
  p.xyz = xyz;
  p.ctr = ctr;
  p.nbr = nbr;
  p.N = N; p.M = M; p.K = K;
  p.out = out;
  return s4g_chain_launch(ch, p, (cudaStream_t)stream);
}
