// One shared-MLP layer on the tensor cores in TF32 — the TIGHT-PARITY mode of the fused path ("a TF32/fp32
// fallback mode for tight parity runs").
//
//   Y[P][N] = act( X[P][K] · W[N][K]^T + shift[N] )        X, W, Y fp32 row-major, fp32 accumulation in TMEM
//
// replaces one Conv{1,2}d(1x1, bias-free) -> BatchNorm(eval, folded on the host) -> ReLU block of the reference
// (nn_utils/conv.py:30-36,70-76) at the arithmetic the UNMODIFIED reference itself gets on this GPU: torch's cuDNN
// convolutions default to TF32 (torch.backends.cudnn.allow_tf32 = True), i.e. 10-bit mantissa operands, fp32 accumulate.
// The product (throughput) path is csrc/mlp_chain.cu in bf16; this kernel trades its fusion for operand precision and is
// what `FusedPointNet2(mlp_backend="tf32")` runs layer by layer.
//
// Shape of the kernel (one 128 x 128 output tile per CTA, 192 threads):
//   warp 0    TMA producer: per K-slab of 32 fp32 (= one 128-byte swizzle row) one X box {32, 128} and one W box
//             {32, 128} into a 4-stage shared-memory ring (SWIZZLE_128B; out-of-range rows / channels are zero-filled by
//             the copy engine, so P, N, K need no padding);
//   warp 1    allocates 128 TMEM columns and issues 4 x tcgen05.mma.kind::tf32 (M = 128, N = 128, K = 8) per slab;
//             tcgen05.commit releases the stage, the last commit publishes the accumulator;
//   warps 2-5 epilogue: tcgen05.ld of the warp's lane quadrant, + shift, ReLU, optional round-to-nearest to TF32 (so the
//             NEXT layer's operand truncation inside the tensor core is exact and unbiased), fp32 rows to global memory.
#include <cuda.h>

#include "common.cuh"

namespace s4g {
namespace lin {

constexpr int kThreads = 192;
constexpr int kStages = 4;
constexpr int kTile = 128;              // rows of X and rows of W (= output columns) per CTA
constexpr int kSlab = 32;               // K elements per stage: 32 fp32 = 128 B = one swizzle row
constexpr int kOperandBytes = kTile * kSlab * 4;  // 16 KB
constexpr int kStageBytes = 2 * kOperandBytes;
constexpr int kSmemBytes = kStages * kStageBytes + 1024;  // + slack to align the ring to 1024 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
    if (!ok && ++spins > (1u << 24)) __trap();  // a protocol bug traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const void* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
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
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
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

__global__ void __launch_bounds__(kThreads, 1)
linear_tf32_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                   const float* __restrict__ shift, float* __restrict__ y, long long ldy, int P, int N, int K, int relu,
                   int round_out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kStages], empty[kStages], acc_full;
  __shared__ uint32_t tmem_slot;
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzle atoms are 1024-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kTile, col0 = blockIdx.y * kTile;
  const int n_slabs = (K + kSlab - 1) / kSlab;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      for (int k = 0; k < n_slabs; ++k) {
        const int s = k % kStages;
        if (k >= kStages) mbar_wait(&empty[s], (unsigned)(k / kStages - 1) & 1u);
        mbar_expect_tx(&full[s], kStageBytes);  // zero-filled out-of-range elements count as transferred bytes
        tma_load_2d(ring + (size_t)s * kStageBytes, &map_x, k * kSlab, row0, &full[s]);
        tma_load_2d(ring + (size_t)s * kStageBytes + kOperandBytes, &map_w, k * kSlab, col0, &full[s]);
      }
    }
  } else if (warp == 1) {
    // instruction descriptor, kind::tf32: D = f32 (bit 4), A = B = tf32 (2 at bits 7, 10), both K-major, N at 17, M at 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
    // shared-memory descriptor of a 128-byte-swizzled K-major tile: SBO = 1024 B (8 rows), version 1, layout 2; LBO unused
    const uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
    for (int k = 0; k < n_slabs; ++k) {
      const int s = k % kStages;
      mbar_wait(&full[s], (unsigned)(k / kStages) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a16 = smem_u32(ring + (size_t)s * kStageBytes) >> 4;
        const uint32_t b16 = a16 + (kOperandBytes >> 4);
#pragma unroll
        for (int j = 0; j < kSlab / 8; ++j)  // K = 8 per MMA = 32 B inside the 128-byte row: +2 in 16-byte units
          umma_tf32(tmem, desc_hi | (uint64_t)((a16 + 2u * j) | (1u << 16)), desc_hi | (uint64_t)((b16 + 2u * j) | (1u << 16)),
                    idesc, (k > 0 || j > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (k == n_slabs - 1) umma_commit(&acc_full);
      }
      __syncwarp();
    }
  } else {
    const int qd = warp & 3;  // the TMEM lane quadrant a warp may read is fixed by warp id % 4
    const long long row = (long long)row0 + qd * 32 + lane;
    mbar_wait(&acc_full, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* out = y + row * ldy + col0;
#pragma unroll 1
    for (int c = 0; c < kTile; c += 32) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)c, v);
      if (row < P) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = col0 + c + e;
          if (col < N) {
            float t = v[e] + __ldg(shift + col);
            if (relu) t = fmaxf(t, 0.f);
            if (round_out) t = round_tf32(t);
            out[c + e] = t;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
}

static int encode_f32_map(CUtensorMap* map, const void* base, long long width, long long ld, long long rows) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    S4G_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    S4G_CHECK_ARG(fn != nullptr && qres == cudaDriverEntryPointSuccess, "linear_tf32: cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)width, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4u};
  const cuuint32_t box[2] = {(cuuint32_t)kSlab, (cuuint32_t)kTile};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  S4G_CHECK_ARG(r == CUDA_SUCCESS, "linear_tf32: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return S4G_OK;
}

}  // namespace lin
}  // namespace s4g

extern "C" int s4g_linear_tf32(const float* x, long long ldx, const float* w, long long ldw, const float* shift, float* y,
                               long long ldy, long long P, int N, int K, int relu, int round_out, void* stream) {
  using namespace s4g::lin;
  S4G_CHECK_ARG(x && w && shift && y, "linear_tf32: null pointer");
  S4G_CHECK_ARG(P >= 0 && P < (1ll << 31) - kTile && N > 0 && K > 0, "linear_tf32: bad shape");
  S4G_CHECK_ARG(ldx >= K && ldw >= K && ldy >= N, "linear_tf32: leading dimension smaller than the row");
  S4G_CHECK_ARG(ldx % 4 == 0 && ldw % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0,
                "linear_tf32: X and W rows must be 16-byte aligned (leading dimensions multiples of 4 floats)");
  if (P == 0) return S4G_OK;
  CUtensorMap mx, mw;
  int rc = encode_f32_map(&mx, x, K, ldx, P);
  if (rc != S4G_OK) return rc;
  rc = encode_f32_map(&mw, w, K, ldw, N);
  if (rc != S4G_OK) return rc;
  static bool attr_set[64] = {};
  if (s4g::first_use_on_device(attr_set)) {
    S4G_CUDA(cudaFuncSetAttribute(linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  dim3 grid((unsigned)((P + kTile - 1) / kTile), (unsigned)((N + kTile - 1) / kTile));
  S4G_CHECK_ARG(grid.y <= 65535, "linear_tf32: too many output channels");
  linear_tf32_kernel<<<grid, kThreads, kSmemBytes, (cudaStream_t)stream>>>(mx, mw, shift, y, ldy, (int)P, N, K, relu,
                                                                          round_out);
  S4G_LAUNCH_CHECK("linear_tf32");
  return S4G_OK;
}
