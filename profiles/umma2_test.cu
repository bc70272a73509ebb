// Functional probe (measurement tool, not product code) of tcgen05.mma.cta_group::2 with the no-swizzle K-major
// layout used by the chain kernel: a CTA pair computes D[256 x N] = A[256 x K] * B[N x K]^T where each CTA holds
// its own 128 rows of A and HALF of B's rows at the same shared-memory offsets; the leader issues, the commit is
// multicast to both CTAs.  Also the transposed use (A = per-CTA weights, B = both CTAs' activations).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma2_test umma2_test.cu && ./umma2_test
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, unsigned par) {
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
  return ok;
}

constexpr int K = 64;

// A: [256][K] bf16 (rows 0-127 -> CTA 0, 128-255 -> CTA 1); B: [N][K] bf16 (rows split in halves); D: [256][N] fp32
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int N) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;               // [K/8][128][16 B]
  uint8_t* sb = smem + 32 * 1024;   // [K/8][N/2][16 B]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  __shared__ uint32_t tslot;
  const unsigned rank = cg::this_cluster().block_rank();
  const int warp = threadIdx.x >> 5;
  const int half = N / 2;
  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    const int r = i / K, c = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sa + ((size_t)(c / 8) * 128 + r) * 16 + (c % 8) * 2) = A[(size_t)(rank * 128 + r) * K + c];
  }
  for (int i = threadIdx.x; i < half * K; i += blockDim.x) {
    const int r = i / K, c = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sb + ((size_t)(c / 8) * half + r) * 16 + (c % 8) * 2) = B[(size_t)(rank * half + r) * K + c];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cg::this_cluster().sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  if (rank == 0 && threadIdx.x == 0) {
    const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a_lo = ((s32(sa) >> 4) + ks * 256u) | (128u << 16);
      const uint32_t b_lo = ((s32(sb) >> 4) + ks * 2u * (uint32_t)half) | ((uint32_t)half << 16);
      asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tm),
                   "l"(hi | a_lo), "l"(hi | b_lo), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)),
                 "h"((unsigned short)3) : "memory");
  }
  while (!try_wait(bar, 0)) {}
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // each warp reads its 32 lanes, N columns
  const int row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int e = 0; e < 16; ++e) D[(size_t)(rank * 128 + row) * N + c0 + e] = __uint_as_float(v[e]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cg::this_cluster().sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  for (int N : {256, 128, 64}) {
    std::vector<__nv_bfloat16> hA(256 * K), hB((size_t)N * K);
    std::vector<float> fA(256 * K), fB((size_t)N * K);
    srand(N);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)((rand() % 7) - 3); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)((rand() % 5) - 2); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 256 * (size_t)N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 256 * (size_t)N * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<2, 128, 100 * 1024>>>(dA, dB, dD, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    std::vector<float> hD(256 * (size_t)N);
    cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int r = 0; r < 256; ++r)
      for (int c = 0; c < N; ++c) {
        float acc = 0;
        for (int kk = 0; kk < K; ++kk) acc += fA[r * K + kk] * fB[(size_t)c * K + kk];
        maxerr = fmax(maxerr, fabs(acc - hD[(size_t)r * N + c]));
      }
    printf("cta_group::2 M=256 N=%d K=%d: max abs error vs CPU %.3g %s\n", N, K, maxerr, maxerr == 0 ? "(exact)" : "MISMATCH");
    fflush(stdout);
  }
  return 0;
}
