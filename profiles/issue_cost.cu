// Micro-benchmark (measurement tool, not product code): cost, for the ISSUING thread, of the instruction
// mix of one MMA job of the fused chain kernel: n_mma x tcgen05.mma (M=128,N) + n_commit x tcgen05.commit
// + n_wait x mbarrier.try_wait on already-completed barriers.  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_cost issue_cost.cu && ./issue_cost
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, unsigned par) {
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}

__global__ void __launch_bounds__(128, 1) k(int N, int jobs, int n_mma, int n_commit, int n_wait, int whole_warp, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 128 * 1024);  // [0..7] commit targets, [8] final, [9] pre-completed
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bars[9])) : "memory");  // phase 0 complete
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  if (warp == 0 && (whole_warp || lane == 0)) {
    const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = s32(smem) >> 4, b0 = s32(smem + 64 * 1024) >> 4;
    const long long t0 = clock64();
    long long t_issue_end = 0;
    for (int j = 0; j < jobs; ++j) {
      for (int w = 0; w < n_wait; ++w) while (!try_wait(&bars[9], 0)) {}
      bool leader = true;
      if (whole_warp) {
        unsigned pred;
        asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
        leader = pred != 0;
      }
      if (leader) {
        for (int i = 0; i < n_mma; ++i) {
          const uint32_t ka = (uint32_t)(i & 7);
          umma(tm + (uint32_t)((j & 1) * 256), hi | ((a0 + ka * 256u) | (128u << 16)), hi | ((b0 + ka * 2u * (uint32_t)N) | ((uint32_t)N << 16)), idesc, 1u);
        }
        for (int c = 0; c < n_commit; ++c) commit(&bars[(j + c) & 7]);
      }
      if (whole_warp) __syncwarp();
    }
    t_issue_end = clock64();
    if (lane == 0) {
      commit(&bars[8]);
      while (!try_wait(&bars[8], 0)) {}
      const long long t1 = clock64();
      out[blockIdx.x * 2 + 0] = t_issue_end - t0;
      out[blockIdx.x * 2 + 1] = t1 - t0;
    }
    if (whole_warp) __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  printf("  N n_mma n_commit n_wait warp | issue cyc/job  total cyc/job\n");
  struct Cfg { int N, n_mma, n_commit, n_wait, ww; };
  Cfg cfgs[] = {{128, 4, 0, 0, 0}, {128, 4, 1, 0, 0}, {128, 4, 2, 0, 0}, {128, 4, 3, 0, 0}, {128, 4, 1, 1, 0}, {128, 4, 1, 3, 0},
                {128, 4, 1, 1, 1}, {128, 4, 1, 3, 1}, {128, 8, 1, 1, 1}, {256, 4, 1, 1, 1}, {256, 4, 1, 3, 1}, {128, 0, 1, 0, 0},
                {128, 0, 0, 1, 0}, {128, 1, 1, 1, 1}, {16, 4, 1, 1, 1}, {64, 4, 1, 1, 1}, {32, 4, 1, 0, 0}, {16, 1, 0, 0, 0}};
  const int jobs = 4000;
  for (const Cfg& c : cfgs) {
    k<<<148, 128, 130 * 1024>>>(c.N, jobs, c.n_mma, c.n_commit, c.n_wait, c.ww, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148 * 2];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[i * 2]; b += h[i * 2 + 1]; }
    printf("%3d %5d %8d %6d %4d | %12.1f  %13.1f\n", c.N, c.n_mma, c.n_commit, c.n_wait, c.ww, a / 148 / jobs, b / 148 / jobs);
    fflush(stdout);
  }
  return 0;
}
