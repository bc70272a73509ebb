// Micro-benchmark (measurement tool, not product code): cost per job of the MMA-issue loop of the fused
// chain kernel in WARP-UNIFORM code (uniform-register descriptors), with its parts switched on one by one:
//   jobs of 4 x tcgen05.mma (N = 256) or 8 x (N = 128); job table in kernel-parameter space; 1..3
//   tcgen05.commit per job; 1..3 mbarrier.try_wait per job on barriers that are already complete.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_loop mma_loop.cu && ./mma_loop
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}

struct Job { uint8_t slot, acc, k16, n8, flags, pad[3]; };
struct Params { Job job[32]; int n_jobs; int tiles; int n_commit; int n_wait; int use_table; int whole_warp; int spinners; int spin_mode; };

__global__ void __launch_bounds__(640, 1) k(const __grid_constant__ Params p, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 192 * 1024);  // [0..7] commit targets, [8] final, [9] pre-completed
  __shared__ uint32_t tslot;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bars[9])) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = __shfl_sync(0xffffffffu, tslot, 0);
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 192 * 1024 + 512);
  if (threadIdx.x == 0) *stop = 0;
  __syncthreads();
  if (warp >= 1 && warp <= p.spinners) {
    // other roles of the chain kernel waiting on mbarriers that are not complete: spin_mode 0 = every lane
    // polls with try_wait (as mbar_wait does), 1 = one lane polls, 2 = test_wait + nanosleep backoff
    while (!*stop) {
      if (p.spin_mode == 0) { try_wait(&bars[10 + (warp & 3)], 0); }
      else if (p.spin_mode == 1) { if ((threadIdx.x & 31) == 0) try_wait(&bars[10 + (warp & 3)], 0); __syncwarp(); }
      else { try_wait(&bars[10 + (warp & 3)], 0); __nanosleep(200); }
    }
  }
  if (warp == 0) {
    const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint32_t a0 = s32(smem) >> 4, b0 = s32(smem + 128 * 1024) >> 4;
    const long long t0 = clock64();
    for (int it = 0; it < p.tiles; ++it) {
      for (int j = 0; j < p.n_jobs; ++j) {
        Job job;
        if (p.use_table) job = p.job[j];
        else { job.slot = j & 3; job.acc = j & 1; job.k16 = p.job[0].k16; job.n8 = p.job[0].n8; job.flags = 0; }
        for (int w = 0; w < p.n_wait; ++w) while (!try_wait(&bars[9], 0)) {}
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t n = (uint32_t)job.n8 * 8u;
          const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
          uint32_t a_lo = (a0 + (uint32_t)job.slot * 2048u) | (128u << 16);
          uint32_t w_lo = (b0 + (uint32_t)(j & 1) * 2048u) | (n << 16);
          const uint32_t d = tm + (uint32_t)job.acc * 256u;
#pragma unroll 1
          for (int q = 0; q < job.k16; ++q) {
            umma(d, hi | a_lo, hi | w_lo, idesc, 1u);
            a_lo += 256u;
            w_lo += 2u * n;
          }
          for (int c = 0; c < p.n_commit; ++c) commit(&bars[(j + c) & 7]);
        }
        __syncwarp();
      }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) {
      commit(&bars[8]);
      while (!try_wait(&bars[8], 0)) {}
      out[blockIdx.x * 2] = t1 - t0;
      out[blockIdx.x * 2 + 1] = clock64() - t0;
      *stop = 1;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
  printf("  N k16 commits waits table | issue cyc/job  total cyc/job  pipe cyc/job\n");
  struct Cfg { int n8, k16, commits, waits, table, spinners, mode; };
  Cfg cfgs[] = {{32, 4, 0, 0, 0}, {32, 4, 1, 0, 0}, {32, 4, 1, 1, 0}, {32, 4, 1, 1, 1}, {32, 4, 2, 2, 1}, {32, 4, 3, 3, 1},
                {16, 8, 1, 1, 1}, {16, 8, 3, 3, 1}, {16, 4, 1, 1, 1},
                {32, 4, 1, 1, 1, 4, 0}, {32, 4, 1, 1, 1, 8, 0}, {32, 4, 1, 1, 1, 19, 0}, {32, 4, 1, 1, 1, 19, 1},
                {32, 4, 1, 1, 1, 19, 2}, {16, 8, 1, 1, 1, 19, 0}, {16, 8, 1, 1, 1, 19, 2}};
  for (const Cfg& c : cfgs) {
    Params p = {};
    p.n_jobs = 23; p.tiles = 200; p.n_commit = c.commits; p.n_wait = c.waits; p.use_table = c.table; p.spinners = c.spinners; p.spin_mode = c.mode;
    for (int j = 0; j < 32; ++j) { p.job[j].slot = j & 3; p.job[j].acc = j & 1; p.job[j].k16 = c.k16; p.job[j].n8 = c.n8; }
    k<<<148, 640, 194 * 1024>>>(p, out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[148 * 2];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < 148; ++i) { a += h[i * 2]; b += h[i * 2 + 1]; }
    const double jobs = 23.0 * 200;
    printf("%3d %3d %7d %5d %5d spin %2d/%d | %13.1f  %13.1f  %12.1f\n", c.n8 * 8, c.k16, c.commits, c.waits, c.table,
           c.spinners, c.mode, a / 148 / jobs,
           b / 148 / jobs, c.k16 * (c.n8 * 8 >= 128 ? c.n8 * 8 / 2.0 : 64.0));
    fflush(stdout);
  }
  return 0;
}
