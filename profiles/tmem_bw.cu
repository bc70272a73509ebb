// Micro-benchmark (measurement tool, not product code): TMEM -> register bandwidth of tcgen05.ld
// (32x32b.x32: 32 lanes x 32 columns x 4 B = 4 KB per warp-instruction) with 4 / 8 / 16 warps per SM and
// 1 or 2 loads in flight per warp; and the cost of fence.proxy.async + a per-warp mbarrier arrive.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(taddr, r)                                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                      \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                       \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"       \
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),        \
                 "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),      \
                 "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),      \
                 "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
               : "r"(taddr))

__global__ void __launch_bounds__(512, 1) k(int warps, int depth, int fence, int iters, long long* out, unsigned* sink) {
  __shared__ uint32_t tslot;
  __shared__ uint64_t bar;
  __shared__ uint4 scratch[512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1024;" ::"r"(s32(&bar)));
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 32);
  unsigned acc = 0;
  const long long t0 = clock64();
  if (warp < warps) {
    uint32_t a[32], b[32];
    for (int i = 0; i < iters; ++i) {
      LD32(tm, a);
      if (depth == 2) LD32(tm + 128, b);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int e = 0; e < 32; ++e) acc += a[e];
      if (depth == 2) {
#pragma unroll
        for (int e = 0; e < 32; ++e) acc += b[e];
      }
      if (fence) {
        scratch[threadIdx.x] = make_uint4(acc, 1, 2, 3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (fence > 1) {
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
        }
      }
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && warp < warps) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tslot), "r"(512u) : "memory");
}

int main() {
  long long* out;
  unsigned* sink;
  cudaMalloc(&out, 148 * 16 * sizeof(long long));
  cudaMalloc(&sink, 4);
  printf("warps depth fence | cyc/iter/warp   TMEM B/cyc/SM\n");
  const int iters = 2000;
  for (int fence : {0, 1, 2})
    for (int depth : {1, 2})
      for (int warps : {4, 8, 16}) {
        cudaMemset(out, 0, 148 * 16 * sizeof(long long));
        k<<<148, 512>>>(warps, depth, fence, iters, out, sink);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long h[148 * 16];
        cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (int i = 0; i < 148; ++i) for (int w = 0; w < warps; ++w) mx += h[i * 16 + w];
        mx /= 148.0 * warps;
        printf("%5d %5d %5d | %13.1f %15.1f\n", warps, depth, fence, mx / iters, (double)warps * depth * 4096.0 * iters / mx);
        fflush(stdout);
      }
  return 0;
}
