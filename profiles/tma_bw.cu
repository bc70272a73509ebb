// Micro-benchmark (measurement tool, not product code): per-SM ingest bandwidth of cp.async.bulk
// (TMA 1-D copies) from an L2-resident buffer as a function of chunk size and chunks in flight.
// Answers: how many bytes must be in flight per SM to stream MLP weights at a given rate?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu && ./tma_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32, 1) bw_kernel(const uint8_t* src, size_t src_bytes, int chunk, int stages, int iters) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)stages * chunk);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t n_chunks = src_bytes / chunk;
    size_t c = (size_t)blockIdx.x * 7;
    for (int i = 0; i < iters + stages; ++i) {
      const int s = i % stages;
      if (i >= stages) {  // wait for the copy issued `stages` iterations ago
        unsigned par = ((i / stages) - 1) & 1, ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(&bar[s])), "r"(par) : "memory");
      }
      if (i < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(smem + (size_t)s * chunk)), "l"(src + (c % n_chunks) * chunk), "r"(chunk), "r"(s32(&bar[s])) : "memory");
        ++c;
      }
    }
  }
}

int main() {
  const size_t src_bytes = 4u << 20;  // 4 MB: L2 resident, like one chain's weights
  uint8_t* src;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  int sms = 148;
  printf("chunk_KB stages grid  GB/s_total  GB/s_per_SM  us_per_chunk\n");
  for (int grid : {1, 148}) {
    for (int chunk : {4096, 8192, 16384, 32768}) {
      for (int stages : {1, 2, 3, 4, 6, 8, 12}) {
        if ((size_t)chunk * stages > 190 * 1024) continue;
        const int iters = 4000;
        size_t smem = (size_t)chunk * stages + 256;
        bw_kernel<<<grid, 32, smem>>>(src, src_bytes, chunk, stages, 200);
        cudaEventRecord(a);
        bw_kernel<<<grid, 32, smem>>>(src, src_bytes, chunk, stages, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double tot = (double)grid * iters * chunk / (ms * 1e-3) / 1e9;
        printf("%7d %6d %4d %11.1f %11.1f %12.3f\n", chunk / 1024, stages, grid, tot, tot / grid, ms * 1e3 / iters);
      }
    }
  }
  (void)sms;
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
