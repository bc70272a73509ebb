// Micro-benchmark (measurement tool, not product code): how fast two loader warps (64 threads) stage a 128-row x
// 128-channel bf16 block (32 KB) from global memory into the [16 pieces][128 rows][16 B] shared-memory layout with
// 16-byte cp.async, for the lane -> (row, piece) mappings the MLP-chain loader could use:
//   map 0: thread = row (rows r and r + 64), 16 pieces each          (32 cache lines per warp instruction)
//   map 1: lane = (row % 8) + 8 * (piece % 4): 8 rows x 64 B          (8 lines per instruction)
//   map 2: lane = (row % 2) + 2 * piece: 2 rows x 256 B                (4 lines per instruction; 16-way bank conflicts)
// rows are `stride` channels apart (256 = the head chain's input); `hot` = the same tile every time (L2 / L1 hits),
// else a fresh tile per iteration from a 1 GB buffer (HBM).  All 148 SMs run the same loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cpasync_bw cpasync_bw.cu && ./cpasync_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(64, 1) k(const uint8_t* src, size_t n_tiles, int stride_bytes, int map, int hot, int iters,
                                           int depth, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];  // depth x 32 KB
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  long long issue = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const size_t tile = hot ? blockIdx.x : ((size_t)blockIdx.x + (size_t)i * gridDim.x) % n_tiles;
    const uint8_t* base = src + tile * 128 * (size_t)stride_bytes;
    uint8_t* dst = smem + (size_t)(i % depth) * 32768;
    const long long a = clock64();
    if (map == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = t + 64 * h;
#pragma unroll 4
        for (int c = 0; c < 16; ++c) cp16(dst + (size_t)c * 2048 + row * 16, base + (size_t)row * stride_bytes + c * 16);
      }
    } else if (map == 1) {
      const int lr = lane & 7, lp = lane >> 3;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const int row = 64 * warp + 8 * g + lr;
#pragma unroll
        for (int c = lp; c < 16; c += 4) cp16(dst + (size_t)c * 2048 + row * 16, base + (size_t)row * stride_bytes + c * 16);
      }
    } else {
      const int lr = lane & 1, c = lane >> 1;
#pragma unroll 4
      for (int g = 0; g < 32; ++g) {
        const int row = 64 * warp + 2 * g + lr;
        cp16(dst + (size_t)c * 2048 + row * 16, base + (size_t)row * stride_bytes + c * 16);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    issue += clock64() - a;
    if (depth == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else asm volatile("cp.async.wait_group 1;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const long long t1 = clock64();
  if (t == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = issue; }
}

int main() {
  const size_t bytes = 1ull << 30;
  uint8_t* src;
  long long* out;
  cudaMalloc(&src, bytes);
  cudaMemset(src, 1, bytes);
  cudaMalloc(&out, 148 * 2 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 32768);
  printf("map stride hot depth | cyc/32KB block  issue cyc/block  B/cyc/SM  TB/s (148 SMs @1.965 GHz)\n");
  const int iters = 200;
  for (int hot = 1; hot >= 0; --hot)
    for (int depth = 1; depth <= 2; ++depth)
      for (int map = 0; map < 3; ++map) {
        const int stride = 512;
        const size_t n_tiles = bytes / (128 * (size_t)stride);
        k<<<148, 64, 2 * 32768>>>(src, n_tiles, stride, map, hot, iters, depth, out);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long h[148 * 2];
        cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
        double a = 0, b = 0;
        for (int i = 0; i < 148; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
        a /= 148.0 * iters;
        b /= 148.0 * iters;
        printf("%3d %6d %3d %5d | %14.0f  %15.0f  %8.1f  %6.2f\n", map, stride, hot, depth, a, b, 32768.0 / a,
               32768.0 / a * 148 * 1.965e9 / 1e12);
        fflush(stdout);
      }
  return 0;
}
