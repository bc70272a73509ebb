// Micro-benchmark (measurement tool, not product code): tcgen05.mma issue rate from shared-memory operands
// (SS mode, bf16, M = 128, no-swizzle K-major layout) alone and under concurrent shared-memory traffic
// (a 1-D TMA weight stream into a ring, and STS.128 epilogue-like stores).  Answers: is the fused MLP
// chain bounded by shared-memory bandwidth rather than by the tensor pipe?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bw umma_bw.cu && ./umma_bw
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

// warp 0: MMA issue; warps 1..tma_warps: TMA streams; warps 8..15: STS traffic (if sts != 0)
__global__ void __launch_bounds__(512, 1) k(const uint8_t* src, size_t src_bytes, int N, int n_mma, int tma_warps, int chunk,
                                            int sts, int same_operand, long long* out, int fill = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [0,64K) A region, [64K,128K) B region, [128K, 128K+64K) TMA ring (4 x 16K), [192K, 208K) STS scratch, then barriers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 208 * 1024);
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  if (fill) {  // operand data: 1 = random bf16 in [-1, 1), 2 = zeros (tensor-pipe power depends on the data)
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    uint32_t x = 0x9E3779B9u * (threadIdx.x + 1) + blockIdx.x;
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) {
      x = x * 1664525u + 1013904223u;
      const uint32_t lo = 0x3F00u | ((x >> 9) & 0x807Fu), hi = 0x3F00u | ((x >> 20) & 0x807Fu);
      w[i] = fill == 1 ? (lo | (hi << 16)) : 0u;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  volatile int* stop = reinterpret_cast<volatile int*>(smem + 208 * 1024 + 512);
  if (threadIdx.x == 0) *stop = 0;
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {
      const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t a0 = s32(smem) >> 4, b0 = s32(smem + 64 * 1024) >> 4;
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t ka = same_operand ? 0u : (uint32_t)(i & 7);
        const uint32_t a_lo = (a0 + ka * 256u) | (128u << 16);            // A: 128 rows, LBO 2048 B
        const uint32_t b_lo = (b0 + ka * 2u * (uint32_t)N) | ((uint32_t)N << 16);  // B: N rows
        umma(tm + (uint32_t)((i & 1) * 256), hi | a_lo, hi | b_lo, idesc, 1u);
      }
      commit(&bars[31]);
      while (!try_wait(&bars[31], 0)) {}
      const long long t1 = clock64();
      out[blockIdx.x * 4 + 0] = t1 - t0;
      *stop = 1;
    }
    __syncwarp();
  } else if (warp <= tma_warps) {
    if (lane == 0) {
      // each TMA warp owns 2 ring stages of `chunk` bytes and keeps both in flight until told to stop
      uint8_t* ring = smem + 128 * 1024 + (size_t)(warp - 1) * 2 * chunk;
      uint64_t* b = bars + (warp - 1) * 2;
      const size_t n_chunks = src_bytes / chunk;
      size_t c = (size_t)blockIdx.x * 7 + warp * 13;
      long long n = 0;
      unsigned par[2] = {0, 0};
      for (int s = 0; s < 2; ++s) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&b[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + (size_t)s * chunk)), "l"(src + (c++ % n_chunks) * chunk), "r"(chunk), "r"(s32(&b[s])) : "memory");
      }
      int s = 0;
      while (!*stop) {
        while (!try_wait(&b[s], par[s])) {}
        par[s] ^= 1;
        ++n;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&b[s])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ring + (size_t)s * chunk)), "l"(src + (c++ % n_chunks) * chunk), "r"(chunk), "r"(s32(&b[s])) : "memory");
        s ^= 1;
      }
      for (int q = 0; q < 2; ++q) { while (!try_wait(&b[s], par[s])) {} s ^= 1; }
      atomicAdd((unsigned long long*)&out[blockIdx.x * 4 + 1], (unsigned long long)n * chunk);
    }
    __syncwarp();
  } else if (warp >= 8 && sts) {
    uint4* dst = reinterpret_cast<uint4*>(smem + 192 * 1024) + (threadIdx.x - 256);
    long long n = 0;
    uint4 v = make_uint4(lane, warp, 3, 4);
    while (!*stop) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { dst[i * 256] = v; v.x += 1; }
      n += 4;
      if (sts > 1) __nanosleep(sts);
    }
    if (lane == 0) atomicAdd((unsigned long long*)&out[blockIdx.x * 4 + 2], (unsigned long long)n * 512);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

int main() {
  const size_t src_bytes = 4u << 20;
  uint8_t* src;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 0, src_bytes);
  long long* out;
  cudaMalloc(&out, 148 * 4 * sizeof(long long));
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("N  same tma_warps chunk sts | cyc/MMA  flop/cyc/SM  tma_B/cyc/SM  sts_B/cyc/SM   (grid 148)\n");
  const int n_mma = 200000;  // ~13 ms per configuration: long enough for the power limiter to act
  struct Cfg { int N, same, tw, chunk, sts, fill; };
  Cfg cfgs[] = {{128, 1, 0, 16384, 0}, {128, 0, 0, 16384, 0}, {256, 0, 0, 16384, 0}, {64, 0, 0, 16384, 0},
                {128, 0, 1, 16384, 0}, {128, 0, 2, 16384, 0}, {128, 0, 4, 16384, 0}, {128, 0, 2, 8192, 0},
                {128, 0, 1, 32768, 0}, {256, 0, 2, 16384, 0}, {256, 0, 4, 16384, 0},
                {128, 0, 0, 16384, 1}, {128, 0, 0, 16384, 200}, {256, 0, 0, 16384, 1}, {128, 0, 2, 16384, 200},
                {256, 0, 2, 16384, 200}, {16, 0, 0, 16384, 0},
                {256, 0, 0, 16384, 0, 2}, {256, 0, 0, 16384, 0, 1}, {128, 0, 0, 16384, 0, 1}, {256, 0, 2, 16384, 0, 1},
                {256, 0, 2, 16384, 200, 1}};
  for (const Cfg& c : cfgs) {
    if (c.tw * 2 * c.chunk > 64 * 1024) continue;
    cudaMemset(out, 0, 148 * 4 * sizeof(long long));
    k<<<148, 512, 210 * 1024>>>(src, src_bytes, c.N, n_mma, c.tw, c.chunk, c.sts, c.same, out, c.fill);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    long long h[148 * 4];
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
    double cyc = 0, tma = 0, sts = 0;
    for (int i = 0; i < 148; ++i) { cyc += h[i * 4]; tma += h[i * 4 + 1]; sts += h[i * 4 + 2]; }
    cyc /= 148;
    printf("%3d %4d %9d %5d %3d %s | %7.1f  %10.0f  %11.1f  %11.1f\n", c.N, c.same, c.tw, c.chunk, c.sts,
           c.fill == 1 ? "random" : c.fill == 2 ? "zeros " : "uninit", cyc / n_mma,
           2.0 * 128 * c.N * 16 * n_mma / cyc, tma / 148 / cyc, sts / 148 / cyc);
  }
  return 0;
}
