// Micro-benchmark (measurement tool, not product code): cost of the hidden-layer epilogue of the MLP chain —
// TMEM accumulator block (128 lanes x 128 fp32 columns) -> + shift, ReLU, bf16 -> 32 KB activation block in shared
// memory ([16 pieces][128 rows][16 B]) -> fence.proxy.async -> per-warp mbarrier arrive — for several codings:
//   variant 0: 8 warps per block, 64 columns per warp as two serial 32-column slabs, shifts by LDG, FADD + F2FP + max
//   variant 1: 16 warps per block, 32 columns per warp (one slab), same arithmetic
//   variant 2: variant 1 with add.f32x2 and cvt.rn.relu.bf16x2.f32 (two instructions per column pair)
//   variant 3: variant 2 with the shifts read from shared memory (LDS.128)
//   variant 4: 8 warps per block, 64 columns per warp, both TMEM loads issued up front, arithmetic of variant 2
//   variant 5: variant 3 without the shift add (shift folded into the GEMM): cvt.relu only
// `active` = warps running the loop (8 or 16); with 16 active warps and an 8-warp variant two blocks are drained
// concurrently (as the two epilogue groups do).  Output: cycles per 128 x 128 block.
// Each variant is also run against a concurrent tensor-core load (operands in shared memory) and bulk-copy load.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_bw epi_bw.cu && ./epi_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(taddr, r)                                                                                       \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                      \
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                       \
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"       \
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), \
                 "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]),        \
                 "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]),      \
                 "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]),      \
                 "=f"(r[29]), "=f"(r[30]), "=f"(r[31])                                                          \
               : "r"(taddr))

__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_old(float a, float b, float ba, float bb) {
  uint32_t r;
  asm("{\n\t.reg .f32 x, y;\n\tadd.f32 x, %1, %3;\n\tadd.f32 y, %2, %4;\n\tcvt.rn.bf16x2.f32 %0, y, x;\n\t"
      "max.bf16x2 %0, %0, %5;\n\t}"
      : "=r"(r) : "f"(a), "f"(b), "f"(ba), "f"(bb), "r"(0u));
  return r;
}
__device__ __forceinline__ uint32_t pack_new(float a, float b, float ba, float bb) {
  uint32_t r;
  asm("{\n\t.reg .b64 u, v, w;\n\t.reg .f32 x, y;\n\tmov.b64 u, {%1, %2};\n\tmov.b64 v, {%3, %4};\n\t"
      "add.rn.f32x2 w, u, v;\n\tmov.b64 {x, y}, w;\n\tcvt.rn.relu.bf16x2.f32 %0, y, x;\n\t}"
      : "=r"(r) : "f"(a), "f"(b), "f"(ba), "f"(bb));
  return r;
}
__device__ __forceinline__ uint32_t pack_nobias(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));
  return r;
}

template <int V>
__device__ __forceinline__ void slab(uint32_t taddr, const float* bias_g, const float* bias_s, uint8_t* dst, bool preissued,
                                     float* v) {
  float4 bv[8];
  if (V <= 1) {
#pragma unroll
    for (int g = 0; g < 8; ++g) bv[g] = __ldg(reinterpret_cast<const float4*>(bias_g) + g);
  }
  if (!preissued) LD32(taddr, v);
  if (V == 2 || V == 4) {
#pragma unroll
    for (int g = 0; g < 8; ++g) bv[g] = __ldg(reinterpret_cast<const float4*>(bias_g) + g);
  }
  if (V == 3) {
#pragma unroll
    for (int g = 0; g < 8; ++g) bv[g] = reinterpret_cast<const float4*>(bias_s)[g];
  }
  if (!preissued) tmem_wait();
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 pk;
    const float* x = v + 8 * g;
    if (V <= 1) {
      pk = make_uint4(pack_old(x[0], x[1], bv[2 * g].x, bv[2 * g].y), pack_old(x[2], x[3], bv[2 * g].z, bv[2 * g].w),
                      pack_old(x[4], x[5], bv[2 * g + 1].x, bv[2 * g + 1].y),
                      pack_old(x[6], x[7], bv[2 * g + 1].z, bv[2 * g + 1].w));
    } else if (V == 5) {
      pk = make_uint4(pack_nobias(x[0], x[1]), pack_nobias(x[2], x[3]), pack_nobias(x[4], x[5]), pack_nobias(x[6], x[7]));
    } else {
      pk = make_uint4(pack_new(x[0], x[1], bv[2 * g].x, bv[2 * g].y), pack_new(x[2], x[3], bv[2 * g].z, bv[2 * g].w),
                      pack_new(x[4], x[5], bv[2 * g + 1].x, bv[2 * g + 1].y),
                      pack_new(x[6], x[7], bv[2 * g + 1].z, bv[2 * g + 1].w));
    }
    *reinterpret_cast<uint4*>(dst + (size_t)g * (128 * 16)) = pk;
  }
}

__device__ __forceinline__ bool try_wait(uint64_t* bar, unsigned par) {
  unsigned ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(s32(bar)), "r"(par) : "memory");
  return ok;
}
// load bit 0: warp 18 issues back-to-back N = 256 tcgen05.mma (operands in shared memory) while the epilogue runs;
// load bit 1: warp 19 streams 32 KB bulk copies (L2 -> shared memory) as the weight producer does
template <int V>
__global__ void __launch_bounds__(640, 1) k(int active, int iters, const float* bias, long long* out, int load,
                                            const uint8_t* wsrc) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tslot;
  __shared__ uint64_t bar;
  __shared__ uint64_t lbar[2];
  __shared__ volatile int stop;
  __shared__ __align__(16) float sbias[512];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1024;" ::"r"(s32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&lbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&lbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  if (threadIdx.x < 512) sbias[threadIdx.x] = bias[threadIdx.x];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int qd = warp & 3;
  const int row = qd * 32 + lane;
  constexpr bool kWide = (V == 0 || V == 4);          // 8 warps per block, 64 columns per warp
  const int grp = kWide ? (warp >> 3) : 0;            // block drained by this warp
  const int c0 = kWide ? 64 * ((warp >> 2) & 1) : 32 * (warp >> 2);
  const long long t0 = clock64();
  if (warp < active && warp < 16) {
    for (int i = 0; i < iters; ++i) {
      const int blk = kWide ? grp : (i & 1);  // TMEM blocks 0, 1 (2, 3 belong to the MMA load); slots 0, 1
      const uint32_t taddr = tslot + ((uint32_t)(qd * 32) << 16) + (uint32_t)(blk * 128 + c0);
      uint8_t* dst = smem + (size_t)(blk & 3) * 32768 + ((size_t)(c0 >> 3) * 128 + row) * 16;
      const float* bg = bias + ((i * 128) & 255) + c0;
      const float* bs = sbias + ((i * 128) & 255) + c0;
      if (V == 4) {
        float a[32], b[32];
        LD32(taddr, a);
        LD32(taddr + 32, b);
        tmem_wait();
        slab<V>(taddr, bg, bs, dst, true, a);
        slab<V>(taddr + 32, bg + 32, bs + 32, dst + 4 * 128 * 16, true, b);
      } else {
        float a[32];
        slab<V>(taddr, bg, bs, dst, false, a);
        if (kWide) slab<V>(taddr + 32, bg + 32, bs + 32, dst + 4 * 128 * 16, false, a);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar)) : "memory");
    }
  }
  const long long t1 = clock64();
  if (lane == 0 && warp < active && warp < 16) out[warp] = t1 - t0;
  if (warp < 16) {
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (threadIdx.x == 0) stop = 1;
  } else if (warp == 18 && (load & 1)) {
    // A = slots 2, 3 (64 KB, K = 256), B = 64 KB behind the slots: [K/8][256 rows][16 B]
    const uint64_t hi = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    const uint32_t a0 = s32(smem + 2 * 32768) >> 4, b0 = s32(smem + 4 * 32768) >> 4;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    unsigned par = 0;
    long long n = 0;
    const long long l0 = clock64();
    while (!stop && clock64() - l0 < 2000000000LL) {
      if (lane == 0) {
        uint32_t a_lo = a0 | (128u << 16), w_lo = b0 | (256u << 16);
#pragma unroll 1
        for (int q = 0; q < 16; ++q) {
          asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tslot + 256u), "l"(hi | a_lo), "l"(hi | w_lo), "r"(idesc), "r"(1u) : "memory");
          a_lo += 256u;
          w_lo = (q == 7) ? (b0 | (256u << 16)) : w_lo + 512u;  // B: 8 K-steps of 8 KB, read twice
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&lbar[0])) : "memory");
        { const long long w0 = clock64(); while (!try_wait(&lbar[0], par)) { if (clock64() - w0 > 100000000LL) { printf("mma wait timeout n=%lld\n", n); break; } } }
      }
      par ^= 1;
      n += 16;
      __syncwarp();
    }
    if (lane == 0) out[16] = n;
  } else if (warp == 19 && (load & 2)) {
    unsigned par = 0;
    long long n = 0;
    const long long l0 = clock64();
    while (!stop && clock64() - l0 < 2000000000LL) {
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&lbar[1])), "r"(32768u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + 6 * 32768)), "l"(wsrc + (n & 7) * 32768), "r"(32768u), "r"(s32(&lbar[1])) : "memory");
        { const long long w0 = clock64(); while (!try_wait(&lbar[1], par)) { if (clock64() - w0 > 100000000LL) { printf("tma wait timeout n=%lld\n", n); break; } } }
      }
      par ^= 1;
      ++n;
      __syncwarp();
    }
    if (lane == 0) out[17] = n;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tslot), "r"(512u) : "memory");
}

template <int V>
static void run(int active, const float* bias, long long* d_out, int load = 0, const uint8_t* wsrc = nullptr) {
  const int iters = 400;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * 32768);
  cudaMemset(d_out, 0, 18 * sizeof(long long));
  k<V><<<148, 640, 7 * 32768>>>(active, 20, bias, d_out, load, wsrc);
  k<V><<<148, 640, 7 * 32768>>>(active, iters, bias, d_out, load, wsrc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d: %s\n", V, cudaGetErrorString(e)); return; }
  long long h[18];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int w = 0; w < active && w < 16; ++w) mx = h[w] > mx ? h[w] : mx;
  // blocks drained during the loop: wide variants drain one block per group per iteration
  const bool wide = (V == 0 || V == 4);
  const double blocks = wide ? (double)iters * (active > 8 ? 2 : 1) : (double)iters;
  printf("%7d %6d %4d | %10.1f %12.1f", V, active, load, (double)mx / iters, (double)mx / blocks);
  if (load & 1) printf("   mma cyc/instr %.1f", (double)mx / (double)h[16]);
  if (load & 2) printf("   cyc/32KB copy %.1f", (double)mx / (double)h[17]);
  printf("\n");
  fflush(stdout);
}

int main(int argc, char** argv) {
  float* bias;
  long long* d_out;
  cudaMalloc(&bias, 4096);
  cudaMemset(bias, 0, 4096);
  cudaMalloc(&d_out, 18 * sizeof(long long));
  uint8_t* wsrc;
  cudaMalloc(&wsrc, 8 * 32768);
  cudaMemset(wsrc, 0, 8 * 32768);
  printf("variant active load | cyc/iter/warp  cyc/128x128 block\n");
  run<0>(8, bias, d_out);
  run<0>(16, bias, d_out);
  run<4>(8, bias, d_out);
  run<4>(16, bias, d_out);
  run<1>(16, bias, d_out);
  run<2>(16, bias, d_out);
  run<3>(16, bias, d_out);
  run<5>(16, bias, d_out);
  // concurrent tensor-core / bulk-copy load: EXPERIMENTAL (hangs on the B200 it was tried on) — ./epi_bw load
  for (int load = 1; load <= 3 && argc > 1; ++load) {
    run<0>(8, bias, d_out, load, wsrc);
    run<0>(16, bias, d_out, load, wsrc);
    run<1>(16, bias, d_out, load, wsrc);
    run<3>(16, bias, d_out, load, wsrc);
    run<5>(16, bias, d_out, load, wsrc);
  }
  printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
