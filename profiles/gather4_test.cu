// Probe (measurement tool, not product code): does cp.async.bulk.tensor.2d ... tile::gather4 deliver four arbitrary rows
// of a [rows][C] bf16 table as four consecutive 128-byte rows of the SWIZZLE_128B K-major layout the MLP chain's
// TMA-input path consumes — and which tensor-map box ({64, 1} or {64, 4}) the instruction wants?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather4_test gather4_test.cu && ./gather4_test
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap map, const int* rows, int col, int dst_off, int n_g4, int bytes,
                  uint16_t* out, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xffff;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    for (int g = 0; g < n_g4; ++g)
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
              "r"(s32(smem + dst_off + g * 512)), "l"(&map), "r"(s32(&bar)), "r"(col), "r"(rows[4 * g]), "r"(rows[4 * g + 1]),
          "r"(rows[4 * g + 2]), "r"(rows[4 * g + 3])
          : "memory");
    unsigned ok = 0;
    long long spins = 0;
    while (!ok && spins < (1 << 22)) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(s32(&bar)) : "memory");
      ++spins;
    }
    *status = ok ? 1 : 0;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 16384 / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  const int R = 4096, C = 256;
  std::vector<uint16_t> h((size_t)R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) h[(size_t)r * C + c] = (uint16_t)((r * 7 + c) & 0xffff);  // raw 16-bit patterns
  uint16_t *d, *out;
  int *rows, *status;
  cudaMalloc(&d, h.size() * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 16384);
  cudaMalloc(&rows, 128 * 4);
  cudaMalloc(&status, 4);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn encode = (EncodeFn)fn;
  std::vector<int> hr(128);
  for (int i = 0; i < 128; ++i) hr[i] = (i * 1237 + 11) % R;
  cudaMemcpy(rows, hr.data(), 128 * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  for (int box_rows : {1, 4}) {
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {64, %d}: encode rc %d\n", box_rows, (int)rc);
    if (rc != CUDA_SUCCESS) continue;
    for (int variant = 0; variant < 3; ++variant) {
      // 0: one gather4 at offset 0;  1: one gather4 at offset 512 (rows 4..7 of the swizzle atom);  2: a whole 128-row half
      const int n_g4 = variant == 2 ? 32 : 1, dst_off = variant == 1 ? 512 : 0, col = 64;
      cudaMemset(status, 0xff, 4);
      k<<<1, 128, 16384>>>(map, rows, col, dst_off, n_g4, n_g4 * 512, out, status);
      const cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
      int st;
      std::vector<uint16_t> ho(8192);
      cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(ho.data(), out, 16384, cudaMemcpyDeviceToHost);
      int bad = 0, untouched = 0;
      for (int rr = 0; rr < 4 * n_g4; ++rr) {
        const int row_in_blk = dst_off / 128 + rr;
        for (int c = 0; c < 64; ++c) {
          const int chunk = (c >> 3) ^ (row_in_blk & 7);  // 16-byte chunk position after the 128-byte swizzle
          const uint16_t got = ho[row_in_blk * 64 + chunk * 8 + (c & 7)];
          const uint16_t want = h[(size_t)hr[rr] * C + col + c];
          if (got != want) ++bad;
          if (got == 0xffff) ++untouched;
        }
      }
      printf("  variant %d (%d gather4, dst +%d): barrier %s, %d mismatches (%d untouched) of %d\n", variant, n_g4, dst_off,
             st == 1 ? "completed" : "TIMED OUT", bad, untouched, 4 * n_g4 * 64);
    }
  }
  return 0;
}
