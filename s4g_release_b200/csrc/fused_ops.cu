// Feature-propagation front end of the fused path (channel-last bf16 features):
//   * three_nn_weights: the reference's PointSearch (interpolate_kernel.cu:33-81) fused with the
//     inverse-squared-distance weights of FeatureInterpolator.forward (pointnet2_utils/modules.py:115-120):
//     w_k = (1/max(d2_k,1e-10)) / sum_k(1/max(d2_k,1e-10)), int32 indices;
//   * interp_concat: InterpolateForward (interpolate_kernel.cu:139-181) + the channel concat of
//     modules.py:124-127 ("interpolated first, then the dense skip feature") written once, as the bf16
//     row matrix the MLP chain consumes.
#include <cuda_bf16.h>

#include "common.cuh"
#include "grid.cuh"

namespace s4g {

constexpr int kNnwThreads = 256;
constexpr int kNnwTile = 2048;

__global__ void __launch_bounds__(kNnwThreads)
three_nn_weights_kernel(const float* __restrict__ query, const float* __restrict__ key, int Nq, int Nk,
                        int* __restrict__ index, float* __restrict__ weight) {
  __shared__ float4 s_key[kNnwTile];
  const int b = blockIdx.y;
  const int i = blockIdx.x * kNnwThreads + threadIdx.x;
  const bool valid = i < Nq;
  const float* Q = query + (size_t)b * 3 * Nq;
  const float* KX = key + (size_t)b * 3 * Nk;
  const float* KY = KX + Nk;
  const float* KZ = KY + Nk;
  const int iq = valid ? i : Nq - 1;
  const float x1 = Q[iq], y1 = Q[Nq + iq], z1 = Q[2 * Nq + iq];
  const float inf = __int_as_float(0x7f800000);
  float d0 = inf, d1 = inf, d2 = inf;
  int i0 = 0, i1 = 0, i2 = 0;
  for (int base = 0; base < Nk; base += kNnwTile) {
    const int n = min(kNnwTile, Nk - base);
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += kNnwThreads)
      s_key[k] = make_float4(__ldg(KX + base + k), __ldg(KY + base + k), __ldg(KZ + base + k), 0.f);
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < n; ++k) {
      const float4 p = s_key[k];
      const float d = sqdist(__fsub_rn(x1, p.x), __fsub_rn(y1, p.y), __fsub_rn(z1, p.z));
      if (d < d2) {
        const int j = base + k;
        if (d < d1) {
          d2 = d1; i2 = i1;
          if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
          else { d1 = d; i1 = j; }
        } else { d2 = d; i2 = j; }
      }
    }
  }
  if (valid) {
    const float v0 = __fdiv_rn(1.0f, fmaxf(d0, 1e-10f));
    const float v1 = __fdiv_rn(1.0f, fmaxf(d1, 1e-10f));
    const float v2 = __fdiv_rn(1.0f, fmaxf(d2, 1e-10f));
    const float norm = __fadd_rn(__fadd_rn(v0, v1), v2);
    int* oi = index + ((size_t)b * Nq + i) * 3;
    float* ow = weight + ((size_t)b * Nq + i) * 3;
    oi[0] = i0; oi[1] = i1; oi[2] = i2;
    ow[0] = __fdiv_rn(v0, norm); ow[1] = __fdiv_rn(v1, norm); ow[2] = __fdiv_rn(v2, norm);
  }
}

// One warp per kInterpRows consecutive query rows of one cloud (blockIdx.y = cloud: no 64-bit divisions); the 3 x 4
// neighbour indices / weights of the rows are one coalesced load by 12 lanes and travel by shuffle; per 16-byte piece
// (8 bf16 channels) the 3 x 4 gathered loads of all rows are issued before any arithmetic, so 12 requests per lane are
// in flight.  (r1 kernel: one row per warp, index arithmetic in 64 bit per lane — 24.7 % of the HBM peak with 61 % of the
// issue slots busy at the finest level, profiles/r02/ncu_geometry.txt.)
constexpr int kInterpRows = 4;

__device__ __forceinline__ uint32_t interp_pair(uint32_t a, uint32_t b, uint32_t c, float w0, float w1, float w2, bool relu) {
  // bf16 -> fp32 is a 16-bit shift; fma(in2,w2, fma(in1,w1, in0*w0)) as interpolate_kernel.cu:167-174
  const float x = __fmaf_rn(__uint_as_float(c << 16), w2, __fmaf_rn(__uint_as_float(b << 16), w1, __fmul_rn(__uint_as_float(a << 16), w0)));
  const float y = __fmaf_rn(__uint_as_float(c & 0xffff0000u), w2,
                            __fmaf_rn(__uint_as_float(b & 0xffff0000u), w1, __fmul_rn(__uint_as_float(a & 0xffff0000u), w0)));
  __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  uint32_t r = *reinterpret_cast<uint32_t*>(&h);
  if (relu) asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(0u));
  return r;
}

__global__ void __launch_bounds__(256)
interp_concat_kernel(const __nv_bfloat16* __restrict__ sparse, const int* __restrict__ index,
                     const float* __restrict__ weight, const __nv_bfloat16* __restrict__ dense, int Nk, int Nq,
                     int C2, int C1, __nv_bfloat16* __restrict__ out, int relu) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kInterpRows;
  if (row0 >= Nq) return;
  const int n_rows = min(kInterpRows, Nq - row0);
  const size_t q0 = (size_t)b * Nq + row0;  // first (batch-global) query row of this warp
  int my_i = 0;
  float my_w = 0.f;
  if (lane < 3 * n_rows) {
    my_i = __ldg(index + q0 * 3 + lane);
    my_w = __ldg(weight + q0 * 3 + lane);
  }
  const uint4* src[kInterpRows][3];
  float w[kInterpRows][3];
#pragma unroll
  for (int r = 0; r < kInterpRows; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int j = __shfl_sync(0xffffffffu, my_i, 3 * r + k);  // rows past n_rows read index 0 (valid) and are not stored
      w[r][k] = __shfl_sync(0xffffffffu, my_w, 3 * r + k);
      src[r][k] = reinterpret_cast<const uint4*>(sparse + ((size_t)b * Nk + j) * C2);
    }
  const int width = C2 + C1;
  for (int c = lane; c < C2 / 8; c += 32) {
    uint4 v[kInterpRows][3];
#pragma unroll
    for (int r = 0; r < kInterpRows; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) v[r][k] = __ldg(src[r][k] + c);
#pragma unroll
    for (int r = 0; r < kInterpRows; ++r) {
      if (r < n_rows) {
        uint4 o;
        o.x = interp_pair(v[r][0].x, v[r][1].x, v[r][2].x, w[r][0], w[r][1], w[r][2], relu);
        o.y = interp_pair(v[r][0].y, v[r][1].y, v[r][2].y, w[r][0], w[r][1], w[r][2], relu);
        o.z = interp_pair(v[r][0].z, v[r][1].z, v[r][2].z, w[r][0], w[r][1], w[r][2], relu);
        o.w = interp_pair(v[r][0].w, v[r][1].w, v[r][2].w, w[r][0], w[r][1], w[r][2], relu);
        reinterpret_cast<uint4*>(out + (q0 + r) * width)[c] = o;
      }
    }
  }
  if (C1 > 0) {
    for (int c = lane; c < C1 / 8; c += 32) {
      uint4 d[kInterpRows];
#pragma unroll
      for (int r = 0; r < kInterpRows; ++r)
        if (r < n_rows) d[r] = __ldg(reinterpret_cast<const uint4*>(dense + (q0 + r) * C1) + c);
#pragma unroll
      for (int r = 0; r < kInterpRows; ++r)
        if (r < n_rows) reinterpret_cast<uint4*>(out + (q0 + r) * width + C2)[c] = d[r];
    }
  }
}

// fp32 channel-first (B,C,N) -> bf16 channel-last [B*N][C] and back (interface <-> fused-path layout)
__global__ void __launch_bounds__(256)
gather_xyz_kernel(const float* __restrict__ xyz, const int* __restrict__ index, int N, int M, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int j = index[(size_t)b * M + m];
  const float* X = xyz + (size_t)b * 3 * N;
  float* O = out + (size_t)b * 3 * M;
  O[m] = __ldg(X + j);
  O[M + m] = __ldg(X + N + j);
  O[2 * M + m] = __ldg(X + 2 * N + j);
}

}  // namespace s4g

extern "C" int s4g_three_nn_weights_f32_i32(const float* query, const float* key, int B, int Nq, int Nk, int* index,
                                            float* weight, void* stream) {
  S4G_CHECK_ARG(query && key && index && weight, "three_nn_weights: null pointer");
  S4G_CHECK_ARG(Nk >= 3, "three_nn_weights: num_key < 3");
  S4G_CHECK_ARG(B >= 0 && B <= 65535 && Nq > 0, "three_nn_weights: bad shape");
  if (B == 0) return S4G_OK;
  if (Nk >= s4g::kGridKnnMinKeys)
    return s4g::three_nn_grid<1>(query, key, B, Nq, Nk, index, weight, (cudaStream_t)stream);
  dim3 grid((Nq + s4g::kNnwThreads - 1) / s4g::kNnwThreads, B);
  s4g::three_nn_weights_kernel<<<grid, s4g::kNnwThreads, 0, (cudaStream_t)stream>>>(query, key, Nq, Nk, index, weight);
  S4G_LAUNCH_CHECK("three_nn_weights");
  return S4G_OK;
}

extern "C" int s4g_interp_concat_act_bf16(const void* sparse, const int* index, const float* weight, const void* dense,
                                          int B, int Nk, int Nq, int C2, int C1, int relu, void* out, void* stream) {
  S4G_CHECK_ARG(sparse && index && weight && out, "interp_concat: null pointer");
  S4G_CHECK_ARG(C2 > 0 && C2 % 8 == 0 && C1 >= 0 && C1 % 8 == 0, "interp_concat: channel counts must be multiples of 8");
  S4G_CHECK_ARG(C1 == 0 || dense != nullptr, "interp_concat: dense feature missing");
  S4G_CHECK_ARG(B >= 0 && B <= 65535 && Nq > 0 && Nk > 0, "interp_concat: bad shape");
  S4G_CHECK_ARG(((uintptr_t)sparse & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)dense & 15) == 0,
                "interp_concat: feature rows must be 16-byte aligned");
  if (B == 0) return S4G_OK;
  const dim3 grid((unsigned)((Nq + 8 * s4g::kInterpRows - 1) / (8 * s4g::kInterpRows)), (unsigned)B);
  s4g::interp_concat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(sparse), index, weight, reinterpret_cast<const __nv_bfloat16*>(dense), Nk,
      Nq, C2, C1, reinterpret_cast<__nv_bfloat16*>(out), relu);
  S4G_LAUNCH_CHECK("interp_concat");
  return S4G_OK;
}

extern "C" int s4g_interp_concat_bf16(const void* sparse, const int* index, const float* weight, const void* dense,
                                      int B, int Nk, int Nq, int C2, int C1, void* out, void* stream) {
  return s4g_interp_concat_act_bf16(sparse, index, weight, dense, B, Nk, Nq, C2, C1, 0, out, stream);
}

extern "C" int s4g_gather_xyz_f32_i32(const float* xyz, const int* index, int B, int N, int M, float* out,
                                      void* stream) {
  S4G_CHECK_ARG(xyz && index && out, "gather_xyz: null pointer");
  S4G_CHECK_ARG(B >= 0 && B <= 65535 && N > 0 && M > 0, "gather_xyz: bad shape");
  if (B == 0) return S4G_OK;
  dim3 grid((M + 255) / 256, B);
  s4g::gather_xyz_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xyz, index, N, M, out);
  S4G_LAUNCH_CHECK("gather_xyz");
  return S4G_OK;
}
