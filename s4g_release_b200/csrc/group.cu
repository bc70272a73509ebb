// gather_points / group_points forward+backward in the reference's interface layout
// (channel-first fp32, int64 indices).  Reference: pointnet2_utils/functions.py:10-25,
// csrc/grouping_kernel.cu:32-54 (forward = expand + at::gather), :57-96,106-152 (backward).
//
// One thread owns one (b, m, k) slot: it reads its int64 index ONCE and walks the C channel planes,
// so index traffic is 8 B per slot instead of 8·C, and the (B,C,M,K) side is fully coalesced.
// These are the drop-in / training-path ops; inference uses the fused gather inside the MLP kernel.
#include "common.cuh"

namespace s4g {

__global__ void __launch_bounds__(256)
group_forward_kernel(const float* __restrict__ input, const int64_t* __restrict__ index, int C, int N, int MK,
                     float* __restrict__ out) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= MK) return;
  const int64_t j = index[(size_t)b * MK + q];
  const float* in = input + (size_t)b * C * N + j;
  float* o = out + (size_t)b * C * MK + q;
#pragma unroll 4
  for (int c = 0; c < C; ++c) o[(size_t)c * MK] = __ldg(in + (size_t)c * N);
}

__global__ void __launch_bounds__(256)
group_backward_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ index, int C, int N, int MK,
                      float* __restrict__ grad_in) {
  const int b = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= MK) return;
  const int64_t j = index[(size_t)b * MK + q];
  float* gi = grad_in + (size_t)b * C * N + j;
  const float* go = grad_out + (size_t)b * C * MK + q;
#pragma unroll 4
  for (int c = 0; c < C; ++c) atomicAdd(gi + (size_t)c * N, __ldg(go + (size_t)c * MK));
}

}  // namespace s4g

extern "C" int s4g_group_points_forward_f32(const float* input, const int64_t* index, int B, int C, int N, int M,
                                            int K, float* out, void* stream) {
  S4G_CHECK_ARG(input && index && out, "group_points_forward: null pointer");
  S4G_CHECK_ARG(B >= 0 && C > 0 && N > 0 && M > 0 && K > 0, "group_points_forward: bad shape");
  S4G_CHECK_ARG(B <= 65535, "group_points_forward: batch too large for one launch");
  S4G_CHECK_ARG((long long)M * K < (1ll << 31), "group_points_forward: M*K too large");
  if (B == 0) return S4G_OK;
  const int MK = M * K;
  dim3 grid((MK + 255) / 256, B);
  s4g::group_forward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(input, index, C, N, MK, out);
  S4G_LAUNCH_CHECK("group_points_forward");
  return S4G_OK;
}

extern "C" int s4g_group_points_backward_f32(const float* grad_out, const int64_t* index, int B, int C, int N, int M,
                                             int K, float* grad_in, void* stream) {
  S4G_CHECK_ARG(grad_out && index && grad_in, "group_points_backward: null pointer");
  S4G_CHECK_ARG(B >= 0 && C > 0 && N > 0 && M > 0 && K > 0, "group_points_backward: bad shape");
  S4G_CHECK_ARG(B <= 65535, "group_points_backward: batch too large for one launch");
  S4G_CHECK_ARG((long long)M * K < (1ll << 31), "group_points_backward: M*K too large");
  if (B == 0) return S4G_OK;
  cudaStream_t s = (cudaStream_t)stream;
  S4G_CUDA(cudaMemsetAsync(grad_in, 0, sizeof(float) * (size_t)B * C * N, s));
  const int MK = M * K;
  dim3 grid((MK + 255) / 256, B);
  s4g::group_backward_kernel<<<grid, 256, 0, s>>>(grad_out, index, C, N, MK, grad_in);
  S4G_LAUNCH_CHECK("group_points_backward");
  return S4G_OK;
}

extern "C" int s4g_gather_points_f32(const float* points, const int64_t* index, int B, int C, int N, int M,
                                     float* out, void* stream) {
  // gather_points is group_points with K = 1
  S4G_CHECK_ARG(points && index && out, "gather_points: null pointer");
  return s4g_group_points_forward_f32(points, index, B, C, N, M, 1, out, stream);
}
