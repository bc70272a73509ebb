// gather_points / group_points forward+backward in the reference's interface layout
// (channel-first fp32, int64 indices).  Reference: pointnet2_utils/functions.py:10-25,
// csrc/grouping_kernel.cu:32-54 (forward = expand + at::gather), :57-96,106-152 (backward).
//
// One thread owns one (b, m, k) slot: it reads its int64 index ONCE and walks the C channel planes,
// so index traffic is 8 B per slot instead of 8·C, and the (B,C,M,K) side is fully coalesced.
// These are the drop-in / module-path ops; inference uses the fused TMA gather inside the MLP kernel.
//
// Forward, staged variant (round 2, feature tensors: >= 32 channels, planes <= 64 KB): the gather reads 4 bytes per
// element out of a 32-byte L2 sector, so the plain kernel is bound by L2 sector traffic (0.26-0.39 of the HBM peak at 64
// clouds, profiles/r01/micro_sweep.md).  A CTA stages up to 16 planes of one cloud with 1-D bulk copies
// (cp.async.bulk, one mbarrier transaction), then gathers from SHARED memory for a chunk of slots — the index is read
// once per slot for all staged planes and the only global traffic left is coalesced: planes in, indices in, output out.
#include "common.cuh"

namespace s4g {

__device__ __forceinline__ unsigned grp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

constexpr int kStageMaxPlanes = 16;
constexpr int kStageSmemBytes = 200 * 1024;

__global__ void __launch_bounds__(256)
group_forward_staged_kernel(const float* __restrict__ input, const int64_t* __restrict__ index, int C, int N, int MK,
                            int planes_per_cta, int slots_per_cta, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_plane[];  // [planes][N]
  __shared__ __align__(8) unsigned long long bar;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * planes_per_cta;
  const int np = min(planes_per_cta, C - c0);
  const unsigned plane_bytes = (unsigned)N * 4u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(grp_smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(grp_smem_u32(&bar)), "r"(plane_bytes * (unsigned)np)
                 : "memory");
    const float* src = input + ((size_t)b * C + c0) * N;
    for (int c = 0; c < np; ++c)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       grp_smem_u32(s_plane + (size_t)c * N)), "l"(src + (size_t)c * N), "r"(plane_bytes), "r"(grp_smem_u32(&bar))
                   : "memory");
  }
  __syncthreads();  // the barrier is initialised before anyone polls it
  unsigned ok = 0;
  while (!ok)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(grp_smem_u32(&bar)) : "memory");
  const int q0 = blockIdx.x * slots_per_cta;
  const int q1 = min(MK, q0 + slots_per_cta);
  const int64_t* idx = index + (size_t)b * MK;
  float* o = out + ((size_t)b * C + c0) * MK;
  for (int q = q0 + threadIdx.x; q < q1; q += 256) {
    const int j = (int)idx[q];
    for (int c = 0; c < np; ++c) o[(size_t)c * MK + q] = s_plane[(size_t)c * N + j];
  }
}

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

static bool g_group_staged = true;
// A/B switch (measurements): 0 = always the plain gather kernel.  Returns the previous setting.
extern "C" int s4g_group_points_set_staged(int on) {
  const int prev = g_group_staged ? 1 : 0;
  g_group_staged = on != 0;
  return prev;
}

extern "C" int s4g_group_points_forward_f32(const float* input, const int64_t* index, int B, int C, int N, int M,
                                            int K, float* out, void* stream) {
  S4G_CHECK_ARG(input && index && out, "group_points_forward: null pointer");
  S4G_CHECK_ARG(B >= 0 && C > 0 && N > 0 && M > 0 && K > 0, "group_points_forward: bad shape");
  S4G_CHECK_ARG(B <= 65535, "group_points_forward: batch too large for one launch");
  S4G_CHECK_ARG((long long)M * K < (1ll << 31), "group_points_forward: M*K too large");
  if (B == 0) return S4G_OK;
  const int MK = M * K;
  // staged variant: planes of N floats in shared memory (16-byte bulk copies need N % 4 == 0 and an aligned base)
  const size_t plane_bytes = (size_t)N * 4;
  // Measured (profiles/r02/group_ab.txt, 64 / 16 clouds): 256 channels x 5 120 points 0.97 -> 0.55 ms (0.34 -> 0.59 of the
  // HBM peak), 512 x 1 024 0.29 -> 0.12 ms; but for the 3 coordinate planes of a large cloud (100 KB each, one CTA per SM,
  // three gathers per staged index) the plain kernel is 2-3x FASTER — so: many channels, small planes only.
  if (g_group_staged && C >= 32 && plane_bytes <= 64 * 1024 && N % 4 == 0 && ((uintptr_t)input & 15) == 0) {
    int planes = (int)(s4g::kStageSmemBytes / plane_bytes);
    planes = planes > s4g::kStageMaxPlanes ? s4g::kStageMaxPlanes : planes;
    planes = planes > C ? C : planes;
    const int c_groups = (C + planes - 1) / planes;
    // slots per CTA: at least twice the staged plane length (the staging is then < 1/3 of the CTA's traffic), shrunk until
    // the launch has >= 2 CTAs per SM; too little work for that -> the plain kernel
    int slots = 2 * N > 4096 ? 2 * N : 4096;
    const long long want = 2LL * s4g::num_sms();
    while (slots > 2048 && (long long)((MK + slots - 1) / slots) * c_groups * B < want) slots >>= 1;
    const long long ctas = (long long)((MK + slots - 1) / slots) * c_groups * B;
    if (ctas >= s4g::num_sms() && c_groups <= 65535 && B <= 65535) {
      static bool attr_set[64] = {};
      if (s4g::first_use_on_device(attr_set)) {
        S4G_CUDA(cudaFuncSetAttribute(s4g::group_forward_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      s4g::kStageSmemBytes));
          }
      dim3 grid((MK + slots - 1) / slots, c_groups, B);
      s4g::group_forward_staged_kernel<<<grid, 256, (size_t)planes * plane_bytes, (cudaStream_t)stream>>>(
          input, index, C, N, MK, planes, slots, out);
      S4G_LAUNCH_CHECK("group_points_forward");
      return S4G_OK;
    }
  }
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
