// Device-side cloud pre-processing — the step right before the forward.
// Reference: GraspDetector._pre_processing / sample_single_cloud (grasp_detector.py:82-105) over
// CloudPreProcessor (cloud_processor/cloud_processor.py:31-42) and transform_numpy_points (utils/math_utils.py:20-24).
//
// What the reference OBSERVABLY does: `voxelize()` and `remove_outliers()` call open3d functions that RETURN a new
// cloud and drop the result, so both are no-ops; the cloud is then multiplied by _REAL2TRAIN (swap x / y, negate z)
// and sub-sampled with np.random.choice to NUM_INPUT points.  s4g_cloud_transform_select_f32 is exactly that
// (indices supplied by the caller: numpy's generator cannot be reproduced on the device), for a batch of clouds.
// The two filters the code INTENDS are provided as well, with our own (stated) definitions, off by default in the
// host mirror:
//   * voxel down-sample: one output point per occupied voxel of side `voxel`, the MEAN of its points, voxels in
//     ascending (z, y, x) cell order (open3d's output order is its hash map's; the set of points is the same);
//   * radius outlier mask: keep a point when MORE than nb_points points (itself included) lie within `radius`
//     (open3d::geometry::PointCloud::RemoveRadiusOutliers) — computed by the ball-query kernel with K = nb_points + 1
//     (csrc/ball_query.cu, csrc/grid.cu), so there is no kernel for it here.
#include "common.cuh"

namespace s4g {

struct Mat34 { float m[12]; };

// out[b, :, i] = (T [cloud[b, :, index[b, i]]; 1])[:3]
__global__ void __launch_bounds__(256)
transform_select_kernel(const float* __restrict__ cloud, int n, const int64_t* __restrict__ index, int m, const Mat34 t,
                        float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float* C = cloud + (size_t)b * 3 * n;
  const int64_t j = index ? index[(size_t)b * m + i] : i;
  const float x = C[j], y = C[n + j], z = C[2 * (size_t)n + j];
  float* O = out + (size_t)b * 3 * m;
#pragma unroll
  for (int r = 0; r < 3; ++r)  // numpy's matmul accumulates left to right: ((a x + b y) + c z) + d
    O[(size_t)r * m + i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.m[4 * r], x), __fmul_rn(t.m[4 * r + 1], y)),
                                                 __fmul_rn(t.m[4 * r + 2], z)), t.m[4 * r + 3]);
}

// voxel keys: cell = floor((p - origin) / voxel) per axis, key = (cz * ny + cy) * nx + cx  (int64)
__global__ void __launch_bounds__(256)
voxel_key_kernel(const float* __restrict__ cloud, int n, float ox, float oy, float oz, float inv_voxel, int nx, int ny,
                 int nz, int64_t* __restrict__ key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int cx = (int)floorf((cloud[i] - ox) * inv_voxel), cy = (int)floorf((cloud[n + i] - oy) * inv_voxel),
      cz = (int)floorf((cloud[2 * (size_t)n + i] - oz) * inv_voxel);
  cx = min(max(cx, 0), nx - 1); cy = min(max(cy, 0), ny - 1); cz = min(max(cz, 0), nz - 1);
  key[i] = ((int64_t)cz * ny + cy) * nx + cx;
}

// points sorted by key (order[i] = source index): one thread per segment start writes the segment's mean (fp64 sum in
// ascending source-index order inside a voxel when the sort is stable) and counts the voxels
__global__ void __launch_bounds__(256)
voxel_mean_kernel(const float* __restrict__ cloud, int n, const int64_t* __restrict__ sorted_key,
                  const int64_t* __restrict__ order, const int* __restrict__ seg_rank, float* __restrict__ out, int cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i > 0 && sorted_key[i] == sorted_key[i - 1]) return;
  double sx = 0.0, sy = 0.0, sz = 0.0;
  int cnt = 0;
  for (int k = i; k < n && sorted_key[k] == sorted_key[i]; ++k) {
    const int64_t j = order[k];
    sx += cloud[j]; sy += cloud[n + j]; sz += cloud[2 * (size_t)n + j];
    ++cnt;
  }
  const int o = seg_rank[i];
  if (o < cap) {
    out[o] = (float)(sx / cnt); out[cap + o] = (float)(sy / cnt); out[2 * (size_t)cap + o] = (float)(sz / cnt);
  }
}

}  // namespace s4g

// cloud (B,3,n) fp32, index (B,m) int64 (NULL = identity, then m must equal n), mat44: 16 floats, HOST memory,
// row-major (only the top three rows are used) -> out (B,3,m).
extern "C" int s4g_cloud_transform_select_f32(const float* cloud, int B, int n, const int64_t* index, int m,
                                              const float* mat44, float* out, void* stream) {
  S4G_CHECK_ARG(cloud && mat44 && out, "cloud_transform_select: null pointer");
  S4G_CHECK_ARG(B >= 0 && n > 0 && m > 0, "cloud_transform_select: bad shape");
  S4G_CHECK_ARG(index != nullptr || m == n, "cloud_transform_select: identity selection needs m == n");
  S4G_CHECK_ARG(B <= 65535, "cloud_transform_select: batch too large for one launch");
  if (B == 0) return S4G_OK;
  s4g::Mat34 t;
  for (int i = 0; i < 12; ++i) t.m[i] = mat44[i];
  dim3 grid((m + 255) / 256, B);
  s4g::transform_select_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cloud, n, index, m, t, out);
  S4G_LAUNCH_CHECK("cloud_transform_select");
  return S4G_OK;
}

extern "C" int s4g_voxel_keys_f32(const float* cloud_3n, int n, const float* origin3, float voxel, const int* dims3,
                                  int64_t* key, void* stream) {
  S4G_CHECK_ARG(cloud_3n && origin3 && dims3 && key, "voxel_keys: null pointer");
  S4G_CHECK_ARG(n > 0 && voxel > 0.f && dims3[0] > 0 && dims3[1] > 0 && dims3[2] > 0, "voxel_keys: bad arguments");
  s4g::voxel_key_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cloud_3n, n, origin3[0], origin3[1], origin3[2],
                                                                        1.0f / voxel, dims3[0], dims3[1], dims3[2], key);
  S4G_LAUNCH_CHECK("voxel_keys");
  return S4G_OK;
}

// sorted_key / order: the keys sorted ascending (stable) and the source index of every sorted position;
// seg_rank[i]: number of segment starts before sorted position i (exclusive scan of the start flags).
extern "C" int s4g_voxel_means_f32(const float* cloud_3n, int n, const int64_t* sorted_key, const int64_t* order,
                                   const int* seg_rank, float* out_3cap, int cap, void* stream) {
  S4G_CHECK_ARG(cloud_3n && sorted_key && order && seg_rank && out_3cap, "voxel_means: null pointer");
  S4G_CHECK_ARG(n > 0 && cap > 0, "voxel_means: bad shape");
  s4g::voxel_mean_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(cloud_3n, n, sorted_key, order, seg_rank,
                                                                         out_3cap, cap);
  S4G_LAUNCH_CHECK("voxel_means");
  return S4G_OK;
}
