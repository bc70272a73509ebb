// Ball query for sm_100a — replaces BallQuery / BallQueryKernel
// (reference: pointnet2_utils/csrc/ball_query_kernel.cu:33-76,89-133).
//
// The reference walks every centroid serially in one thread.  Here one WARP owns kCentroids
// centroids and sweeps the cloud in index order, 32 consecutive points per step (one coalesced
// 128-byte line per coordinate plane, read straight from the channel-first input — no transposed
// copy), four steps in flight.  Hits are compacted in order with a ballot + prefix popcount, so the
// "first K in index order" rule and the strict d2 < r*r test are preserved exactly; a warp stops as
// soon as all of its centroids are full.  Every output slot is written exactly once (padding with
// the first neighbour, zeros when there is none), so the outputs need no memset.
#include "common.cuh"
#include "grid.cuh"

namespace s4g {

constexpr int kBqWarps = 8;
constexpr int kBqCentroids = 4;  // centroids per warp (register-resident), reuses each point load
constexpr int kBqUnroll = 4;     // 32-point groups per loop trip

template <typename IndexT>
__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(const float* __restrict__ points, const float* __restrict__ centroids, int N, int M, float r2,
                  int K, IndexT* __restrict__ index, IndexT* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int m0 = (blockIdx.x * kBqWarps + warp) * kBqCentroids;
  if (m0 >= M) return;
  const float* X = points + (size_t)b * 3 * N;
  const float* Y = X + N;
  const float* Z = Y + N;
  const float* CX = centroids + (size_t)b * 3 * M;

  float cx[kBqCentroids], cy[kBqCentroids], cz[kBqCentroids];
  int cnt[kBqCentroids], first[kBqCentroids];
  bool done[kBqCentroids];
#pragma unroll
  for (int c = 0; c < kBqCentroids; ++c) {
    const int m = min(m0 + c, M - 1);
    cx[c] = CX[m];
    cy[c] = CX[M + m];
    cz[c] = CX[2 * M + m];
    cnt[c] = 0;
    first[c] = 0;
    done[c] = (m0 + c >= M);
  }
  const unsigned lt = (1u << lane) - 1u;

  for (int base = 0; base < N; base += 32 * kBqUnroll) {
    float x[kBqUnroll], y[kBqUnroll], z[kBqUnroll];
#pragma unroll
    for (int u = 0; u < kBqUnroll; ++u) {
      const int j = base + 32 * u + lane;
      const bool v = j < N;
      x[u] = v ? __ldg(X + j) : 0.f;
      y[u] = v ? __ldg(Y + j) : 0.f;
      z[u] = v ? __ldg(Z + j) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kBqUnroll; ++u) {
      const int j = base + 32 * u + lane;
      const bool v = j < N;
#pragma unroll
      for (int c = 0; c < kBqCentroids; ++c) {
        if (done[c]) continue;  // warp-uniform
        const float d = sqdist(__fsub_rn(x[u], cx[c]), __fsub_rn(y[u], cy[c]), __fsub_rn(z[u], cz[c]));
        const bool hit = v && (d < r2);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
          const int pos = cnt[c] + __popc(bal & lt);
          if (hit && pos < K) index[((size_t)b * M + m0 + c) * K + pos] = (IndexT)j;
          if (cnt[c] == 0) first[c] = base + 32 * u + __ffs(bal) - 1;
          cnt[c] += __popc(bal);
          if (cnt[c] >= K) { cnt[c] = K; done[c] = true; }
        }
      }
    }
    bool all = true;
#pragma unroll
    for (int c = 0; c < kBqCentroids; ++c) all = all && done[c];
    if (all) break;
  }
#pragma unroll
  for (int c = 0; c < kBqCentroids; ++c) {
    if (m0 + c >= M) continue;
    IndexT* out = index + ((size_t)b * M + m0 + c) * K;
    const IndexT fill = (IndexT)first[c];  // 0 when there was no hit: the reference's zero-init
    for (int k = cnt[c] + lane; k < K; k += 32) out[k] = fill;
    if (lane == 0 && count != nullptr) count[(size_t)b * M + m0 + c] = (IndexT)cnt[c];
  }
}

template <typename IndexT>
static int ball_query_entry(const float* points, const float* centroids, int B, int N, int M, float radius, int K,
                            IndexT* index, IndexT* count, cudaStream_t stream) {
  S4G_CHECK_ARG(points && centroids && index, "ball_query: null pointer");
  S4G_CHECK_ARG(B >= 0 && N > 0 && M > 0 && K > 0, "ball_query: bad shape B=%d N=%d M=%d K=%d", B, N, M, K);
  S4G_CHECK_ARG(B <= 65535, "ball_query: batch too large for one launch");
  if (B == 0) return S4G_OK;
  if (N >= kGridBallMinPoints && K <= kGridBallMaxK && radius > 0.f)  // output-sensitive exact path (grid.cu)
    return ball_query_grid<IndexT>(points, centroids, B, N, M, radius, K, index, count, stream);
  const float r2 = radius * radius;  // fp32 product, ball_query_kernel.cu:48
  const int per_cta = kBqWarps * kBqCentroids;
  dim3 grid((M + per_cta - 1) / per_cta, B);
  ball_query_kernel<IndexT><<<grid, kBqWarps * 32, 0, stream>>>(points, centroids, N, M, r2, K, index, count);
  S4G_LAUNCH_CHECK("ball_query");
  return S4G_OK;
}

}  // namespace s4g

extern "C" int s4g_ball_query_f32(const float* points, const float* centroids, int B, int N, int M, float radius,
                                  int K, int64_t* index, int64_t* count, void* stream) {
  return s4g::ball_query_entry<int64_t>(points, centroids, B, N, M, radius, K, index, count, (cudaStream_t)stream);
}

extern "C" int s4g_ball_query_f32_i32(const float* points, const float* centroids, int B, int N, int M, float radius,
                                      int K, int32_t* index, int32_t* count, void* stream) {
  return s4g::ball_query_entry<int32_t>(points, centroids, B, N, M, radius, K, index, count, (cudaStream_t)stream);
}
