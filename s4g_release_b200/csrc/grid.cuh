// Uniform-grid spatial index over one batch of clouds, used to make ball query and 3-NN search
// output-sensitive while keeping the reference's results bit-for-bit:
//   * ball query = "the K smallest indices among points with d2 < r2" (the reference's "first K hits in
//     index order", ball_query_kernel.cu:57-71) — every such point lies in the 3x3x3 cell block around the
//     centroid when the cell edge is >= r * 1.01;
//   * 3-NN = the 3 smallest (d2, index) pairs — found inside the block whenever the third distance is
//     smaller than the cell edge (else the query goes to an exact brute-force pass).
// The distance arithmetic is the reference's (common.cuh::sqdist) in both the grid and fallback paths.
//
// Build (per call, on the stream): bbox -> per-cloud GridDesc -> histogram of cell keys -> exclusive scan
// -> scatter of (x, y, z, index) as one float4 per point into cell order (order inside a cell is
// irrelevant: both queries order their own candidates).
#pragma once
#include "common.cuh"

namespace s4g {

constexpr int kGridCells = 32768;  // table stride: max cells per cloud

struct GridDesc {
  float ox, oy, oz;  // origin = bbox minimum
  float inv_s;       // 1 / cell edge
  float s;           // cell edge
  int dx, dy, dz;    // cells per axis (dx * dy * dz <= kGridCells)
};

struct Grid {
  GridDesc* desc;    // [B]
  int* start;        // [B * kGridCells + 1] exclusive prefix of the per-cell counts
  float4* sorted;    // [B * N] (x, y, z, __int_as_float(index)) in cell order
  void* arena;       // one cudaMallocAsync block holding all of the above + temporaries
  int B, N;
};

enum GridMode { GRID_BALL = 0, GRID_KNN = 1 };

// Builds the index over points (B,3,N).  mode GRID_BALL: cell edge >= radius * 1.01;
// GRID_KNN: cell edge ~ 2 * sqrt(max projected bbox area / N) (twice the mean spacing of a surface cloud).
int grid_build(const float* points, int B, int N, int mode, float radius, Grid* grid, cudaStream_t stream);
int grid_free(Grid* grid, cudaStream_t stream);

// exact grid-accelerated queries (grid.cu); same outputs as the linear-scan kernels
template <typename IndexT>
int ball_query_grid(const float* points, const float* centroids, int B, int N, int M, float radius, int K,
                    IndexT* index, IndexT* count, cudaStream_t stream);
// MODE 0: int64 index + squared distances; MODE 1: int32 index + normalised interpolation weights
template <int MODE>
int three_nn_grid(const float* query, const float* key, int B, int Nq, int Nk, void* index, float* out,
                  cudaStream_t stream);
constexpr int kGridBallMinPoints = 8192;  // below this the linear scan is as fast
constexpr int kGridBallMaxK = 256;
constexpr int kGridKnnMinKeys = 1024;

__device__ __forceinline__ int cell_coord(float v, float o, float inv_s, int n) {
  const int c = (int)floorf((v - o) * inv_s);
  return min(max(c, 0), n - 1);
}

}  // namespace s4g
