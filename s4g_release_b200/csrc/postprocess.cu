// Device-side grasp post-processing for sm_100a — replaces the numpy / python-loop tail of the reference's
// GraspDetector (grasp_detector.py:124-185 post_processing + orthogonalization, :214-232 collision loop over
// view_collision_checker.py:37-65, :235-251 importance sampling) and adds the translation de-duplication the
// reference only sketches (utils/file_logger_cls.py:220-225; README.md:58 says NMS is omitted).
//
// The reference's selection has two indexing quirks that are part of its observable behaviour and are
// reproduced here (oracle/model_cpu.py::post_processing restates them):
//   * `frame_R[:, index_high2low]` (:153) indexes the FULL prediction with ranks inside the filtered set, and the
//     following `.transpose(0, 1)` is numpy's (a no-op on a 2-D array), so `.reshape([-1, 3, 3])` cuts the
//     (9, n_high) array row by row: element e of candidate k is Rsel.flat[9 k + e], Rsel = frame_R[:, ranks];
//   * `high_score_index[index_good_direction]` (:160) indexes the UNSORTED filtered set with sorted positions,
//     so candidate k is paired with the k-th filtered point in index order (and its score / translation).
// Precision follows the reference: fp32 softmax and Gram-Schmidt, fp64 score expectation / translation / 4x4.
#include <math.h>

#include "common.cuh"

namespace s4g {

constexpr int kPostThreads = 1024;

// ------------------------------------------------------------------------------------------------
// 1. per-point grasp score: softmax over the score classes (fp32), expectation with linspace weights (fp64)
//    grasp_detector.py:142-145
// ------------------------------------------------------------------------------------------------
__global__ void grasp_scores_kernel(const float* __restrict__ logits, int B, int C, int N, double* __restrict__ score) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * N) return;
  const int b = (int)(i / N), n = (int)(i - (long long)b * N);
  const float* x = logits + (long long)b * C * N + n;
  float m = x[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, x[(long long)c * N]);
  float e[8], sum = 0.f;
  for (int c = 0; c < C; ++c) {
    e[c] = expf(x[(long long)c * N] - m);
    sum += e[c];
  }
  double s = 0.0;
  for (int c = 0; c < C; ++c) {
    const double value = (double)(c + 1) / (double)C;  // np.linspace(0, 1, C + 1)[1:]
    s += value * (double)(e[c] / sum);
  }
  score[i] = s;
}

// ------------------------------------------------------------------------------------------------
// 2. selection: threshold -> rank (descending score) -> verticalness filter.  One block per scene.
//    grasp_detector.py:148-160
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;           // exclusive prefix of the warp totals
    if (lane == 31) s_warp[32] = wi;  // block total
  }
  __syncthreads();
  const int res = s_warp[warp] + incl - v;
  total = s_warp[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(kPostThreads, 1)
grasp_select_kernel(const double* __restrict__ score, const float* __restrict__ frame_R, int N, double score_thr,
                    double vert_thr, double a20, double a21, double a22, int* __restrict__ high,  // [B][N] scratch
                    int* __restrict__ perm,                                                       // [B][N] scratch
                    int max_out, int* __restrict__ out_point, int* __restrict__ out_rot, int* __restrict__ n_out,
                    int* __restrict__ n_high_out) {
  __shared__ int s_warp[33];
  __shared__ double s_tile[kPostThreads];
  const int b = blockIdx.x;
  const double* sc = score + (long long)b * N;
  int* H = high + (long long)b * N;
  int* P = perm + (long long)b * N;
  // ---- ordered compaction of the high-score points (np.nonzero) ----
  int n_high = 0;
  for (int base = 0; base < N; base += kPostThreads) {
    const int i = base + threadIdx.x;
    const int f = (i < N && sc[i] > score_thr) ? 1 : 0;
    int total;
    const int pos = block_exclusive_scan(f, s_warp, total);
    if (f) H[n_high + pos] = i;
    n_high += total;
  }
  __syncthreads();
  // ---- rank by counting: np.argsort(scores[high])[::-1]; ties: later position first ----
  for (int a0 = 0; a0 < n_high; a0 += kPostThreads) {
    const int a = a0 + threadIdx.x;
    const double sa = a < n_high ? sc[H[a]] : 0.0;
    int rank = 0;
    for (int t0 = 0; t0 < n_high; t0 += kPostThreads) {
      __syncthreads();
      if (t0 + threadIdx.x < n_high) s_tile[threadIdx.x] = sc[H[t0 + threadIdx.x]];
      __syncthreads();
      const int lim = min(kPostThreads, n_high - t0);
      for (int q = 0; q < lim; ++q) {
        const double sb = s_tile[q];
        rank += (sb > sa || (sb == sa && t0 + q > a)) ? 1 : 0;
      }
    }
    if (a < n_high) P[rank] = a;  // index_high2low[rank] = position inside the filtered set
  }
  __syncthreads();
  // ---- verticalness of the approach axis (first column) of candidate k's rotation (quirk 1) ----
  const float* R = frame_R + (long long)b * 9 * N;
  auto rot_elem = [&](int k, int e) -> float {
    const long long f = 9LL * k + e;  // position in the row-major (9, n_high) array frame_R[:, index_high2low]
    return R[(f / n_high) * N + P[f % n_high]];
  };
  int n_sel = 0;
  for (int base = 0; base < n_high; base += kPostThreads) {
    const int k = base + threadIdx.x;
    int f = 0;
    if (k < n_high) {
      // x_direction = -(camera2base_R @ TRAIN2REAL_R) @ rotation[:, :, 0]; vertical degree = its z component
      const double vd = a20 * (double)rot_elem(k, 0) + a21 * (double)rot_elem(k, 3) + a22 * (double)rot_elem(k, 6);
      f = vd > vert_thr ? 1 : 0;
    }
    int total;
    const int pos = block_exclusive_scan(f, s_warp, total);
    if (f && n_sel + pos < max_out) {
      out_point[(long long)b * max_out + n_sel + pos] = H[k];  // quirk 2: k-th filtered point in index order
      out_rot[(long long)b * max_out + n_sel + pos] = k;       // sorted position whose (scrambled) rotation is used
    }
    n_sel += total;
  }
  if (threadIdx.x == 0) {
    n_out[b] = n_sel;  // may exceed max_out: the caller re-runs with a larger capacity
    n_high_out[b] = n_high;
  }
}

// ------------------------------------------------------------------------------------------------
// 3. poses: translation decode, Gram-Schmidt, camera-frame 4x4.  grasp_detector.py:124-135,161-180
// ------------------------------------------------------------------------------------------------
struct Mat4 { double m[16]; };
struct TScore { double v[8]; };

__global__ void grasp_pose_kernel(const float* __restrict__ points, const float* __restrict__ frame_R,
                                  const float* __restrict__ frame_t, const double* __restrict__ score, int N, int T,
                                  const int* __restrict__ out_point, const int* __restrict__ out_rot,
                                  const int* __restrict__ n_out, const int* __restrict__ n_high_arr,
                                  const int* __restrict__ perm, int max_out, const Mat4 train2real, const TScore ts,
                                  double* __restrict__ poses, double* __restrict__ out_score) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = min(n_out[b], max_out);
  if (j >= n) return;
  const int pt = out_point[(long long)b * max_out + j];
  const int k = out_rot[(long long)b * max_out + j];
  const int n_high = n_high_arr[b];
  const int* P = perm + (long long)b * N;
  const float* R = frame_R + (long long)b * 9 * N;
  float r[9];  // rotation[k] = (frame_R[:, index_high2low]).reshape(-1, 3, 3)[k]: 9 consecutive elements, row-major
#pragma unroll
  for (int e = 0; e < 9; ++e) {
    const long long f = 9LL * k + e;
    r[e] = R[(f / n_high) * N + P[f % n_high]];
  }
  // translation = softmax(frame_t[:, valid_index]) (fp32); offset = sum(translation * t_score) (fp64)
  const float* Tl = frame_t + (long long)b * T * N + pt;
  float m = Tl[0];
  for (int c = 1; c < T; ++c) m = fmaxf(m, Tl[(long long)c * N]);
  float e4[8], sum = 0.f;
  for (int c = 0; c < T; ++c) {
    e4[c] = expf(Tl[(long long)c * N] - m);
    sum += e4[c];
  }
  double off = 0.0;
  for (int c = 0; c < T; ++c) off += (double)(e4[c] / sum) * ts.v[c];  // t_score = [0.08, 0.06, 0.04, 0.02] (:177)
  const float* Pp = points + (long long)b * 3 * N + pt;
  const double tx = -off * (double)r[0] + (double)Pp[0];
  const double ty = -off * (double)r[3] + (double)Pp[(long long)N];
  const double tz = -off * (double)r[6] + (double)Pp[2LL * N];
  // Gram-Schmidt in fp32, numpy's operation order
  float x0 = r[0], x1 = r[3], x2 = r[6];
  float nx = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
  x0 = x0 / nx; x1 = x1 / nx; x2 = x2 / nx;
  float y0 = r[1], y1 = r[4], y2 = r[7];
  const float d = __fadd_rn(__fadd_rn(__fmul_rn(x0, y0), __fmul_rn(x1, y1)), __fmul_rn(x2, y2));
  y0 = __fsub_rn(y0, __fmul_rn(d, x0)); y1 = __fsub_rn(y1, __fmul_rn(d, x1)); y2 = __fsub_rn(y2, __fmul_rn(d, x2));
  const float ny = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(y0, y0), __fmul_rn(y1, y1)), __fmul_rn(y2, y2)));
  y0 = y0 / ny; y1 = y1 / ny; y2 = y2 / ny;
  const float z0 = __fsub_rn(__fmul_rn(x1, y2), __fmul_rn(x2, y1));
  const float z1 = __fsub_rn(__fmul_rn(x2, y0), __fmul_rn(x0, y2));
  const float z2 = __fsub_rn(__fmul_rn(x0, y1), __fmul_rn(x1, y0));
  const double G[16] = {x0, y0, z0, tx, x1, y1, z1, ty, x2, y2, z2, tz, 0.0, 0.0, 0.0, 1.0};
  double* out = poses + ((long long)b * max_out + j) * 16;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += train2real.m[a * 4 + k] * G[k * 4 + c];
      out[a * 4 + c] = acc;
    }
  out_score[(long long)b * max_out + j] = score[(long long)b * N + pt];
}

// ------------------------------------------------------------------------------------------------
// 4. gripper / cloud collision test of every pose against every point (view_collision_checker.py:37-65).
//    One block per pose; fp32 like the reference's torch path.
// ------------------------------------------------------------------------------------------------
struct Gripper {
  float finger_length, bottom_length, half_hand_thickness, half_bottom_width, half_bottom_space, back_margin;
  float back_threshold, finger_threshold;
};

__global__ void __launch_bounds__(256)
grasp_collision_kernel(const double* __restrict__ poses, int n_poses, const float* __restrict__ cloud, int n_points,
                       const Gripper g, unsigned char* __restrict__ ok, int* __restrict__ counts) {
  const int p = blockIdx.x;
  if (p >= n_poses) return;
  __shared__ float gl[12];
  __shared__ int s_back, s_finger;
  if (threadIdx.x == 0) {
    // torch_batch_transformation_inv (utils/math_utils.py:27-40) of the fp32 copy of the pose
    float Rm[9], t[3];
    for (int a = 0; a < 3; ++a) {
      for (int c = 0; c < 3; ++c) Rm[a * 3 + c] = (float)poses[(long long)p * 16 + a * 4 + c];
      t[a] = (float)poses[(long long)p * 16 + a * 4 + 3];
    }
    for (int a = 0; a < 3; ++a) {
      for (int c = 0; c < 3; ++c) gl[a * 4 + c] = Rm[c * 3 + a];
      float acc = 0.f;
      for (int c = 0; c < 3; ++c) acc = __fmaf_rn(-Rm[c * 3 + a], t[c], acc);
      gl[a * 4 + 3] = acc;
    }
    s_back = 0;
    s_finger = 0;
  }
  __syncthreads();
  int back = 0, finger = 0;
  for (int i = threadIdx.x; i < n_points; i += blockDim.x) {
    const float px = cloud[3LL * i], py = cloud[3LL * i + 1], pz = cloud[3LL * i + 2];
    const float lx = gl[0] * px + gl[1] * py + gl[2] * pz + gl[3];
    if (!(lx < g.finger_length && lx > -g.bottom_length)) continue;
    const float ly = gl[4] * px + gl[5] * py + gl[6] * pz + gl[7];
    const float lz = gl[8] * px + gl[9] * py + gl[10] * pz + gl[11];
    const bool zc = lz < g.half_hand_thickness && lz > -g.half_hand_thickness;
    if (!zc) continue;
    if (ly < g.half_bottom_width && ly > -g.half_bottom_width && lx < -g.back_margin) ++back;
    const bool left = ly < g.half_bottom_width && ly > g.half_bottom_space;
    const bool right = ly > -g.half_bottom_width && ly < -g.half_bottom_space;
    if (left || right) ++finger;
  }
  back = __reduce_add_sync(0xffffffffu, back);
  finger = __reduce_add_sync(0xffffffffu, finger);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_back, back);
    atomicAdd(&s_finger, finger);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // back collisions are tested first and return early; the finger test follows (:50-65)
    ok[p] = ((float)s_back > g.back_threshold || (float)s_finger > g.finger_threshold) ? 0 : 1;
    if (counts) { counts[2 * p] = s_back; counts[2 * p + 1] = s_finger; }
  }
}

// ------------------------------------------------------------------------------------------------
// 5. greedy translation de-duplication in descending score order (the check sketched at
//    utils/file_logger_cls.py:220-225): a pose is dropped when the L1 distance of its translation to an
//    already kept pose is below `min_dist`.  One block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grasp_nms_kernel(const double* __restrict__ poses, const int* __restrict__ order, int n, double min_dist,
                 int* __restrict__ kept, int* __restrict__ n_kept) {
  __shared__ int s_n, s_hit;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    const int c = order[i];
    const double cx = poses[16LL * c + 3], cy = poses[16LL * c + 7], cz = poses[16LL * c + 11];
    if (threadIdx.x == 0) s_hit = 0;
    __syncthreads();
    const int nk = s_n;
    bool hit = false;
    for (int q = threadIdx.x; q < nk && !hit; q += blockDim.x) {
      const int k = kept[q];
      const double d = fabs(poses[16LL * k + 3] - cx) + fabs(poses[16LL * k + 7] - cy) + fabs(poses[16LL * k + 11] - cz);
      hit = d < min_dist;
    }
    if (hit) s_hit = 1;
    __syncthreads();
    if (threadIdx.x == 0 && !s_hit) kept[s_n++] = c;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_kept = s_n;
}

// ------------------------------------------------------------------------------------------------
// 6. importance sampling (grasp_detector.py:235-251): cum = cumsum(exp(5 s)); for each sorted uniform u the
//    first index with cum[idx] >= u * cum[-1].  Sequential fp64 cumsum like numpy's.
// ------------------------------------------------------------------------------------------------
__global__ void grasp_sample_kernel(const double* __restrict__ scores, int n, const double* __restrict__ sorted_u, int m,
                                    double* __restrict__ cum, int* __restrict__ picked) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
      acc += exp(5.0 * scores[i]);
      cum[i] = acc;
    }
    int idx = 0;
    for (int i = 0; i < m; ++i) {
      const double target = sorted_u[i] * cum[n - 1];
      while (idx < n - 1 && cum[idx] < target) ++idx;
      picked[i] = idx;
    }
  }
}


// ------------------------------------------------------------------------------------------------
// 7. the same tail for a whole batch without a host round trip (BASELINE config 5: thousands of scenes streamed
//    through one GPU): collision test of every candidate of every scene, then — one block per scene — ordered
//    compaction of the collision-free candidates, optional translation de-duplication in descending score
//    order, importance sampling (or "all of them" when at most m are left, grasp_detector.py:235), and the
//    gather of the selected poses.  Candidate counts stay on the device.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grasp_collision_batch_kernel(const double* __restrict__ poses, const int* __restrict__ n_cand, int cap,
                             const float* __restrict__ cloud_b3n, int n_points, const Gripper g,
                             unsigned char* __restrict__ ok) {
  const int b = blockIdx.y;
  const int n = min(n_cand[b], cap);
  __shared__ float gl[12];
  __shared__ int s_back, s_finger;
  const float* X = cloud_b3n + (long long)b * 3 * n_points;
  const float* Y = X + n_points;
  const float* Z = Y + n_points;
  for (int p = blockIdx.x; p < n; p += gridDim.x) {
    const double* P = poses + ((long long)b * cap + p) * 16;
    __syncthreads();
    if (threadIdx.x == 0) {
      float Rm[9], t[3];
      for (int a = 0; a < 3; ++a) {
        for (int c = 0; c < 3; ++c) Rm[a * 3 + c] = (float)P[a * 4 + c];
        t[a] = (float)P[a * 4 + 3];
      }
      for (int a = 0; a < 3; ++a) {
        for (int c = 0; c < 3; ++c) gl[a * 4 + c] = Rm[c * 3 + a];
        float acc = 0.f;
        for (int c = 0; c < 3; ++c) acc = __fmaf_rn(-Rm[c * 3 + a], t[c], acc);
        gl[a * 4 + 3] = acc;
      }
      s_back = 0;
      s_finger = 0;
    }
    __syncthreads();
    int back = 0, finger = 0;
    for (int i = threadIdx.x; i < n_points; i += blockDim.x) {
      const float px = X[i], py = Y[i], pz = Z[i];
      const float lx = gl[0] * px + gl[1] * py + gl[2] * pz + gl[3];
      if (!(lx < g.finger_length && lx > -g.bottom_length)) continue;
      const float ly = gl[4] * px + gl[5] * py + gl[6] * pz + gl[7];
      const float lz = gl[8] * px + gl[9] * py + gl[10] * pz + gl[11];
      const bool zc = lz < g.half_hand_thickness && lz > -g.half_hand_thickness;
      if (!zc) continue;
      if (ly < g.half_bottom_width && ly > -g.half_bottom_width && lx < -g.back_margin) ++back;
      const bool left = ly < g.half_bottom_width && ly > g.half_bottom_space;
      const bool right = ly > -g.half_bottom_width && ly < -g.half_bottom_space;
      if (left || right) ++finger;
    }
    back = __reduce_add_sync(0xffffffffu, back);
    finger = __reduce_add_sync(0xffffffffu, finger);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&s_back, back);
      atomicAdd(&s_finger, finger);
    }
    __syncthreads();
    if (threadIdx.x == 0)
      ok[(long long)b * cap + p] = ((float)s_back > g.back_threshold || (float)s_finger > g.finger_threshold) ? 0 : 1;
  }
}

constexpr int kFinishThreads = 1024;

__global__ void __launch_bounds__(kFinishThreads, 1)
grasp_finish_batch_kernel(const double* __restrict__ poses, const double* __restrict__ scores, const int* __restrict__ n_cand,
                          int cap, const unsigned char* __restrict__ ok, double nms_min_dist,
                          const double* __restrict__ sorted_u, int m, int* __restrict__ work,  // [B][3*cap]
                          double* __restrict__ cum,                                          // [B][cap]
                          int* __restrict__ out_index, int* __restrict__ out_n, double* __restrict__ out_poses,
                          double* __restrict__ out_scores) {
  const int b = blockIdx.x;
  const int n = min(n_cand[b], cap);
  const double* P = poses + (long long)b * cap * 16;
  const double* S = scores + (long long)b * cap;
  int* list = work + (long long)b * 3 * cap;  // collision-free candidates, candidate order
  int* order = list + cap;                    // ... in descending score order (stable)
  int* kept = order + cap;                    // ... surviving the de-duplication
  __shared__ int s_warp[32];
  __shared__ int s_base, s_n, s_hit;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // (a) ordered compaction
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < n; c0 += kFinishThreads) {
    const int i = c0 + threadIdx.x;
    const bool f = i < n && (ok == nullptr || ok[(long long)b * cap + i] != 0);
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (f) list[s_base + before + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  int L = s_base;
  const int* seq = list;
  // (b) translation de-duplication in descending score order (ties: lower index first)
  if (nms_min_dist > 0.0 && L > 0) {
    for (int i = threadIdx.x; i < L; i += kFinishThreads) {
      const double si = S[list[i]];
      int rank = 0;
      for (int j = 0; j < L; ++j) {
        const double sj = S[list[j]];
        rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
      }
      order[rank] = list[i];
    }
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    for (int i = 0; i < L; ++i) {
      const int c = order[i];
      const double cx = P[16LL * c + 3], cy = P[16LL * c + 7], cz = P[16LL * c + 11];
      if (threadIdx.x == 0) s_hit = 0;
      __syncthreads();
      const int nk = s_n;
      bool hit = false;
      for (int q = threadIdx.x; q < nk && !hit; q += kFinishThreads) {
        const int k = kept[q];
        hit = fabs(P[16LL * k + 3] - cx) + fabs(P[16LL * k + 7] - cy) + fabs(P[16LL * k + 11] - cz) < nms_min_dist;
      }
      if (hit) s_hit = 1;
      __syncthreads();
      if (threadIdx.x == 0 && !s_hit) kept[s_n++] = c;
      __syncthreads();
    }
    L = s_n;
    seq = kept;
  }
  // (c) importance sampling when more than m are left (grasp_detector.py:235-251), else all of them
  int n_sel = L < m ? L : m;
  int* oi = out_index + (long long)b * m;
  if (L > m && sorted_u != nullptr) {
    if (threadIdx.x == 0) {
      double* cm = cum + (long long)b * cap;
      double acc = 0.0;
      for (int i = 0; i < L; ++i) {
        acc += exp(5.0 * S[seq[i]]);
        cm[i] = acc;
      }
      int idx = 0;
      for (int i = 0; i < m; ++i) {
        const double target = sorted_u[(long long)b * m + i] * cm[L - 1];
        while (idx < L - 1 && cm[idx] < target) ++idx;
        oi[i] = seq[idx];
      }
    }
  } else {
    for (int i = threadIdx.x; i < n_sel; i += kFinishThreads) oi[i] = seq[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) out_n[b] = n_sel;
  // (d) gather
  for (int e = threadIdx.x; e < n_sel * 16; e += kFinishThreads)
    out_poses[((long long)b * m) * 16 + e] = P[16LL * oi[e >> 4] + (e & 15)];
  for (int i = threadIdx.x; i < n_sel; i += kFinishThreads) out_scores[(long long)b * m + i] = S[oi[i]];
}

}  // namespace s4g

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" int s4g_grasp_scores_f32(const float* score_logits, int B, int C, int N, double* score, void* stream) {
  S4G_CHECK_ARG(score_logits && score, "grasp_scores: null pointer");
  S4G_CHECK_ARG(B >= 0 && N > 0 && C >= 1 && C <= 8, "grasp_scores: bad shape B=%d C=%d N=%d", B, C, N);
  if (B == 0) return S4G_OK;
  const long long total = (long long)B * N;
  s4g::grasp_scores_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(score_logits, B, C, N, score);
  S4G_LAUNCH_CHECK("grasp_scores");
  return S4G_OK;
}

// approach_row: third row of -(camera2base[:3,:3] @ TRAIN2REAL[:3,:3]) (fp64, 3 values).
// workspace: 2 * B * N int32.  Outputs per scene: n_out (may exceed max_out -> enlarge and call again),
// n_high, out_point / out_rot [B][max_out].
extern "C" int s4g_grasp_select(const double* score, const float* frame_R, int B, int N, double score_threshold,
                                double vertical_threshold, const double* approach_row, int* workspace, int max_out,
                                int* out_point, int* out_rot, int* n_out, int* n_high, void* stream) {
  S4G_CHECK_ARG(score && frame_R && approach_row && workspace && out_point && out_rot && n_out && n_high,
                "grasp_select: null pointer");
  S4G_CHECK_ARG(B >= 0 && N > 0 && max_out > 0, "grasp_select: bad shape");
  if (B == 0) return S4G_OK;
  s4g::grasp_select_kernel<<<B, s4g::kPostThreads, 0, (cudaStream_t)stream>>>(
      score, frame_R, N, score_threshold, vertical_threshold, approach_row[0], approach_row[1], approach_row[2], workspace,
      workspace + (size_t)B * N, max_out, out_point, out_rot, n_out, n_high);
  S4G_LAUNCH_CHECK("grasp_select");
  return S4G_OK;
}

// n_high / workspace: as left by s4g_grasp_select.  train2real: 4x4 fp64 row-major, t_score: T fp64 offsets (both
// host memory).  poses [B][max_out][16] fp64, out_score [B][max_out] fp64.
extern "C" int s4g_grasp_poses(const float* points, const float* frame_R, const float* frame_t, const double* score, int B,
                               int N, int T, const int* out_point, const int* out_rot, const int* n_out, const int* n_high,
                               const int* workspace, int max_out,
                               const double* train2real, const double* t_score, double* poses, double* out_score,
                               void* stream) {
  S4G_CHECK_ARG(points && frame_R && frame_t && score && out_point && out_rot && n_out && n_high && workspace && train2real &&
                    t_score && poses && out_score,
                "grasp_poses: null pointer");
  S4G_CHECK_ARG(B >= 0 && N > 0 && T >= 1 && T <= 8 && max_out > 0, "grasp_poses: bad shape");
  if (B == 0) return S4G_OK;
  s4g::Mat4 m;
  for (int i = 0; i < 16; ++i) m.m[i] = train2real[i];
  s4g::TScore ts = {};
  for (int i = 0; i < T; ++i) ts.v[i] = t_score[i];
  dim3 grid((unsigned)((max_out + 127) / 128), (unsigned)B);
  s4g::grasp_pose_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(points, frame_R, frame_t, score, N, T, out_point, out_rot,
                                                                 n_out, n_high, workspace + (size_t)B * N, max_out, m, ts, poses, out_score);
  S4G_LAUNCH_CHECK("grasp_poses");
  return S4G_OK;
}

// gripper: 8 floats {finger_length, bottom_length, half_hand_thickness, half_bottom_width, half_bottom_space,
// back_margin, back_threshold, finger_threshold} (configs/gripper_config.py:9-21, processing_config.py:37-40).
extern "C" int s4g_grasp_collision_f32(const double* poses, int n_poses, const float* cloud_n3, int n_points,
                                       const float* gripper, unsigned char* ok, int* counts, void* stream) {
  S4G_CHECK_ARG(poses && cloud_n3 && gripper && ok, "grasp_collision: null pointer");
  S4G_CHECK_ARG(n_poses >= 0 && n_points >= 0, "grasp_collision: bad shape");
  if (n_poses == 0) return S4G_OK;
  s4g::Gripper g = {gripper[0], gripper[1], gripper[2], gripper[3], gripper[4], gripper[5], gripper[6], gripper[7]};
  s4g::grasp_collision_kernel<<<n_poses, 256, 0, (cudaStream_t)stream>>>(poses, n_poses, cloud_n3, n_points, g, ok, counts);
  S4G_LAUNCH_CHECK("grasp_collision");
  return S4G_OK;
}

extern "C" int s4g_grasp_nms(const double* poses, const int* order, int n, double min_dist, int* kept, int* n_kept,
                             void* stream) {
  S4G_CHECK_ARG(poses && order && kept && n_kept, "grasp_nms: null pointer");
  S4G_CHECK_ARG(n >= 0, "grasp_nms: bad shape");
  s4g::grasp_nms_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(poses, order, n, min_dist, kept, n_kept);
  S4G_LAUNCH_CHECK("grasp_nms");
  return S4G_OK;
}

extern "C" int s4g_grasp_importance_sample(const double* scores, int n, const double* sorted_uniform, int m, double* cum,
                                           int* picked, void* stream) {
  S4G_CHECK_ARG(scores && sorted_uniform && cum && picked, "grasp_importance_sample: null pointer");
  S4G_CHECK_ARG(n > 0 && m >= 0, "grasp_importance_sample: bad shape");
  s4g::grasp_sample_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scores, n, sorted_uniform, m, cum, picked);
  S4G_LAUNCH_CHECK("grasp_importance_sample");
  return S4G_OK;
}

// Batched tail (BASELINE config 5).  poses [B][cap][16] / scores [B][cap] / n_cand [B] as written by s4g_grasp_poses and
// s4g_grasp_select (cap = their max_out); cloud (B,3,n_points) fp32 channel-first, NULL = no collision test;
// nms_min_dist <= 0 = no de-duplication; sorted_uniform [B][m] fp64 ascending per scene (device), NULL = keep the first
// m.  workspace: B * cap * (3 * 4 + 8 + 1) bytes.  Outputs: out_index [B][m] (candidate index), out_n [B],
// out_poses [B][m][16], out_scores [B][m]; entries past out_n[b] are undefined.
extern "C" size_t s4g_grasp_finish_batch_workspace(int B, int cap) { return (size_t)B * cap * (3 * 4 + 8 + 1) + 64; }

extern "C" int s4g_grasp_finish_batch(const double* poses, const double* scores, const int* n_cand, int B, int cap,
                                      const float* cloud_b3n, int n_points, const float* gripper, double nms_min_dist,
                                      const double* sorted_uniform, int m, void* workspace, size_t workspace_bytes,
                                      int* out_index, int* out_n, double* out_poses, double* out_scores, void* stream) {
  S4G_CHECK_ARG(poses && scores && n_cand && workspace && out_index && out_n && out_poses && out_scores,
                "grasp_finish_batch: null pointer");
  S4G_CHECK_ARG(B >= 0 && cap > 0 && m > 0, "grasp_finish_batch: bad shape");
  S4G_CHECK_ARG(workspace_bytes >= s4g_grasp_finish_batch_workspace(B, cap), "grasp_finish_batch: workspace too small");
  S4G_CHECK_ARG(B <= 65535, "grasp_finish_batch: batch too large for one launch");
  if (B == 0) return S4G_OK;
  cudaStream_t s = (cudaStream_t)stream;
  double* cum = reinterpret_cast<double*>(workspace);
  int* work = reinterpret_cast<int*>(cum + (size_t)B * cap);
  unsigned char* ok = reinterpret_cast<unsigned char*>(work + (size_t)B * cap * 3);
  if (cloud_b3n) {
    S4G_CHECK_ARG(gripper && n_points > 0, "grasp_finish_batch: collision test needs the gripper and a cloud");
    s4g::Gripper g = {gripper[0], gripper[1], gripper[2], gripper[3], gripper[4], gripper[5], gripper[6], gripper[7]};
    dim3 grid((unsigned)(cap < 128 ? cap : 128), (unsigned)B);
    s4g::grasp_collision_batch_kernel<<<grid, 256, 0, s>>>(poses, n_cand, cap, cloud_b3n, n_points, g, ok);
    S4G_LAUNCH_CHECK("grasp_collision_batch");
  }
  s4g::grasp_finish_batch_kernel<<<B, s4g::kFinishThreads, 0, s>>>(poses, scores, n_cand, cap, cloud_b3n ? ok : nullptr,
                                                                  nms_min_dist, sorted_uniform, m, work, cum, out_index,
                                                                  out_n, out_poses, out_scores);
  S4G_LAUNCH_CHECK("grasp_finish_batch");
  return S4G_OK;
}
