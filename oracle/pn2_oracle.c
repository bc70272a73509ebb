/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C) of the seven `pn2_ext` operators of
 * yzqin/s4g-release, fp32, for parity checking of the sm_100a kernels in
 * s4g_release_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity pinning: the reference ships no golden vectors for these ops
 * (SURVEY.md §8c).  This file is pinned against the reference's own CUDA
 * kernels compiled unmodified for sm_100a (oracle/_ref, see build_ref.py) by
 * tests/test_ref_cuda_parity.py on the GPU box, and against fixtures captured
 * from such a run (tests/golden/ref_cuda_*.npz).
 *
 * Every distance follows the arithmetic nvcc emits for the reference
 * expression (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1) under the default
 * -fmad=true:  d = fmaf(dz, dz, fmaf(dx, dx, dy*dy)).   Build with
 * -ffp-contract=off so the host compiler does not re-contract anything.
 *
 * All tensors use the reference's interface layout: channel-first (B, C, N)
 * fp32, int64 indices.   Paths below are relative to
 * inference/grasp_proposal/network_models/models/pointnet2_utils/ .
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float sqdist(float dx, float dy, float dz) {
  /* FMUL on dy, then two FFMAs (SASS of the reference objects) */
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

int pn2o_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* csrc/sampling_kernel.cu:34-42  get_block(): next pow2 of N, capped at 512;
 * the switch at :150-167 falls back to 16 for anything smaller. */
static int fps_block(int64_t n) {
  int cnt = 0;
  int64_t x = n - 1;
  while (x > 0) { x >>= 1; cnt++; }
  int64_t b = (int64_t)1 << cnt;
  if (b > 512) b = 512;
  if (b < 16) b = 16;
  return (int)b;
}

/*
 * csrc/sampling_kernel.cu:49-119 FarthestPointSampleKernel, simulated literally:
 * BLOCK virtual threads, each scanning j = t, t+BLOCK, ... with a strict '>'
 * against (max_dist = 0, max_ind = cur), then the shared-memory tree reduction
 * offset = BLOCK/2 .. 1 that keeps the LOWER slot on ties (:100-113).
 *   points (B,3,N) fp32  ->  index (B,M) int64.   temp starts at -1 (:144).
 */
void pn2o_farthest_point_sample(const float* points, int64_t B, int64_t N, int64_t M,
                                int64_t* index) {
  const int block = fps_block(N);
#pragma omp parallel for schedule(dynamic, 1) if (B >= 4)
  for (int64_t b = 0; b < B; ++b) {
    const float* px = points + b * 3 * N;
    const float* py = px + N;
    const float* pz = py + N;
    float* temp = (float*)malloc(sizeof(float) * (size_t)N);
    float* sd = (float*)malloc(sizeof(float) * (size_t)block);
    int32_t* si = (int32_t*)malloc(sizeof(int32_t) * (size_t)block);
    for (int64_t j = 0; j < N; ++j) temp[j] = -1.0f;
    int64_t* out = index + b * M;
    int32_t cur = 0;
    out[0] = 0;
    for (int64_t i = 1; i < M; ++i) {
      const float x1 = px[cur], y1 = py[cur], z1 = pz[cur];
#pragma omp parallel for schedule(static) if (B < 4 && N >= 4096)
      for (int t = 0; t < block; ++t) {
        float max_dist = 0.0f;
        int32_t max_ind = cur;
        for (int64_t j = t; j < N; j += block) {
          float dist = sqdist(px[j] - x1, py[j] - y1, pz[j] - z1);
          float last = temp[j];
          if (last > dist || last < 0) temp[j] = dist; else dist = last;
          if (dist > max_dist) { max_dist = dist; max_ind = (int32_t)j; }
        }
        sd[t] = max_dist;
        si[t] = max_ind;
      }
      for (int offset = block / 2; offset > 0; offset /= 2) {
        for (int t = 0; t < offset; ++t) {
          if (sd[t] < sd[t + offset]) { sd[t] = sd[t + offset]; si[t] = si[t + offset]; }
        }
      }
      cur = si[0];
      out[i] = cur;
    }
    free(temp); free(sd); free(si);
  }
}

/* functions.py:10-25 gather_points: out[b,c,m] = points[b,c,index[b,m]] */
void pn2o_gather_points(const float* points, const int64_t* index, int64_t B, int64_t C,
                        int64_t N, int64_t M, float* out) {
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c)
      for (int64_t m = 0; m < M; ++m)
        out[(b * C + c) * M + m] = points[(b * C + c) * N + index[b * M + m]];
}

/*
 * csrc/ball_query_kernel.cu:33-76 BallQueryKernel.  Ascending j, strict
 * d < r*r (r*r in fp32, :48), first hit pre-fills all K slots (:64-67), stop at
 * K hits (:57); index zero-initialised (:109), count = number of hits <= K.
 */
void pn2o_ball_query(const float* points, const float* centroids, int64_t B, int64_t N,
                     int64_t M, float radius, int64_t K, int64_t* index, int64_t* count) {
  const float r2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(dynamic, 64)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t i = 0; i < M; ++i) {
      const float* px = points + b * 3 * N;
      const float* py = px + N;
      const float* pz = py + N;
      const float* cx = centroids + b * 3 * M;
      const float x1 = cx[i], y1 = cx[M + i], z1 = cx[2 * M + i];
      int64_t* idx = index + (b * M + i) * K;
      for (int64_t k = 0; k < K; ++k) idx[k] = 0;
      int64_t cnt = 0;
      for (int64_t j = 0; j < N && cnt < K; ++j) {
        float d = sqdist(px[j] - x1, py[j] - y1, pz[j] - z1);
        if (d < r2) {
          if (cnt == 0) { for (int64_t k = 0; k < K; ++k) idx[k] = j; }
          else idx[cnt] = j;
          ++cnt;
        }
      }
      count[b * M + i] = cnt;
    }
  }
}

/* csrc/grouping_kernel.cu:32-54 GroupPointsForward: out[b,c,m,k] = in[b,c,idx[b,m,k]] */
void pn2o_group_points_forward(const float* input, const int64_t* index, int64_t B, int64_t C,
                               int64_t N, int64_t M, int64_t K, float* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      const float* in = input + (b * C + c) * N;
      const int64_t* idx = index + b * M * K;
      float* o = out + (b * C + c) * M * K;
      for (int64_t q = 0; q < M * K; ++q) o[q] = in[idx[q]];
    }
}

/* csrc/grouping_kernel.cu:57-96 GroupPointsBackwardKernel: scatter-add (atomicAdd
 * in the reference, so its summation order is unspecified; this restatement
 * sums in (m,k) order -> compare with a tolerance, not bit-exactly). */
void pn2o_group_points_backward(const float* grad_out, const int64_t* index, int64_t B, int64_t C,
                                int64_t N, int64_t M, int64_t K, float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)(B * C * N));
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      float* gi = grad_in + (b * C + c) * N;
      const int64_t* idx = index + b * M * K;
      const float* go = grad_out + (b * C + c) * M * K;
      for (int64_t q = 0; q < M * K; ++q) gi[idx[q]] += go[q];
    }
}

/*
 * csrc/interpolate_kernel.cu:33-81 PointSearchKernel (K = 3 only, :24,105).
 * min_dist[3] = {1e40} -> {+inf, 0, 0} in fp32, min_ind[3] = {-1, 0, 0}
 * (:53-54); strict '<' insertion in key order (:64), so earlier keys win ties.
 * The distance is computed query-minus-key (:62); the squares are identical.
 * Outputs index (B,Nq,3) int64 and SQUARED distance (B,Nq,3).
 */
void pn2o_point_search(const float* query, const float* key, int64_t B, int64_t Nq, int64_t Nk,
                       int64_t* index, float* distance) {
#pragma omp parallel for collapse(2) schedule(dynamic, 64)
  for (int64_t b = 0; b < B; ++b) {
    for (int64_t i = 0; i < Nq; ++i) {
      const float* q = query + b * 3 * Nq;
      const float* kx = key + b * 3 * Nk;
      const float* ky = kx + Nk;
      const float* kz = ky + Nk;
      const float x1 = q[i], y1 = q[Nq + i], z1 = q[2 * Nq + i];
      float md[3] = {(float)1e40, 0.0f, 0.0f};
      int mi[3] = {-1, 0, 0};
      for (int64_t j = 0; j < Nk; ++j) {
        float d = sqdist(x1 - kx[j], y1 - ky[j], z1 - kz[j]);
        for (int k = 0; k < 3; ++k) {
          if (d < md[k]) {
            for (int l = 2; l > k; --l) { md[l] = md[l - 1]; mi[l] = mi[l - 1]; }
            md[k] = d; mi[k] = (int)j;
            break;
          }
        }
      }
      for (int k = 0; k < 3; ++k) {
        index[(b * Nq + i) * 3 + k] = mi[k];
        distance[(b * Nq + i) * 3 + k] = md[k];
      }
    }
  }
}

/* csrc/interpolate_kernel.cu:139-181 InterpolateForwardKernel:
 * out[b,c,n] = fma(in2,w2, fma(in1,w1, fma(in0,w0, 0)))  (k = 0,1,2 in order) */
void pn2o_interpolate_forward(const float* input, const int64_t* index, const float* weight,
                              int64_t B, int64_t C, int64_t Nk, int64_t Nq, float* out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      const float* in = input + (b * C + c) * Nk;
      for (int64_t n = 0; n < Nq; ++n) {
        const int64_t* idx = index + (b * Nq + n) * 3;
        const float* w = weight + (b * Nq + n) * 3;
        float v = 0.0f;
        for (int k = 0; k < 3; ++k) v = fmaf(in[idx[k]], w[k], v);
        out[(b * C + c) * Nq + n] = v;
      }
    }
}

/* csrc/interpolate_kernel.cu:243-286 InterpolateBackwardKernel: scatter grad*w
 * to the 3 sources (atomicAdd in the reference; sequential order here). */
void pn2o_interpolate_backward(const float* grad_out, const int64_t* index, const float* weight,
                               int64_t B, int64_t C, int64_t Nk, int64_t Nq, float* grad_in) {
  memset(grad_in, 0, sizeof(float) * (size_t)(B * C * Nk));
#pragma omp parallel for collapse(2) schedule(static)
  for (int64_t b = 0; b < B; ++b)
    for (int64_t c = 0; c < C; ++c) {
      float* gi = grad_in + (b * C + c) * Nk;
      for (int64_t n = 0; n < Nq; ++n) {
        const int64_t* idx = index + (b * Nq + n) * 3;
        const float* w = weight + (b * Nq + n) * 3;
        const float g = grad_out[(b * C + c) * Nq + n];
        for (int k = 0; k < 3; ++k) gi[idx[k]] += g * w[k];
      }
    }
}

/*
 * Closed form of the FPS tie rule used by the sm_100a kernel, exposed so the
 * CPU tests can check it against the literal simulation above on tie-heavy
 * inputs: among points whose running min-distance equals the (positive)
 * maximum, the reference picks the one minimising
 *     ( bitreverse_{log2 BLOCK}(j mod BLOCK),  j ),
 * and repeats the previous index when the maximum is 0.
 */
void pn2o_farthest_point_sample_keyed(const float* points, int64_t B, int64_t N, int64_t M,
                                      int64_t* index) {
  const int block = fps_block(N);
  int L = 0;
  while ((1 << L) < block) ++L;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < B; ++b) {
    const float* px = points + b * 3 * N;
    const float* py = px + N;
    const float* pz = py + N;
    float* temp = (float*)malloc(sizeof(float) * (size_t)N);
    for (int64_t j = 0; j < N; ++j) temp[j] = INFINITY;
    int64_t* out = index + b * M;
    int64_t cur = 0;
    out[0] = 0;
    for (int64_t i = 1; i < M; ++i) {
      const float x1 = px[cur], y1 = py[cur], z1 = pz[cur];
      float best = 0.0f;
      uint64_t best_tb = ~(uint64_t)0;
      int64_t best_j = cur;
      for (int64_t j = 0; j < N; ++j) {
        float d = sqdist(px[j] - x1, py[j] - y1, pz[j] - z1);
        if (d < temp[j]) temp[j] = d; else d = temp[j];
        uint32_t slot = (uint32_t)(j & (block - 1)), rev = 0;
        for (int k = 0; k < L; ++k) rev |= ((slot >> k) & 1u) << (L - 1 - k);
        uint64_t tb = ((uint64_t)rev << 40) | (uint64_t)j;
        if (d > best || (d == best && d > 0.0f && tb < best_tb)) { best = d; best_tb = tb; best_j = j; }
      }
      cur = best_j;
      out[i] = cur;
    }
    free(temp);
  }
}
