// Farthest point sampling for sm_100a — replaces FarthestPointSample / FarthestPointSampleKernel
// (reference: pointnet2_utils/csrc/sampling_kernel.cu:49-119,128-172).
//
// Design (B200-first, not a port of the reference's global-memory loop):
//   * one thread-block CLUSTER per cloud (1, 2, 4 or 8 CTAs of 512 threads, chosen so that
//     B x CLUSTER fills the 148 SMs and the cloud fits on chip);
//   * every point lives in REGISTERS for the whole kernel (x, y, z and its running min-distance,
//     P points per thread) — the M-1 dependent iterations never touch HBM or L2 again;
//   * per iteration: P fused distance/min/argmax updates per thread, a two-instruction warp argmax
//     (REDUX.MAX on the distance bits, REDUX.MIN on a tie-break key), one 24-byte record per warp
//     pushed into every CTA of the cluster through distributed shared memory, ONE barrier
//     (__syncthreads or barrier.cluster), and a second warp-REDUX over the <=128 records that every
//     warp performs redundantly, so there is no broadcast step;
//   * the winner's coordinates travel with its record (read from a shared-memory copy of the cloud
//     by the owning lane), so the next iteration starts without a global load.
//
// Tie-breaking.  The reference reduces BLOCK = min(nextpow2(N),512) per-thread candidates with a
// shared-memory tree (offset = BLOCK/2 .. 1) that keeps the LOWER slot on ties, after each thread
// kept the first strict maximum of its strided points.  That tournament is a total order:
//     larger distance first, then smaller bitreverse_{log2 BLOCK}(j mod BLOCK), then smaller j,
// and if the maximum distance is 0 the previous index repeats.  The key below encodes exactly that
// order, so any reduction shape gives the reference's answer bit-for-bit:
//     tb(j) = __brev(j & (BLOCK-1)) | (j >> log2 BLOCK)        (minimised among equal distances)
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "grid.cuh"

namespace cg = cooperative_groups;

namespace s4g {

__device__ __forceinline__ unsigned fps_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fps_mbar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(fps_smem_u32(bar)));
}
__device__ __forceinline__ void fps_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fps_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fps_mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok = 0, spins = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(fps_smem_u32(bar)), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) __trap();  // a protocol bug traps instead of hanging the GPU
  }
}
__device__ __forceinline__ unsigned fps_mapa(unsigned addr, int rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void fps_st_async_v4(unsigned raddr, uint4 v, unsigned rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(rbar) : "memory");
}
__device__ __forceinline__ void fps_st_async_b32(unsigned raddr, unsigned v, unsigned rbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(v), "r"(rbar)
               : "memory");
}

constexpr int kFpsThreads = 512;
constexpr int kFpsWarps = kFpsThreads / 32;

// SMEM_XYZ: clouds too large for the register file (N > 102 400) keep only the running distances in registers
// and re-read the coordinates from the CTA's shared-memory copy every iteration (clusters of up to 16 CTAs).
template <int P, int CLUSTER, bool SMEM_XYZ, typename IndexT>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_kernel(const float* __restrict__ points, int N, int M, int L, IndexT* __restrict__ index) {
  extern __shared__ float s_xyz[];  // [3][kFpsThreads * P]: this CTA's slice of the cloud
  constexpr int kLocal = kFpsThreads * P;
  // Clusters of 4+ CTAs reduce inside the CTA first and exchange ONE record per CTA: the receiving CTA retires
  // remote stores one at a time, and 16 x CLUSTER x 2 of them per iteration cost more than the distance pass
  // (measured per iteration at N = 25 600: cluster 8, 1.85 -> 1.07 us; cluster 4, 0.90 -> 0.86 us; cluster 2 is
  // faster with the direct per-warp exchange, 1.04 vs 1.12 us: profiles/r01/fps_probe.txt).
  constexpr bool kHier = CLUSTER >= 4;
  constexpr int kEntries = kHier ? CLUSTER : kFpsWarps * CLUSTER;
  // one 32-byte record per warp (per CTA when kHier) of the cluster and iteration parity:
  // {distance bits, tie-break key, x, y | z}
  __shared__ __align__(16) uint4 s_rec[2][kEntries][2];
  __shared__ __align__(16) uint4 s_wrec[kHier ? 2 : 1][kHier ? kFpsWarps : 1][2];  // kHier: this CTA's warp records
  __shared__ __align__(8) unsigned long long s_bar[2];  // clusters: records of parity b have all landed

  const int t = threadIdx.x;
  const int lane = t & 31;
  const int warp = t >> 5;
  unsigned rank = 0;
  if constexpr (CLUSTER > 1) rank = cg::this_cluster().block_rank();
  const int cloud = blockIdx.x / CLUSTER;
  const float* X = points + (size_t)cloud * 3 * N;
  const float* Y = X + N;
  const float* Z = Y + N;
  IndexT* out = index + (size_t)cloud * M;

  float* sx = s_xyz;
  float* sy = s_xyz + kLocal;
  float* sz = s_xyz + 2 * kLocal;

  // Thread t of CTA `rank` owns points j = t + 512 * (rank * P + p), p = 0..P-1: increasing j and
  // (for BLOCK = 512) a constant reduction slot, so "first strict maximum" inside the thread is
  // the reference's per-thread rule.
  const int chunk = rank * P;
  // (a packed fp32x2 distance update — FADD2 / FMUL2 / FFMA2 — was measured and is bit-exact but not faster: the
  // packed instructions issue at half rate, so the scalar form is kept)
  constexpr int PR = SMEM_XYZ ? 1 : P;
  float px[PR], py[PR], pz[PR], dist[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int j = t + kFpsThreads * (chunk + p);
    const bool valid = j < N;
    const float x = valid ? X[j] : 0.f, y = valid ? Y[j] : 0.f, z = valid ? Z[j] : 0.f;
    if constexpr (!SMEM_XYZ) { px[p] = x; py[p] = y; pz[p] = z; }
    dist[p] = valid ? __int_as_float(0x7f800000) : 0.f;  // padding can never exceed a real point
    sx[p * kFpsThreads + t] = x;
    sy[p * kFpsThreads + t] = y;
    sz[p * kFpsThreads + t] = z;
  }
  const unsigned bmask = (1u << L) - 1u;
  const unsigned lowmask = (L == 0) ? 0xffffffffu : ((1u << (32 - L)) - 1u);

  int cur = 0;
  float cx = X[0], cy = Y[0], cz = Z[0];
  if (rank == 0 && t == 0) out[0] = 0;

  if constexpr (CLUSTER > 1) {
    if (t == 0) {
      fps_mbar_init(&s_bar[0]);
      fps_mbar_init(&s_bar[1]);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cg::this_cluster().sync();  // peers are resident and their barriers initialised before any remote store
  } else {
    __syncthreads();
  }

  for (int i = 1; i < M; ++i) {
    const int buf = i & 1;
    float best = 0.f;
    int bi = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      float x, y, z;
      if constexpr (SMEM_XYZ) { x = sx[p * kFpsThreads + t]; y = sy[p * kFpsThreads + t]; z = sz[p * kFpsThreads + t]; }
      else { x = px[p]; y = py[p]; z = pz[p]; }
      const float d = sqdist(__fsub_rn(x, cx), __fsub_rn(y, cy), __fsub_rn(z, cz));
      const float dd = fminf(dist[p], d);
      dist[p] = dd;
      if (dd > best) { best = dd; bi = p; }
    }
    // ---- warp argmax: 2 REDUX ----
    const unsigned db = __float_as_uint(best);  // distances are >= 0: bit order == value order
    const unsigned wmax = __reduce_max_sync(0xffffffffu, db);
    const unsigned j = (unsigned)(t + kFpsThreads * (chunk + bi));
    const unsigned tb = (db == wmax) ? (__brev(j & bmask) | (j >> L)) : 0xffffffffu;
    const unsigned wtb = __reduce_min_sync(0xffffffffu, tb);
    constexpr unsigned kRecBytes = 20;  // 16-byte + 4-byte remote store per record
    if constexpr (kHier) {
      if (tb == wtb) {  // exactly one lane: keys are distinct per point
        const int lp = bi * kFpsThreads + t;
        s_wrec[buf][warp][0] = make_uint4(wmax, wtb, __float_as_uint(sx[lp]), __float_as_uint(sy[lp]));
        s_wrec[buf][warp][1].x = __float_as_uint(sz[lp]);
      }
      __syncthreads();
      if (warp == 0) {
        // (a record that lands before the barrier is armed just makes the transaction count negative for a moment)
        if (lane == 0) fps_mbar_expect_tx(&s_bar[buf], CLUSTER * kRecBytes);
        uint2 kv = make_uint2(0u, 0xffffffffu);
        if (lane < kFpsWarps) kv = *reinterpret_cast<const uint2*>(&s_wrec[buf][lane][0]);
        const unsigned cmax = __reduce_max_sync(0xffffffffu, kv.x);
        const unsigned ck = __reduce_min_sync(0xffffffffu, (kv.x == cmax) ? kv.y : 0xffffffffu);
        if (lane < kFpsWarps && kv.x == cmax && kv.y == ck) {
          const uint4 lo = s_wrec[buf][lane][0];
          const unsigned zb = s_wrec[buf][lane][1].x;
          const unsigned rec = fps_smem_u32(&s_rec[buf][rank][0]);
          const unsigned bar = fps_smem_u32(&s_bar[buf]);
#pragma unroll
          for (int q = 0; q < CLUSTER; ++q) {
            const unsigned rrec = fps_mapa(rec, q), rbar = fps_mapa(bar, q);
            fps_st_async_v4(rrec, lo, rbar);
            fps_st_async_b32(rrec + 16, zb, rbar);
          }
        }
      }
    } else {
    if constexpr (CLUSTER > 1) {
      // arm this parity's barrier for the kEntries records of this iteration (a record that lands first just
      // makes the transaction count negative for a moment: the phase cannot complete before this arrival)
      if (t == 0) fps_mbar_expect_tx(&s_bar[buf], kEntries * kRecBytes);
    }
    if (tb == wtb) {  // exactly one lane: keys are distinct per point
      const int lp = bi * kFpsThreads + t;
      const int e = rank * kFpsWarps + warp;
      const uint4 lo = make_uint4(wmax, wtb, __float_as_uint(sx[lp]), __float_as_uint(sy[lp]));
      const unsigned zb = __float_as_uint(sz[lp]);
      if constexpr (CLUSTER > 1) {
        // push the record into every CTA of the cluster; each store signals the destination's barrier, so
        // nobody waits for a cluster-wide barrier (no barrier.cluster, no L1 flush): DSMEM latency only
        const unsigned rec = fps_smem_u32(&s_rec[buf][e][0]);
        const unsigned bar = fps_smem_u32(&s_bar[buf]);
#pragma unroll
        for (int q = 0; q < CLUSTER; ++q) {
          const unsigned rrec = fps_mapa(rec, q), rbar = fps_mapa(bar, q);
          fps_st_async_v4(rrec, lo, rbar);
          fps_st_async_b32(rrec + 16, zb, rbar);
        }
      } else {
        s_rec[buf][e][0] = lo;
        s_rec[buf][e][1] = make_uint4(zb, 0u, 0u, 0u);
      }
    }
    }
    if constexpr (CLUSTER > 1) fps_mbar_wait(&s_bar[buf], (unsigned)((i - 1) >> 1) & 1u);  // ((i-1)/2)-th use of this parity
    else __syncthreads();
    // ---- every warp reduces the cluster's records redundantly ----
    unsigned d = 0u, k = 0xffffffffu;
    int e = 0;
#pragma unroll
    for (int q = lane; q < kEntries; q += 32) {  // (kHier: at most 16 records, one per CTA)
      const uint2 kv = *reinterpret_cast<const uint2*>(&s_rec[buf][q][0]);
      if (kv.x > d || (kv.x == d && kv.y < k)) { d = kv.x; k = kv.y; e = q; }
    }
    const unsigned gmax = __reduce_max_sync(0xffffffffu, d);
    const unsigned gk = __reduce_min_sync(0xffffffffu, (d == gmax) ? k : 0xffffffffu);
    const unsigned who = __ballot_sync(0xffffffffu, d == gmax && k == gk);
    const int src = __ffs(who) - 1;
    const int ge = __shfl_sync(0xffffffffu, e, src);
    if (gmax != 0u) {  // all remaining distances 0 -> the reference repeats the previous index
      const uint4 lo = s_rec[buf][ge][0];
      cx = __uint_as_float(lo.z); cy = __uint_as_float(lo.w); cz = __uint_as_float(s_rec[buf][ge][1].x);
      cur = (int)(__brev(gk & ~lowmask) + ((gk & lowmask) << L));
    }
    if (rank == 0 && t == 0) out[i] = (IndexT)cur;
  }
  if constexpr (CLUSTER > 1) cg::this_cluster().sync();  // no CTA exits while a peer may still read or write it
}

// ------------------------------------------------------------------------------------------------
// Bucket FPS — exact, output-sensitive.  The kernels above recompute all N distances in each of the M - 1
// iterations; but a point's running minimum can only change when the new centroid is closer than that minimum,
// and late in the process the minima are a few point spacings.  The cloud is cut into 1024 spatial buckets
// (equal chunks of the cell-sorted point array of csrc/grid.cu); each bucket keeps the tight bounding box of its points and its current
// (max running-minimum, tie key, coordinates of that point).  In an iteration a bucket is touched only if the
// centroid is closer to its box than its current maximum — otherwise none of its minima can change and its record
// stands.  The argmax over bucket records with the same tie key is the argmax over points, so the result is the
// reference's bit for bit (tests: every FPS parity case).  The test is conservative by 1e-5 relative, far above the
// rounding of either side; a bucket that is not skipped is recomputed exactly.
// One CTA of 1024 threads per cloud — no cluster, no remote exchange: lane l of warp w owns bucket 32 l + w (its
// record lives in that lane's registers; spatial neighbours land in different warps), the running minima of all
// points live in shared memory (cell order), coordinates are re-read from the cell-sorted copy in L2 only for the
// buckets that are touched.  Per iteration: box test, the touched buckets (usually 0-2 per warp), warp argmax
// (2 REDUX), ONE __syncthreads, and the 32-record reduction every warp repeats.
// ------------------------------------------------------------------------------------------------
constexpr int kBucketThreads = 1024;

template <typename IndexT>
__global__ void __launch_bounds__(kBucketThreads, 1)
fps_bucket_kernel(const int* __restrict__ start, const float4* __restrict__ sorted,
                  const float* __restrict__ points, int N, int M, int L, IndexT* __restrict__ index) {
  extern __shared__ float s_temp[];  // [N] running minimum of every point, in cell order
  __shared__ __align__(16) uint4 s_rec[2][32][2];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int cloud = blockIdx.x;
  // buckets = 1024 equal chunks of the cell-sorted point array (cells are visited x-fastest, so a chunk is a short
  // run of neighbouring cells: compact, and balanced whatever the density — a uniform grid of 1024 cells leaves a
  // table-top scene with ~200 occupied cells of > 100 points)
  const int bucket = lane * 32 + warp;
  const int per = (N + kBucketThreads - 1) / kBucketThreads;
  const int base = start[(size_t)cloud * kGridCells];  // start[] is a batch-global prefix
  const int b_start = min(bucket * per, N);
  const int b_count = min(per, N - b_start);
  const float4* pts = sorted + base;
  IndexT* out = index + (size_t)cloud * M;
  const unsigned bmask = (1u << L) - 1u;
  const unsigned lowmask = (L == 0) ? 0xffffffffu : ((1u << (32 - L)) - 1u);
  const float inf = __int_as_float(0x7f800000);

  for (int j = t; j < N; j += kBucketThreads) s_temp[j] = inf;
  // tight box of this lane's bucket
  float lox = inf, loy = inf, loz = inf, hix = -inf, hiy = -inf, hiz = -inf;
  for (int k = 0; k < b_count; ++k) {
    const float4 p = __ldg(pts + b_start + k);
    lox = fminf(lox, p.x); loy = fminf(loy, p.y); loz = fminf(loz, p.z);
    hix = fmaxf(hix, p.x); hiy = fmaxf(hiy, p.y); hiz = fmaxf(hiz, p.z);
  }
  float r_best = b_count > 0 ? inf : 0.f;  // bucket record: max running minimum, its tie key, its coordinates
  unsigned r_key = 0xffffffffu;
  float r_x = 0.f, r_y = 0.f, r_z = 0.f;

  const float* X = points + (size_t)cloud * 3 * N;
  float cx = X[0], cy = X[N], cz = X[2 * (size_t)N];
  int cur = 0;
  if (t == 0) out[0] = 0;
  __syncthreads();

  for (int i = 1; i < M; ++i) {
    const int buf = i & 1;
    // ---- which of this warp's buckets can change? ----
    const float ex = fmaxf(fmaxf(lox - cx, cx - hix), 0.f);
    const float ey = fmaxf(fmaxf(loy - cy, cy - hiy), 0.f);
    const float ez = fmaxf(fmaxf(loz - cz, cz - hiz), 0.f);
    const float dmin2 = (ex * ex + ey * ey + ez * ez) * 0.99999f;
    const bool touched = b_count > 0 && !(dmin2 >= r_best);
    unsigned todo = __ballot_sync(0xffffffffu, touched);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const int s0 = __shfl_sync(0xffffffffu, b_start, src);
      const int cnt = __shfl_sync(0xffffffffu, b_count, src);
      float best = 0.f, bx = 0.f, by = 0.f, bz = 0.f;
      unsigned bkey = 0xffffffffu;
      for (int k = lane; k < cnt; k += 32) {
        const float4 p = __ldg(pts + s0 + k);
        const float d = sqdist(__fsub_rn(p.x, cx), __fsub_rn(p.y, cy), __fsub_rn(p.z, cz));
        const float dd = fminf(s_temp[s0 + k], d);
        s_temp[s0 + k] = dd;
        const unsigned j = (unsigned)__float_as_int(p.w);
        const unsigned key = __brev(j & bmask) | (j >> L);
        if (dd > best || (dd == best && key < bkey)) { best = dd; bkey = key; bx = p.x; by = p.y; bz = p.z; }
      }
      const unsigned db = __float_as_uint(best);
      const unsigned wmax = __reduce_max_sync(0xffffffffu, db);
      const unsigned wkey = __reduce_min_sync(0xffffffffu, db == wmax ? bkey : 0xffffffffu);
      const unsigned who = __ballot_sync(0xffffffffu, db == wmax && bkey == wkey);
      const int win = __ffs(who) - 1;
      const float wx = __shfl_sync(0xffffffffu, bx, win), wy = __shfl_sync(0xffffffffu, by, win),
                  wz = __shfl_sync(0xffffffffu, bz, win);
      if (lane == src) { r_best = __uint_as_float(wmax); r_key = wkey; r_x = wx; r_y = wy; r_z = wz; }
    }
    // ---- warp argmax over its 32 bucket records ----
    const unsigned rb = __float_as_uint(r_best);
    const unsigned wmax = __reduce_max_sync(0xffffffffu, rb);
    const unsigned wkey = __reduce_min_sync(0xffffffffu, rb == wmax ? r_key : 0xffffffffu);
    if (wmax == 0u) {  // every point of this warp's buckets already has distance 0: it cannot win
      if (lane == 0) s_rec[buf][warp][0] = make_uint4(0u, 0xffffffffu, 0u, 0u);
    } else if (rb == wmax && r_key == wkey) {  // exactly one lane: keys are distinct per point
      s_rec[buf][warp][0] = make_uint4(wmax, wkey, __float_as_uint(r_x), __float_as_uint(r_y));
      s_rec[buf][warp][1].x = __float_as_uint(r_z);
    }
    __syncthreads();
    const uint2 kv = *reinterpret_cast<const uint2*>(&s_rec[buf][lane][0]);
    const unsigned gmax = __reduce_max_sync(0xffffffffu, kv.x);
    const unsigned gk = __reduce_min_sync(0xffffffffu, kv.x == gmax ? kv.y : 0xffffffffu);
    if (gmax != 0u) {  // all remaining distances 0 -> the reference repeats the previous index
      const unsigned who = __ballot_sync(0xffffffffu, kv.x == gmax && kv.y == gk);
      const int ge = __ffs(who) - 1;
      const uint4 lo = s_rec[buf][ge][0];
      cx = __uint_as_float(lo.z); cy = __uint_as_float(lo.w); cz = __uint_as_float(s_rec[buf][ge][1].x);
      cur = (int)(__brev(gk & ~lowmask) + ((gk & lowmask) << L));
    }
    if (t == 0) out[i] = (IndexT)cur;
  }
}

// OFF by default: measured at 64 x 25 600 -> 5 120 on table-top scenes it takes 5.66 ms against 5.56 ms for the
// register-resident cluster kernel, and 5.1 ms against 3.9 ms for a single scene — the distance work drops ~20x, but an
// iteration is still a dependent chain (box test -> L2 reads of the touched buckets -> warp argmax -> barrier -> block
// argmax) of ~2 000 cycles, and the first ~100 iterations touch every bucket.  Kept (s4g_fps_set_bucket_mode) with its
// parity tests as the starting point for a version that keeps the coordinates on chip.
static int g_fps_bucket_mode = 0;
constexpr int kBucketMinPoints = 4096;    // below this the register-resident kernels are as fast
constexpr int kBucketMaxPoints = 55000;   // running minima must fit in shared memory (4 B per point)

template <typename IndexT>
static int launch_fps_bucket(const float* points, int B, int N, int M, int L, IndexT* index, cudaStream_t stream) {
  Grid g = {};
  int rc = grid_build(points, B, N, GRID_KNN, 0.f, &g, stream);
  if (rc != S4G_OK) return rc;
  auto kern = fps_bucket_kernel<IndexT>;
  const size_t smem = sizeof(float) * (size_t)N;
  S4G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<B, kBucketThreads, smem, stream>>>(g.start, g.sorted, points, N, M, L, index);
  count_launch();
  cudaError_t e = cudaGetLastError();
  rc = grid_free(&g, stream);
  if (e != cudaSuccess) return set_error((int)e, "fps_bucket launch: %s", cudaGetErrorString(e));
  return rc;
}

template <int P, int CLUSTER, bool SMEM_XYZ, typename IndexT>
static int launch_fps(const float* points, int B, int N, int M, int L, IndexT* index, cudaStream_t stream) {
  auto kern = fps_kernel<P, CLUSTER, SMEM_XYZ, IndexT>;
  const size_t smem = (size_t)3 * kFpsThreads * P * sizeof(float);
  S4G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CLUSTER > 8) S4G_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CLUSTER));
  cfg.blockDim = dim3(kFpsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  S4G_CUDA(cudaLaunchKernelEx(&cfg, kern, points, N, M, L, index));
  count_launch();
  return S4G_OK;
}

template <int CLUSTER, typename IndexT>
static int dispatch_p(int P, const float* points, int B, int N, int M, int L, IndexT* index, cudaStream_t stream) {
#define S4G_FPS_CASE(PP) \
  if (P <= PP) return launch_fps<PP, CLUSTER, false, IndexT>(points, B, N, M, L, index, stream);
  S4G_FPS_CASE(1)
  S4G_FPS_CASE(2)
  S4G_FPS_CASE(4)
  S4G_FPS_CASE(7)
  S4G_FPS_CASE(10)
  S4G_FPS_CASE(13)
  S4G_FPS_CASE(16)
  S4G_FPS_CASE(20)
  S4G_FPS_CASE(25)
#undef S4G_FPS_CASE
  return set_error(S4G_E_UNSUPPORTED, "farthest_point_sample: %d points per thread exceeds the register budget", P);
}

constexpr int kFpsMaxP = 25;

template <typename IndexT>
static int fps_entry(const float* points, int B, int N, int M, IndexT* index, cudaStream_t stream) {
  S4G_CHECK_ARG(points != nullptr && index != nullptr, "farthest_point_sample: null pointer");
  S4G_CHECK_ARG(B >= 0 && N > 0, "farthest_point_sample: bad shape B=%d N=%d", B, N);
  S4G_CHECK_ARG(M > 0, "farthest_point_sample: num_centroids <= 0");           // CHECK_GT, sampling_kernel.cu:138
  S4G_CHECK_ARG(N >= M, "farthest_point_sample: num_points < num_centroids");  // CHECK_GE, :139
  if (B == 0) return S4G_OK;
  // BLOCK of the reference (sampling_kernel.cu:34-42,150-167) only enters through the tie rule.
  int L = 0;
  while ((1 << L) < N && L < 9) ++L;
  if (L < 4) L = 4;
  if (g_fps_bucket_mode && N >= kBucketMinPoints && N <= kBucketMaxPoints && M >= 64)
    return launch_fps_bucket<IndexT>(points, B, N, M, L, index, stream);
  // clouds beyond the register-resident capacity (8 CTAs x 512 threads x 25 points): coordinates in shared memory,
  // 32 running distances per thread, clusters of 8 or 16 CTAs (up to 262 144 points)
  if (N > kFpsThreads * 8 * kFpsMaxP) {
    constexpr int kBigP = 32;
    if (N <= kFpsThreads * 8 * kBigP) return launch_fps<kBigP, 8, true, IndexT>(points, B, N, M, L, index, stream);
    if (N <= kFpsThreads * 16 * kBigP) return launch_fps<kBigP, 16, true, IndexT>(points, B, N, M, L, index, stream);
    return set_error(S4G_E_UNSUPPORTED, "farthest_point_sample: N=%d exceeds the on-chip capacity (%d points)", N,
                     kFpsThreads * 16 * kBigP);
  }
  // Cluster size from a measured cost model (profiles/r01/fps_probe2.txt): one iteration costs about
  // a(cluster) + 0.021 us x points per thread, a = 0.29 / 0.50 / 0.63 / 0.64 us for clusters of 1 / 2 / 4 / 8 CTAs
  // (block barrier vs distributed-shared-memory exchange), times the number of waves the batch needs on the SMs.
  // A single CTA wins whenever the cloud is small (<= ~10 points per thread); clusters of 8 only pay for a handful
  // of clouds (16 clouds x 8 CTAs measured 1.5 us per iteration against 0.88 us with clusters of 4).
  const int sms = num_sms();
  int cluster = 0;
  double best_cost = 0.0;
  for (int c = 1; c <= 8; c *= 2) {
    const int Pc = (N + kFpsThreads * c - 1) / (kFpsThreads * c);
    if (Pc > kFpsMaxP || (c == 8 && B > 4 && cluster != 0)) continue;
    static const double a[9] = {0, 0.29, 0.50, 0, 0.63, 0, 0, 0, 0.64};
    const long long ctas = (long long)B * c;
    const double waves = (double)((ctas + sms - 1) / sms);
    const double cost = waves * (a[c] + 0.021 * Pc);
    if (cluster == 0 || cost < best_cost) { cluster = c; best_cost = cost; }
  }
  if (cluster == 0) cluster = 8;
  if (const char* e = getenv("S4G_FPS_CLUSTER")) {  // experiments: force the cluster size where the cloud still fits
    const int c = atoi(e);
    if ((c == 1 || c == 2 || c == 4 || c == 8) && (N + kFpsThreads * c - 1) / (kFpsThreads * c) <= kFpsMaxP) cluster = c;
  }
  const int P = (N + kFpsThreads * cluster - 1) / (kFpsThreads * cluster);
  switch (cluster) {
    case 1: return dispatch_p<1, IndexT>(P, points, B, N, M, L, index, stream);
    case 2: return dispatch_p<2, IndexT>(P, points, B, N, M, L, index, stream);
    case 4: return dispatch_p<4, IndexT>(P, points, B, N, M, L, index, stream);
    default: return dispatch_p<8, IndexT>(P, points, B, N, M, L, index, stream);
  }
}

}  // namespace s4g

extern "C" int s4g_fps_set_bucket_mode(int on) {
  const int old = s4g::g_fps_bucket_mode;
  s4g::g_fps_bucket_mode = on ? 1 : 0;
  return old;
}

extern "C" int s4g_farthest_point_sample_f32(const float* points, int B, int N, int M, int64_t* index, void* stream) {
  return s4g::fps_entry<int64_t>(points, B, N, M, index, (cudaStream_t)stream);
}

extern "C" int s4g_farthest_point_sample_f32_i32(const float* points, int B, int N, int M, int32_t* index,
                                                 void* stream) {
  return s4g::fps_entry<int32_t>(points, B, N, M, index, (cudaStream_t)stream);
}
