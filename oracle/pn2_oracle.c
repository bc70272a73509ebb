/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C) of the seven `pn2_ext` operators of
 * yzqin/s4g-release, fp32 and fp64, for parity checking of the sm_100a kernels in
 * s4g_release_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity pinning: the reference ships no golden vectors for these ops
 * (SURVEY.md §8c).  This file is pinned against the reference's own CUDA
 * kernels compiled unmodified for sm_100a (oracle/_ref, see build_ref.py) by
 * tests/test_ref_cuda_parity.py on the GPU box (reference CUDA vs this file vs the
 * sm_100a kernels, three ways, on the same seeded inputs).
 *
 * Every distance follows the arithmetic nvcc emits for the reference
 * expression (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1) under the default
 * -fmad=true:  d = fmaf(dz, dz, fmaf(dx, dx, dy*dy)).   Build with
 * -ffp-contract=off so the host compiler does not re-contract anything.
 *
 * All tensors use the reference's interface layout: channel-first (B, C, N)
 * fp32 (pn2o_*) or fp64 (pn2o_*_f64), int64 indices.   Paths below are relative to
 * inference/grasp_proposal/network_models/models/pointnet2_utils/ .
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

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

#define REAL float
#define FN(name) pn2o_##name
#define FMA fmaf
#define BIG_DIST ((float)1e40)
#include "pn2_oracle_body.inc"
#undef REAL
#undef FN
#undef FMA
#undef BIG_DIST

/* the double instantiation: the SASS of the reference's double kernels shows the same sequence
 * DMUL dy*dy ; DFMA dx*dx + . ; DFMA dz*dz + .  (oracle/_ref, cuobjdump) */
#define REAL double
#define FN(name) pn2o_##name##_f64
#define FMA fma
#define BIG_DIST 1e40
#include "pn2_oracle_body.inc"
