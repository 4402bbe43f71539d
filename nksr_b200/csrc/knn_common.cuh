// Warp-level k-nearest selection shared by the kNN-PCA normals (normals.cu) and the GT-SDF generator (sdfgen.cu):
// candidates closer than the current bound go into a 128-entry shared-memory buffer of packed (squared distance bits,
// index) words; whenever it would overflow the warp sorts it (bitonic network), keeps the k best and tightens the
// bound to the k-th distance.  Ties go to the lower index.
#pragma once
#include "common.cuh"

namespace {

constexpr int kKnnBuf = 128;
constexpr unsigned long long kKnnInf = 0xffffffffffffffffull;

__device__ __forceinline__ void knn_sort128(unsigned long long* __restrict__ key, int lane) {
  // ascending bitonic sort of 128 packed words by one warp
  for (int k = 2; k <= kKnnBuf; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int q0 = 0; q0 < kKnnBuf / 2; q0 += 32) {
        const int q = q0 + lane;
        const int lo = ((q & ~(j - 1)) << 1) | (q & (j - 1));
        const int hi = lo | j;
        const unsigned long long a = key[lo], b = key[hi];
        if ((a > b) == ((lo & k) == 0)) { key[lo] = b; key[hi] = a; }
      }
      __syncwarp();
    }
  }
}

// `bound0`: candidates at or beyond this squared distance are of no interest (a level's answer is only accepted when
// the k-th distance stays within the cell size, so nothing further than that needs to be kept, let alone sorted)
__device__ __forceinline__ void knn_reset(unsigned long long* __restrict__ key, int& fill, float& bound, int lane,
                                          float bound0 = 3.0e38f) {
  for (int t = lane; t < kKnnBuf; t += 32) key[t] = kKnnInf;
  __syncwarp();
  fill = 0;
  bound = bound0;
}

// one chunk of <= 32 candidates (lane: squared distance d2 of candidate q, `valid`); all lanes must call
__device__ __forceinline__ void knn_push(unsigned long long* __restrict__ key, int& fill, float& bound, const int k,
                                         const float d2, const int q, const bool valid, const int lane) {
  const bool in = valid && d2 < bound;
  const unsigned bm = __ballot_sync(0xffffffffu, in);
  if (fill + __popc(bm) > kKnnBuf) {       // no room: keep the k best, tighten the bound
    knn_sort128(key, lane);
    for (int t = k + lane; t < kKnnBuf; t += 32) key[t] = kKnnInf;
    fill = fill < k ? fill : k;
    if (fill == k) bound = __uint_as_float((unsigned)(key[k - 1] >> 32));
    __syncwarp();
  }
  const bool in2 = in && d2 < bound;
  const unsigned bm2 = __ballot_sync(0xffffffffu, in2);
  if (in2) key[fill + __popc(bm2 & ((1u << lane) - 1u))] = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)q;
  fill += __popc(bm2);
  __syncwarp();
}

// candidates = the contiguous point range [sb, se) of the sorted cloud
__device__ __forceinline__ void knn_scan_range(unsigned long long* __restrict__ key, int& fill, float& bound, const int k,
                                               const float* __restrict__ xyz, const int64_t sb, const int64_t se,
                                               const float px, const float py, const float pz, const int lane) {
  for (int64_t q0 = sb; q0 < se; q0 += 32) {
    const int64_t q = q0 + lane;
    float d2 = 3.0e38f;
    if (q < se) {
      const float dx = __ldg(xyz + 3 * q) - px, dy = __ldg(xyz + 3 * q + 1) - py, dz = __ldg(xyz + 3 * q + 2) - pz;
      d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
    }
    knn_push(key, fill, bound, k, d2, (int)q, q < se, lane);
  }
}

// final sort; returns the number of neighbours found (<= k) and their k-th squared distance
__device__ __forceinline__ int knn_finish(unsigned long long* __restrict__ key, const int fill, const int k,
                                          float& dk2, const int lane) {
  knn_sort128(key, lane);
  const int got = fill < k ? fill : k;
  dk2 = got > 0 ? __uint_as_float((unsigned)(key[got - 1] >> 32)) : 0.f;
  return got;
}

}  // namespace
