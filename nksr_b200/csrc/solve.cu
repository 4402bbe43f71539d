// CSR SpMV and Jacobi-preconditioned conjugate gradients.
// Replaces the solve inside KernelField.solve* (models/nksr_net.py:105-112; tolerance knob
// solver_tol at examples/recons_waymo.py:33; verbose hook models/nksr_net.py:97-98).
//
// All dot products are two-stage and deterministic: every block writes one fp64 partial, the
// consumer kernels re-reduce the (fixed-length) partial array in a fixed order.  No host
// round trip per iteration: alpha and beta are formed on the device from the partials; the host
// only reads ||r||^2 every `check_every` iterations.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;
constexpr int kWarpsPerBlock = kBlock / 32;
constexpr int kGrid = 148 * 8;  // persistent-style grid: 8 blocks per SM (B200: 148 SMs)

// y = A x, one warp per row, grid-stride over rows.  DOT: also partial[blockIdx] = sum x_i * y_i
template <bool DOT>
__global__ void __launch_bounds__(kBlock)
k_spmv(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
       const float* __restrict__ x, float* __restrict__ y, int64_t n, double* __restrict__ partial) {
  __shared__ double wsum[kWarpsPerBlock];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  double local = 0.0;
  for (int64_t row = blockIdx.x * (int64_t)kWarpsPerBlock + wid; row < n; row += nwarps) {
    const int64_t b = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
    // matrix stream: evict-first loads (read once per SpMV); x: read-only path, stays in L1/L2.
    // Four independent 128 B column + value requests per lane keep ~1 KB per warp in flight.
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int64_t p = b + lane;
    for (; p < e; p += 128) {   // predicated tail: absent entries read as (col 0, value 0)
      const bool q1 = p + 32 < e, q2 = p + 64 < e, q3 = p + 96 < e;
      const int c0 = __ldcs(col + p), c1 = q1 ? __ldcs(col + p + 32) : 0, c2 = q2 ? __ldcs(col + p + 64) : 0,
                c3 = q3 ? __ldcs(col + p + 96) : 0;
      const float v0 = __ldcs(val + p), v1 = q1 ? __ldcs(val + p + 32) : 0.f, v2 = q2 ? __ldcs(val + p + 64) : 0.f,
                  v3 = q3 ? __ldcs(val + p + 96) : 0.f;
      s0 = fmaf(v0, __ldg(x + c0), s0);
      s1 = fmaf(v1, __ldg(x + c1), s1);
      s2 = fmaf(v2, __ldg(x + c2), s2);
      s3 = fmaf(v3, __ldg(x + c3), s3);
    }
    float s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) {
      y[row] = s;
      if (DOT) local += (double)s * (double)__ldg(x + row);
    }
  }
  if (DOT) {
    if (lane == 0) wsum[wid] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < kWarpsPerBlock; ++w) t += wsum[w];
      partial[blockIdx.x] = t;
    }
  }
}

// deterministic block-wide reduction of a fixed-length fp64 array; result broadcast to all threads
__device__ __forceinline__ double reduce_partials(const double* __restrict__ arr, int len, double* sh) {
  double t = 0.0;
  for (int i = threadIdx.x; i < len; i += kBlock) t += arr[i];
  t = warp_sum_d(t);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < kWarpsPerBlock; ++w) r += sh[w];
  __syncthreads();
  return r;
}

__device__ __forceinline__ void block_store_partial(double local, double* sh, double* dst) {
  local = warp_sum_d(local);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kWarpsPerBlock; ++w) t += sh[w];
    *dst = t;
  }
  __syncthreads();
}

// x = 0, r = b, z = r/diag, p = z; partials: rz, bb
__global__ void __launch_bounds__(kBlock)
k_pcg_init(const float* __restrict__ b, const float* __restrict__ diag, float* __restrict__ x, float* __restrict__ r,
           float* __restrict__ p, int64_t n, double* __restrict__ part_rz, double* __restrict__ part_bb) {
  __shared__ double sh[kWarpsPerBlock];
  double rz = 0.0, bb = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const float bi = b[i], d = diag[i];
    const float zi = d > 0.f ? bi / d : 0.f;
    x[i] = 0.f;
    r[i] = bi;
    p[i] = zi;
    rz += (double)bi * zi;
    bb += (double)bi * bi;
  }
  block_store_partial(rz, sh, part_rz + blockIdx.x);
  block_store_partial(bb, sh, part_bb + blockIdx.x);
}

// alpha = rz/pAp; x += alpha p; r -= alpha Ap; z = r/diag (kept in ap); partials rz_new, rr
__global__ void __launch_bounds__(kBlock)
k_pcg_update(const float* __restrict__ diag, const float* __restrict__ p, float* __restrict__ ap,
             float* __restrict__ x, float* __restrict__ r, int64_t n, const double* __restrict__ part_rz,
             const double* __restrict__ part_pap, double* __restrict__ part_rz_new, double* __restrict__ part_rr) {
  __shared__ double sh[kWarpsPerBlock];
  const double rz = reduce_partials(part_rz, kGrid, sh);
  const double pap = reduce_partials(part_pap, kGrid, sh);
  const float alpha = pap != 0.0 ? (float)(rz / pap) : 0.f;
  double rzn = 0.0, rr = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock) {
    const float pi = p[i], api = ap[i], d = diag[i];
    x[i] = fmaf(alpha, pi, x[i]);
    const float ri = fmaf(-alpha, api, r[i]);
    r[i] = ri;
    const float zi = d > 0.f ? ri / d : 0.f;
    ap[i] = zi;  // Ap is dead after this point: reuse its storage for z
    rzn += (double)ri * zi;
    rr += (double)ri * ri;
  }
  block_store_partial(rzn, sh, part_rz_new + blockIdx.x);
  block_store_partial(rr, sh, part_rr + blockIdx.x);
}

// beta = rz_new/rz; p = z + beta p
__global__ void __launch_bounds__(kBlock)
k_pcg_direction(const float* __restrict__ z, float* __restrict__ p, int64_t n, const double* __restrict__ part_rz,
                const double* __restrict__ part_rz_new) {
  __shared__ double sh[kWarpsPerBlock];
  const double rz = reduce_partials(part_rz, kGrid, sh);
  const double rzn = reduce_partials(part_rz_new, kGrid, sh);
  const float beta = rz != 0.0 ? (float)(rzn / rz) : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)kBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kBlock)
    p[i] = fmaf(beta, p[i], z[i]);
}

struct PcgWs {
  float *r, *p, *ap;
  double *rz0, *rz1, *pap, *rr, *bb;
};

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static PcgWs carve(void* ws, int64_t n) {
  unsigned char* c = reinterpret_cast<unsigned char*>(ws);
  PcgWs w;
  size_t vec = align256((size_t)n * sizeof(float));
  w.r = reinterpret_cast<float*>(c); c += vec;
  w.p = reinterpret_cast<float*>(c); c += vec;
  w.ap = reinterpret_cast<float*>(c); c += vec;
  size_t part = align256(kGrid * sizeof(double));
  w.rz0 = reinterpret_cast<double*>(c); c += part;
  w.rz1 = reinterpret_cast<double*>(c); c += part;
  w.pap = reinterpret_cast<double*>(c); c += part;
  w.rr = reinterpret_cast<double*>(c); c += part;
  w.bb = reinterpret_cast<double*>(c); c += part;
  return w;
}

}  // namespace

extern "C" {

int nksr_spmv(const int64_t* rowptr, const int32_t* col, const float* val, const float* x, float* y, int64_t n,
              void* stream) {
  if (n <= 0) return n == 0 ? NKSR_OK : NKSR_E_INVALID;
  int grid = (int)((n + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (grid > kGrid) grid = kGrid;
  k_spmv<false><<<grid, kBlock, 0, as_stream(stream)>>>(rowptr, col, val, x, y, n, nullptr);
  NKSR_CHECK_LAUNCH();
  return NKSR_OK;
}

size_t nksr_pcg_workspace_bytes(int64_t n) {
  return 3 * align256((size_t)(n > 0 ? n : 1) * sizeof(float)) + 5 * align256(kGrid * sizeof(double)) + 256;
}

int nksr_pcg_solve(const int64_t* rowptr, const int32_t* col, const float* val, const float* diag, const float* b,
                   float* x, int64_t n, float tol, int max_iter, int check_every, int profile, void* ws,
                   size_t ws_bytes, double* info, void* stream) {
  if (n <= 0 || !info || max_iter < 0) return NKSR_E_INVALID;
  if (ws_bytes < nksr_pcg_workspace_bytes(n)) return NKSR_E_WORKSPACE;
  if (check_every < 1) check_every = 1;
  cudaStream_t s = as_stream(stream);
  PcgWs w = carve(ws, n);
  double host_part[kGrid];
  // optional profiling: CUDA events around every SpMV launch on this stream (info[2], info[3])
  const int kMaxEv = 512;
  cudaEvent_t ev[2 * kMaxEv];
  int n_ev = 0;
  if (profile)
    for (int i = 0; i < 2 * kMaxEv; ++i) cudaEventCreate(&ev[i]);
  // all partial arrays start at zero (blocks beyond a short grid never write)
  if (cudaMemsetAsync(w.rz0, 0, 5 * align256(kGrid * sizeof(double)), s) != cudaSuccess) return NKSR_E_CUDA;
  k_pcg_init<<<kGrid, kBlock, 0, s>>>(b, diag, x, w.r, w.p, n, w.rz0, w.bb);
  NKSR_CHECK_LAUNCH();
  if (cudaMemcpyAsync(host_part, w.bb, kGrid * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess)
    return NKSR_E_CUDA;
  if (cudaStreamSynchronize(s) != cudaSuccess) return NKSR_E_CUDA;
  double bb = 0.0;
  for (int i = 0; i < kGrid; ++i) bb += host_part[i];
  info[0] = 0.0;
  info[1] = 0.0;
  if (profile) { info[2] = 0.0; info[3] = 0.0; }
  if (!(bb > 0.0)) {
    if (profile) for (int i = 0; i < 2 * kMaxEv; ++i) cudaEventDestroy(ev[i]);
    return NKSR_OK;
  }
  const double target = (double)tol * (double)tol * bb;
  double* rz_cur = w.rz0;
  double* rz_new = w.rz1;
  int it = 0;
  double rr = bb;
  while (it < max_iter) {
    const bool timed = profile && n_ev < kMaxEv;
    if (timed) cudaEventRecord(ev[2 * n_ev], s);
    k_spmv<true><<<kGrid, kBlock, 0, s>>>(rowptr, col, val, w.p, w.ap, n, w.pap);
    if (timed) { cudaEventRecord(ev[2 * n_ev + 1], s); ++n_ev; }
    k_pcg_update<<<kGrid, kBlock, 0, s>>>(diag, w.p, w.ap, x, w.r, n, rz_cur, w.pap, rz_new, w.rr);
    k_pcg_direction<<<kGrid, kBlock, 0, s>>>(w.ap, w.p, n, rz_cur, rz_new);
    double* t = rz_cur; rz_cur = rz_new; rz_new = t;
    ++it;
    if (it % check_every == 0 || it == max_iter) {
      if (cudaMemcpyAsync(host_part, w.rr, kGrid * sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess)
        return NKSR_E_CUDA;
      if (cudaStreamSynchronize(s) != cudaSuccess) return NKSR_E_CUDA;
      rr = 0.0;
      for (int i = 0; i < kGrid; ++i) rr += host_part[i];
      if (!(rr == rr)) break;  // NaN guard
      if (rr <= target) break;
    }
  }
  NKSR_CHECK_LAUNCH();
  info[0] = (double)it;
  info[1] = sqrt(rr / bb);
  if (profile) {
    cudaStreamSynchronize(s);
    double ms = 0.0;
    for (int i = 0; i < n_ev; ++i) {
      float t = 0.f;
      cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]);
      ms += t;
    }
    info[2] = ms;
    info[3] = (double)n_ev;
    for (int i = 0; i < 2 * kMaxEv; ++i) cudaEventDestroy(ev[i]);
  }
  return NKSR_OK;
}

}  // extern "C"
