// the sparse convolution of the NKSRNetwork encoder / U-Net (models/nksr_net.py:73-78 call it,
// configs/default/train.yaml:17-18 size it: unet.f_maps = 32) as ONE gather-GEMM kernel over the hierarchy's index tables:
//
//     y[i, :] = act( bias + res[i, :] + sum_k [idx[i,k] >= 0]  x[idx[i,k], :] . W[k] )          W[k]: Cin x Cout
//
//   * 3x3x3 sparse convolution on level l:     idx = nbr27[l]   (K = 27)
//   * stride-2 convolution level l -> l+1:     idx = child8[l+1] (K = 8: one weight per octant)
// The tables are the ones the Gram assembly already uses (csrc/svh.cu), so no hash lookups happen here.
//
// Tiling: a CTA owns 128 output voxels x TN output channels (TN = 64 or 32); per (k, 32-channel chunk) the 128 gathered
// input rows (128 B each, coalesced) and the 32 x TN slice of W[k] are staged in shared memory.  A (k, tile) pair whose
// 128 sources are all absent is skipped (borders of the hierarchy).
//   k_gather_gemm_f32 : fp32 FFMA, 8 x TN/16 outputs per thread (A tile transposed in smem: two LDS.128 + one LDS.128/64
//                       per 32 / 16 FMAs) -- bit-for-bit an fp32 sum, the parity kernel
//   k_gather_gemm_tf32: mma.sync.m16n8k8 TF32 (fp32 accumulate), one 16 x TN strip per warp, operands staged by a
//                       two-stage cp.async pipeline; inputs are rounded to TF32 (10-bit mantissa, cvt.rna), so results
//                       differ from fp32 by ~1e-3 relative
//   k_gather_gemm_tc  : tcgen05.mma kind::tf32, 128 x TN accumulator in TMEM (TN = 32 / 64 / 128), operands gathered by
//                       cp.async into SWIZZLE_128B tiles, three-stage mbarrier ring -- the fast kernel (see below)
#include "common.cuh"

namespace {

constexpr int kTM = 128;      // output rows per CTA
constexpr int kKC = 32;       // input channels per staged chunk
constexpr int kThreads = 256;

template <int TN>
__global__ void __launch_bounds__(kThreads, 2)
k_gather_gemm_f32(const float* __restrict__ x, const int32_t* __restrict__ idx, int64_t n_out, int K,
                  const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ res,
                  float* __restrict__ y, int Cin, int Cout, int relu) {
  constexpr int CN = TN / 16;                       // output channels per thread
  __shared__ __align__(16) float As[kKC][kTM + 4];  // transposed: [channel][row]
  __shared__ __align__(16) float Bs[kKC][TN];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * kTM;
  const int n0 = blockIdx.y * TN;
  float acc[8][CN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;

  for (int k = 0; k < K; ++k) {
    // sources of this warp's 16 rows
    int src_l = -1;
    if (lane < 16) {
      const int64_t r = row0 + wid * 16 + lane;
      if (r < n_out) src_l = __ldg(idx + r * K + k);
    }
    if (!__syncthreads_or(src_l >= 0)) continue;    // nothing to gather for this offset in the whole tile (uniform)
    for (int c0 = 0; c0 < Cin; c0 += kKC) {
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const int s = __shfl_sync(0xffffffffu, src_l, j);
        As[lane][wid * 16 + j] = s >= 0 ? __ldg(x + (int64_t)s * Cin + c0 + lane) : 0.f;
      }
      const float* wp = W + ((int64_t)k * Cin + c0) * Cout + n0;
      for (int t = tid; t < kKC * TN / 4; t += kThreads) {
        const int r = t / (TN / 4), q = t % (TN / 4);
        *reinterpret_cast<float4*>(&Bs[r][q * 4]) = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)r * Cout) + q);
      }
      __syncthreads();
#pragma unroll 8
      for (int c = 0; c < kKC; ++c) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[c][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[c][ty * 8 + 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float b[CN];
        if (CN == 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&Bs[c][tx * 4]);
          b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
        } else {
          const float2 bv = *reinterpret_cast<const float2*>(&Bs[c][tx * 2]);
          b[0] = bv.x; b[1] = bv.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  // epilogue: bias, residual, activation
  float bv[CN];
#pragma unroll
  for (int j = 0; j < CN; ++j) bv[j] = bias ? __ldg(bias + n0 + tx * CN + j) : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = row0 + ty * 8 + i;
    if (r >= n_out) continue;
    float* yp = y + r * Cout + n0 + tx * CN;
    const float* rp = res ? res + r * Cout + n0 + tx * CN : nullptr;
#pragma unroll
    for (int j = 0; j < CN; ++j) {
      float v = acc[i][j] + bv[j];
      if (rp) v += __ldg(rp + j);
      if (relu) v = fmaxf(v, 0.f);
      yp[j] = v;
    }
  }
}

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col).  Fragments (g = lane / 4, t = lane % 4):
//   a0 = A[g][t]  a1 = A[g+8][t]  a2 = A[g][t+4]  a3 = A[g+8][t+4];   b0 = B[t][g]  b1 = B[t+4][g]
//   d0 = D[g][2t] d1 = D[g][2t+1] d2 = D[g+8][2t] d3 = D[g+8][2t+1]
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, const int src_bytes) {
  // 16-byte asynchronous copy global -> shared (LDGSTS); src_bytes = 0 fills the destination with zeros
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Two-stage pipeline: while the MMAs of one (offset k, 32-channel chunk) step run, the gathered rows and the weight
// slice of the next step are already on their way into the other shared-memory buffer (cp.async, zero-fill for absent
// sources).  The tile's 128 x K source indices are staged once (they are one contiguous block of `idx`), which also
// tells which offsets have no source in the whole tile -- those steps are skipped.
// WROUNDED: the weights were rounded to TF32 by the caller (no conversion of the B fragments here).
template <int TN, bool WROUNDED>
__global__ void __launch_bounds__(kThreads, 2)
k_gather_gemm_tf32(const float* __restrict__ x, const int32_t* __restrict__ idx, int64_t n_out, int K,
                   const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ res,
                   float* __restrict__ y, int Cin, int Cout, int relu) {
  constexpr int NT = TN / 8;                  // 8-column mma tiles per warp strip
  constexpr int AS = kKC + 4;                 // A row stride (floats): fragment loads hit 32 distinct banks
  constexpr int BS = TN + 8;                  // B row stride = 8 mod 32: b0 / b1 loads hit 32 distinct banks
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* As = reinterpret_cast<float*>(smem_raw);                    // [2][kTM][AS]
  float* Bs = As + 2 * kTM * AS;                                     // [2][kKC][BS]
  int32_t* src_s = reinterpret_cast<int32_t*>(Bs + 2 * kKC * BS);    // [kTM][K]
  __shared__ unsigned kmask;                                         // bit k: some row of the tile has a source at k
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t row0 = (int64_t)blockIdx.x * kTM;
  const int n0 = blockIdx.y * TN;
  const int rows_here = (int)min((int64_t)kTM, n_out - row0);
  if (tid == 0) kmask = 0u;
  __syncthreads();
  {
    unsigned mine = 0u;
    const int32_t* ip = idx + row0 * K;
    for (int e = tid; e < kTM * K; e += kThreads) {
      const int r = e / K;
      const int v = r < rows_here ? __ldg(ip + e) : -1;
      src_s[e] = v;
      if (v >= 0) mine |= 1u << (e - r * K);
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicOr(&kmask, mine);
  }
  __syncthreads();
  const unsigned km = kmask;
  const int nchunk = Cin / kKC;
  const int nsteps = __popc(km) * nchunk;

  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

  // step s -> (k = the (s / nchunk)-th set bit of km, c0 = (s % nchunk) * 32)
  auto issue = [&](const int s, const int buf) {
    const int k = __fns(km, 0, s / nchunk + 1);
    const int c0 = (s - (s / nchunk) * nchunk) * kKC;
    float* a = As + buf * kTM * AS;
#pragma unroll
    for (int j = 0; j < kTM * 8 / kThreads; ++j) {
      const int e = tid + j * kThreads;
      const int r = e >> 3, seg = e & 7;
      const int sidx = src_s[r * K + k];
      const float* gp = x + (int64_t)(sidx >= 0 ? sidx : 0) * Cin + c0 + seg * 4;
      cp_async16(a + r * AS + seg * 4, gp, sidx >= 0 ? 16 : 0);
    }
    float* b = Bs + buf * kKC * BS;
    const float* wp = W + ((int64_t)k * Cin + c0) * Cout + n0;
    for (int e = tid; e < kKC * TN / 4; e += kThreads) {
      const int r = e / (TN / 4), q = e % (TN / 4);
      cp_async16(b + r * BS + q * 4, wp + (int64_t)r * Cout + q * 4, 16);
    }
    cp_async_commit();
  };

  if (nsteps > 0) issue(0, 0);
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s & 1;
    if (s + 1 < nsteps) {
      issue(s + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* Aw = As + buf * kTM * AS + wid * 16 * AS;      // this warp's 16 rows
    const float* Bb = Bs + buf * kKC * BS;
#pragma unroll
    for (int ks = 0; ks < kKC; ks += 8) {
      uint32_t a[4];
      a[0] = to_tf32(Aw[g * AS + ks + t]);
      a[1] = to_tf32(Aw[(g + 8) * AS + ks + t]);
      a[2] = to_tf32(Aw[g * AS + ks + t + 4]);
      a[3] = to_tf32(Aw[(g + 8) * AS + ks + t + 4]);
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const float b0 = Bb[(ks + t) * BS + j * 8 + g], b1 = Bb[(ks + t + 4) * BS + j * 8 + g];
        mma_tf32(acc[j], a, WROUNDED ? __float_as_uint(b0) : to_tf32(b0), WROUNDED ? __float_as_uint(b1) : to_tf32(b1));
      }
    }
    __syncthreads();                                            // the buffer is refilled two steps ahead
  }
  // epilogue: rows g / g+8 of the warp's strip, columns 2t / 2t+1 of every 8-column tile
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t r = row0 + wid * 16 + g + 8 * h;
    if (r >= n_out) continue;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int cidx = n0 + j * 8 + 2 * t;
      float v0 = acc[j][2 * h], v1 = acc[j][2 * h + 1];
      if (bias) { v0 += __ldg(bias + cidx); v1 += __ldg(bias + cidx + 1); }
      if (res) {
        const float2 rv = __ldg(reinterpret_cast<const float2*>(res + r * Cout + cidx));
        v0 += rv.x; v1 += rv.y;
      }
      if (relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
      *reinterpret_cast<float2*>(y + r * Cout + cidx) = make_float2(v0, v1);
    }
  }
}

template <int TN>
size_t tf32_smem_bytes(int K) {
  return (size_t)(2 * kTM * (kKC + 4) + 2 * kKC * (TN + 8)) * sizeof(float) + (size_t)kTM * K * sizeof(int32_t);
}

template <int TN, bool WROUNDED>
int launch_tf32(dim3 grid, cudaStream_t s, const float* x, const int32_t* idx, int64_t n_out, int K, const float* W,
                const float* bias, const float* res, float* y, int c_in, int c_out, int relu) {
  const size_t smem = tf32_smem_bytes<TN>(K);
  if (smem > 200 * 1024) return NKSR_E_INVALID;
  if (cudaFuncSetAttribute(k_gather_gemm_tf32<TN, WROUNDED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
      cudaSuccess)
    return NKSR_E_CUDA;
  k_gather_gemm_tf32<TN, WROUNDED><<<grid, kThreads, smem, s>>>(x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu);
  return NKSR_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// k_gather_gemm_tc: the same gather-GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulator in TMEM).
//
//   * A tile (128 gathered rows x 32 channels = 128 B per row) and B tile (TN output channels x 32 input channels; the
//     caller passes W transposed to [K][c_out][c_in], "K-major" for the MMA) are written by cp.async straight into the
//     canonical K-major SWIZZLE_128B shared-memory layout: row r at r * 128 B, its 16-byte chunk c at position
//     c ^ (r & 7); 8-row groups 1024 B apart (the descriptor's stride byte offset).  Absent sources are zero-filled.
//   * one thread issues 4 x tcgen05.mma (M = 128, N = TN, K = 8) per (offset, 32-channel chunk) step; the step's
//     tcgen05.commit arrives on the stage's mbarrier, which is what lets the gather refill that stage: a three-stage
//     ring, two steps of gathers in flight under the MMAs.
//   * the accumulator (128 lanes x TN fp32 columns of TMEM) is read once, at the end: tcgen05.ld 32x32b (one row of 32
//     columns per thread), bias / residual / ReLU in registers, 128-bit stores.
// The operands are fp32 in shared memory; the tensor core reads their upper 19 bits (TF32).  W is rounded by the caller.
constexpr int kTcStages = 3;
constexpr int kTcATile = kTM * kKC * 4;        // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major SWIZZLE_128B (PTX ISA "tcgen05 matrix descriptor"): start address >> 4 in
// bits [0,14), leading byte offset (unused for swizzled K-major; 1) in [16,30), stride byte offset 1024 >> 4 in [32,46),
// version 1 in [46,48), base offset 0 (tiles are 1024-byte aligned), layout type 2 = SWIZZLE_128B in [61,64)
__device__ __forceinline__ uint64_t tc_smem_desc(const uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}

// instruction descriptor of kind::tf32: D fp32 (bits [4,6) = 1), A and B TF32 ([7,10) = [10,13) = 2), both K-major
// (bits 15, 16 = 0), N >> 3 in [17,23), M >> 4 in [24,29)
template <int TN>
__device__ __forceinline__ constexpr uint32_t tc_instr_desc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(kTM >> 4) << 24);
}

__device__ __forceinline__ void tc_mma_tf32(const uint32_t tmem_d, const uint64_t adesc, const uint64_t bdesc,
                                            const uint32_t idesc, const uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait_or_trap(const uint32_t bar, const uint32_t parity) {
  // bounded: a commit that never arrives must end as a launch failure, not as a hung GPU
  for (int i = 0; i < (1 << 24); ++i) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return;
  }
  __trap();
}

template <int TN>
__global__ void __launch_bounds__(kThreads, 2)
k_gather_gemm_tc(const float* __restrict__ x, const int32_t* __restrict__ idx, int64_t n_out, int K,
                 const float* __restrict__ Wt, const float* __restrict__ bias, const float* __restrict__ res,
                 float* __restrict__ y, int Cin, int Cout, int relu) {
  constexpr int NS = kTcStages;
  constexpr int kBTile = TN * kKC * 4;
  extern __shared__ unsigned char smem_dyn[];
  // SWIZZLE_128B atoms repeat every 1024 bytes and the descriptors carry base offset 0: align the tiles by hand
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t a_s = base;                                  // [NS][128 rows][128 B]
  const uint32_t b_s = base + NS * kTcATile;                  // [NS][TN rows][128 B]
  int32_t* src_s = reinterpret_cast<int32_t*>(sm + NS * (kTcATile + kBTile));   // [128][K]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + NS * (kTcATile + kBTile) + ((kTM * K * 4 + 15) & ~15));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NS);
  unsigned* kmask_s = reinterpret_cast<unsigned*>(tmem_slot + 1);

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int64_t row0 = (int64_t)blockIdx.x * kTM;
  const int n0 = blockIdx.y * TN;
  const int rows_here = (int)min((int64_t)kTM, n_out - row0);

  if (tid == 0) {
    *kmask_s = 0u;
    for (int i = 0; i < NS; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wid == 0) {                              // one warp allocates the accumulator's TMEM columns (and frees them)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot;
  {
    unsigned mine = 0u;
    const int32_t* ip = idx + row0 * K;
    for (int e = tid; e < kTM * K; e += kThreads) {
      const int r = e / K;
      const int v = r < rows_here ? __ldg(ip + e) : -1;
      src_s[e] = v;
      if (v >= 0) mine |= 1u << (e - r * K);
    }
    mine = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0 && mine) atomicOr(kmask_s, mine);
  }
  __syncthreads();
  const unsigned km = *kmask_s;
  const int nchunk = Cin / kKC;
  const int nsteps = __popc(km) * nchunk;

  auto issue = [&](const int s) {
    const int buf = s % NS;
    const int k = __fns(km, 0, s / nchunk + 1);
    const int c0 = (s - (s / nchunk) * nchunk) * kKC;
    const uint32_t a = a_s + buf * kTcATile;
#pragma unroll
    for (int j = 0; j < kTM * 8 / kThreads; ++j) {
      const int e = tid + j * kThreads;
      const int r = e >> 3, seg = e & 7;
      const int sidx = src_s[r * K + k];
      const float* gp = x + (int64_t)(sidx >= 0 ? sidx : 0) * Cin + c0 + seg * 4;
      const uint32_t d = a + r * 128 + ((seg ^ (r & 7)) << 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gp), "r"(sidx >= 0 ? 16 : 0) : "memory");
    }
    const uint32_t b = b_s + buf * kBTile;
    const float* wp = Wt + ((int64_t)k * Cout + n0) * Cin + c0;
#pragma unroll
    for (int j = 0; j < TN * 8 / kThreads; ++j) {
      const int e = tid + j * kThreads;
      const int r = e >> 3, seg = e & 7;
      const uint32_t d = b + r * 128 + ((seg ^ (r & 7)) << 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, 16;" ::"r"(d), "l"(wp + (int64_t)r * Cin + seg * 4) : "memory");
    }
  };

  for (int p = 0; p < NS - 1; ++p) {
    if (p < nsteps) issue(p);
    cp_async_commit();
  }
  constexpr uint32_t idesc = tc_instr_desc<TN>();
  for (int s = 0; s < nsteps; ++s) {
    const int pf = s + NS - 1;                 // its stage was read by the MMAs of step s - 1
    if (pf < nsteps) {
      if (s >= 1) mbar_wait_or_trap(smem_u32(bars + (s - 1) % NS), ((s - 1) / NS) & 1);
      issue(pf);
    }
    cp_async_commit();
    cp_async_wait<NS - 1>();                   // this thread's copies of step s have landed ...
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ... and are visible to the tensor core's proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int buf = s % NS;
      const uint64_t ad = tc_smem_desc(a_s + buf * kTcATile), bd = tc_smem_desc(b_s + buf * kBTile);
#pragma unroll
      for (int kk = 0; kk < kKC / 8; ++kk)     // 8 TF32 = 32 bytes along K inside the 128-byte swizzle row
        tc_mma_tf32(tmem_d, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, (s | kk) != 0);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(bars + buf))
                   : "memory");
    }
  }
  cp_async_wait<0>();
  // epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 (its sub-partition); warps 0-3 take the first half of the
  // columns, warps 4-7 the second (TN = 32: warps 0-3 take all), 32 columns = one 128-byte row segment at a time
  constexpr int kColsPerWarp = TN >= 64 ? TN / 2 : 32;
  const int colw = (wid >> 2) * kColsPerWarp;
  if (nsteps > 0) {
    mbar_wait_or_trap(smem_u32(bars + (nsteps - 1) % NS), ((nsteps - 1) / NS) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  const int64_t r = row0 + (wid & 3) * 32 + lane;
  if (colw < TN) {
#pragma unroll 1
    for (int cc = 0; cc < kColsPerWarp; cc += 32) {
      const int col0 = colw + cc;
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 0.f;
      if (nsteps > 0) {
        uint32_t u[32];
        const uint32_t taddr = tmem_d + ((uint32_t)((wid & 3) * 32) << 16) + (uint32_t)col0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
              "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
              "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
              "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u[i]);
      }
      if (r < n_out) {
        float* yp = y + r * Cout + n0 + col0;
        const float* rp = res ? res + r * Cout + n0 + col0 : nullptr;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          if (bias) {
            const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + n0 + col0) + q);
            o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
          }
          if (rp) {
            const float4 rq = __ldg(reinterpret_cast<const float4*>(rp) + q);
            o.x += rq.x; o.y += rq.y; o.z += rq.z; o.w += rq.w;
          }
          if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          reinterpret_cast<float4*>(yp)[q] = o;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wid == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TN) : "memory");
}

template <int TN>
int launch_tc(dim3 grid, cudaStream_t s, const float* x, const int32_t* idx, int64_t n_out, int K, const float* Wt,
              const float* bias, const float* res, float* y, int c_in, int c_out, int relu) {
  const size_t smem = 1024 + (size_t)kTcStages * (kTcATile + TN * kKC * 4) + (((size_t)kTM * K * 4 + 15) & ~(size_t)15) +
                      kTcStages * 8 + 16;
  if (smem > 113 * 1024) return NKSR_E_INVALID;     // two CTAs per SM
  if (cudaFuncSetAttribute(k_gather_gemm_tc<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return NKSR_E_CUDA;
  k_gather_gemm_tc<TN><<<grid, kThreads, smem, s>>>(x, idx, n_out, K, Wt, bias, res, y, c_in, c_out, relu);
  return NKSR_OK;
}

}  // namespace

extern "C" {

int nksr_gather_gemm(const float* x, const int32_t* idx, int64_t n_out, int K, const float* W, const float* bias,
                     const float* res, float* y, int c_in, int c_out, int relu, int tf32, void* stream) {
  if (tf32 < 0 || tf32 > 3) return NKSR_E_INVALID;
  if (n_out < 0 || K < 1 || c_in < kKC || c_in % kKC != 0 || c_out < 32 || c_out % 32 != 0) return NKSR_E_INVALID;
  if (n_out == 0) return NKSR_OK;
  if (!x || !idx || !W || !y) return NKSR_E_INVALID;
  cudaStream_t s = as_stream(stream);
  const int tn = c_out % 64 == 0 ? 64 : 32;
  const dim3 grid((unsigned)((n_out + kTM - 1) / kTM), (unsigned)(c_out / tn));
  if (K > 32 && tf32) return NKSR_E_INVALID;           // the tile's offset mask is one 32-bit word
  if (tf32 == 3) {                                     // tcgen05: W is [K][c_out][c_in], rounded to TF32
    // 128 output channels per CTA where the layer is that wide: the gathered rows are fetched once per 128 columns
    // (r2y: 100 K rows, 128 -> 128 channels: 0.36 ms against 0.60 ms with 64-column tiles, 0.89 ms for mma.sync)
    int rc;
    if (c_out % 128 == 0) {
      const dim3 g128(grid.x, (unsigned)(c_out / 128));
      rc = launch_tc<128>(g128, s, x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu);
    } else {
      rc = tn == 64 ? launch_tc<64>(grid, s, x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu)
                    : launch_tc<32>(grid, s, x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu);
    }
    if (rc != NKSR_OK) return rc;
  } else if (tf32) {
    int rc;
    if (tn == 64)
      rc = tf32 == 2 ? launch_tf32<64, true>(grid, s, x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu)
                     : launch_tf32<64, false>(grid, s, x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu);
    else
      rc = tf32 == 2 ? launch_tf32<32, true>(grid, s, x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu)
                     : launch_tf32<32, false>(grid, s, x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu);
    if (rc != NKSR_OK) return rc;
  } else {
    if (tn == 64)
      k_gather_gemm_f32<64><<<grid, kThreads, 0, s>>>(x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu);
    else
      k_gather_gemm_f32<32><<<grid, kThreads, 0, s>>>(x, idx, n_out, K, W, bias, res, y, c_in, c_out, relu);
  }
  NKSR_CHECK_LAUNCH();
  return NKSR_OK;
}

}  // extern "C"
