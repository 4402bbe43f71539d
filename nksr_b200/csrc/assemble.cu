// Gram-matrix assembly A = E^T diag(w) E + reg*R into CSR.
// Replaces the matrix build inside KernelField.solve_non_fused / fused_mode
// (models/nksr_net.py:100-112, examples/recons_waymo.py:33).
//
// Layout (DESIGN.md SPEC S6): unknowns are ordered level-major, Morton inside a level.  Row
// (l,i) stores, in this order, [same-level 125-stencil | coarser level l+1 (<=64) | ... | level
// L-1 | finer-level entries (transposes)].  Only ACTIVE column voxels are stored.
//
// Numeric phase: one warp per row.  For each of the 27 voxels u around i, the constraint rows
// whose containing voxel is u form one contiguous range (locations are Morton sorted); every
// such row r contributes  w * E[r,i] * E[r, :]  and all rows of u share the same 27-stencils on
// level l and on every coarser level, so lane s accumulates stencil slot s in a register and
// the warp flushes once per u into a per-warp shared-memory tile indexed by structural slot --
// no atomics, deterministic summation order.  The round-1 ncu capture showed the kernel to be
// instruction-issue bound in the gradient-row loop (profiles/r1_*), hence: neighbour indices
// and row ranges are fetched lane-parallel once per row, row pointers advance by one add per
// level, and with approx_kernel_grad the three gradient rows are rebuilt from ONE 128-byte
// line (<phi,z_s> + tau) with nine FMAs instead of being loaded.
#include <cstdlib>

#include <cub/cub.cuh>

#include "common.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxSlots = 125 + 64 * (NKSR_MAX_DEPTH - 1);

__constant__ signed char c_d27[27][3] = {
    {-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0}, {-1, 0, 1}, {-1, 1, -1}, {-1, 1, 0}, {-1, 1, 1},
    {0, -1, -1},  {0, -1, 0},  {0, -1, 1},  {0, 0, -1},  {0, 0, 0},  {0, 0, 1},  {0, 1, -1},  {0, 1, 0},  {0, 1, 1},
    {1, -1, -1},  {1, -1, 0},  {1, -1, 1},  {1, 0, -1},  {1, 0, 0},  {1, 0, 1},  {1, 1, -1},  {1, 1, 0},  {1, 1, 1}};

// same-level voxel at offset-space coords (nx,ny,nz) in the 125-neighbourhood of voxel i
// (coords ux,uy,uz): through the parent's 27-stencil and its child table; the top level owns an
// explicit 125-neighbour table.
__device__ __forceinline__ int lookup_near(const nksr_svh_t& svh, int l, int i, int ux, int uy, int uz, int nx,
                                           int ny, int nz) {
  if (svh.parent[l] != nullptr) {  // also true for the top level when the virtual level exists
    const int p = __ldg(svh.parent[l] + i);
    if (p < 0) return -1;
    const int ex = (nx >> 1) - (ux >> 1), ey = (ny >> 1) - (uy >> 1), ez = (nz >> 1) - (uz >> 1);
    const int pn = __ldg(svh.nbr27[l + 1] + (int64_t)p * 27 + (ex + 1) * 9 + (ey + 1) * 3 + (ez + 1));
    if (pn < 0) return -1;
    return __ldg(svh.child8[l + 1] + (int64_t)pn * 8 + (((nx & 1) << 2) | ((ny & 1) << 1) | (nz & 1)));
  }
  return __ldg(svh.nbr125_top + (int64_t)i * 125 + (nx - ux + 2) * 25 + (ny - uy + 2) * 5 + (nz - uz + 2));
}

struct RowGeom {
  int ux, uy, uz;           // offset-space coords of the row voxel
  int anc[NKSR_MAX_DEPTH];  // ancestor index at level l+k (anc[0] = i)
};

__device__ __forceinline__ void row_geom(const nksr_svh_t& svh, int l, int i, RowGeom& g) {
  morton3_decode(__ldg(svh.keys[l] + i), g.ux, g.uy, g.uz);
  g.anc[0] = i;
  int a = i;
#pragma unroll
  for (int k = 1; k < NKSR_MAX_DEPTH; ++k) {
    if (l + k < svh.depth) a = a >= 0 ? __ldg(svh.parent[l + k - 1] + a) : -1;
    g.anc[k] = a;
  }
}

// column voxel (index at its level) of structural slot t of row (l,i); -1 when inactive.
// t < 125: same level; else k = 1 + (t-125)/64 levels up, 4x4x4 candidate box from lo.
__device__ __forceinline__ int slot_column(const nksr_svh_t& svh, int l, const RowGeom& g, int t, int& k_out) {
  if (t < 125) {
    k_out = 0;
    const int dx = t / 25 - 2, dy = (t / 5) % 5 - 2, dz = t % 5 - 2;
    return lookup_near(svh, l, g.anc[0], g.ux, g.uy, g.uz, g.ux + dx, g.uy + dy, g.uz + dz);
  }
  int q = t - 125;
  const int k = 1 + (q >> 6);
  k_out = k;
  q &= 63;
  const int ox = q >> 4, oy = (q >> 2) & 3, oz = q & 3;
  const int cx = (((g.ux - 1) >> k) - 1) + ox, cy = (((g.uy - 1) >> k) - 1) + oy, cz = (((g.uz - 1) >> k) - 1) + oz;
  if (cx > ((g.ux + 1) >> k) + 1 || cy > ((g.uy + 1) >> k) + 1 || cz > ((g.uz + 1) >> k) + 1) return -1;
  int a = g.anc[0];
#pragma unroll
  for (int j = 1; j < NKSR_MAX_DEPTH; ++j)
    if (j == k) a = g.anc[j];
  if (a < 0) return -1;
  return lookup_near(svh, l + k, a, g.ux >> k, g.uy >> k, g.uz >> k, cx, cy, cz);
}

__device__ __forceinline__ void row_of_warp(const nksr_svh_t& svh, int64_t row, int& l, int& i) {
  l = 0;
  while (l + 1 < svh.depth && row >= svh.offset[l + 1]) ++l;
  i = (int)(row - svh.offset[l]);
}

__global__ void __launch_bounds__(kWarps * 32)
k_gram_count(nksr_svh_t svh, int64_t n_total, int32_t* __restrict__ cnt, int32_t* __restrict__ cnt_down) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5);
  if (row >= n_total) return;
  int l, i;
  row_of_warp(svh, row, l, i);
  RowGeom g;
  row_geom(svh, l, i, g);
  const int nslots = 125 + 64 * (svh.depth - 1 - l);
  int c = 0;
  for (int t0 = 0; t0 < nslots; t0 += 32) {
    int t = t0 + lane, k = 0;
    int col = t < nslots ? slot_column(svh, l, g, t, k) : -1;
    if (col >= 0 && k > 0) atomicAdd(cnt_down + svh.offset[l + k] + col, 1);
    c += __popc(__ballot_sync(0xffffffffu, col >= 0));
  }
  if (lane == 0) cnt[row] = c;
}

struct AddPair {
  const int32_t* a;
  const int32_t* b;
  __host__ __device__ __forceinline__ int64_t operator()(const int64_t& i) const { return (int64_t)a[i] + b[i]; }
};

__global__ void k_set_last(const int32_t* cnt, const int32_t* cnt_down, int64_t n, int64_t* rowptr) {
  if (threadIdx.x == 0 && blockIdx.x == 0) rowptr[n] = rowptr[n - 1] + cnt[n - 1] + cnt_down[n - 1];
}

template <bool COMPACT, int MAXL>
__global__ void __launch_bounds__(kWarps * 32, MAXL <= 4 ? 4 : 2)
k_gram_fill(nksr_svh_t svh, nksr_feat_t feat, nksr_constraints_t cs, int64_t n_total,
            const int32_t* __restrict__ cnt, const int64_t* __restrict__ rowptr, int32_t* __restrict__ col_out,
            float* __restrict__ val_out, float* __restrict__ rhs, float* __restrict__ diag,
            int32_t* __restrict__ cursor) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int64_t row = blockIdx.x * (int64_t)kWarps + wid;
  if (row >= n_total) return;
  int l, i;
  row_of_warp(svh, row, l, i);
  const int L = svh.depth;
  const int nup = L - 1 - l;
  const int nslots = 125 + 64 * nup;
  float* acc = smem + wid * kMaxSlots;
  for (int t = lane; t < nslots; t += 32) acc[t] = 0.f;
  RowGeom g;
  row_geom(svh, l, i, g);
  const int64_t N = cs.n_pos, K = cs.n_nrm;
  const int32_t* rp = cs.range_pos ? cs.range_pos + 2 * svh.offset[l] : nullptr;
  const int32_t* rn = cs.range_nrm ? cs.range_nrm + 2 * svh.offset[l] : nullptr;
  float bsum = 0.f;
  const int sl = lane < 27 ? lane : 13;
  const int ldx = c_d27[sl][0], ldy = c_d27[sl][1], ldz = c_d27[sl][2];
  // lane-parallel prefetch of the 27 neighbour voxels and their constraint-row ranges
  const int my_u = lane < 27 ? __ldg(svh.nbr27[l] + (int64_t)i * 27 + lane) : -1;
  int my_pb = 0, my_pe = 0, my_nb = 0, my_ne = 0;
  if (my_u >= 0) {
    if (rp) { my_pb = __ldg(rp + 2 * (int64_t)my_u); my_pe = __ldg(rp + 2 * (int64_t)my_u + 1); }
    if (rn) { my_nb = __ldg(rn + 2 * (int64_t)my_u); my_ne = __ldg(rn + 2 * (int64_t)my_u + 1); }
  }
  // quadratic B-spline as polynomials in tau for this lane's offset d (SPEC S4):
  // b = c0 + tau (c1 + c2 tau), db = c1 + 2 c2 tau
  const float cx0 = ldx == 0 ? 0.75f : 0.125f, cx1 = 0.5f * (float)ldx, cx2 = ldx == 0 ? -1.f : 0.5f;
  const float cy0 = ldy == 0 ? 0.75f : 0.125f, cy1 = 0.5f * (float)ldy, cy2 = ldy == 0 ? -1.f : 0.5f;
  const float cz0 = ldz == 0 ? 0.75f : 0.125f, cz1 = 0.5f * (float)ldz, cz2 = ldz == 0 ? -1.f : 0.5f;
  const float inv_wl = 1.f / (svh.voxel_size * (float)(1 << l));
  const int64_t pos_level = N * NKSR_ROW_STRIDE;
  const int64_t nrm_level = K * NKSR_ROW_STRIDE * (COMPACT ? 1 : 3);
  __syncwarp();

  for (int us = 0; us < 27; ++us) {
    const int u = __shfl_sync(0xffffffffu, my_u, us);
    if (u < 0) continue;
    const int pb = __shfl_sync(0xffffffffu, my_pb, us), pe = __shfl_sync(0xffffffffu, my_pe, us);
    const int nb = __shfl_sync(0xffffffffu, my_nb, us), ne = __shfl_sync(0xffffffffu, my_ne, us);
    if (pb == pe && nb == ne) continue;
    const int si = 26 - us;  // slot of i inside u's stencil
    float r[MAXL];
#pragma unroll
    for (int k = 0; k < MAXL; ++k) r[k] = 0.f;
    for (int q = pb; q < pe; ++q) {
      const float* p0 = cs.e_pos + ((int64_t)l * N + q) * NKSR_ROW_STRIDE;
      const float a = cs.w_pos * __ldg(p0 + si);
      const float* pk = p0 + lane;
#pragma unroll
      for (int k = 0; k < MAXL; ++k)
        if (k <= nup) { r[k] = fmaf(a, __ldg(pk), r[k]); pk += pos_level; }
    }
    if (COMPACT) {
      // one line per (location, level): <phi,z_s> in slots 0..26, tau in 27..29;
      // E_a[s] = dB_a B_b B_c <phi,z_s> / W_level
      for (int q = nb; q < ne; ++q) {
        const float* pk = cs.e_nrm + ((int64_t)l * K + q) * NKSR_ROW_STRIDE + lane;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, iw = inv_wl;
#pragma unroll
        for (int k = 0; k < MAXL; ++k) {
          if (k <= nup) {
            const float line = __ldg(pk);
            pk += nrm_level;
            const float tx = __shfl_sync(0xffffffffu, line, 27), ty = __shfl_sync(0xffffffffu, line, 28),
                        tz = __shfl_sync(0xffffffffu, line, 29);
            const float bx = fmaf(fmaf(cx2, tx, cx1), tx, cx0), dbx = fmaf(2.f * cx2, tx, cx1);
            const float by = fmaf(fmaf(cy2, ty, cy1), ty, cy0), dby = fmaf(2.f * cy2, ty, cy1);
            const float bz = fmaf(fmaf(cz2, tz, cz1), tz, cz0), dbz = fmaf(2.f * cz2, tz, cz1);
            const float sc = (lane < 27 ? line : 0.f) * iw;
            const float e0 = dbx * by * bz * sc, e1 = bx * dby * bz * sc, e2 = bx * by * dbz * sc;
            if (k == 0) {
              a0 = cs.w_nrm * __shfl_sync(0xffffffffu, e0, si);
              a1 = cs.w_nrm * __shfl_sync(0xffffffffu, e1, si);
              a2 = cs.w_nrm * __shfl_sync(0xffffffffu, e2, si);
              const float* t = cs.t_nrm + (int64_t)q * 3;
              bsum = fmaf(a0, __ldg(t), fmaf(a1, __ldg(t + 1), fmaf(a2, __ldg(t + 2), bsum)));
            }
            r[k] = fmaf(a0, e0, fmaf(a1, e1, fmaf(a2, e2, r[k])));
            iw *= 0.5f;
          }
        }
      }
    } else {
      for (int q = nb; q < ne; ++q) {
        const float* p0 = cs.e_nrm + ((int64_t)l * K + q) * (3 * NKSR_ROW_STRIDE);
        const float* t = cs.t_nrm + (int64_t)q * 3;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          const float a = cs.w_nrm * __ldg(p0 + ax * NKSR_ROW_STRIDE + si);
          bsum = fmaf(a, __ldg(t + ax), bsum);
          const float* pk = p0 + ax * NKSR_ROW_STRIDE + lane;
#pragma unroll
          for (int k = 0; k < MAXL; ++k)
            if (k <= nup) { r[k] = fmaf(a, __ldg(pk), r[k]); pk += nrm_level; }
        }
      }
    }
    // flush: every lane < 27 owns a distinct structural slot per level
    if (lane < 27) {
      const int udx = c_d27[us][0], udy = c_d27[us][1], udz = c_d27[us][2];
      acc[(udx + ldx + 2) * 25 + (udy + ldy + 2) * 5 + (udz + ldz + 2)] += r[0];
      const int vx = g.ux + udx, vy = g.uy + udy, vz = g.uz + udz;  // coords of u
#pragma unroll
      for (int k = 1; k < MAXL; ++k) {
        if (k <= nup) {
          const int ox = ((vx >> k) + ldx) - (((g.ux - 1) >> k) - 1);
          const int oy = ((vy >> k) + ldy) - (((g.uy - 1) >> k) - 1);
          const int oz = ((vz >> k) + ldz) - (((g.uz - 1) >> k) - 1);
          acc[125 + 64 * (k - 1) + (ox << 4) + (oy << 2) + oz] += r[k];
        }
      }
    }
    __syncwarp();
  }
  // regulariser: R_{i,i+d} = w_reg * B3(d) * <z_i, z_{i+d}>  (SPEC S5)
  if (cs.w_reg != 0.f && my_u >= 0) {
    const int C = feat.channels;
    const float* zi = feat.z[l] + (int64_t)i * C;
    const float* zn = feat.z[l] + (int64_t)my_u * C;
    float d = 0.f;
    for (int c = 0; c < C; ++c) d = fmaf(__ldg(zi + c), __ldg(zn + c), d);
    const float bw = (ldx == 0 ? 0.75f : 0.125f) * (ldy == 0 ? 0.75f : 0.125f) * (ldz == 0 ? 0.75f : 0.125f);
    acc[(ldx + 2) * 25 + (ldy + 2) * 5 + (ldz + 2)] += cs.w_reg * bw * d;
  }
  __syncwarp();
  // write-out in structural order
  const int64_t p0 = rowptr[row];
  int written = 0;
  for (int t0 = 0; t0 < nslots; t0 += 32) {
    const int t = t0 + lane;
    int k = 0;
    const int c = t < nslots ? slot_column(svh, l, g, t, k) : -1;
    const unsigned m = __ballot_sync(0xffffffffu, c >= 0);
    if (c >= 0) {
      const int64_t p = p0 + written + __popc(m & ((1u << lane) - 1u));
      const float v = acc[t];
      const int64_t gc = svh.offset[l + k] + c;
      col_out[p] = (int32_t)gc;
      val_out[p] = v;
      if (k == 0 && c == i) diag[row] = v;
      if (k > 0) {  // transposed copy into the coarse row's finer-level segment
        const int64_t q = rowptr[gc] + cnt[gc] + atomicAdd(cursor + gc, 1);
        col_out[q] = (int32_t)row;
        val_out[q] = v;
      }
    }
    written += __popc(m);
  }
  if (lane == 0) rhs[row] = bsum;
}

// ---------------------------------------------------------------------------------------------
// Grouped numeric assembly (used whenever the hierarchy carries the virtual top level).
// One CTA per PARENT voxel p (level l+1): its <= 8 children are the rows, one warp each.  The 27-
// neighbourhoods of the children cover a 4x4x4 block of level-l voxels, so every constraint row
// in that block is staged ONCE in shared memory (all eight warps issue the loads: many bytes in
// flight, none of the per-row latency chains of the warp-per-row kernel) and, for compact
// gradient rows, expanded once into its three gradient rows; each row-warp then consumes the
// staged rows of the cells within its own 27-stencil with shared-memory loads only.
constexpr int kChunk = 8;

template <bool COMPACT, int MAXL>
__global__ void __launch_bounds__(kWarps * 32, MAXL <= 4 ? 3 : 2)
k_gram_fill_group(nksr_svh_t svh, nksr_feat_t feat, nksr_constraints_t cs, const int32_t* __restrict__ cnt,
                  const int64_t* __restrict__ rowptr, int32_t* __restrict__ col_out, float* __restrict__ val_out,
                  float* __restrict__ rhs, float* __restrict__ diag, int32_t* __restrict__ cursor) {
  constexpr int SLOTS = 125 + 64 * (MAXL - 1);
  constexpr int ESTRIDE = MAXL * 96;  // floats per staged entry: [level][3 rows][32]
  extern __shared__ float smem[];
  float* s_acc = smem;                   // [8][SLOTS]
  float* s_exp = smem + kWarps * SLOTS;  // [2][kChunk][MAXL][3][32]
  __shared__ int s_u[64], s_pb[64], s_np[64], s_nb[64], s_nn[64], s_off[65];
  __shared__ int s_mq[3][kChunk], s_mcell[3][kChunk], s_mtype[3][kChunk];
  __shared__ float s_t[2][kChunk][3];

  int64_t b = blockIdx.x;
  int l = 0;
  while (l + 1 < svh.depth && b >= svh.n[l + 1]) { b -= svh.n[l + 1]; ++l; }
  const int p = (int)b;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nup = svh.depth - 1 - l;
  const int64_t N = cs.n_pos, K = cs.n_nrm;
  const int32_t* rp = cs.range_pos ? cs.range_pos + 2 * svh.offset[l] : nullptr;
  const int32_t* rn = cs.range_nrm ? cs.range_nrm + 2 * svh.offset[l] : nullptr;
  const int i = __ldg(svh.child8[l + 1] + (int64_t)p * 8 + wid);  // this warp's row voxel, -1 = none

  if (tid < 64) {
    const int ax = tid >> 4, ay = (tid >> 2) & 3, az = tid & 3;
    const int ex = ax == 0 ? -1 : (ax == 3 ? 1 : 0), ey = ay == 0 ? -1 : (ay == 3 ? 1 : 0),
              ez = az == 0 ? -1 : (az == 3 ? 1 : 0);
    const int pn = __ldg(svh.nbr27[l + 1] + (int64_t)p * 27 + (ex + 1) * 9 + (ey + 1) * 3 + (ez + 1));
    int u = -1, pb = 0, np = 0, nb = 0, nn = 0;
    if (pn >= 0)
      u = __ldg(svh.child8[l + 1] + (int64_t)pn * 8 + ((((ax + 1) & 1) << 2) | (((ay + 1) & 1) << 1) | ((az + 1) & 1)));
    if (u >= 0) {
      if (rp) { pb = __ldg(rp + 2 * (int64_t)u); np = __ldg(rp + 2 * (int64_t)u + 1) - pb; }
      if (rn) { nb = __ldg(rn + 2 * (int64_t)u); nn = __ldg(rn + 2 * (int64_t)u + 1) - nb; }
    }
    s_u[tid] = u; s_pb[tid] = pb; s_np[tid] = np; s_nb[tid] = nb; s_nn[tid] = nn;
  }
  for (int t = tid; t < kWarps * SLOTS; t += kWarps * 32) s_acc[t] = 0.f;
  __syncthreads();
  if (tid < 64) {
    int off = 0;
    for (int j = 0; j < tid; ++j) off += s_np[j] + s_nn[j];
    s_off[tid] = off;
    if (tid == 63) s_off[64] = off + s_np[63] + s_nn[63];
  }
  __syncthreads();
  const int T = s_off[64];

  // per-entry metadata of chunk `ci` (entries [ci*kChunk, ...)) into slot ci % 3
  auto make_meta = [&](int ci) {
    const int ge = ci * kChunk + tid;
    if (tid < kChunk && ge < T) {
      int lo = 0, hi = 64;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= ge) lo = mid; else hi = mid;
      }
      const int j = ge - s_off[lo];
      const bool isn = j >= s_np[lo];
      s_mcell[ci % 3][tid] = lo;
      s_mtype[ci % 3][tid] = isn ? 1 : 0;
      s_mq[ci % 3][tid] = isn ? s_nb[lo] + j - s_np[lo] : s_pb[lo] + j;
    }
  };
  make_meta(0);
  __syncthreads();

  RowGeom g;
  g.ux = g.uy = g.uz = 0;
#pragma unroll
  for (int k = 0; k < NKSR_MAX_DEPTH; ++k) g.anc[k] = -1;
  if (i >= 0) {
    morton3_decode(__ldg(svh.keys[l] + i), g.ux, g.uy, g.uz);
    g.anc[0] = i;
    int a = p;
#pragma unroll
    for (int k = 1; k < NKSR_MAX_DEPTH; ++k) {
      if (l + k < svh.depth) { g.anc[k] = a; a = (a >= 0 && l + k + 1 < svh.depth) ? __ldg(svh.parent[l + k] + a) : -1; }
    }
  }
  const int sl = lane < 27 ? lane : 13;
  const int ldx = c_d27[sl][0], ldy = c_d27[sl][1], ldz = c_d27[sl][2];
  const float cx0 = ldx == 0 ? 0.75f : 0.125f, cx1 = 0.5f * (float)ldx, cx2 = ldx == 0 ? -1.f : 0.5f;
  const float cy0 = ldy == 0 ? 0.75f : 0.125f, cy1 = 0.5f * (float)ldy, cy2 = ldy == 0 ? -1.f : 0.5f;
  const float cz0 = ldz == 0 ? 0.75f : 0.125f, cz1 = 0.5f * (float)ldz, cz2 = ldz == 0 ? -1.f : 0.5f;
  const float inv_wl = 1.f / (svh.voxel_size * (float)(1 << l));
  const int ccx = (wid >> 2) & 1, ccy = (wid >> 1) & 1, ccz = wid & 1;  // child position of this row
  float* acc = s_acc + wid * SLOTS;
  float r[MAXL];
#pragma unroll
  for (int k = 0; k < MAXL; ++k) r[k] = 0.f;
  float bsum = 0.f;
  int cur_cell = -1, si = 0, udx = 0, udy = 0, udz = 0;
  bool member = false;

  auto flush = [&]() {
    if (member && lane < 27) {
      acc[(udx + ldx + 2) * 25 + (udy + ldy + 2) * 5 + (udz + ldz + 2)] += r[0];
      const int vx = g.ux + udx, vy = g.uy + udy, vz = g.uz + udz;
#pragma unroll
      for (int k = 1; k < MAXL; ++k) {
        if (k <= nup) {
          const int ox = ((vx >> k) + ldx) - (((g.ux - 1) >> k) - 1);
          const int oy = ((vy >> k) + ldy) - (((g.uy - 1) >> k) - 1);
          const int oz = ((vz >> k) + ldz) - (((g.uz - 1) >> k) - 1);
          acc[125 + 64 * (k - 1) + (ox << 4) + (oy << 2) + oz] += r[k];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < MAXL; ++k) r[k] = 0.f;
  };

  const int nlev = nup + 1;
  for (int c0 = 0, it = 0; c0 < T; c0 += kChunk, ++it) {
    const int buf = it & 1, mb = it % 3;
    make_meta(it + 1);
    const int ne = min(kChunk, T - c0);
    // ---- stage: one warp per (entry, level) line
    for (int ln = wid; ln < ne * nlev; ln += kWarps) {
      const int e = ln / nlev, k = ln - e * nlev;
      const int q = s_mq[mb][e];
      float* dst = s_exp + ((size_t)(buf * kChunk + e) * MAXL + k) * 96;
      if (!s_mtype[mb][e]) {
        dst[lane] = __ldg(cs.e_pos + ((int64_t)(l + k) * N + q) * NKSR_ROW_STRIDE + lane);
      } else if (COMPACT) {
        const float line = __ldg(cs.e_nrm + ((int64_t)(l + k) * K + q) * NKSR_ROW_STRIDE + lane);
        const float tx = __shfl_sync(0xffffffffu, line, 27), ty = __shfl_sync(0xffffffffu, line, 28),
                    tz = __shfl_sync(0xffffffffu, line, 29);
        const float bx = fmaf(fmaf(cx2, tx, cx1), tx, cx0), dbx = fmaf(2.f * cx2, tx, cx1);
        const float by = fmaf(fmaf(cy2, ty, cy1), ty, cy0), dby = fmaf(2.f * cy2, ty, cy1);
        const float bz = fmaf(fmaf(cz2, tz, cz1), tz, cz0), dbz = fmaf(2.f * cz2, tz, cz1);
        const float sc = (lane < 27 ? line : 0.f) * (inv_wl / (float)(1 << k));
        dst[lane] = dbx * by * bz * sc;
        dst[32 + lane] = bx * dby * bz * sc;
        dst[64 + lane] = bx * by * dbz * sc;
        if (k == 0 && lane < 3) s_t[buf][e][lane] = __ldg(cs.t_nrm + (int64_t)q * 3 + lane);
      } else {
        const float* src = cs.e_nrm + ((int64_t)(l + k) * K + q) * (3 * NKSR_ROW_STRIDE);
        dst[lane] = __ldg(src + lane);
        dst[32 + lane] = __ldg(src + 32 + lane);
        dst[64 + lane] = __ldg(src + 64 + lane);
        if (k == 0 && lane < 3) s_t[buf][e][lane] = __ldg(cs.t_nrm + (int64_t)q * 3 + lane);
      }
    }
    __syncthreads();
    // ---- consume
    if (i >= 0) {
      for (int e = 0; e < ne; ++e) {
        const int cell = s_mcell[mb][e];
        if (cell != cur_cell) {
          flush();
          cur_cell = cell;
          udx = (cell >> 4) - 1 - ccx; udy = ((cell >> 2) & 3) - 1 - ccy; udz = (cell & 3) - 1 - ccz;
          member = udx >= -1 && udx <= 1 && udy >= -1 && udy <= 1 && udz >= -1 && udz <= 1;
          si = 26 - ((udx + 1) * 9 + (udy + 1) * 3 + (udz + 1));
        }
        if (!member) continue;
        const float* ex = s_exp + (size_t)(buf * kChunk + e) * ESTRIDE;
        if (!s_mtype[mb][e]) {
          const float a = cs.w_pos * ex[si];
#pragma unroll
          for (int k = 0; k < MAXL; ++k)
            if (k <= nup) r[k] = fmaf(a, ex[k * 96 + lane], r[k]);
        } else {
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            const float a = cs.w_nrm * ex[ax * 32 + si];
            bsum = fmaf(a, s_t[buf][e][ax], bsum);
#pragma unroll
            for (int k = 0; k < MAXL; ++k)
              if (k <= nup) r[k] = fmaf(a, ex[k * 96 + ax * 32 + lane], r[k]);
          }
        }
      }
    }
  }
  flush();
  if (i < 0) return;
  __syncwarp();
  const int64_t row = svh.offset[l] + i;
  // regulariser: R_{i,i+d} = w_reg * B3(d) * <z_i, z_{i+d}>  (SPEC S5)
  if (cs.w_reg != 0.f && lane < 27) {
    const int nbv = __ldg(svh.nbr27[l] + (int64_t)i * 27 + lane);
    if (nbv >= 0) {
      const int C = feat.channels;
      const float* zi = feat.z[l] + (int64_t)i * C;
      const float* zn = feat.z[l] + (int64_t)nbv * C;
      float d = 0.f;
      for (int c = 0; c < C; ++c) d = fmaf(__ldg(zi + c), __ldg(zn + c), d);
      const float bw = (ldx == 0 ? 0.75f : 0.125f) * (ldy == 0 ? 0.75f : 0.125f) * (ldz == 0 ? 0.75f : 0.125f);
      acc[(ldx + 2) * 25 + (ldy + 2) * 5 + (ldz + 2)] += cs.w_reg * bw * d;
    }
  }
  __syncwarp();
  const int nslots = 125 + 64 * nup;
  const int64_t p0 = rowptr[row];
  int written = 0;
  for (int t0 = 0; t0 < nslots; t0 += 32) {
    const int t = t0 + lane;
    int k = 0;
    const int c = t < nslots ? slot_column(svh, l, g, t, k) : -1;
    const unsigned m = __ballot_sync(0xffffffffu, c >= 0);
    if (c >= 0) {
      const int64_t pp = p0 + written + __popc(m & ((1u << lane) - 1u));
      const float v = acc[t];
      const int64_t gc = svh.offset[l + k] + c;
      col_out[pp] = (int32_t)gc;
      val_out[pp] = v;
      if (k == 0 && c == i) diag[row] = v;
      if (k > 0) {
        const int64_t q = rowptr[gc] + cnt[gc] + atomicAdd(cursor + gc, 1);
        col_out[q] = (int32_t)row;
        val_out[q] = v;
      }
    }
    written += __popc(m);
  }
  if (lane == 0) rhs[row] = bsum;
}

// bitonic sort of the finer-level segment of each listed row by column (block per row)
__global__ void k_sort_down(const int32_t* __restrict__ cnt, const int32_t* __restrict__ cnt_down,
                            const int64_t* __restrict__ rowptr, const int32_t* __restrict__ rows, int64_t n_rows,
                            int32_t* __restrict__ col, float* __restrict__ val, int cap) {
  extern __shared__ unsigned char raw[];
  int32_t* sc = reinterpret_cast<int32_t*>(raw);
  float* sv = reinterpret_cast<float*>(raw) + cap;
  if (blockIdx.x >= n_rows) return;
  const int64_t row = rows[blockIdx.x];
  const int m = cnt_down[row];
  if (m <= 1 || m > cap) return;
  const int64_t p0 = rowptr[row] + cnt[row];
  int m2 = 1;
  while (m2 < m) m2 <<= 1;
  for (int t = threadIdx.x; t < m2; t += blockDim.x) {
    sc[t] = t < m ? col[p0 + t] : 0x7fffffff;
    sv[t] = t < m ? val[p0 + t] : 0.f;
  }
  __syncthreads();
  for (int k = 2; k <= m2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < m2; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const bool up = (t & k) == 0;
          const int a = sc[t], b = sc[p];
          if ((a > b) == up) {
            sc[t] = b; sc[p] = a;
            const float x = sv[t]; sv[t] = sv[p]; sv[p] = x;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < m; t += blockDim.x) {
    col[p0 + t] = sc[t];
    val[p0 + t] = sv[t];
  }
}

static int64_t total_unknowns(const nksr_svh_t* svh) {
  return svh->offset[svh->depth - 1] + svh->n[svh->depth - 1];
}

}  // namespace

extern "C" {

int nksr_gram_count(const nksr_svh_t* svh, int32_t* cnt, int32_t* cnt_down, void* stream) {
  if (!svh || svh->depth < 1 || svh->depth > NKSR_MAX_DEPTH) return NKSR_E_INVALID;
  if (!svh->nbr125_top && !svh->parent[svh->depth - 1]) return NKSR_E_INVALID;
  const int64_t n = total_unknowns(svh);
  if (n == 0) return NKSR_OK;
  if (cudaMemsetAsync(cnt_down, 0, (size_t)n * sizeof(int32_t), as_stream(stream)) != cudaSuccess) return NKSR_E_CUDA;
  k_gram_count<<<grid_for(n, kWarps), kWarps * 32, 0, as_stream(stream)>>>(*svh, n, cnt, cnt_down);
  NKSR_CHECK_LAUNCH();
  return NKSR_OK;
}

size_t nksr_scan_workspace_bytes(int64_t n) {
  size_t bytes = 0;
  cub::CountingInputIterator<int64_t> cnt_it(0);
  cub::TransformInputIterator<int64_t, AddPair, cub::CountingInputIterator<int64_t>> it(cnt_it,
                                                                                        AddPair{nullptr, nullptr});
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (int64_t*)nullptr, n);
  return bytes + 256;
}

int nksr_gram_rowptr(const int32_t* cnt, const int32_t* cnt_down, int64_t n, int64_t* rowptr, void* ws,
                     size_t ws_bytes, void* stream) {
  if (n <= 0) return NKSR_E_INVALID;
  cub::CountingInputIterator<int64_t> cnt_it(0);
  cub::TransformInputIterator<int64_t, AddPair, cub::CountingInputIterator<int64_t>> it(cnt_it,
                                                                                        AddPair{cnt, cnt_down});
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, it, rowptr, n);
  if (need > ws_bytes) return NKSR_E_WORKSPACE;
  if (cub::DeviceScan::ExclusiveSum(ws, need, it, rowptr, n, as_stream(stream)) != cudaSuccess) return NKSR_E_CUDA;
  k_set_last<<<1, 32, 0, as_stream(stream)>>>(cnt, cnt_down, n, rowptr);
  NKSR_CHECK_LAUNCH();
  return NKSR_OK;
}

int nksr_gram_fill(const nksr_svh_t* svh, const nksr_feat_t* feat, const nksr_constraints_t* c, const int32_t* cnt,
                   const int64_t* rowptr, int32_t* col, float* val, float* rhs, float* diag, int32_t* cursor,
                   void* stream) {
  if (!svh || !feat || !c || svh->depth < 1 || svh->depth > NKSR_MAX_DEPTH) return NKSR_E_INVALID;
  if (!svh->nbr125_top && !svh->parent[svh->depth - 1]) return NKSR_E_INVALID;
  const int64_t n = total_unknowns(svh);
  if (n == 0) return NKSR_OK;
  cudaStream_t s = as_stream(stream);
  const char* variant = getenv("NKSR_FILL_VARIANT");  // "group" selects the staged kernel (A/B switch)
  if (variant && variant[0] == 'g' && svh->parent[svh->depth - 1] && svh->depth < NKSR_MAX_DEPTH) {
    // grouped kernel: one CTA per parent voxel of every level (the virtual level parents the top)
    int64_t groups = 0;
    for (int l = 0; l < svh->depth; ++l) groups += svh->n[l + 1];
    if (groups <= 0 || groups > 0x7fffffff) return NKSR_E_INVALID;
#define NKSR_FILLG(COMPACT, MAXL)                                                                              \
  do {                                                                                                         \
    const size_t sm = (size_t)(kWarps * (125 + 64 * (MAXL - 1)) + 2 * kChunk * MAXL * 96) * sizeof(float);     \
    if (cudaFuncSetAttribute(k_gram_fill_group<COMPACT, MAXL>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                             (int)sm) != cudaSuccess)                                                         \
      return NKSR_E_CUDA;                                                                                      \
    k_gram_fill_group<COMPACT, MAXL><<<(unsigned)groups, kWarps * 32, sm, s>>>(*svh, *feat, *c, cnt, rowptr,   \
                                                                               col, val, rhs, diag, cursor);  \
  } while (0)
    if (svh->depth <= 4) {
      if (c->nrm_compact) NKSR_FILLG(true, 4); else NKSR_FILLG(false, 4);
    } else {
      if (c->nrm_compact) NKSR_FILLG(true, NKSR_MAX_DEPTH); else NKSR_FILLG(false, NKSR_MAX_DEPTH);
    }
#undef NKSR_FILLG
    NKSR_CHECK_LAUNCH();
    return NKSR_OK;
  }
  const size_t smem = (size_t)kWarps * kMaxSlots * sizeof(float);
  const int grid = grid_for(n, kWarps);
#define NKSR_FILL(COMPACT, MAXL) \
  k_gram_fill<COMPACT, MAXL><<<grid, kWarps * 32, smem, s>>>(*svh, *feat, *c, n, cnt, rowptr, col, val, rhs, diag, cursor)
  if (svh->depth <= 4) {
    if (c->nrm_compact) NKSR_FILL(true, 4); else NKSR_FILL(false, 4);
  } else {
    if (c->nrm_compact) NKSR_FILL(true, NKSR_MAX_DEPTH); else NKSR_FILL(false, NKSR_MAX_DEPTH);
  }
#undef NKSR_FILL
  NKSR_CHECK_LAUNCH();
  return NKSR_OK;
}

int nksr_gram_sort_down(const int32_t* cnt, const int32_t* cnt_down, const int64_t* rowptr, const int32_t* rows,
                        int64_t n_rows, int cap, int32_t* col, float* val, void* stream) {
  // one block per listed row; segments longer than `cap` (a power of two <= 16384) are left in
  // insertion order.
  if (n_rows <= 0) return NKSR_OK;
  if (cap < 2 || cap > 16384 || (cap & (cap - 1))) return NKSR_E_INVALID;
  if (cap * 8 > 48 * 1024 &&
      cudaFuncSetAttribute(k_sort_down, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 8) != cudaSuccess)
    return NKSR_E_CUDA;
  const int threads = cap >= 2048 ? 256 : (cap >= 256 ? 128 : 32);
  k_sort_down<<<(unsigned)n_rows, threads, cap * 8, as_stream(stream)>>>(cnt, cnt_down, rowptr, rows, n_rows, col,
                                                                         val, cap);
  NKSR_CHECK_LAUNCH();
  return NKSR_OK;
}

}  // extern "C"
