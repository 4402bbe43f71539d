// Per-(location, level) neural-kernel evaluation, one warp per location, lane = stencil slot.
// K_l(x, i) = B3((x - c_i)/W_l) * <phi_l(x), z_i>,  phi_l(x) = trilinear interpolation of z
// (DESIGN.md SPEC S4).  Shared by row building (Gram assembly) and field evaluation.
#pragma once
#include "common.cuh"

struct LaneKernel {
  int nb;       // neighbour voxel index of this lane's slot (-1: none / lane >= 27)
  float k;      // K_l(x, nb)
  float dk[3];  // grad_x K_l(x, nb)
  float dot;    // <phi_l(x), z_nb>  (0 when nb < 0)
  float tau[3]; // local coordinate of x in the containing voxel
};

// weights of neighbour d in {-1,0,1} along one axis at local coordinate tau in [-.5,.5)
// Tent derivative (SPEC S4): one-sided derivative of the trilinear cell containing x, except in
// the snap zone |tau| < 2^-12 around a voxel centre, where the symmetric derivative is used --
// the reference places its normal constraints exactly AT voxel centres (models/nksr_net.py:100),
// where a one-sided rule would depend on the last rounding bit of the coordinate.
#define NKSR_TENT_SNAP 0.000244140625f
__device__ __forceinline__ void axis_weights(float tau, int d, float& b, float& db, float& t, float& dt) {
  const bool mid = fabsf(tau) < NKSR_TENT_SNAP;
  if (d == 0) {
    b = 0.75f - tau * tau;
    db = -2.f * tau;
    t = tau >= 0.f ? 1.f - tau : 1.f + tau;
    dt = mid ? 0.f : (tau >= 0.f ? -1.f : 1.f);
  } else if (d < 0) {
    float h = 0.5f - tau;
    b = 0.5f * h * h;
    db = -h;
    t = tau >= 0.f ? 0.f : -tau;
    dt = mid ? -0.5f : (tau >= 0.f ? 0.f : -1.f);
  } else {
    float h = 0.5f + tau;
    b = 0.5f * h * h;
    db = h;
    t = tau >= 0.f ? tau : 0.f;
    dt = mid ? 0.5f : (tau >= 0.f ? 1.f : 0.f);
  }
}

// All 32 lanes must call.  base >= 0.  GRAD: also the gradient; FULLGRAD: include the
// grad(phi) term (approx_kernel_grad == false).
// (ux,uy,uz): offset-space coordinates of the containing voxel `base` -- the caller already has them from the
// point's own quantisation ((h + 2^20) >> (level+1), SPEC S1), so the key is neither loaded nor decoded;
// inv0 = 1 / voxel_size in fp64, computed once per location (1/(W 2^l) = inv0 * 2^-l exactly).
template <bool GRAD>
__device__ __forceinline__ LaneKernel eval_level_lane(const int32_t* __restrict__ nbr27,
                                                      const float* __restrict__ z, int C, int level,
                                                      float wl, double inv0, float px, float py, float pz, int base,
                                                      int ux, int uy, int uz, bool fullgrad, int lane) {
  LaneKernel r;
  const int off = level_offset(level);
  // local coordinate in voxel units; the subtraction is done in fp64 to avoid cancellation
  const double inv = inv0 * (1.0 / (double)(1 << level));
  float tx = (float)((double)px * inv - ((double)(ux - off) + 0.5));
  float ty = (float)((double)py * inv - ((double)(uy - off) + 0.5));
  float tz = (float)((double)pz * inv - ((double)(uz - off) + 0.5));
  int dx, dy, dz;
  slot_to_d(lane < 27 ? lane : 13, dx, dy, dz);
  float bx, dbx, ttx, dtx, by, dby, tty, dty, bz, dbz, ttz, dtz;
  axis_weights(tx, dx, bx, dbx, ttx, dtx);
  axis_weights(ty, dy, by, dby, tty, dty);
  axis_weights(tz, dz, bz, dbz, ttz, dtz);
  r.nb = lane < 27 ? __ldg(nbr27 + (int64_t)base * 27 + lane) : -1;
  const bool ok = r.nb >= 0;
  const float B3 = bx * by * bz;
  const float T3 = ok ? ttx * tty * ttz : 0.f;
  float dT3[3];
  if (GRAD) {
    dT3[0] = ok ? dtx * tty * ttz : 0.f;
    dT3[1] = ok ? ttx * dty * ttz : 0.f;
    dT3[2] = ok ? ttx * tty * dtz : 0.f;
  }
  float dot = 0.f, ddot[3] = {0.f, 0.f, 0.f};
  const float* zr = z + (int64_t)(ok ? r.nb : 0) * C;
  auto channel = [&](const float zc) {
    float phi = warp_sum(T3 * zc);
    dot = fmaf(phi, zc, dot);
    if (GRAD && fullgrad) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        float dphi = warp_sum(dT3[a] * zc);
        ddot[a] = fmaf(dphi, zc, ddot[a]);
      }
    }
  };
  if ((C & 3) == 0) {
    // the 27 lanes read 27 different feature rows: every load instruction is a 27-wavefront gather in L1, so fetch
    // four channels per instruction (rows of C = 4, 8, 16 ... floats are 16-byte aligned); same arithmetic, same order
    for (int c4 = 0; c4 < C; c4 += 4) {
      const float4 v = ok ? __ldg(reinterpret_cast<const float4*>(zr + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      channel(v.x); channel(v.y); channel(v.z); channel(v.w);
    }
  } else {
    for (int c = 0; c < C; ++c) channel(ok ? __ldg(zr + c) : 0.f);
  }
  r.k = ok ? B3 * dot : 0.f;
  r.dot = ok ? dot : 0.f;
  r.tau[0] = tx; r.tau[1] = ty; r.tau[2] = tz;
  if (GRAD) {
    const float iw = 1.f / wl;
    r.dk[0] = ok ? (dbx * by * bz * dot + B3 * ddot[0]) * iw : 0.f;
    r.dk[1] = ok ? (bx * dby * bz * dot + B3 * ddot[1]) * iw : 0.f;
    r.dk[2] = ok ? (bx * by * dbz * dot + B3 * ddot[2]) * iw : 0.f;
  } else {
    r.dk[0] = r.dk[1] = r.dk[2] = 0.f;
  }
  return r;
}
