// C++/OpenMP CPU restatement of the NKSR reconstruction hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.
//
// PARITY UNPINNED (see oracle/nksr_oracle.py and DESIGN.md section 1): the reference ships this path as
// a closed wheel; this file restates DESIGN.md's SPEC S1-S7 a SECOND time, in plain C++/OpenMP and
// written independently of oracle/nksr_oracle.py (different data structures: binary searches instead of
// tables, a Gustavson-style E^T W E instead of scipy's sparse product).  tests/test_cpu_port.py requires
// the two restatements to agree (keys bit for bit, matrix entries to fp32 rounding), which is the
// strongest pin available for an oracle without reference vectors.  Never linked or imported by the
// product (nksr_b200/).  Call sites it follows:
//   SparseFeatureHierarchy(...).build_point_splatting          models/nksr_net.py:57-62
//   KernelField(...).solve_non_fused(pos, nrm_xyz, nrm_val, w)  models/nksr_net.py:91-112
//   CPU plumbing of the reference run                          examples/recons_waymo_cpu.py:48-63
//
// Build (oracle/Makefile):  g++ -O2 -fopenmp -shared -fPIC -o oracle/_build/libnksr_oracle_cpu.so oracle/nksr_oracle_cpu.cpp
// (no -ffast-math: the voxel quantisation relies on IEEE fp32 division.)
#include <algorithm>
#include <parallel/algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int kMaxDepth = 8;

inline uint64_t part1by2(uint64_t v) {
  v &= 0x1FFFFFull;
  v = (v | (v << 32)) & 0x1F00000000FFFFull;
  v = (v | (v << 16)) & 0x1F0000FF0000FFull;
  v = (v | (v << 8)) & 0x100F00F00F00F00Full;
  v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
  v = (v | (v << 2)) & 0x1249249249249249ull;
  return v;
}
inline uint32_t compact1by2(uint64_t v) {
  v &= 0x1249249249249249ull;
  v = (v | (v >> 2)) & 0x10C30C30C30C30C3ull;
  v = (v | (v >> 4)) & 0x100F00F00F00F00Full;
  v = (v | (v >> 8)) & 0x1F0000FF0000FFull;
  v = (v | (v >> 16)) & 0x1F00000000FFFFull;
  v = (v | (v >> 32)) & 0x1FFFFFull;
  return (uint32_t)v;
}
inline int64_t morton3(int x, int y, int z) {
  return (int64_t)((part1by2((uint32_t)x) << 2) | (part1by2((uint32_t)y) << 1) | part1by2((uint32_t)z));
}
inline void demorton3(int64_t k, int& x, int& y, int& z) {
  x = (int)compact1by2((uint64_t)k >> 2);
  y = (int)compact1by2((uint64_t)k >> 1);
  z = (int)compact1by2((uint64_t)k);
}
inline int level_offset(int l) { return 1 << (19 - l); }
constexpr int kHalfOffset = 1 << 20;

struct Svh {
  float w;
  int depth;
  std::vector<int64_t> keys[kMaxDepth];
  std::vector<int64_t> offset;  // first unknown of each level (+ total)
  int64_t total() const { return offset.back(); }
  int find(int l, int64_t key) const {
    const auto& k = keys[l];
    auto it = std::lower_bound(k.begin(), k.end(), key);
    return (it != k.end() && *it == key) ? (int)(it - k.begin()) : -1;
  }
};

// SPEC S1: h = floor(x / (W/2)) with IEEE fp32 division; +2^20 offset
inline void half_coords(const float* p, float half_w, int h[3]) {
  for (int a = 0; a < 3; ++a) {
    volatile float q = p[a] / half_w;  // volatile: keep the fp32 rounding of the quotient
    h[a] = (int)std::floor(q) + kHalfOffset;
  }
}

struct Entry {
  int col;
  float val;
};

// kernel row of one location on one level (SPEC S4); returns the number of entries written (<= 27).
// grad == nullptr: value row into `val`; else three gradient rows into grad[0..2].
int level_row(const Svh& s, int l, const float* const* feat, int C, const float* p, bool want_grad, bool approx,
              Entry* val, Entry* gx, Entry* gy, Entry* gz) {
  const float half_w = s.w * 0.5f;
  int h[3];
  half_coords(p, half_w, h);
  const int bx = h[0] >> (l + 1), by = h[1] >> (l + 1), bz = h[2] >> (l + 1);
  const int base = s.find(l, morton3(bx, by, bz));
  if (base < 0) return 0;
  const double wl = (double)(s.w * (float)(1 << l));
  const int off = level_offset(l);
  const double tau[3] = {(double)p[0] / wl - ((double)(bx - off) + 0.5), (double)p[1] / wl - ((double)(by - off) + 0.5),
                         (double)p[2] / wl - ((double)(bz - off) + 0.5)};
  double B[3][3], dB[3][3], T[3][3], dT[3][3];
  for (int a = 0; a < 3; ++a) {
    const double t = tau[a];
    B[a][0] = 0.5 * (0.5 - t) * (0.5 - t); B[a][1] = 0.75 - t * t; B[a][2] = 0.5 * (0.5 + t) * (0.5 + t);
    dB[a][0] = -(0.5 - t); dB[a][1] = -2.0 * t; dB[a][2] = 0.5 + t;
    const bool pos = t >= 0, mid = std::fabs(t) < 1.0 / 4096.0;
    T[a][0] = pos ? 0.0 : -t; T[a][1] = pos ? 1.0 - t : 1.0 + t; T[a][2] = pos ? t : 0.0;
    dT[a][0] = mid ? -0.5 : (pos ? 0.0 : -1.0); dT[a][1] = mid ? 0.0 : (pos ? -1.0 : 1.0);
    dT[a][2] = mid ? 0.5 : (pos ? 1.0 : 0.0);
  }
  int nb[27];
  double phi[32] = {0}, dphi[3][32] = {{0}};
  for (int sl = 0; sl < 27; ++sl) {
    const int dx = sl / 9 - 1, dy = (sl / 3) % 3 - 1, dz = sl % 3 - 1;
    nb[sl] = s.find(l, morton3(bx + dx, by + dy, bz + dz));
    if (nb[sl] < 0) continue;
    const float* z = feat[l] + (int64_t)nb[sl] * C;
    const double t3 = T[0][dx + 1] * T[1][dy + 1] * T[2][dz + 1];
    for (int c = 0; c < C; ++c) phi[c] += t3 * z[c];
    if (want_grad && !approx) {
      const double g0 = dT[0][dx + 1] * T[1][dy + 1] * T[2][dz + 1], g1 = T[0][dx + 1] * dT[1][dy + 1] * T[2][dz + 1],
                   g2 = T[0][dx + 1] * T[1][dy + 1] * dT[2][dz + 1];
      for (int c = 0; c < C; ++c) { dphi[0][c] += g0 * z[c]; dphi[1][c] += g1 * z[c]; dphi[2][c] += g2 * z[c]; }
    }
  }
  int cnt = 0;
  for (int sl = 0; sl < 27; ++sl) {
    if (nb[sl] < 0) continue;
    const int dx = sl / 9, dy = (sl / 3) % 3, dz = sl % 3;
    const float* z = feat[l] + (int64_t)nb[sl] * C;
    double dot = 0, dd[3] = {0, 0, 0};
    for (int c = 0; c < C; ++c) {
      dot += phi[c] * z[c];
      if (want_grad && !approx) { dd[0] += dphi[0][c] * z[c]; dd[1] += dphi[1][c] * z[c]; dd[2] += dphi[2][c] * z[c]; }
    }
    const double b3 = B[0][dx] * B[1][dy] * B[2][dz];
    const int col = (int)(s.offset[l] + nb[sl]);
    if (!want_grad) {
      val[cnt] = {col, (float)(b3 * dot)};
    } else {
      gx[cnt] = {col, (float)((dB[0][dx] * B[1][dy] * B[2][dz] * dot + b3 * dd[0]) / wl)};
      gy[cnt] = {col, (float)((B[0][dx] * dB[1][dy] * B[2][dz] * dot + b3 * dd[1]) / wl)};
      gz[cnt] = {col, (float)((B[0][dx] * B[1][dy] * dB[2][dz] * dot + b3 * dd[2]) / wl)};
    }
    ++cnt;
  }
  return cnt;
}

struct System {
  std::vector<int64_t> rowptr;
  std::vector<int32_t> col;
  std::vector<float> val;
  std::vector<float> rhs;
};

}  // namespace

extern "C" {

struct nksr_cpu_svh;  // opaque = Svh

void* nksr_cpu_svh_build(const float* xyz, int64_t n, float voxel_size, int depth) {
  Svh* s = new Svh();
  s->w = voxel_size;
  s->depth = depth;
  const float half_w = voxel_size * 0.5f;
  for (int l = 0; l < depth; ++l) {
    std::vector<int64_t>& k = s->keys[l];
    k.resize((size_t)n * 8);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      int h[3];
      half_coords(xyz + 3 * i, half_w, h);
      const int bx = ((h[0] >> l) - 1) >> 1, by = ((h[1] >> l) - 1) >> 1, bz = ((h[2] >> l) - 1) >> 1;
      for (int a = 0; a < 8; ++a) k[(size_t)i * 8 + a] = morton3(bx + ((a >> 2) & 1), by + ((a >> 1) & 1), bz + (a & 1));
    }
    __gnu_parallel::sort(k.begin(), k.end());      // OpenMP multiway mergesort of libstdc++'s parallel mode
    k.erase(std::unique(k.begin(), k.end()), k.end());
  }
  s->offset.assign(depth + 1, 0);
  for (int l = 0; l < depth; ++l) s->offset[l + 1] = s->offset[l] + (int64_t)s->keys[l].size();
  return s;
}
void nksr_cpu_svh_free(void* h) { delete static_cast<Svh*>(h); }
int64_t nksr_cpu_svh_count(void* h, int l) { return (int64_t) static_cast<Svh*>(h)->keys[l].size(); }
void nksr_cpu_svh_keys(void* h, int l, int64_t* out) {
  const auto& k = static_cast<Svh*>(h)->keys[l];
  std::memcpy(out, k.data(), k.size() * sizeof(int64_t));
}
// voxel centres of level l: (ijk + 0.5) * W_l in fp32
void nksr_cpu_svh_centers(void* h, int l, float* out) {
  const Svh* s = static_cast<Svh*>(h);
  const float wl = s->w * (float)(1 << l);
  const int off = level_offset(l);
  for (size_t i = 0; i < s->keys[l].size(); ++i) {
    int x, y, z;
    demorton3(s->keys[l][i], x, y, z);
    out[3 * i] = ((float)(x - off) + 0.5f) * wl;
    out[3 * i + 1] = ((float)(y - off) + 0.5f) * wl;
    out[3 * i + 2] = ((float)(z - off) + 0.5f) * wl;
  }
}

// A = E^T diag(w) E + reg R (SPEC S5) as CSR with sorted columns; returns an opaque system handle.
void* nksr_cpu_build_system(void* h, const float* const* feat, int C, const float* pos_xyz, int64_t n_pos,
                            const float* nrm_xyz, const float* nrm_val, int64_t n_nrm, float w_pos, float w_nrm,
                            float w_reg, int approx) {
  const Svh& s = *static_cast<Svh*>(h);
  const int L = s.depth;
  const int64_t n = s.total();
  const int64_t M = n_pos + 3 * n_nrm;
  const int maxrow = 27 * L;
  // E in padded row storage
  std::vector<Entry> E((size_t)M * maxrow);
  std::vector<int> Elen((size_t)M, 0);
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t r = 0; r < n_pos; ++r) {
    int len = 0;
    for (int l = 0; l < L; ++l)
      len += level_row(s, l, feat, C, pos_xyz + 3 * r, false, false, &E[(size_t)r * maxrow + len], nullptr, nullptr, nullptr);
    Elen[r] = len;
  }
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t k = 0; k < n_nrm; ++k) {
    int len = 0;
    const size_t r0 = (size_t)(n_pos + 3 * k) * maxrow;
    for (int l = 0; l < L; ++l)
      len += level_row(s, l, feat, C, nrm_xyz + 3 * k, true, approx != 0, nullptr, &E[r0 + len], &E[r0 + maxrow + len],
                       &E[r0 + 2 * (size_t)maxrow + len]);
    Elen[n_pos + 3 * k] = Elen[n_pos + 3 * k + 1] = Elen[n_pos + 3 * k + 2] = len;
  }
  // transpose index: for every unknown the (row, value) pairs touching it (built in parallel; the order of the
  // pairs inside one unknown's list is scheduling dependent, the fp64 accumulation below makes that immaterial)
  std::vector<int64_t> tptr(n + 1, 0);
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < M; ++r)
    for (int e = 0; e < Elen[r]; ++e) {
#pragma omp atomic
      ++tptr[E[(size_t)r * maxrow + e].col + 1];
    }
  for (int64_t i = 0; i < n; ++i) tptr[i + 1] += tptr[i];
  std::vector<int64_t> trow(tptr[n]);
  std::vector<float> tval(tptr[n]);
  {
    std::vector<int64_t> cur(tptr.begin(), tptr.end() - 1);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < M; ++r)
      for (int e = 0; e < Elen[r]; ++e) {
        const Entry& en = E[(size_t)r * maxrow + e];
        int64_t p;
#pragma omp atomic capture
        p = cur[en.col]++;
        trow[p] = r;
        tval[p] = en.val;
      }
  }
  System* sys = new System();
  sys->rhs.assign(n, 0.f);
  std::vector<std::vector<Entry>> rows((size_t)n);
  const double b1[3] = {0.125, 0.75, 0.125};
#pragma omp parallel
  {
    std::vector<double> acc((size_t)n, 0.0);     // dense per-thread accumulator + touched list
    std::vector<char> seen((size_t)n, 0);
    std::vector<int> touched;
    auto add = [&](int col, double v) {
      if (!seen[col]) { seen[col] = 1; touched.push_back(col); }
      acc[col] += v;
    };
#pragma omp for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; ++i) {
      touched.clear();
      double b = 0;
      for (int64_t p = tptr[i]; p < tptr[i + 1]; ++p) {
        const int64_t r = trow[p];
        const double wr = r < n_pos ? (double)w_pos : (double)w_nrm;
        const double a = wr * tval[p];
        if (r >= n_pos) b += a * nrm_val[r - n_pos];
        for (int e = 0; e < Elen[r]; ++e) {
          const Entry& en = E[(size_t)r * maxrow + e];
          add(en.col, a * en.val);
        }
      }
      // regulariser: same-level 27-neighbourhood, K_l(c_j, i) = B3(d) <z_i, z_j>
      int l = 0;
      while (l + 1 < L && i >= s.offset[l + 1]) ++l;
      const int ii = (int)(i - s.offset[l]);
      int x, y, z;
      demorton3(s.keys[l][ii], x, y, z);
      for (int sl = 0; sl < 27 && w_reg != 0.f; ++sl) {
        const int dx = sl / 9 - 1, dy = (sl / 3) % 3 - 1, dz = sl % 3 - 1;
        const int j = s.find(l, morton3(x + dx, y + dy, z + dz));
        if (j < 0) continue;
        double d = 0;
        for (int c = 0; c < C; ++c) d += (double)feat[l][(int64_t)ii * C + c] * feat[l][(int64_t)j * C + c];
        const int col = (int)(s.offset[l] + j);
        add(col, (double)w_reg * b1[dx + 1] * b1[dy + 1] * b1[dz + 1] * d);
      }
      std::sort(touched.begin(), touched.end());
      auto& out = rows[i];
      out.reserve(touched.size());
      for (int c : touched) {
        out.push_back({c, (float)acc[c]});
        acc[c] = 0.0;
        seen[c] = 0;
      }
      sys->rhs[i] = (float)b;
    }
  }
  sys->rowptr.assign(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) sys->rowptr[i + 1] = sys->rowptr[i] + (int64_t)rows[i].size();
  sys->col.resize(sys->rowptr[n]);
  sys->val.resize(sys->rowptr[n]);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    for (size_t e = 0; e < rows[i].size(); ++e) {
      sys->col[sys->rowptr[i] + e] = rows[i][e].col;
      sys->val[sys->rowptr[i] + e] = rows[i][e].val;
    }
  return sys;
}
void nksr_cpu_system_free(void* h) { delete static_cast<System*>(h); }
int64_t nksr_cpu_system_n(void* h) { return (int64_t) static_cast<System*>(h)->rhs.size(); }
int64_t nksr_cpu_system_nnz(void* h) { return (int64_t) static_cast<System*>(h)->col.size(); }
void nksr_cpu_system_copy(void* h, int64_t* rowptr, int32_t* col, float* val, float* rhs) {
  const System* s = static_cast<System*>(h);
  std::memcpy(rowptr, s->rowptr.data(), s->rowptr.size() * sizeof(int64_t));
  std::memcpy(col, s->col.data(), s->col.size() * sizeof(int32_t));
  std::memcpy(val, s->val.data(), s->val.size() * sizeof(float));
  std::memcpy(rhs, s->rhs.data(), s->rhs.size() * sizeof(float));
}

// Jacobi-PCG in fp32 with fp64 dot products (SPEC S7); returns iterations, *relres = ||r||/||b||
int nksr_cpu_pcg(void* h, float tol, int max_iter, float* x, double* relres) {
  const System& s = *static_cast<System*>(h);
  const int64_t n = (int64_t)s.rhs.size();
  std::vector<float> dinv(n), r(n), z(n), p(n), ap(n);
  double rz = 0, bb = 0;
#pragma omp parallel for reduction(+ : rz, bb)
  for (int64_t i = 0; i < n; ++i) {
    float d = 0.f;
    for (int64_t q = s.rowptr[i]; q < s.rowptr[i + 1]; ++q)
      if (s.col[q] == i) d = s.val[q];
    dinv[i] = d > 0.f ? 1.f / d : 0.f;
    x[i] = 0.f;
    r[i] = s.rhs[i];
    z[i] = r[i] * dinv[i];
    p[i] = z[i];
    rz += (double)r[i] * z[i];
    bb += (double)r[i] * r[i];
  }
  *relres = 0;
  if (!(bb > 0)) return 0;
  int it = 0;
  double rr = bb;
  while (it < max_iter && rr > (double)tol * tol * bb) {
    double pap = 0;
#pragma omp parallel for reduction(+ : pap) schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; ++i) {
      float acc = 0.f;
      for (int64_t q = s.rowptr[i]; q < s.rowptr[i + 1]; ++q) acc += s.val[q] * p[s.col[q]];
      ap[i] = acc;
      pap += (double)acc * p[i];
    }
    const float alpha = (float)(rz / pap);
    double rzn = 0;
    rr = 0;
#pragma omp parallel for reduction(+ : rzn, rr)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * ap[i];
      z[i] = r[i] * dinv[i];
      rzn += (double)r[i] * z[i];
      rr += (double)r[i] * r[i];
    }
    const float beta = (float)(rzn / rz);
    rz = rzn;
#pragma omp parallel for
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    ++it;
  }
  *relres = std::sqrt(rr / bb);
  return it;
}


// f(x) = sum_l sum_{i in N27(b_l(x))} alpha_i K_l(x,i) and (optionally) its gradient (SPEC S3/S4), fp64 sums.
// Call site restated: field.evaluate_f(xyz, grad) (models/loss.py:189-198,225).
void nksr_cpu_evaluate(void* h, const float* const* feat, int C, const float* alpha, const float* xyz, int64_t m,
                       int want_grad, int approx, double* f, double* g) {
  const Svh& s = *static_cast<Svh*>(h);
  const int L = s.depth;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t q = 0; q < m; ++q) {
    Entry v[27], gx[27], gy[27], gz[27];
    double acc = 0, ax = 0, ay = 0, az = 0;
    for (int l = 0; l < L; ++l) {
      const int c = level_row(s, l, feat, C, xyz + 3 * q, false, false, v, nullptr, nullptr, nullptr);
      for (int e = 0; e < c; ++e) acc += (double)alpha[v[e].col] * v[e].val;
      if (want_grad) {
        const int cg = level_row(s, l, feat, C, xyz + 3 * q, true, approx != 0, nullptr, gx, gy, gz);
        for (int e = 0; e < cg; ++e) {
          const double a = alpha[gx[e].col];
          ax += a * gx[e].val; ay += a * gy[e].val; az += a * gz[e].val;
        }
      }
    }
    f[q] = acc;
    if (want_grad) { g[3 * q] = ax; g[3 * q + 1] = ay; g[3 * q + 2] = az; }
  }
}

// hierarchy from given (sorted, unique) keys per level -- for pruned / adaptive hierarchies
void* nksr_cpu_svh_from_keys(const int64_t* const* keys, const int64_t* counts, float voxel_size, int depth) {
  Svh* s = new Svh();
  s->w = voxel_size;
  s->depth = depth;
  s->offset.assign(depth + 1, 0);
  for (int l = 0; l < depth; ++l) {
    s->keys[l].assign(keys[l], keys[l] + counts[l]);
    s->offset[l + 1] = s->offset[l] + counts[l];
  }
  return s;
}

// containing voxel index per level (-1 when inactive), (L, m) row-major
void nksr_cpu_locate(void* h, const float* xyz, int64_t m, int32_t* base) {
  const Svh& s = *static_cast<Svh*>(h);
  const float half_w = s.w * 0.5f;
#pragma omp parallel for schedule(static)
  for (int64_t q = 0; q < m; ++q) {
    int hh[3];
    half_coords(xyz + 3 * q, half_w, hh);
    for (int l = 0; l < s.depth; ++l)
      base[(int64_t)l * m + q] = s.find(l, morton3(hh[0] >> (l + 1), hh[1] >> (l + 1), hh[2] >> (l + 1)));
  }
}

// 27-neighbour table of level l: out[i*27 + s], s = (dx+1)*9 + (dy+1)*3 + (dz+1), -1 where the neighbour is inactive
// (the table the stand-in network's smoothing and the CPU baseline read; one binary search per entry)
void nksr_cpu_nbr27(void* h, int l, int32_t* out) {
  const Svh& s = *static_cast<Svh*>(h);
  const int64_t n = (int64_t)s.keys[l].size();
  const int lim = 1 << 21;      // 21 bits per axis in a key (KEY_BITS of the numpy restatement)
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    int x, y, z;
    demorton3(s.keys[l][i], x, y, z);
    int t = 0;
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz, ++t) {
          const int a = x + dx, b = y + dy, c = z + dz;
          const bool ok = a >= 0 && b >= 0 && c >= 0 && a < lim && b < lim && c < lim;
          out[i * 27 + t] = ok ? s.find(l, morton3(a, b, c)) : -1;
        }
  }
}

// out[i][c] = sum over the slots s = 0..26 (in this order, fp64) of acc[nbr27[i][s]][c]: the 27-neighbourhood pooling of
// the stand-in network (nksr_b200/network.py, csrc: nksr_pool27), with the summation order of the numpy restatement
void nksr_cpu_pool27(const int32_t* nbr27, int64_t n, const double* acc, int C, double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double* o = out + i * C;
    for (int c = 0; c < C; ++c) o[c] = 0.0;
    for (int t = 0; t < 27; ++t) {
      const int j = nbr27[i * 27 + t];
      if (j < 0) continue;
      const double* a = acc + (int64_t)j * C;
      for (int c = 0; c < C; ++c) o[c] += a[c];
    }
  }
}

// SPEC S6 storage pattern, restated independently of the numpy oracle: row (l,i) stores every ACTIVE column in
// the same-level 5^3 stencil, in the box [((u-1)>>k)-1, ((u+1)>>k)+1]^3 of every coarser level l+k, and the
// transposes of the latter.  cnt[row] = stored entries of the row (own + transposed).
void nksr_cpu_structural_counts(void* h, int32_t* cnt) {
  const Svh& s = *static_cast<Svh*>(h);
  const int L = s.depth;
  const int64_t n = s.total();
  for (int64_t i = 0; i < n; ++i) cnt[i] = 0;
  for (int l = 0; l < L; ++l) {
    const int64_t nl = (int64_t)s.keys[l].size();
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < nl; ++i) {
      int x, y, z;
      demorton3(s.keys[l][i], x, y, z);
      int own = 0;
      for (int dx = -2; dx <= 2; ++dx)
        for (int dy = -2; dy <= 2; ++dy)
          for (int dz = -2; dz <= 2; ++dz)
            if (s.find(l, morton3(x + dx, y + dy, z + dz)) >= 0) ++own;
      for (int k = 1; l + k < L; ++k) {
        const int lo[3] = {((x - 1) >> k) - 1, ((y - 1) >> k) - 1, ((z - 1) >> k) - 1};
        const int hi[3] = {((x + 1) >> k) + 1, ((y + 1) >> k) + 1, ((z + 1) >> k) + 1};
        for (int cx = lo[0]; cx <= hi[0]; ++cx)
          for (int cy = lo[1]; cy <= hi[1]; ++cy)
            for (int cz = lo[2]; cz <= hi[2]; ++cz) {
              const int c = s.find(l + k, morton3(cx, cy, cz));
              if (c < 0) continue;
              ++own;
#pragma omp atomic
              ++cnt[s.offset[l + k] + c];
            }
      }
#pragma omp atomic
      cnt[s.offset[l] + i] += own;
    }
  }
}

// sorted global column indices of one row of the SPEC S6 pattern; returns their number (cols holds >= cnt[row])
int64_t nksr_cpu_structural_row(void* h, int64_t row, int32_t* cols) {
  const Svh& s = *static_cast<Svh*>(h);
  const int L = s.depth;
  int l = 0;
  while (l + 1 < L && row >= s.offset[l + 1]) ++l;
  const int64_t i = row - s.offset[l];
  int x, y, z;
  demorton3(s.keys[l][i], x, y, z);
  std::vector<int32_t> out;
  for (int dx = -2; dx <= 2; ++dx)
    for (int dy = -2; dy <= 2; ++dy)
      for (int dz = -2; dz <= 2; ++dz) {
        const int j = s.find(l, morton3(x + dx, y + dy, z + dz));
        if (j >= 0) out.push_back((int32_t)(s.offset[l] + j));
      }
  for (int k = 1; l + k < L; ++k)
    for (int cx = ((x - 1) >> k) - 1; cx <= ((x + 1) >> k) + 1; ++cx)
      for (int cy = ((y - 1) >> k) - 1; cy <= ((y + 1) >> k) + 1; ++cy)
        for (int cz = ((z - 1) >> k) - 1; cz <= ((z + 1) >> k) + 1; ++cz) {
          const int c = s.find(l + k, morton3(cx, cy, cz));
          if (c >= 0) out.push_back((int32_t)(s.offset[l + k] + c));
        }
  // transposes: finer voxels j (level l-k) whose box on this level contains (x,y,z): j's coords u satisfy
  // ((u-1)>>k)-1 <= x <= ((u+1)>>k)+1  <=>  u in [((x-1)<<k) - 1, ((x+2)<<k)] per axis (checked exactly below)
  for (int k = 1; l - k >= 0; ++k) {
    const int lf = l - k;
    const int span = 1 << k;
    for (int ux = ((x - 2) << k); ux < ((x + 3) << k); ++ux) {
      if (!(((ux - 1) >> k) - 1 <= x && x <= ((ux + 1) >> k) + 1)) continue;
      for (int uy = ((y - 2) << k); uy < ((y + 3) << k); ++uy) {
        if (!(((uy - 1) >> k) - 1 <= y && y <= ((uy + 1) >> k) + 1)) continue;
        for (int uz = ((z - 2) << k); uz < ((z + 3) << k); ++uz) {
          if (!(((uz - 1) >> k) - 1 <= z && z <= ((uz + 1) >> k) + 1)) continue;
          const int j = s.find(lf, morton3(ux, uy, uz));
          if (j >= 0) out.push_back((int32_t)(s.offset[lf] + j));
        }
      }
    }
    (void)span;
  }
  std::sort(out.begin(), out.end());
  std::memcpy(cols, out.data(), out.size() * sizeof(int32_t));
  return (int64_t)out.size();
}

int nksr_cpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
