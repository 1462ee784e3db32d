/*
 * sgtd_oracle.cpp -- CPU oracle for stages 2-4 of the SGTD hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see sgtd_oracle.h).  PARITY UNPINNED: the
 * reference has no tests/golden vectors for this path; the restatement is
 * anchored on the reference source lines cited at each function.
 *
 * Build:  g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC  (oracle/Makefile)
 * The reference is built with plain "-O3" for baseline x86-64 (no -march, no
 * FMA; R/CMakeLists.txt:5-7), so every float/double expression below is
 * evaluated operation by operation in the order the reference source (and
 * Eigen's fixed-size unrolled evaluators) performs it.
 *
 * R = /root/reference/src/sgtd
 */
#include "sgtd_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Key4 {
  int64_t x, y, z, a; /* STDesc_LOC: equality on x,y,z,a only (STDesc.h:229-236) */
  bool operator==(const Key4 &o) const {
    return x == o.x && y == o.y && z == o.z && a == o.a;
  }
};
struct Key4Hash {
  /* STDesc.h:241-246.  The value of the hash never reaches a result. */
  size_t operator()(const Key4 &s) const {
    const int64_t HASH_P = 116101, MAX_N = 10000000000LL;
    return (size_t)(((((((s.z * HASH_P) % MAX_N + s.y) * HASH_P) % MAX_N + s.x) *
                      HASH_P) % MAX_N + s.a));
  }
};
struct Key3 {
  int64_t x, y, z; /* VOXEL_LOC (STDesc.h:126-136) */
  bool operator==(const Key3 &o) const { return x == o.x && y == o.y && z == o.z; }
};
struct Key3Hash {
  size_t operator()(const Key3 &s) const {
    const int64_t HASH_P = 116101, MAX_N = 10000000000LL;
    return (size_t)((((s.z * HASH_P) % MAX_N + s.y) * HASH_P) % MAX_N + s.x);
  }
};

/* Combinatorial_Binary_Encoding, R/src/STDesc.cpp:3-16: three 4-bit fields. */
inline int cbe(int a, int b, int c) {
  return ((a & 15) << 8) | ((b & 15) << 4) | (c & 15);
}

inline double norm3(double x, double y, double z) {
  /* Eigen Vector3d::norm() = sqrt(cwiseAbs2().sum()); a fixed-size reduction of three terms is
   * unrolled as e0 + (e1 + e2) (Eigen 3.3 Redux.h, redux_novec_unroller: halves of len/2 and
   * len - len/2) -- the order oracle/shim/Eigen/Core gives the reference build in oracle/_ref. */
  return std::sqrt(x * x + (y * y + z * z));
}

} // namespace

struct orc_handle {
  orc_config cfg;
  uint32_t current_frame_id = 0;
  std::vector<orc_desc> db; /* insertion order == global index g */
  std::unordered_map<Key4, std::vector<uint32_t>, Key4Hash> buckets;
};

/* ------------------------------------------------------------------------ */
/* 3x3 JacobiSVD restated from Eigen 3.3 (Eigen/src/SVD/JacobiSVD.h,
 * Eigen/src/Jacobi/Jacobi.h) -- third-party, not under /root/reference; call
 * site R/src/STDesc.cpp:560-563.  Square input => no QR preconditioner.
 * Matrices are row-major double[9].                                        */
namespace {

struct Rot {
  double c, s;
}; /* J = [c s; -s c] */

inline Rot rot_transpose(Rot j) { return Rot{j.c, -j.s}; }
inline Rot rot_mul(Rot a, Rot b) {
  /* JacobiRotation::operator* (real) */
  return Rot{a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c};
}
/* apply_rotation_in_the_plane(x, y, j): x' = c x + s y ; y' = -s x + c y */
inline void apply_rows(double *M, int n, int p, int q, Rot j) {
  if (j.c == 1.0 && j.s == 0.0) return;
  for (int i = 0; i < n; ++i) {
    double xi = M[p * n + i], yi = M[q * n + i];
    M[p * n + i] = j.c * xi + j.s * yi;
    M[q * n + i] = -j.s * xi + j.c * yi;
  }
}
inline void apply_cols(double *M, int n, int p, int q, Rot jr) {
  /* applyOnTheRight(p,q,j) == apply_rotation_in_the_plane(col p, col q, j^T) */
  Rot j = rot_transpose(jr);
  if (j.c == 1.0 && j.s == 0.0) return;
  for (int i = 0; i < n; ++i) {
    double xi = M[i * n + p], yi = M[i * n + q];
    M[i * n + p] = j.c * xi + j.s * yi;
    M[i * n + q] = -j.s * xi + j.c * yi;
  }
}
/* JacobiRotation::makeJacobi(x, y, z) for the symmetric 2x2 [x y; y z] */
inline Rot make_jacobi(double x, double y, double z) {
  double deno = 2.0 * std::fabs(y);
  if (deno < DBL_MIN) return Rot{1.0, 0.0};
  double tau = (x - z) / deno;
  double w = std::sqrt(tau * tau + 1.0);
  double t = (tau > 0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
  double sign_t = t > 0 ? 1.0 : -1.0;
  double n = 1.0 / std::sqrt(t * t + 1.0);
  Rot r;
  r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  r.c = n;
  return r;
}
/* internal::real_2x2_jacobi_svd */
inline void real_2x2_jacobi_svd(const double *W, int n, int p, int q, Rot *jl,
                                Rot *jr) {
  double m[4] = {W[p * n + p], W[p * n + q], W[q * n + p], W[q * n + q]};
  Rot rot1;
  double t = m[0] + m[3];
  double d = m[2] - m[1];
  if (std::fabs(d) < DBL_MIN) {
    rot1.s = 0.0;
    rot1.c = 1.0;
  } else {
    double u = t / d;
    double tmp = std::sqrt(1.0 + u * u);
    rot1.s = 1.0 / tmp;
    rot1.c = u / tmp;
  }
  apply_rows(m, 2, 0, 1, rot1);
  *jr = make_jacobi(m[0], m[1], m[3]);
  *jl = rot_mul(rot1, rot_transpose(*jr));
}

void jacobi_svd3(const double A[9], double U[9], double sv[3], double V[9]) {
  const int n = 3;
  const double precision = 2.0 * DBL_EPSILON;
  const double considerAsZero = DBL_MIN;
  double scale = 0.0;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(A[i]));
  if (scale == 0.0) scale = 1.0;
  double W[9];
  for (int i = 0; i < 9; ++i) W[i] = A[i] / scale;
  for (int i = 0; i < 9; ++i) U[i] = V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  double maxDiag = std::max(std::fabs(W[0]), std::max(std::fabs(W[4]), std::fabs(W[8])));
  bool finished = false;
  while (!finished) {
    finished = true;
    for (int p = 1; p < n; ++p) {
      for (int q = 0; q < p; ++q) {
        double threshold = std::max(considerAsZero, precision * maxDiag);
        if (std::fabs(W[p * n + q]) > threshold || std::fabs(W[q * n + p]) > threshold) {
          finished = false;
          Rot jl, jr;
          real_2x2_jacobi_svd(W, n, p, q, &jl, &jr);
          apply_rows(W, n, p, q, jl);                   /* W.applyOnTheLeft(p,q,jl)   */
          {                                             /* U.applyOnTheRight(p,q,jl^T) */
            Rot jt = rot_transpose(jl);
            apply_cols(U, n, p, q, jt);
          }
          apply_cols(W, n, p, q, jr);                   /* W.applyOnTheRight(p,q,jr)  */
          apply_cols(V, n, p, q, jr);                   /* V.applyOnTheRight(p,q,jr)  */
          maxDiag = std::max(maxDiag, std::max(std::fabs(W[p * n + p]), std::fabs(W[q * n + q])));
        }
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double a = std::fabs(W[i * n + i]);
    sv[i] = a;
    if (a != 0.0) {
      double f = W[i * n + i] / a;
      for (int r = 0; r < n; ++r) U[r * n + i] *= f;
    }
  }
  for (int i = 0; i < n; ++i) sv[i] *= scale;
  for (int i = 0; i < n; ++i) {
    int pos = i;
    double mx = sv[i];
    for (int k = i + 1; k < n; ++k)
      if (sv[k] > mx) { mx = sv[k]; pos = k; }
    if (mx == 0.0) break;
    if (pos != i) {
      std::swap(sv[i], sv[pos]);
      for (int r = 0; r < n; ++r) {
        std::swap(U[r * n + i], U[r * n + pos]);
        std::swap(V[r * n + i], V[r * n + pos]);
      }
    }
  }
}

/* C = A * B^T: Eigen's coefficient-based small product, each coefficient the reduction
 * a0*b0 + (a1*b1 + a2*b2) (same unrolled tree as norm3). */
inline void mul_abt(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[i * 3 + j] = A[i * 3 + 0] * B[j * 3 + 0] +
                     (A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2]);
}
inline double det3(const double m[9]) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
         m[2] * (m[3] * m[7] - m[4] * m[6]);
}
inline void matvec(const double R[9], const double v[3], double o[3]) {
  for (int i = 0; i < 3; ++i)
    o[i] = R[i * 3 + 0] * v[0] + (R[i * 3 + 1] * v[1] + R[i * 3 + 2] * v[2]);
}

/* triangle_solver, R/src/STDesc.cpp:549-571 */
void triangle_solver(const orc_desc &s, const orc_desc &r, double R[9], double t[3]) {
  double sv[9], rv[9], sc[3], rc[3];
  for (int i = 0; i < 9; ++i) { sv[i] = (double)s.vert[i]; rv[i] = (double)r.vert[i]; }
  /* center_ = (A + B + C) / 3   (STDesc.cpp:296) */
  for (int k = 0; k < 3; ++k) {
    sc[k] = ((sv[k] + sv[3 + k]) + sv[6 + k]) / 3.0;
    rc[k] = ((rv[k] + rv[3 + k]) + rv[6 + k]) / 3.0;
  }
  double src[9], ref[9]; /* row-major 3x3, column c = vertex c - center */
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < 3; ++k) {
      src[k * 3 + c] = sv[c * 3 + k] - sc[k];
      ref[k * 3 + c] = rv[c * 3 + k] - rc[k];
    }
  double cov[9];
  mul_abt(src, ref, cov); /* covariance = src * ref^T */
  double U[9], S[3], V[9];
  jacobi_svd3(cov, U, S, V);
  mul_abt(V, U, R); /* rot = V * U^T */
  if (det3(R) < 0) {
    double VK[9];
    for (int i = 0; i < 3; ++i) {
      VK[i * 3 + 0] = V[i * 3 + 0];
      VK[i * 3 + 1] = V[i * 3 + 1];
      VK[i * 3 + 2] = -V[i * 3 + 2];
    }
    mul_abt(VK, U, R); /* V * K * U^T */
  }
  /* t = -rot * center_src + center_ref */
  double nR[9], tmp[3];
  for (int i = 0; i < 9; ++i) nR[i] = -R[i];
  matvec(nR, sc, tmp);
  for (int k = 0; k < 3; ++k) t[k] = tmp[k] + rc[k];
}

/* the residual test of candidate_verify, R/src/STDesc.cpp:487-501 */
inline bool pair_is_inlier(const double R[9], const double t[3], const orc_desc &a,
                           const orc_desc &b) {
  const double dis_threshold = 3.0;
  for (int v = 0; v < 3; ++v) {
    double p[3] = {(double)a.vert[v * 3], (double)a.vert[v * 3 + 1], (double)a.vert[v * 3 + 2]};
    double q[3];
    matvec(R, p, q);
    double dx = (q[0] + t[0]) - (double)b.vert[v * 3];
    double dy = (q[1] + t[1]) - (double)b.vert[v * 3 + 1];
    double dz = (q[2] + t[2]) - (double)b.vert[v * 3 + 2];
    if (!(norm3(dx, dy, dz) < dis_threshold)) return false;
  }
  return true;
}

struct Match {
  int32_t q;
  uint8_t cell;
  uint32_t g;
};

} // namespace

/* ------------------------------------------------------------------------ */
extern "C" {

orc_handle *orc_create(const orc_config *cfg) {
  orc_handle *h = new orc_handle();
  h->cfg = *cfg;
  return h;
}
void orc_destroy(orc_handle *h) { delete h; }
uint32_t orc_current_frame_id(const orc_handle *h) { return h->current_frame_id; }
int64_t orc_db_size(const orc_handle *h) { return (int64_t)h->db.size(); }

void orc_jacobi_svd3(const double A[9], double U[9], double s[3], double V[9]) {
  jacobi_svd3(A, U, s, V);
}
void orc_triangle_solver(const orc_desc *src, const orc_desc *ref, double R[9], double t[3]) {
  triangle_solver(*src, *ref, R, t);
}

/* BuildSingleScanSTD, R/src/STDesc.cpp:174-315.
 * kNN: PCL KdTreeFLANN -> FLANN KDTreeSingleIndex, L2_Simple<float> (third
 * party, absent): exact k nearest, float32 accumulate (dx*dx + dy*dy) + dz*dz,
 * ascending.  Tie order is FLANN-traversal dependent and therefore unpinned;
 * this oracle breaks ties by lower node index (SURVEY 8a note ii).          */
int64_t orc_build(orc_handle *h, const float *xyz, const uint32_t *label, int32_t K,
                  orc_desc *out, int64_t cap) {
  const orc_config &c = h->cfg;
  const double scale = 1.0 / c.std_side_resolution;
  const int near_num = c.descriptor_near_num;
  const double max_dis = c.descriptor_max_len, min_dis = c.descriptor_min_len;
  if (K < near_num) return -1; /* reference reads stale indices here (UB) */
  std::unordered_map<Key3, bool, Key3Hash> feat_map;
  std::vector<std::pair<float, int>> dist(K);
  std::vector<int> nn(near_num);
  int64_t n_out = 0;
  for (int i = 0; i < K; ++i) {
    const float qx = xyz[i * 3], qy = xyz[i * 3 + 1], qz = xyz[i * 3 + 2];
    for (int j = 0; j < K; ++j) {
      float dx = qx - xyz[j * 3], dy = qy - xyz[j * 3 + 1], dz = qz - xyz[j * 3 + 2];
      float r = dx * dx;
      r += dy * dy;
      r += dz * dz;
      dist[j] = std::make_pair(r, j);
    }
    std::partial_sort(dist.begin(), dist.begin() + near_num, dist.end());
    for (int k = 0; k < near_num; ++k) nn[k] = dist[k].second;
    for (int m = 1; m < near_num - 1; ++m) {
      for (int n = m + 1; n < near_num; ++n) {
        const float *p1 = &xyz[i * 3], *p2 = &xyz[nn[m] * 3], *p3 = &xyz[nn[n] * 3];
        const uint32_t lab1 = label[i], lab2 = label[nn[m]], lab3 = label[nn[n]];
        /* float subtraction, then pow(double,2) == exact square (:198-203) */
        auto side = [](const float *u, const float *v) {
          double dx = (double)(u[0] - v[0]), dy = (double)(u[1] - v[1]),
                 dz = (double)(u[2] - v[2]);
          return std::sqrt((dx * dx + dy * dy) + dz * dz);
        };
        double a = side(p1, p2), b = side(p1, p3), cc = side(p3, p2);
        if (a > max_dis || b > max_dis || cc > max_dis || a < min_dis || b < min_dis ||
            cc < min_dis)
          continue;
        int l1[3] = {1, 2, 0}, l2[3] = {1, 0, 3}, l3[3] = {0, 2, 3};
        auto swp = [](int *u, int *v) { for (int k = 0; k < 3; ++k) std::swap(u[k], v[k]); };
        if (a > b) { std::swap(a, b); swp(l1, l2); }
        if (b > cc) { std::swap(b, cc); swp(l2, l3); }
        if (a > b) { std::swap(a, b); swp(l1, l2); }
        /* pcl::PointXYZ d_p (float) then (int64_t) truncation (:244-248) */
        float fx = (float)(a * 1000), fy = (float)(b * 1000), fz = (float)(cc * 1000);
        Key3 pos{(int64_t)fx, (int64_t)fy, (int64_t)fz};
        if (feat_map.find(pos) != feat_map.end()) continue;
        const float *A, *B, *C;
        uint32_t la, lb, lc;
        if (l1[0] == l2[0]) { A = p1; la = lab1; }
        else if (l1[1] == l2[1]) { A = p2; la = lab2; }
        else { A = p3; la = lab3; }
        if (l1[0] == l3[0]) { B = p1; lb = lab1; }
        else if (l1[1] == l3[1]) { B = p2; lb = lab2; }
        else { B = p3; lb = lab3; }
        if (l2[0] == l3[0]) { C = p1; lc = lab1; }
        else if (l2[1] == l3[1]) { C = p2; lc = lab2; }
        else { C = p3; lc = lab3; }
        feat_map[pos] = true;
        if (n_out < cap) {
          orc_desc &d = out[n_out];
          std::memset(&d, 0, sizeof(d));
          d.side[0] = scale * a; d.side[1] = scale * b; d.side[2] = scale * cc;
          for (int k = 0; k < 3; ++k) { d.vert[k] = A[k]; d.vert[3 + k] = B[k]; d.vert[6 + k] = C[k]; }
          d.frame = h->current_frame_id;
          d.lab[0] = (uint8_t)la; d.lab[1] = (uint8_t)lb; d.lab[2] = (uint8_t)lc;
          d.anchor = (uint16_t)i; d.m = (uint8_t)m; d.n = (uint8_t)n;
        }
        ++n_out;
      }
    }
  }
  return n_out;
}

void orc_db_key(const orc_desc *d, int32_t out[4]) {
  /* R/src/STDesc.cpp:155-161 */
  out[0] = (int)(d->side[0] + 0.5);
  out[1] = (int)(d->side[1] + 0.5);
  out[2] = (int)(d->side[2] + 0.5);
  out[3] = cbe(d->lab[0], d->lab[1], d->lab[2]);
}

/* AddSTDescs, R/src/STDesc.cpp:149-172 */
void orc_add(orc_handle *h, const orc_desc *d, int64_t n) {
  h->current_frame_id++;
  for (int64_t i = 0; i < n; ++i) {
    int32_t k[4];
    orc_db_key(&d[i], k);
    Key4 pos{k[0], k[1], k[2], k[3]};
    uint32_t g = (uint32_t)h->db.size();
    h->db.push_back(d[i]);
    h->buckets[pos].push_back(g);
  }
}

/* The map phase of the node (R/src/semantic_graph_localization.cpp:455-458) for many keyframes:
 * BuildSingleScanSTD of every scan (independent: done on `nthreads` threads, each scan stamped with the
 * frame id it will be added as), then AddSTDescs in scan order.  A scan with fewer nodes than
 * descriptor_near_num is added as an empty keyframe.  Returns the number of descriptors added. */
int64_t orc_build_add_many(orc_handle *h, const float *xyz, const uint32_t *label, const int64_t *off,
                           int32_t nscans, int32_t nthreads) {
  std::vector<std::vector<orc_desc>> per((size_t)nscans);
  const uint32_t base = h->current_frame_id;
  const int npairs = (h->cfg.descriptor_near_num - 1) * (h->cfg.descriptor_near_num - 2) / 2;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1)
  for (int32_t s = 0; s < nscans; ++s) {
    orc_handle tmp;
    tmp.cfg = h->cfg;
    tmp.current_frame_id = base + (uint32_t)s;
    const int32_t K = (int32_t)(off[s + 1] - off[s]);
    std::vector<orc_desc> &v = per[(size_t)s];
    v.resize((size_t)std::max(K, 1) * npairs);
    const int64_t n = orc_build(&tmp, xyz + 3 * off[s], label + off[s], K, v.data(), (int64_t)v.size());
    v.resize((size_t)std::max<int64_t>(n, 0));
  }
  int64_t total = 0;
  for (int32_t s = 0; s < nscans; ++s) {
    orc_add(h, per[(size_t)s].data(), (int64_t)per[(size_t)s].size());
    total += (int64_t)per[(size_t)s].size();
    std::vector<orc_desc>().swap(per[(size_t)s]);
  }
  return total;
}

/* SearchLoop, R/src/STDesc.cpp:84-147 */
int32_t orc_search(orc_handle *h, const orc_desc *q, int64_t nq, orc_cand *cands,
                   int32_t cap_cand, int32_t *m_q, uint8_t *m_cell, uint32_t *m_g,
                   int32_t *inl, int64_t cap_match, int32_t *votes_out, int64_t n_frames,
                   double best[2], orc_search_stats *stats, int32_t nthreads) {
  const orc_config &c = h->cfg;
  best[0] = -1; best[1] = 0;
  if (stats) std::memset(stats, 0, sizeof(*stats));
  if (nq == 0) return -1; /* "No STDescs!" (:89-93) */
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
  omp_set_num_threads(nthreads);
#endif
  /* ---- candidate_selector, R/src/STDesc.cpp:318-460 ---- */
  std::vector<std::vector<Match>> per_q(nq);
  int64_t sP = 0, sPf = 0, sE = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : sP, sPf, sE)
  for (int64_t i = 0; i < nq; ++i) {
    const orc_desc &s = q[i];
    const double dis_threshold = norm3(s.side[0], s.side[1], s.side[2]) * c.rough_dis_threshold;
    const int code = cbe(s.lab[0], s.lab[1], s.lab[2]);
    int ordinal = 0;
    for (int x = -1; x <= 1; ++x)
      for (int y = -1; y <= 1; ++y)
        for (int z = -1; z <= 1; ++z, ++ordinal) {
          Key4 pos;
          pos.x = (int)(s.side[0] + x); /* truncation toward zero (:359-361) */
          pos.y = (int)(s.side[1] + y);
          pos.z = (int)(s.side[2] + z);
          pos.a = code;
          double cx = (double)pos.x + 0.5, cy = (double)pos.y + 0.5, cz = (double)pos.z + 0.5;
          if (!(norm3(s.side[0] - cx, s.side[1] - cy, s.side[2] - cz) < 1.5)) continue;
          ++sP;
          auto it = h->buckets.find(pos);
          if (it == h->buckets.end()) continue;
          ++sPf;
          const std::vector<uint32_t> &bk = it->second;
          sE += (int64_t)bk.size();
          for (size_t j = 0; j < bk.size(); ++j) {
            const orc_desc &d = h->db[bk[j]];
            if ((uint32_t)(s.frame - d.frame) > 0) { /* unsigned: frame ids differ (:373) */
              double dis = norm3(s.side[0] - d.side[0], s.side[1] - d.side[1], s.side[2] - d.side[2]);
              if (dis < dis_threshold) per_q[i].push_back(Match{(int32_t)i, (uint8_t)ordinal, bk[j]});
            }
          }
        }
  }
  /* votes (:405-420); MAX_FRAME_N replaced by the real frame count (SURVEY 8a iii) */
  const int64_t F = (int64_t)h->current_frame_id;
  std::vector<int32_t> votes((size_t)std::max<int64_t>(F, 1), 0);
  std::vector<Match> all;
  std::vector<int32_t> match_index_vec; /* frame id of every match, as the reference keeps it (:415-417) */
  for (int64_t i = 0; i < nq; ++i)
    for (const Match &m : per_q[i]) {
      const int32_t f = (int32_t)h->db[m.g].frame;
      votes[f] += 1;
      all.push_back(m);
      match_index_vec.push_back(f);
    }
  if (stats) { stats->Q = nq; stats->P = sP; stats->Pfound = sPf; stats->E = sE; stats->M = (int64_t)all.size(); }
  if (votes_out)
    for (int64_t f = 0; f < std::min(F, n_frames); ++f) votes_out[f] = votes[f];
  /* ranking (:423-453) */
  int32_t ncand = 0;
  int64_t moff = 0, ioff = 0;
  for (int cnt = 0; cnt < c.candidate_num; ++cnt) {
    int32_t max_vote = 1, max_idx = -1;
    for (int64_t f = 0; f < F; ++f)
      if (votes[f] > max_vote) { max_vote = votes[f]; max_idx = (int32_t)f; }
    if (!(max_idx >= 0 && max_vote >= 5)) break;
    votes[max_idx] = 0;
    if (ncand >= cap_cand) return -2;
    orc_cand &cd = cands[ncand];
    std::memset(&cd, 0, sizeof(cd));
    cd.frame = max_idx; cd.votes = max_vote; cd.match_off = (int32_t)moff;
    int32_t nm = 0;
    for (size_t mi = 0; mi < all.size(); ++mi)
      if (match_index_vec[mi] == max_idx) { /* :437-447 */
        const Match &m = all[mi];
        if (moff + nm >= cap_match) return -2;
        m_q[moff + nm] = m.q; m_cell[moff + nm] = m.cell; m_g[moff + nm] = m.g;
        ++nm;
      }
    cd.nmatch = nm;
    moff += nm;
    ++ncand;
  }
  /* ---- candidate_verify for every candidate (:108-131, :462-547) ---- */
  double best_score = 0;
  int best_id = -1;
  for (int32_t ci = 0; ci < ncand; ++ci) {
    orc_cand &cd = cands[ci];
    const int32_t M = cd.nmatch;
    const int32_t *mq = m_q + cd.match_off;
    const uint32_t *mg = m_g + cd.match_off;
    const int skip_len = (int)(M / 50) + 1;
    const int use_size = M / skip_len;
    std::vector<int> vote_list(use_size, 0);
#pragma omp parallel for schedule(static)
    for (int hy = 0; hy < use_size; ++hy) {
      double R[9], t[3];
      triangle_solver(q[mq[hy * skip_len]], h->db[mg[hy * skip_len]], R, t);
      int vote = 0;
      for (int j = 0; j < M; ++j)
        if (pair_is_inlier(R, t, q[mq[j]], h->db[mg[j]])) vote++;
      vote_list[hy] = vote;
    }
    int max_vote_index = 0, max_vote = 0;
    for (int hy = 0; hy < use_size; ++hy)
      if (max_vote < vote_list[hy]) { max_vote_index = hy; max_vote = vote_list[hy]; }
    cd.inlier_off = (int32_t)ioff;
    double verify_score = -1;
    if (max_vote >= 4) {
      cd.best_hyp = max_vote_index;
      triangle_solver(q[mq[max_vote_index * skip_len]], h->db[mg[max_vote_index * skip_len]], cd.R, cd.t);
      int32_t ni = 0;
      for (int j = 0; j < M; ++j)
        if (pair_is_inlier(cd.R, cd.t, q[mq[j]], h->db[mg[j]])) inl[ioff + ni++] = j;
      cd.ninlier = ni;
      ioff += ni;
      verify_score = ni;
    } else {
      cd.best_hyp = -1;
      cd.ninlier = 0;
      /* relative_pose is left default-constructed by the reference; report identity */
      cd.R[0] = cd.R[4] = cd.R[8] = 1.0;
    }
    cd.score = (int)verify_score;
    if (verify_score > best_score) { best_score = verify_score; best_id = cd.frame; }
  }
  if (best_score > c.icp_threshold) { best[0] = best_id; best[1] = best_score; }
  return ncand;
}

} /* extern "C" */
