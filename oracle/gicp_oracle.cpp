/*
 * gicp_oracle.cpp -- CPU oracle for the GICP refinement of the verified candidates.
 *
 * TEST INFRASTRUCTURE ONLY (see sgtd_oracle.h).  PARITY UNPINNED for this part: the reference's
 * fast_gicp sits on PCL (kd-tree search, Registration base class, transformPointCloud,
 * getFitnessScore) and Eigen (JacobiSVD, LDLT, 4x4 inverse), none of which is installed, and no
 * fixture exists.  Restated from the cited lines, floating point: the CUDA path is compared with
 * this file within the tolerance written in tests/test_gpu_gicp.py.
 *
 *   FastGICP::calculate_covariances   R/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp:251-301
 *   FastGICP::update_correspondences  :119-156
 *   FastGICP::linearize / compute_error :158-247
 *   LsqRegistration::computeTransformation / is_converged / step_lm
 *                                      R/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:53-166
 *   se3_exp / so3_exp / skewd          R/include/fast_gicp/so3/so3.hpp:21-104
 *   the node's multi-candidate loop    R/src/semantic_graph_localization.cpp:651-721
 *   pcl::Registration::getFitnessScore (PCL 1.12, restated: mean squared nearest-neighbour distance of
 *                                      the transformed source, float distances accumulated in double)
 * R = /root/reference/src/sgtd.  Nearest-neighbour searches are exact (brute force); distance ties go
 * to the lower index (kd-tree traversal order is not reproducible).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "sgtd_oracle.h"

namespace {

struct Iso { double R[9]; double t[3]; };  // x -> R x + t

Iso iso_identity() { Iso T{}; T.R[0] = T.R[4] = T.R[8] = 1.0; return T; }
Iso iso_mul(const Iso &A, const Iso &B) {  // A * B
  Iso C{};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) C.R[i * 3 + j] += A.R[i * 3 + k] * B.R[k * 3 + j];
    C.t[i] = A.t[i];
    for (int k = 0; k < 3; ++k) C.t[i] += A.R[i * 3 + k] * B.t[k];
  }
  return C;
}

/* float L2 (FLANN L2_Simple: ((dx*dx)+dy*dy)+dz*dz), exact k nearest, ascending (distance, index) */
void knn(const float *pts, int64_t n, const float q[3], int k, std::vector<std::pair<float, int>> &out) {
  out.resize((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    const float dx = q[0] - pts[3 * i], dy = q[1] - pts[3 * i + 1], dz = q[2] - pts[3 * i + 2];
    float d = dx * dx; d += dy * dy; d += dz * dz;
    out[(size_t)i] = std::make_pair(d, (int)i);
  }
  const size_t kk = (size_t)std::min<int64_t>(k, n);
  std::partial_sort(out.begin(), out.begin() + kk, out.end());
  out.resize(kk);
}

/* calculate_covariances with RegularizationMethod::PLANE (fast_gicp_impl.hpp:251-301) -> 3x3 blocks */
void covariances(const float *pts, int64_t n, int k, std::vector<double> &cov) {
  cov.assign((size_t)n * 9, 0.0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; ++i) {
    std::vector<std::pair<float, int>> nn;
    knn(pts, n, pts + 3 * i, k, nn);
    /* neighbors is 4 x k_correspondences_; columns beyond the neighbours found stay zero (:263-266 leaves
     * them uninitialised in the reference; clouds here always have at least k points) */
    double mean[3] = {0, 0, 0};
    for (auto &p : nn) for (int a = 0; a < 3; ++a) mean[a] += (double)pts[3 * p.second + a];
    for (int a = 0; a < 3; ++a) mean[a] /= (double)k;
    double c[9] = {0};
    for (auto &p : nn) {
      double d[3];
      for (int a = 0; a < 3; ++a) d[a] = (double)pts[3 * p.second + a] - mean[a];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) c[a * 3 + b] += d[a] * d[b];
    }
    for (int a = 0; a < 9; ++a) c[a] /= (double)k;
    double U[9], s[3], V[9];
    orc_jacobi_svd3(c, U, s, V);
    const double vals[3] = {1.0, 1.0, 1e-3}; /* PLANE */
    double *o = &cov[(size_t)i * 9];
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) {
        double v = 0;
        for (int m = 0; m < 3; ++m) v += U[a * 3 + m] * vals[m] * V[b * 3 + m];
        o[a * 3 + b] = v;
      }
  }
}

bool inv3(const double m[9], double o[9]) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return det != 0.0;
}

struct Gicp {
  const float *src; int64_t ns;
  const float *tgt; int64_t nt;
  std::vector<double> cov_s, cov_t;
  std::vector<int> corr;
  std::vector<double> mahal; /* 9 per source point */

  /* update_correspondences, fast_gicp_impl.hpp:119-156 (corr_dist_threshold_ = float max) */
  void update_correspondences(const Iso &T) {
    float Rf[9], tf[3];
    for (int i = 0; i < 9; ++i) Rf[i] = (float)T.R[i];
    for (int i = 0; i < 3; ++i) tf[i] = (float)T.t[i];
    corr.assign((size_t)ns, -1);
    mahal.assign((size_t)ns * 9, 0.0);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < ns; ++i) {
      const float *p = src + 3 * i;
      float q[3];
      for (int a = 0; a < 3; ++a) q[a] = ((Rf[a * 3] * p[0] + Rf[a * 3 + 1] * p[1]) + Rf[a * 3 + 2] * p[2]) + tf[a];
      float best = 0; int bi = -1;
      for (int64_t j = 0; j < nt; ++j) {
        const float dx = q[0] - tgt[3 * j], dy = q[1] - tgt[3 * j + 1], dz = q[2] - tgt[3 * j + 2];
        float d = dx * dx; d += dy * dy; d += dz * dz;
        if (bi < 0 || d < best) { best = d; bi = (int)j; }
      }
      corr[(size_t)i] = bi;
      if (bi < 0) continue;
      /* RCR = cov_B + T cov_A T^T ; the 4th row / column only carries the 1 that makes it invertible */
      const double *A = &cov_s[(size_t)i * 9], *B = &cov_t[(size_t)bi * 9];
      double RA[9], RCR[9];
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { double v = 0; for (int m = 0; m < 3; ++m) v += T.R[a * 3 + m] * A[m * 3 + b]; RA[a * 3 + b] = v; }
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { double v = 0; for (int m = 0; m < 3; ++m) v += RA[a * 3 + m] * T.R[b * 3 + m]; RCR[a * 3 + b] = B[a * 3 + b] + v; }
      inv3(RCR, &mahal[(size_t)i * 9]);
    }
  }

  /* linearize / compute_error, :158-247.  H, b may be null. */
  double linearize(const Iso &T, bool update, double *H, double *b) {
    if (update) update_correspondences(T);
    double sum = 0.0;
    if (H) { std::memset(H, 0, 36 * sizeof(double)); std::memset(b, 0, 6 * sizeof(double)); }
    for (int64_t i = 0; i < ns; ++i) {
      const int j = corr[(size_t)i];
      if (j < 0) continue;
      double a[3], e[3];
      for (int r = 0; r < 3; ++r)
        a[r] = T.R[r * 3] * (double)src[3 * i] + T.R[r * 3 + 1] * (double)src[3 * i + 1] + T.R[r * 3 + 2] * (double)src[3 * i + 2] + T.t[r];
      for (int r = 0; r < 3; ++r) e[r] = (double)tgt[3 * j + r] - a[r];
      const double *M = &mahal[(size_t)i * 9];
      double Me[3];
      for (int r = 0; r < 3; ++r) Me[r] = M[r * 3] * e[0] + M[r * 3 + 1] * e[1] + M[r * 3 + 2] * e[2];
      sum += e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
      if (!H) continue;
      /* J = [ skew(T a) | -I ]  (3 x 6) */
      double J[18] = {0, -a[2], a[1], -1, 0, 0, a[2], 0, -a[0], 0, -1, 0, -a[1], a[0], 0, 0, 0, -1};
      double MJ[18];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 6; ++c) MJ[r * 6 + c] = M[r * 3] * J[c] + M[r * 3 + 1] * J[6 + c] + M[r * 3 + 2] * J[12 + c];
      for (int r = 0; r < 6; ++r) {
        for (int c = 0; c < 6; ++c) H[r * 6 + c] += J[r] * MJ[c] + J[6 + r] * MJ[6 + c] + J[12 + r] * MJ[12 + c];
        b[r] += J[r] * Me[0] + J[6 + r] * Me[1] + J[12 + r] * Me[2];
      }
    }
    return sum;
  }
};

/* (H) d = rhs for a symmetric positive definite 6x6: LDL^T */
void solve6(const double *Hin, const double *rhs, double *d) {
  double L[36] = {0}, D[6];
  for (int j = 0; j < 6; ++j) {
    double v = Hin[j * 6 + j];
    for (int k = 0; k < j; ++k) v -= L[j * 6 + k] * L[j * 6 + k] * D[k];
    D[j] = v;
    L[j * 6 + j] = 1.0;
    for (int i = j + 1; i < 6; ++i) {
      double w = Hin[i * 6 + j];
      for (int k = 0; k < j; ++k) w -= L[i * 6 + k] * L[j * 6 + k] * D[k];
      L[i * 6 + j] = w / D[j];
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) { double v = rhs[i]; for (int k = 0; k < i; ++k) v -= L[i * 6 + k] * y[k]; y[i] = v; }
  for (int i = 0; i < 6; ++i) y[i] /= D[i];
  for (int i = 5; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 6; ++k) v -= L[k * 6 + i] * d[k]; d[i] = v; }
}

/* se3_exp, so3.hpp:59-104 (rotation first) */
Iso se3_exp(const double a[6]) {
  const double w[3] = {a[0], a[1], a[2]};
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double q = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * q;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * q;
  } else {
    const double th = std::sqrt(theta_sq), h = 0.5 * th;
    imag = std::sin(h) / th;
    real = std::cos(h);
  }
  /* Eigen::Quaterniond(w, x, y, z).toRotationMatrix() (no normalisation) */
  const double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  Iso T{};
  T.R[0] = 1 - (tyy + tzz); T.R[1] = txy - twz; T.R[2] = txz + twy;
  T.R[3] = txy + twz; T.R[4] = 1 - (txx + tzz); T.R[5] = tyz - twx;
  T.R[6] = txz - twy; T.R[7] = tyz + twx; T.R[8] = 1 - (txx + tyy);
  const double theta = std::sqrt(theta_sq);
  double V[9];
  if (theta < 1e-10) {
    std::memcpy(V, T.R, sizeof(V));
  } else {
    const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double O2[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double v = 0; for (int k = 0; k < 3; ++k) v += O[i * 3 + k] * O[k * 3 + j]; O2[i * 3 + j] = v; }
    const double c1 = (1.0 - std::cos(theta)) / theta_sq, c2 = (theta - std::sin(theta)) / (theta_sq * theta);
    for (int i = 0; i < 9; ++i) V[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
  }
  for (int i = 0; i < 3; ++i) T.t[i] = V[i * 3] * a[3] + V[i * 3 + 1] * a[4] + V[i * 3 + 2] * a[5];
  return T;
}

bool is_converged(const Iso &d, double rot_eps, double trans_eps) { /* lsq_registration_impl.hpp:83-93 */
  double m = 0;
  for (int i = 0; i < 9; ++i) m = std::max(m, std::fabs(d.R[i] - ((i % 4 == 0) ? 1.0 : 0.0)) / rot_eps);
  for (int i = 0; i < 3; ++i) m = std::max(m, std::fabs(d.t[i]) / trans_eps);
  return m < 1;
}

} // namespace

extern "C" {

/* reg.setInputTarget(target); reg.setInputSource(init * source); reg.align() with the identity guess;
 * final16 = reg.getFinalTransformation() (row-major 4x4, values of the float matrix), *fitness =
 * reg.getFitnessScore().  src / tgt: n x 3 floats. */
int32_t orc_gicp_align(const float *src_in, int64_t ns, const float *tgt, int64_t nt, const double init12[12],
                       int32_t k, int32_t max_iterations, double rot_eps, double trans_eps, double final16[16],
                       double *fitness, int32_t *iterations, int32_t *converged) {
  if (ns < k || nt < k || k < 1 || k > 64) return -1;
  /* pcl::transformPointCloud(*src_cloud, *src_cloud, new_trans1): float */
  float Mf[12];
  for (int i = 0; i < 12; ++i) Mf[i] = (float)init12[i];
  std::vector<float> src((size_t)ns * 3);
  for (int64_t i = 0; i < ns; ++i)
    for (int a = 0; a < 3; ++a)
      src[3 * i + a] = ((Mf[a * 4] * src_in[3 * i] + Mf[a * 4 + 1] * src_in[3 * i + 1]) + Mf[a * 4 + 2] * src_in[3 * i + 2]) + Mf[a * 4 + 3];
  Gicp g;
  g.src = src.data(); g.ns = ns; g.tgt = tgt; g.nt = nt;
  covariances(src.data(), ns, k, g.cov_s);
  covariances(tgt, nt, k, g.cov_t);
  /* LsqRegistration::computeTransformation, LM (:53-81, :123-166) */
  Iso x0 = iso_identity();
  double lambda = -1.0;
  bool conv = false;
  int it = 0;
  for (int i = 0; i < max_iterations && !conv; ++i) {
    it = i;
    double H[36], b[6];
    const double y0 = g.linearize(x0, true, H, b);
    if (lambda < 0.0) { double m = 0; for (int d = 0; d < 6; ++d) m = std::max(m, std::fabs(H[d * 6 + d])); lambda = 1e-9 * m; }
    double nu = 2.0;
    bool ok = false;
    Iso delta = iso_identity();
    for (int j = 0; j < 10; ++j) {
      double Hl[36], nb[6], d[6];
      std::memcpy(Hl, H, sizeof(Hl));
      for (int q = 0; q < 6; ++q) { Hl[q * 6 + q] += lambda; nb[q] = -b[q]; }
      solve6(Hl, nb, d);
      delta = se3_exp(d);
      const Iso xi = iso_mul(delta, x0);
      const double yi = g.linearize(xi, false, nullptr, nullptr);
      double den = 0;
      for (int q = 0; q < 6; ++q) den += d[q] * (lambda * d[q] - b[q]);
      const double rho = (y0 - yi) / den;
      if (rho < 0) {
        if (is_converged(delta, rot_eps, trans_eps)) { ok = true; break; }
        lambda = nu * lambda;
        nu = 2 * nu;
        continue;
      }
      x0 = xi;
      lambda = lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      ok = true;
      break;
    }
    if (!ok) break; /* "lm not converged!!" */
    conv = is_converged(delta, rot_eps, trans_eps);
  }
  /* final_transformation_ = x0.cast<float>().matrix() */
  float Ff[12];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Ff[r * 4 + c] = (float)x0.R[r * 3 + c]; Ff[r * 4 + 3] = (float)x0.t[r]; }
  for (int i = 0; i < 12; ++i) final16[i] = (double)Ff[i];
  final16[12] = final16[13] = final16[14] = 0.0; final16[15] = 1.0;
  /* getFitnessScore(): mean squared distance of the transformed source to its nearest target point */
  double sum = 0.0;
#pragma omp parallel for reduction(+ : sum) schedule(dynamic, 64)
  for (int64_t i = 0; i < ns; ++i) {
    float q[3];
    for (int a = 0; a < 3; ++a) q[a] = ((Ff[a * 4] * src[3 * i] + Ff[a * 4 + 1] * src[3 * i + 1]) + Ff[a * 4 + 2] * src[3 * i + 2]) + Ff[a * 4 + 3];
    float best = 0; bool have = false;
    for (int64_t j = 0; j < nt; ++j) {
      const float dx = q[0] - tgt[3 * j], dy = q[1] - tgt[3 * j + 1], dz = q[2] - tgt[3 * j + 2];
      float d = dx * dx; d += dy * dy; d += dz * dz;
      if (!have || d < best) { best = d; have = true; }
    }
    sum += (double)best;
  }
  *fitness = sum / (double)ns;
  *iterations = it;
  *converged = conv ? 1 : 0;
  return 0;
}

} /* extern "C" */
