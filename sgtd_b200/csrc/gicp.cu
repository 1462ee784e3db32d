// gicp.cu -- GICP refinement of the verified candidates on sm_100a.
//
// Replaces fast_gicp::FastGICP as the node uses it (R/src/semantic_graph_localization.cpp:651-721):
//   calculate_covariances   R/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp:251-301  -> k_gicp_knn_cov
//   update_correspondences  :119-156                                                   -> k_gicp_correspond
//   linearize/compute_error :158-247                                                   -> k_gicp_linearize
//   LM driver               R/include/fast_gicp/gicp/impl/lsq_registration_impl.hpp:53-166 (host, 6x6)
//   getFitnessScore         (PCL)                                                      -> k_gicp_fitness
// The reference searches kd-trees; clouds here are a few thousand (down-sampled source) to ~1e5 (target)
// points, so exact brute-force searches over shared-memory tiles are both simpler and far faster on a
// B200 (an LM iteration tests n_src x n_tgt pairs, ~1e9 at most: < 1 ms).  Distance ties go to the lower
// index.  Arithmetic: float distances in FLANN's L2_Simple order, everything else in double like the
// reference; compiled -fmad=false.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "internal.cuh"
#include "svd3.cuh"

namespace sgtd {

constexpr int kGicpThreads = 128;
constexpr int kGicpMaxK = 64;

__device__ __forceinline__ float l2f(float qx, float qy, float qz, const float4 &p) {
  const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// pcl::transformPointCloud with the float 3x4 matrix M (row-major)
__global__ void k_gicp_transform(const float *src, int64_t n, const float *M, float4 *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
  float4 o;
  o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[0], x), __fmul_rn(M[1], y)), __fmul_rn(M[2], z)), M[3]);
  o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[4], x), __fmul_rn(M[5], y)), __fmul_rn(M[6], z)), M[7]);
  o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(M[8], x), __fmul_rn(M[9], y)), __fmul_rn(M[10], z)), M[11]);
  o.w = 1.f;
  out[i] = o;
}
__global__ void k_gicp_pack(const float *src, int64_t n, float4 *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 1.f);
}

// One thread per point: exact k nearest neighbours (the point itself included) over shared-memory tiles,
// kept as an ascending (distance, index) list; then mean, covariance / k, SVD and the PLANE
// regularisation U diag(1, 1, 1e-3) V^T.  cov: 9 doubles per point (row-major 3x3).
__global__ void __launch_bounds__(kGicpThreads) k_gicp_knn_cov(const float4 *pts, int64_t n, int k, double *cov) {
  __shared__ float4 s_tile[kGicpThreads];
  const int64_t i = (int64_t)blockIdx.x * kGicpThreads + threadIdx.x;
  const float4 q = pts[i < n ? i : n - 1];
  float dk[kGicpMaxK];
  int ik[kGicpMaxK];
  for (int j = 0; j < k; ++j) { dk[j] = __int_as_float(0x7f800000); ik[j] = 0x7fffffff; }
  float worst = __int_as_float(0x7f800000);
  for (int64_t t0 = 0; t0 < n; t0 += kGicpThreads) {
    __syncthreads();
    if (t0 + threadIdx.x < n) s_tile[threadIdx.x] = pts[t0 + threadIdx.x];
    __syncthreads();
    const int m = (int)min((int64_t)kGicpThreads, n - t0);
    for (int j = 0; j < m; ++j) {
      const float d = l2f(q.x, q.y, q.z, s_tile[j]);
      if (d < worst) {  // ascending scan of indices: an equal distance never displaces an earlier index
        int p = k - 1;
        while (p > 0 && dk[p - 1] > d) { dk[p] = dk[p - 1]; ik[p] = ik[p - 1]; --p; }
        dk[p] = d; ik[p] = (int)(t0 + j);
        worst = dk[k - 1];
      }
    }
  }
  if (i >= n) return;
  double mean[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < k; ++j) { const float4 p = pts[ik[j]]; mean[0] += (double)p.x; mean[1] += (double)p.y; mean[2] += (double)p.z; }
  mean[0] /= (double)k; mean[1] /= (double)k; mean[2] /= (double)k;
  double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < k; ++j) {
    const float4 p = pts[ik[j]];
    const double d[3] = {(double)p.x - mean[0], (double)p.y - mean[1], (double)p.z - mean[2]};
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) c[a * 3 + b] += d[a] * d[b];
  }
#pragma unroll
  for (int a = 0; a < 9; ++a) c[a] /= (double)k;
  double U[9], V[9];
  svd3(c, U, V);
  const double vals[3] = {1.0, 1.0, 1e-3};
  double *o = cov + 9 * i;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double v = 0.0;
#pragma unroll
      for (int m = 0; m < 3; ++m) v += U[a * 3 + m] * vals[m] * V[b * 3 + m];
      o[a * 3 + b] = v;
    }
}

struct Pose12 { double R[9], t[3]; };

// nearest target point of every (float-)transformed source point; optionally the Mahalanobis matrix
// (cov_B + R cov_A R^T)^-1 of the pair (update_correspondences), optionally only the squared distance
// (getFitnessScore).
__global__ void __launch_bounds__(kGicpThreads) k_gicp_correspond(const float4 *src, int64_t ns, const float4 *tgt, int64_t nt,
                                                                   Pose12 T, const double *cov_s, const double *cov_t, int *corr,
                                                                   double *mahal, float *sqd) {
  __shared__ float4 s_tile[kGicpThreads];
  const int64_t i = (int64_t)blockIdx.x * kGicpThreads + threadIdx.x;
  const float4 p = src[i < ns ? i : ns - 1];
  float Rf[9], tf[3];
#pragma unroll
  for (int a = 0; a < 9; ++a) Rf[a] = (float)T.R[a];
#pragma unroll
  for (int a = 0; a < 3; ++a) tf[a] = (float)T.t[a];
  const float qx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Rf[0], p.x), __fmul_rn(Rf[1], p.y)), __fmul_rn(Rf[2], p.z)), tf[0]);
  const float qy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Rf[3], p.x), __fmul_rn(Rf[4], p.y)), __fmul_rn(Rf[5], p.z)), tf[1]);
  const float qz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(Rf[6], p.x), __fmul_rn(Rf[7], p.y)), __fmul_rn(Rf[8], p.z)), tf[2]);
  float best = __int_as_float(0x7f800000);
  int bi = -1;
  for (int64_t t0 = 0; t0 < nt; t0 += kGicpThreads) {
    __syncthreads();
    if (t0 + threadIdx.x < nt) s_tile[threadIdx.x] = tgt[t0 + threadIdx.x];
    __syncthreads();
    const int m = (int)min((int64_t)kGicpThreads, nt - t0);
    for (int j = 0; j < m; ++j) {
      const float d = l2f(qx, qy, qz, s_tile[j]);
      if (d < best) { best = d; bi = (int)(t0 + j); }
    }
  }
  if (i >= ns) return;
  if (sqd) sqd[i] = best;
  if (!corr) return;
  corr[i] = bi;
  if (bi < 0) return;
  const double *A = cov_s + 9 * i, *B = cov_t + 9 * (int64_t)bi;
  double RA[9], m[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) RA[a * 3 + b] = T.R[a * 3] * A[b] + T.R[a * 3 + 1] * A[3 + b] + T.R[a * 3 + 2] * A[6 + b];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) m[a * 3 + b] = B[a * 3 + b] + (RA[a * 3] * T.R[b * 3] + RA[a * 3 + 1] * T.R[b * 3 + 1] + RA[a * 3 + 2] * T.R[b * 3 + 2]);
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double id = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
  double *o = mahal + 9 * i;
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// sum of e^T M e and, if want_h, of J^T M J (upper triangle, 21 values) and J^T M e (6): one partial per
// block, summed by k_gicp_reduce in block order (deterministic).
constexpr int kLinVals = 28;
__global__ void __launch_bounds__(kGicpThreads) k_gicp_linearize(const float4 *src, int64_t ns, const float4 *tgt, const int *corr,
                                                                  const double *mahal, Pose12 T, int want_h, double *partial) {
  __shared__ double s_red[kGicpThreads / 32][kLinVals];
  const int64_t i = (int64_t)blockIdx.x * kGicpThreads + threadIdx.x;
  double v[kLinVals];
#pragma unroll
  for (int a = 0; a < kLinVals; ++a) v[a] = 0.0;
  const int j = i < ns ? corr[i] : -1;
  if (j >= 0) {
    const float4 p = src[i], q = tgt[j];
    double a[3], e[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) a[r] = T.R[r * 3] * (double)p.x + T.R[r * 3 + 1] * (double)p.y + T.R[r * 3 + 2] * (double)p.z + T.t[r];
    e[0] = (double)q.x - a[0]; e[1] = (double)q.y - a[1]; e[2] = (double)q.z - a[2];
    const double *M = mahal + 9 * i;
    double Me[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) Me[r] = M[r * 3] * e[0] + M[r * 3 + 1] * e[1] + M[r * 3 + 2] * e[2];
    v[0] = e[0] * Me[0] + e[1] * Me[1] + e[2] * Me[2];
    if (want_h) {
      // J = [ skew(T a) | -I ]
      const double J[18] = {0, -a[2], a[1], -1, 0, 0, a[2], 0, -a[0], 0, -1, 0, -a[1], a[0], 0, 0, 0, -1};
      double MJ[18];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) MJ[r * 6 + c] = M[r * 3] * J[c] + M[r * 3 + 1] * J[6 + c] + M[r * 3 + 2] * J[12 + c];
      int w = 1;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) v[w++] = J[r] * MJ[c] + J[6 + r] * MJ[6 + c] + J[12 + r] * MJ[12 + c];
#pragma unroll
      for (int r = 0; r < 6; ++r) v[22 + r] = J[r] * Me[0] + J[6 + r] * Me[1] + J[12 + r] * Me[2];
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < kLinVals; ++a) {
    double x = v[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[wid][a] = x;
  }
  __syncthreads();
  if (threadIdx.x < kLinVals) {
    double x = 0.0;
    for (int w = 0; w < kGicpThreads / 32; ++w) x += s_red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * kLinVals + threadIdx.x] = x;
  }
}
__global__ void k_gicp_reduce(const double *partial, int nblocks, double *out) {
  const int a = threadIdx.x;
  if (a >= kLinVals) return;
  double x = 0.0;
  for (int b = 0; b < nblocks; ++b) x += partial[(size_t)b * kLinVals + a];
  out[a] = x;
}
__global__ void k_gicp_sum_f(const float *v, int64_t n, double *out) {  // one block, fixed order
  __shared__ double s[kGicpThreads];
  double x = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += kGicpThreads) x += (double)v[i];
  s[threadIdx.x] = x;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < kGicpThreads; ++i) t += s[i]; *out = t; }
}

// ---- host side: 6x6 LDL^T, se3 exponential, LM loop ----------------------------------------------------
static void solve6(const double *Hin, const double *rhs, double *d) {
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

static Pose12 pose_identity() { Pose12 T{}; T.R[0] = T.R[4] = T.R[8] = 1.0; return T; }
static Pose12 pose_mul(const Pose12 &A, const Pose12 &B) {
  Pose12 C{};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) C.R[i * 3 + j] += A.R[i * 3 + k] * B.R[k * 3 + j];
    C.t[i] = A.t[i];
    for (int k = 0; k < 3; ++k) C.t[i] += A.R[i * 3 + k] * B.t[k];
  }
  return C;
}
// se3_exp (R/include/fast_gicp/so3/so3.hpp:59-104): quaternion exponential of the rotation part, V * translation
static Pose12 se3_exp(const double a[6]) {
  const double w[3] = {a[0], a[1], a[2]};
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double q = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * q;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * q;
  } else {
    const double th = std::sqrt(theta_sq), hh = 0.5 * th;
    imag = std::sin(hh) / th;
    real = std::cos(hh);
  }
  const double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  Pose12 T{};
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
static bool is_converged(const Pose12 &d, double rot_eps, double trans_eps) {
  double m = 0;
  for (int i = 0; i < 9; ++i) m = std::max(m, std::fabs(d.R[i] - ((i % 4 == 0) ? 1.0 : 0.0)) / rot_eps);
  for (int i = 0; i < 3; ++i) m = std::max(m, std::fabs(d.t[i]) / trans_eps);
  return m < 1;
}

struct GicpPool {
  DevBuf<float4> src, tgt;
  DevBuf<float> in_s, in_t, sqd, M12;
  DevBuf<double> cov_s, cov_t, mahal, partial, out;
  DevBuf<int> corr;
  const float *tgt_key = nullptr; int64_t tgt_n = 0; int tgt_k = 0;  // target covariances are reused while the target stays
  ~GicpPool() {
    src.release(); tgt.release(); in_s.release(); in_t.release(); sqd.release(); M12.release();
    cov_s.release(); cov_t.release(); mahal.release(); partial.release(); out.release(); corr.release();
  }
};
static void gicp_pool_free(void *p) { delete static_cast<GicpPool *>(p); }

}  // namespace sgtd

using namespace sgtd;

static bool gicp_is_device_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

extern "C" int sgtd_gicp_params_default(sgtd_gicp_params *p) {
  if (!p) return SGTD_E_INVALID;
  p->num_neighbors = 20; p->max_iterations = 10;  // R/config/SG_localization.yaml:24-25
  p->rotation_epsilon = 2e-3; p->transformation_epsilon = 5e-4;  // lsq_registration_impl.hpp:11-12
  p->best_fitness = 15.0;  // SG_data/best_fitness, R/config/SG_localization.yaml:15
  return SGTD_OK;
}

extern "C" int sgtd_gicp_align(sgtd_handle *h, const float *source_xyz, int64_t n_source, const float *target_xyz,
                               int64_t n_target, const double *init12, const sgtd_gicp_params *prm, double *final12,
                               double *fitness, int32_t *iterations, int32_t *converged) {
  if (!h || !source_xyz || !target_xyz || !prm || !final12 || !fitness) return sgtd_fail(h, SGTD_E_INVALID, "null argument", __FILE__, __LINE__);
  const int k = prm->num_neighbors;
  if (k < 1 || k > kGicpMaxK || n_source < k || n_target < k || prm->max_iterations < 0)
    return sgtd_fail(h, SGTD_E_INVALID, "gicp: num_neighbors must be in [1,64] and both clouds need at least that many points", __FILE__, __LINE__);
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  cudaStream_t st = h->stream;
  if (!h->gicp_pool) { h->gicp_pool = new GicpPool(); h->gicp_pool_free = gicp_pool_free; }
  GicpPool &P = *static_cast<GicpPool *>(h->gicp_pool);
  int rc = SGTD_OK;
#define G_CUDA(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rc = sgtd_fail(h, SGTD_E_CUDA, #expr, __FILE__, __LINE__, _e); goto done; } } while (0)
  {
    const int64_t ns = n_source, nt = n_target;
    const unsigned bs = (unsigned)((ns + kGicpThreads - 1) / kGicpThreads), bt = (unsigned)((nt + kGicpThreads - 1) / kGicpThreads);
    G_CUDA(P.src.reserve((size_t)ns, st, false)); G_CUDA(P.tgt.reserve((size_t)nt, st, false));
    G_CUDA(P.cov_s.reserve((size_t)ns * 9, st, false)); G_CUDA(P.cov_t.reserve((size_t)nt * 9, st, false));
    G_CUDA(P.mahal.reserve((size_t)ns * 9, st, false)); G_CUDA(P.corr.reserve((size_t)ns, st, false));
    G_CUDA(P.sqd.reserve((size_t)ns, st, false)); G_CUDA(P.partial.reserve((size_t)bs * kLinVals, st, false));
    G_CUDA(P.out.reserve(kLinVals + 1, st, false)); G_CUDA(P.M12.reserve(12, st, false));
    const float *d_s = source_xyz, *d_t = target_xyz;
    if (!gicp_is_device_ptr(source_xyz)) {
      G_CUDA(P.in_s.reserve((size_t)ns * 3, st, false));
      G_CUDA(cudaMemcpyAsync(P.in_s.p, source_xyz, (size_t)ns * 12, cudaMemcpyHostToDevice, st));
      d_s = P.in_s.p;
    }
    // the node transforms the source with the candidate's loop transform first (:692-696), guess = identity
    float M[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    if (init12) for (int i = 0; i < 12; ++i) M[i] = (float)init12[i];
    G_CUDA(cudaMemcpyAsync(P.M12.p, M, sizeof(M), cudaMemcpyHostToDevice, st));
    k_gicp_transform<<<bs, kGicpThreads, 0, st>>>(d_s, ns, P.M12.p, P.src.p);
    k_gicp_knn_cov<<<bs, kGicpThreads, 0, st>>>(P.src.p, ns, k, P.cov_s.p);
    h->launches += 2;
    // the target (map keyframe cloud) and its covariances are kept while the same buffer is passed again
    // (reg.clearTarget() in the reference recomputes them for every candidate)
    if (!(P.tgt_key == target_xyz && P.tgt_n == nt && P.tgt_k == k && prm->reuse_target)) {
      if (!gicp_is_device_ptr(target_xyz)) {
        G_CUDA(P.in_t.reserve((size_t)nt * 3, st, false));
        G_CUDA(cudaMemcpyAsync(P.in_t.p, target_xyz, (size_t)nt * 12, cudaMemcpyHostToDevice, st));
        d_t = P.in_t.p;
      }
      k_gicp_pack<<<bt, kGicpThreads, 0, st>>>(d_t, nt, P.tgt.p);
      k_gicp_knn_cov<<<bt, kGicpThreads, 0, st>>>(P.tgt.p, nt, k, P.cov_t.p);
      h->launches += 2;
      P.tgt_key = target_xyz; P.tgt_n = nt; P.tgt_k = k;
    }
    G_CUDA(cudaGetLastError());
    // LsqRegistration::computeTransformation with Levenberg-Marquardt (lsq_registration_impl.hpp:53-81,123-166)
    Pose12 x0 = pose_identity();
    double lambda = -1.0;
    bool conv = false;
    int it = 0;
    double vals[kLinVals];
    auto linearize = [&](const Pose12 &T, bool update, bool want_h) -> cudaError_t {
      if (update) {
        k_gicp_correspond<<<bs, kGicpThreads, 0, st>>>(P.src.p, ns, P.tgt.p, nt, T, P.cov_s.p, P.cov_t.p, P.corr.p, P.mahal.p, nullptr);
        h->launches++;
      }
      k_gicp_linearize<<<bs, kGicpThreads, 0, st>>>(P.src.p, ns, P.tgt.p, P.corr.p, P.mahal.p, T, want_h ? 1 : 0, P.partial.p);
      k_gicp_reduce<<<1, 32, 0, st>>>(P.partial.p, (int)bs, P.out.p);
      h->launches += 2;
      cudaError_t e = cudaMemcpyAsync(vals, P.out.p, sizeof(vals), cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return e;
      return cudaStreamSynchronize(st);
    };
    for (int i = 0; i < prm->max_iterations && !conv; ++i) {
      it = i;
      G_CUDA(linearize(x0, true, true));
      const double y0 = vals[0];
      double H[36], b[6];
      int w = 1;
      for (int r = 0; r < 6; ++r) for (int c = r; c < 6; ++c) { H[r * 6 + c] = vals[w]; H[c * 6 + r] = vals[w]; ++w; }
      for (int r = 0; r < 6; ++r) b[r] = vals[22 + r];
      if (lambda < 0.0) { double m = 0; for (int d = 0; d < 6; ++d) m = std::max(m, std::fabs(H[d * 6 + d])); lambda = 1e-9 * m; }
      double nu = 2.0;
      bool ok = false;
      Pose12 delta = pose_identity();
      for (int j = 0; j < 10; ++j) {
        double Hl[36], nb[6], d[6];
        std::memcpy(Hl, H, sizeof(Hl));
        for (int q = 0; q < 6; ++q) { Hl[q * 6 + q] += lambda; nb[q] = -b[q]; }
        solve6(Hl, nb, d);
        delta = se3_exp(d);
        const Pose12 xi = pose_mul(delta, x0);
        G_CUDA(linearize(xi, false, false));
        const double yi = vals[0];
        double den = 0;
        for (int q = 0; q < 6; ++q) den += d[q] * (lambda * d[q] - b[q]);
        const double rho = (y0 - yi) / den;
        if (rho < 0) {
          if (is_converged(delta, prm->rotation_epsilon, prm->transformation_epsilon)) { ok = true; break; }
          lambda = nu * lambda;
          nu = 2 * nu;
          continue;
        }
        x0 = xi;
        lambda = lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
        ok = true;
        break;
      }
      if (!ok) break;  // "lm not converged!!"
      conv = is_converged(delta, prm->rotation_epsilon, prm->transformation_epsilon);
    }
    // final_transformation_ = x0.cast<float>().matrix(); fitness = mean squared NN distance of the aligned source
    Pose12 xf{};
    for (int r = 0; r < 9; ++r) xf.R[r] = (double)(float)x0.R[r];
    for (int r = 0; r < 3; ++r) xf.t[r] = (double)(float)x0.t[r];
    k_gicp_correspond<<<bs, kGicpThreads, 0, st>>>(P.src.p, ns, P.tgt.p, nt, xf, nullptr, nullptr, nullptr, nullptr, P.sqd.p);
    k_gicp_sum_f<<<1, kGicpThreads, 0, st>>>(P.sqd.p, ns, P.out.p + kLinVals);
    h->launches += 2;
    double sum = 0.0;
    G_CUDA(cudaMemcpyAsync(&sum, P.out.p + kLinVals, 8, cudaMemcpyDeviceToHost, st));
    G_CUDA(cudaStreamSynchronize(st));
    *fitness = sum / (double)ns;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) final12[r * 4 + c] = xf.R[r * 3 + c]; final12[r * 4 + 3] = xf.t[r]; }
    if (iterations) *iterations = it;
    if (converged) *converged = conv ? 1 : 0;
  }
done:
#undef G_CUDA
  if (prev >= 0 && prev != h->device) cudaSetDevice(prev);
  return rc;
}

// The node's multi-candidate refinement (R/src/semantic_graph_localization.cpp:603,651-721): the candidates
// are visited in order of match_fitness descending (`order`, e.g. from sgtd_recall_rank); the first one whose
// GICP fitness drops below best_fitness ends the search, otherwise the lowest fitness below 100 wins.
extern "C" int sgtd_gicp_refine_candidates(sgtd_handle *h, const float *source_xyz, int64_t n_source,
                                           const float *const *targets_xyz, const int64_t *n_targets,
                                           const sgtd_candidate *cands, const int32_t *order, int32_t ncand,
                                           const sgtd_gicp_params *prm, int32_t *chosen, double *transformation12,
                                           double *fitness_out, int32_t *n_aligned) {
  if (!h || !source_xyz || !targets_xyz || !n_targets || !cands || !prm || !chosen || !transformation12 || !fitness_out)
    return sgtd_fail(h, SGTD_E_INVALID, "null argument", __FILE__, __LINE__);
  double bitness = 100.0;  // :650
  *chosen = -1;
  *fitness_out = bitness;
  const double I12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  std::memcpy(transformation12, I12, sizeof(I12));
  int aligned = 0;
  for (int32_t i = 0; i < ncand; ++i) {
    const int32_t c = order ? order[i] : i;
    const sgtd_candidate &cd = cands[c];
    if (cd.frame < 0 || !targets_xyz[c]) continue;
    double init12[12], fin[12], fit = 0.0;
    for (int r = 0; r < 3; ++r) { for (int k2 = 0; k2 < 3; ++k2) init12[r * 4 + k2] = cd.R[r * 3 + k2]; init12[r * 4 + 3] = cd.t[r]; }
    const int rc = sgtd_gicp_align(h, source_xyz, n_source, targets_xyz[c], n_targets[c], init12, prm, fin, &fit, nullptr, nullptr);
    if (rc != SGTD_OK) return rc;
    ++aligned;
    if (fit < bitness) { bitness = fit; *chosen = c; *fitness_out = fit; std::memcpy(transformation12, fin, sizeof(fin)); }
    if (fit < prm->best_fitness) { *chosen = c; *fitness_out = fit; std::memcpy(transformation12, fin, sizeof(fin)); break; }
  }
  if (n_aligned) *n_aligned = aligned;
  return SGTD_OK;
}
