// svd3.cuh -- two-sided Jacobi SVD of a 3x3 (the algorithm of Eigen::JacobiSVD for square real input),
// shared by the verification kernels (triangle_solver, R/src/STDesc.cpp:560) and the GICP covariance
// regularisation (R/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp:277).
#pragma once
#include <cuda_runtime.h>

namespace sgtd {

struct Rot2 { double c, s; };
__device__ __forceinline__ Rot2 rot_t(Rot2 j) { Rot2 r; r.c = j.c; r.s = -j.s; return r; }

// rows p,q of a row-major 3x3: x' = c x + s y ; y' = -s x + c y
__device__ __forceinline__ void rot_rows(double *M, int p, int q, Rot2 j) {
  if (j.c == 1.0 && j.s == 0.0) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double xi = M[p * 3 + i], yi = M[q * 3 + i];
    M[p * 3 + i] = __dadd_rn(__dmul_rn(j.c, xi), __dmul_rn(j.s, yi));
    M[q * 3 + i] = __dadd_rn(__dmul_rn(-j.s, xi), __dmul_rn(j.c, yi));
  }
}
__device__ __forceinline__ void rot_cols(double *M, int p, int q, Rot2 jr) {
  const Rot2 j = rot_t(jr);
  if (j.c == 1.0 && j.s == 0.0) return;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double xi = M[i * 3 + p], yi = M[i * 3 + q];
    M[i * 3 + p] = __dadd_rn(__dmul_rn(j.c, xi), __dmul_rn(j.s, yi));
    M[i * 3 + q] = __dadd_rn(__dmul_rn(-j.s, xi), __dmul_rn(j.c, yi));
  }
}

// Two-sided Jacobi SVD of a 3x3 (the algorithm of Eigen::JacobiSVD for square
// real input, which triangle_solver calls at STDesc.cpp:560): W = U S V^T.
__device__ inline void svd3(const double *A, double *U, double *V) {
  const double kMin = 2.2250738585072014e-308, kPrec = 2.0 * 2.220446049250313e-16;
  double scale = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) scale = fmax(scale, fabs(A[i]));
  if (scale == 0.0) scale = 1.0;
  double W[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) { W[i] = A[i] / scale; U[i] = V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  double maxDiag = fmax(fabs(W[0]), fmax(fabs(W[4]), fabs(W[8])));
  bool finished = false;
  for (int sweep = 0; sweep < 64 && !finished; ++sweep) {
    finished = true;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 0) ? 1 : 2, q = (pq == 2) ? 1 : 0;  // (1,0) (2,0) (2,1)
      const double threshold = fmax(kMin, __dmul_rn(kPrec, maxDiag));
      if (fabs(W[p * 3 + q]) > threshold || fabs(W[q * 3 + p]) > threshold) {
        finished = false;
        // real_2x2_jacobi_svd
        double m00 = W[p * 3 + p], m01 = W[p * 3 + q], m10 = W[q * 3 + p], m11 = W[q * 3 + q];
        Rot2 rot1;
        const double t = __dadd_rn(m00, m11), d = __dsub_rn(m10, m01);
        if (fabs(d) < kMin) { rot1.s = 0.0; rot1.c = 1.0; }
        else {
          const double u = t / d;
          const double tmp = __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(u, u)));
          rot1.s = 1.0 / tmp; rot1.c = u / tmp;
        }
        if (!(rot1.c == 1.0 && rot1.s == 0.0)) {
          const double a0 = __dadd_rn(__dmul_rn(rot1.c, m00), __dmul_rn(rot1.s, m10));
          const double a1 = __dadd_rn(__dmul_rn(rot1.c, m01), __dmul_rn(rot1.s, m11));
          const double b0 = __dadd_rn(__dmul_rn(-rot1.s, m00), __dmul_rn(rot1.c, m10));
          const double b1 = __dadd_rn(__dmul_rn(-rot1.s, m01), __dmul_rn(rot1.c, m11));
          m00 = a0; m01 = a1; m10 = b0; m11 = b1;
        }
        (void)m10;
        // makeJacobi(m00, m01, m11)
        Rot2 jr;
        const double deno = __dmul_rn(2.0, fabs(m01));
        if (deno < kMin) { jr.c = 1.0; jr.s = 0.0; }
        else {
          const double tau = __dsub_rn(m00, m11) / deno;
          const double w = __dsqrt_rn(__dadd_rn(__dmul_rn(tau, tau), 1.0));
          const double tt = (tau > 0.0) ? 1.0 / __dadd_rn(tau, w) : 1.0 / __dsub_rn(tau, w);
          const double sign_t = tt > 0.0 ? 1.0 : -1.0;
          const double n = 1.0 / __dsqrt_rn(__dadd_rn(__dmul_rn(tt, tt), 1.0));
          jr.s = __dmul_rn(__dmul_rn(__dmul_rn(-sign_t, m01 / fabs(m01)), fabs(tt)), n);
          jr.c = n;
        }
        // j_left = rot1 * j_right^T
        const Rot2 jrt = rot_t(jr);
        Rot2 jl;
        jl.c = __dsub_rn(__dmul_rn(rot1.c, jrt.c), __dmul_rn(rot1.s, jrt.s));
        jl.s = __dadd_rn(__dmul_rn(rot1.c, jrt.s), __dmul_rn(rot1.s, jrt.c));
        rot_rows(W, p, q, jl);
        rot_cols(U, p, q, rot_t(jl));
        rot_cols(W, p, q, jr);
        rot_cols(V, p, q, jr);
        maxDiag = fmax(maxDiag, fmax(fabs(W[p * 3 + p]), fabs(W[q * 3 + q])));
      }
    }
  }
  double sv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double a = fabs(W[i * 3 + i]);
    sv[i] = a;
    if (a != 0.0) { const double f = W[i * 3 + i] / a; U[0 * 3 + i] *= f; U[1 * 3 + i] *= f; U[2 * 3 + i] *= f; }
  }
  // sort singular values descending (selection, swapping columns of U and V)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int pos = i; double mx = sv[i];
#pragma unroll
    for (int kk = i + 1; kk < 3; ++kk) if (sv[kk] > mx) { mx = sv[kk]; pos = kk; }
    if (mx == 0.0) break;
    if (pos != i) {
      double t = sv[i]; sv[i] = sv[pos]; sv[pos] = t;
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        t = U[rr * 3 + i]; U[rr * 3 + i] = U[rr * 3 + pos]; U[rr * 3 + pos] = t;
        t = V[rr * 3 + i]; V[rr * 3 + i] = V[rr * 3 + pos]; V[rr * 3 + pos] = t;
      }
    }
  }
}

}  // namespace sgtd
