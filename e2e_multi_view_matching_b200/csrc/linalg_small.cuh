// Register/shared-memory resident fp64 linear algebra for the pose kernels: cyclic Jacobi
// eigen-solvers (3x3, 4x4 per thread; 9x9 warp-cooperative), 3x3 inverse, 6x6 / NxN Cholesky.
// These replace the reference's torch.svd / torch.lu calls (cuSOLVER/MAGMA) on kilobyte-sized
// problems (estimate_relative_pose.py:72,76; kornia triangulate_points).
#pragma once
#include <cuda_runtime.h>

// Per-thread cyclic Jacobi for a symmetric NxN matrix held in registers.
// On exit a[i][i] are the eigenvalues and column j of v the eigenvector of a[j][j].
template <int N, int SWEEPS>
__device__ __forceinline__ void jacobi_eig_reg(double (&a)[N][N], double (&v)[N][N]) {
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < SWEEPS; ++sweep) {
#pragma unroll
    for (int p = 0; p < N - 1; ++p) {
#pragma unroll
      for (int q = p + 1; q < N; ++q) {
        const double apq = a[p][q];
        if (fabs(apq) > 1e-300) {
          const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = rsqrt(t * t + 1.0);
          const double s = t * c;
          a[p][p] -= t * apq;
          a[q][q] += t * apq;
          a[p][q] = 0.0;
          a[q][p] = 0.0;
#pragma unroll
          for (int k = 0; k < N; ++k) {
            if (k != p && k != q) {
              const double akp = a[k][p], akq = a[k][q];
              a[k][p] = c * akp - s * akq;
              a[p][k] = a[k][p];
              a[k][q] = s * akp + c * akq;
              a[q][k] = a[k][q];
            }
            const double vkp = v[k][p], vkq = v[k][q];
            v[k][p] = c * vkp - s * vkq;
            v[k][q] = s * vkp + c * vkq;
          }
        }
      }
    }
  }
}

// index of the smallest diagonal entry
template <int N>
__device__ __forceinline__ int argmin_diag(const double (&a)[N][N]) {
  int m = 0;
#pragma unroll
  for (int i = 1; i < N; ++i)
    if (a[i][i] < a[m][m]) m = i;
  return m;
}

// Warp-cooperative cyclic Jacobi on a symmetric 9x9 in shared memory (A, V: 81 doubles each).
// Lane k < 9 owns index k of the row/column updates.  All 32 lanes must call.
__device__ __forceinline__ void jacobi_eig9_warp(double* A, double* V, int lane) {
  constexpr int N = 9;
  for (int e = lane; e < N * N; e += 32) V[e] = (e / N == e % N) ? 1.0 : 0.0;
  __syncwarp();
  for (int sweep = 0; sweep < 14; ++sweep) {
    double off = 0.0;
    for (int e = lane; e < N * N; e += 32)
      if (e / N != e % N) off += A[e] * A[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) off += __shfl_xor_sync(0xffffffffu, off, o);
    double diag = 0.0;
    if (lane < N) diag = A[lane * N + lane] * A[lane * N + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) diag += __shfl_xor_sync(0xffffffffu, diag, o);
    if (off <= 1e-32 * diag) break;
    for (int p = 0; p < N - 1; ++p) {
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        const double app = A[p * N + p], aqq = A[q * N + q];
        __syncwarp();
        if (fabs(apq) > 1e-300) {
          const double theta = (aqq - app) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = rsqrt(t * t + 1.0);
          const double s = t * c;
          if (lane < N) {
            const int k = lane;
            // column update (A <- A J) and eigenvectors
            const double akp = A[k * N + p], akq = A[k * N + q];
            A[k * N + p] = c * akp - s * akq;
            A[k * N + q] = s * akp + c * akq;
            const double vkp = V[k * N + p], vkq = V[k * N + q];
            V[k * N + p] = c * vkp - s * vkq;
            V[k * N + q] = s * vkp + c * vkq;
          }
          __syncwarp();
          if (lane < N) {
            const int k = lane;
            // row update (A <- J^T A)
            const double apk = A[p * N + k], aqk = A[q * N + k];
            A[p * N + k] = c * apk - s * aqk;
            A[q * N + k] = s * apk + c * aqk;
          }
          __syncwarp();
          if (lane == 0) { A[p * N + q] = 0.0; A[q * N + p] = 0.0; }
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ bool inv3_sym(const double m[6], double inv[6]) {
  // m = [a00,a01,a02,a11,a12,a22]
  const double c00 = m[3] * m[5] - m[4] * m[4];
  const double c01 = m[2] * m[4] - m[1] * m[5];
  const double c02 = m[1] * m[4] - m[2] * m[3];
  const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  if (!(fabs(det) > 0.0)) return false;
  const double id = 1.0 / det;
  inv[0] = c00 * id; inv[1] = c01 * id; inv[2] = c02 * id;
  inv[3] = (m[0] * m[5] - m[2] * m[2]) * id;
  inv[4] = (m[1] * m[2] - m[0] * m[4]) * id;
  inv[5] = (m[0] * m[3] - m[1] * m[1]) * id;
  return true;
}

// Dense solve A x = b for small N (row-major A, overwritten) with partial pivoting
// (the reference uses torch.lu/lu_solve, bundle_adjust_gauss_newton_2_view.py:184-190).
template <int N>
__device__ inline bool lu_solve_small(double* A, double* b) {
  for (int k = 0; k < N; ++k) {
    int piv = k;
    double best = fabs(A[k * N + k]);
    for (int i = k + 1; i < N; ++i)
      if (fabs(A[i * N + k]) > best) { best = fabs(A[i * N + k]); piv = i; }
    if (!(best > 0.0)) return false;
    if (piv != k) {
      for (int j = 0; j < N; ++j) { const double t = A[k * N + j]; A[k * N + j] = A[piv * N + j]; A[piv * N + j] = t; }
      const double t = b[k]; b[k] = b[piv]; b[piv] = t;
    }
    const double inv = 1.0 / A[k * N + k];
    for (int i = k + 1; i < N; ++i) {
      const double f = A[i * N + k] * inv;
      for (int j = k + 1; j < N; ++j) A[i * N + j] -= f * A[k * N + j];
      b[i] -= f * b[k];
    }
  }
  for (int i = N - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < N; ++j) s -= A[i * N + j] * b[j];
    b[i] = s / A[i * N + i];
  }
  return true;
}

// DLT triangulation (kornia triangulate_points): smallest right-singular vector of the 4x4
// system built from P1 = [I|0] and P2 = [R|t], de-homogenised with kornia's eps rule.
__device__ __forceinline__ void triangulate_dlt(const double R[9], const double t[3], double x1,
                                                double y1, double x2, double y2, double X[3]) {
  double A[4][4];
  // rows: x1*P1[2]-P1[0], y1*P1[2]-P1[1], x2*P2[2]-P2[0], y2*P2[2]-P2[1]
  A[0][0] = -1.0; A[0][1] = 0.0;  A[0][2] = x1; A[0][3] = 0.0;
  A[1][0] = 0.0;  A[1][1] = -1.0; A[1][2] = y1; A[1][3] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    A[2][i] = x2 * R[6 + i] - R[i];
    A[3][i] = y2 * R[6 + i] - R[3 + i];
  }
  A[2][3] = x2 * t[2] - t[0];
  A[3][3] = y2 * t[2] - t[1];
  double M[4][4], V[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += A[k][i] * A[k][j];
      M[i][j] = s;
    }
  jacobi_eig_reg<4, 8>(M, V);
  const int m = argmin_diag<4>(M);
  double h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = V[i][0];
    if (m == 1) h[i] = V[i][1];
    if (m == 2) h[i] = V[i][2];
    if (m == 3) h[i] = V[i][3];
  }
  const double scale = fabs(h[3]) > 1e-8 ? 1.0 / (h[3] + 1e-8) : 1.0;
  X[0] = h[0] * scale; X[1] = h[1] * scale; X[2] = h[2] * scale;
}

// ---- rotations (ceres/rotation.h semantics) ----------------------------------------------------
__device__ inline void aa_to_R(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2 > 2.220446049250313e-16) {
    const double th = sqrt(th2), k0 = w[0] / th, k1 = w[1] / th, k2 = w[2] / th;
    const double c = cos(th), s = sin(th), v = 1.0 - c;
    // I + s K + (1-c) K^2, K = hat(k)
    R[0] = 1 + v * (-(k1 * k1 + k2 * k2)); R[1] = -s * k2 + v * k0 * k1; R[2] = s * k1 + v * k0 * k2;
    R[3] = s * k2 + v * k0 * k1; R[4] = 1 + v * (-(k0 * k0 + k2 * k2)); R[5] = -s * k0 + v * k1 * k2;
    R[6] = -s * k1 + v * k0 * k2; R[7] = s * k0 + v * k1 * k2; R[8] = 1 + v * (-(k0 * k0 + k1 * k1));
  } else {
    R[0] = 1; R[1] = -w[2]; R[2] = w[1]; R[3] = w[2]; R[4] = 1; R[5] = -w[0]; R[6] = -w[1]; R[7] = w[0]; R[8] = 1;
  }
}

__device__ inline void R_to_aa(const double* R, double* w) {
  // ceres::RotationMatrixToAngleAxis via the quaternion
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr >= 0.0) {
    double t = sqrt(tr + 1.0);
    q[0] = 0.5 * t; t = 0.5 / t;
    q[1] = (R[7] - R[5]) * t; q[2] = (R[2] - R[6]) * t; q[3] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    q[i + 1] = 0.5 * t; t = 0.5 / t;
    q[0] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j + 1] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k + 1] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const double s = sqrt(s2);
    const double th = q[0] < 0.0 ? 2.0 * atan2(-s, -q[0]) : 2.0 * atan2(s, q[0]);
    const double k = th / s;
    w[0] = q[1] * k; w[1] = q[2] * k; w[2] = q[3] * k;
  } else {
    w[0] = q[1] * 2.0; w[1] = q[2] * 2.0; w[2] = q[3] * 2.0;
  }
}


constexpr int MAXU = 43;   // leading dimension of the small dense systems (6 x 7 free cameras = 42 unknowns, +1: odd
                           // in 8-byte banks, and room for the augmented right-hand-side row of chol_solve_warp_rows)

// Warp Cholesky solve with one lane per matrix row (n + 1 <= 32): left-looking factorisation, lane i forms
// L[i][j] = (A[i][j] - sum_{k<j} L[i][k] L[j][k]) / L[j][j] from its OWN row (conflict-free: the leading
// dimension MAXU is odd in 8-byte banks) and row j (a broadcast read).  The right-hand side rides along as
// row n of the matrix, so the forward substitution is part of the factorisation; the backward substitution is
// column oriented.  Short loops on purpose: the fully unrolled register version of this (7 k straight-line
// instructions, executed once per LM iteration) was slower than the shared-memory loops -- instruction fetch.
// Measured (tools/mvba_timing.py, 24 x 24): 39 us for chol_solve_warp below, ~5 us here.
// A: symmetric, row-major, ld = MAXU, rows 0..n used (row n = scratch), overwritten.  x: rhs in, solution out.
__device__ inline bool chol_solve_warp_rows(double* A, double* x, int n, int lane) {
  if (lane < n) A[n * MAXU + lane] = x[lane];
  __syncwarp();
  double* row = A + (lane <= n ? lane : 0) * MAXU;
  bool ok = true;
  for (int j = 0; j < n; ++j) {
    const double* rj = A + j * MAXU;
    double s = 0.0;
    if (lane >= j && lane <= n) {
      s = row[j];
#pragma unroll 4
      for (int k = 0; k < j; ++k) s -= row[k] * rj[k];
    }
    double d = __shfl_sync(0xffffffffu, s, j);
    if (!(d > 0.0)) { ok = false; break; }     // uniform
    d = sqrt(d);
    __syncwarp();
    if (lane == j) row[j] = d;
    else if (lane > j && lane <= n) row[j] = s / d;
    __syncwarp();
  }
  if (!ok) return false;
  double y = lane < n ? A[n * MAXU + lane] : 0.0;      // L y = b solved above: y sits in row n
  for (int k = n - 1; k >= 0; --k) {
    const double xk = __shfl_sync(0xffffffffu, y, k) / A[k * MAXU + k];
    if (lane == k) y = xk;
    else if (lane < k) y -= A[k * MAXU + lane] * xk;
  }
  if (lane < n) x[lane] = y;
  __syncwarp();
  return true;
}

// warp-cooperative Cholesky solve of the n x n SPD system in shared memory (n <= 42).
// A (row-major, ld = MAXU) is overwritten, x holds rhs on entry / solution on exit.
__device__ inline bool chol_solve_warp(double* A, double* x, int n, int lane) {
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    __syncwarp();
    double d = A[k * MAXU + k];
    if (!(d > 0.0)) { ok = false; break; }
    d = sqrt(d);
    __syncwarp();
    if (lane == 0) A[k * MAXU + k] = d;
    for (int i = k + 1 + lane; i < n; i += 32) A[i * MAXU + k] /= d;
    __syncwarp();
    for (int j = k + 1; j < n; ++j) {
      const double ljk = A[j * MAXU + k];
      for (int i = j + lane; i < n; i += 32) A[i * MAXU + j] -= A[i * MAXU + k] * ljk;
    }
  }
  __syncwarp();
  if (!ok) return false;
  if (lane == 0) {
    for (int i = 0; i < n; ++i) {
      double s = x[i];
      for (int j = 0; j < i; ++j) s -= A[i * MAXU + j] * x[j];
      x[i] = s / A[i * MAXU + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double s = x[i];
      for (int j = i + 1; j < n; ++j) s -= A[j * MAXU + i] * x[j];
      x[i] = s / A[i * MAXU + i];
    }
  }
  __syncwarp();
  return true;
}

