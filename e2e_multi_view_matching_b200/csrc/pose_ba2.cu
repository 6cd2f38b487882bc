// Two-view Gauss-Newton / Levenberg-Marquardt bundle adjustment, one CTA per pair problem.
// Reference: BundleAdjustGaussNewton2View.run and compute_A_b / fill_J,
// pose_optimization/two_view/bundle_adjust_gauss_newton_2_view.py:50-201.
//
// Same iteration as the reference -- 11 evaluations, 10 steps, lambda0 = 0.1, /3.5 on a new best
// residual, x1.5 otherwise, Jacobi scaling (A + lambda diag A) delta = b, step ALWAYS applied,
// best iterate returned -- but the (6+3n)^2 dense J^T J + LU of the reference (38 MB and ~78
// GFLOP per iteration on structural zeros at n = 1024) is replaced by its Schur complement:
// every 3x3 point block is eliminated in registers and only a 6x6 camera system is solved.
// Algebraically the same step; fp64 throughout (the reference is fp32 with a dense LU).
#include "../../include/mvm_b200.h"
#include "common.cuh"
#include "linalg_small.cuh"

namespace {

constexpr int NT = 256;
constexpr int NRED = 28;   // 21 (6x6 upper) + 6 (rhs) + 1 (residual norm / flag)

struct Ba2Args {
  const float* k0n; const float* k1n;   // [B,N,2] normalised image coordinates
  const float* conf;                    // [B,N]; <= 0 marks an invalid match
  const float* T_init;                  // [B,16]
  const int* n_valid;                   // [B] effective keypoints per item or null
  const unsigned char* mask;            // [B,N] or null: conf is treated as 0 where mask == 0
  int N, n_iter;
  float* T_out;                         // [B,16]
  unsigned char* valid_batch;           // [B]
  double* pts;                          // [B,N,3] scratch (3-D points)
  float* trace;                         // [B, n_iter+1] residual norms or null
};

__device__ __forceinline__ void block_reduce(double* v, int n, double (*s_red)[NRED], double* s_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = 0; e < n; ++e) v[e] = warp_sum_d(v[e]);
  __syncthreads();
  if (lane == 0)
    for (int e = 0; e < n; ++e) s_red[warp][e] = v[e];
  __syncthreads();
  if (threadIdx.x < n) {
    double s = 0.0;
    for (int w = 0; w < NT / 32; ++w) s += s_red[w][threadIdx.x];
    s_out[threadIdx.x] = s;
  }
  __syncthreads();
}

// se3 exponential as used by the reference: pytorch3d se3_exp_map(.)^T with the eps = 1e-4
// clamp on |w|^2 (bundle_adjust_gauss_newton_2_view.py:194), then T1 <- Exp(delta) T1 (:195).
__device__ void apply_se3_update(const double d[6], double T[12]) {
  const double v[3] = {d[0], d[1], d[2]}, w[3] = {d[3], d[4], d[5]};
  const double nr = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = sqrt(fmax(nr, 1e-4));
  const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double K2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += K[i * 3 + k] * K[k * 3 + j];
      K2[i * 3 + j] = s;
    }
  const double f1 = sin(th) / th, f2 = (1.0 - cos(th)) / (th * th), f3 = (th - sin(th)) / (th * th * th);
  double R[9], V[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = f1 * K[i] + f2 * K2[i] + I;
    V[i] = I + f2 * K[i] + f3 * K2[i];
  }
  double tt[3];
  for (int i = 0; i < 3; ++i) tt[i] = V[i * 3] * v[0] + V[i * 3 + 1] * v[1] + V[i * 3 + 2] * v[2];
  double N[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * T[k * 4 + j];
      N[i * 4 + j] = s;
    }
    N[i * 4 + 3] += tt[i];
  }
  for (int i = 0; i < 12; ++i) T[i] = N[i];
}

// Per-point normal-equation blocks.  Unknown order per point: camera-1 [v | omega] (6), point (3).
struct PointBlocks {
  double App[6];     // 3x3 symmetric: 00,01,02,11,12,22
  double Acp[18];    // 6x3
  double bp[3];
  double Acc[21];    // 6x6 upper
  double bc[6];
  double rho;
};

__device__ __forceinline__ void point_blocks(const double T[12], const double p[3], double x0,
                                             double y0, double x1, double y1, double w,
                                             PointBlocks& o) {
  // camera 0 (fixed, identity): q = p
  const double iz0 = 1.0 / p[2];
  const double J0[2][3] = {{w * iz0, 0.0, -w * p[0] * iz0 * iz0}, {0.0, w * iz0, -w * p[1] * iz0 * iz0}};
  const double r0[2] = {w * (p[0] * iz0 - x0), w * (p[1] * iz0 - y0)};
  // camera 1: q = R p + t
  double q[3];
  for (int i = 0; i < 3; ++i) q[i] = T[i * 4] * p[0] + T[i * 4 + 1] * p[1] + T[i * 4 + 2] * p[2] + T[i * 4 + 3];
  const double iz1 = 1.0 / q[2];
  const double Jpi[2][3] = {{w * iz1, 0.0, -w * q[0] * iz1 * iz1}, {0.0, w * iz1, -w * q[1] * iz1 * iz1}};
  const double r1[2] = {w * (q[0] * iz1 - x1), w * (q[1] * iz1 - y1)};
  // point block of camera 1: Jpi R ; camera block: Jpi [I | -hat(q)]
  double J1[2][3], Jc[2][6];
  for (int r = 0; r < 2; ++r) {
    for (int c = 0; c < 3; ++c)
      J1[r][c] = Jpi[r][0] * T[c] + Jpi[r][1] * T[4 + c] + Jpi[r][2] * T[8 + c];
    Jc[r][0] = Jpi[r][0]; Jc[r][1] = Jpi[r][1]; Jc[r][2] = Jpi[r][2];
    // -hat(q) = [[0, q2, -q1], [-q2, 0, q0], [q1, -q0, 0]]
    Jc[r][3] = -Jpi[r][1] * q[2] + Jpi[r][2] * q[1];
    Jc[r][4] = Jpi[r][0] * q[2] - Jpi[r][2] * q[0];
    Jc[r][5] = -Jpi[r][0] * q[1] + Jpi[r][1] * q[0];
  }
  int e = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j)
      o.App[e++] = J0[0][i] * J0[0][j] + J0[1][i] * J0[1][j] + J1[0][i] * J1[0][j] + J1[1][i] * J1[1][j];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 3; ++j) o.Acp[i * 3 + j] = Jc[0][i] * J1[0][j] + Jc[1][i] * J1[1][j];
  for (int j = 0; j < 3; ++j)
    o.bp[j] = -(J0[0][j] * r0[0] + J0[1][j] * r0[1] + J1[0][j] * r1[0] + J1[1][j] * r1[1]);
  e = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) o.Acc[e++] = Jc[0][i] * Jc[0][j] + Jc[1][i] * Jc[1][j];
  for (int i = 0; i < 6; ++i) o.bc[i] = -(Jc[0][i] * r1[0] + Jc[1][i] * r1[1]);
  o.rho = r0[0] * r0[0] + r0[1] * r0[1] + r1[0] * r1[0] + r1[1] * r1[1];
}

__global__ void __launch_bounds__(NT) ba2_kernel(Ba2Args a) {
  __shared__ double s_red[NT / 32][NRED];
  __shared__ double s_sum[NRED];
  __shared__ double s_T[12], s_best[12], s_delta[6];
  __shared__ double s_lambda, s_best_r;
  __shared__ int s_ok, s_precond;

  const int b = blockIdx.x, NS = a.N, tid = threadIdx.x;
  const int N = a.n_valid ? min(max(a.n_valid[b], 0), NS) : NS;
  const float* k0 = a.k0n + (long long)b * NS * 2;
  const float* k1 = a.k1n + (long long)b * NS * 2;
  const float* cf_raw = a.conf + (long long)b * NS;
  const unsigned char* mk = a.mask ? a.mask + (long long)b * NS : nullptr;
  auto CF = [&](int i) -> float { return (mk && !mk[i]) ? 0.f : cf_raw[i]; };
  double* P = a.pts + (long long)b * NS * 3;
  float* Tout = a.T_out + b * 16;

  // valid = conf > 0 (:129-131); weights conf / (0.5 * sum over the 2n observations) (:44-48)
  double v2[NRED];
  v2[0] = 0.0; v2[1] = 0.0;
  for (int i = tid; i < N; i += NT)
    if (CF(i) > 0.f) { v2[0] += 1.0; v2[1] += (double)CF(i); }
  block_reduce(v2, 2, s_red, s_sum);
  const int n_matches = (int)(s_sum[0] + 0.5);
  const double sum_conf = fmax(2.0 * s_sum[1], 1e-6);
  const double w_scale = 1.0 / (0.5 * sum_conf);
  __syncthreads();
  if (n_matches <= 6) {            // excluded batch item (:134-138)
    if (tid < 16) Tout[tid] = a.T_init[b * 16 + tid];
    if (tid == 0) a.valid_batch[b] = 0;
    return;
  }
  if (tid == 0) a.valid_batch[b] = 1;
  if (tid < 12) { s_T[tid] = (double)a.T_init[b * 16 + tid]; s_best[tid] = s_T[tid]; }
  if (tid == 0) { s_lambda = 0.1; s_best_r = 0.0; }
  __syncthreads();

  // initial points: DLT triangulation with P0 = [I|0], P1 = T_init[:3] (:115-125)
  {
    double R[9], t[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) R[i * 3 + j] = s_T[i * 4 + j];
      t[i] = s_T[i * 4 + 3];
    }
    for (int i = tid; i < N; i += NT) {
      if (!(CF(i) > 0.f)) continue;
      double X[3];
      triangulate_dlt(R, t, k0[2 * i], k0[2 * i + 1], k1[2 * i], k1[2 * i + 1], X);
      P[3 * i] = X[0]; P[3 * i + 1] = X[1]; P[3 * i + 2] = X[2];
    }
  }
  __syncthreads();

  for (int it = 0; it <= a.n_iter; ++it) {
    double T[12];
    for (int i = 0; i < 12; ++i) T[i] = s_T[i];
    // ---- pass 1: camera block, rhs, residual norm, preconditioner validity ----
    double acc[NRED];
    for (int e = 0; e < NRED; ++e) acc[e] = 0.0;
    int diag_ok = 1;
    for (int i = tid; i < N; i += NT) {
      if (!(CF(i) > 0.f)) continue;
      PointBlocks pb;
      const double p[3] = {P[3 * i], P[3 * i + 1], P[3 * i + 2]};
      point_blocks(T, p, k0[2 * i], k0[2 * i + 1], k1[2 * i], k1[2 * i + 1], (double)CF(i) * w_scale, pb);
      for (int e = 0; e < 21; ++e) acc[e] += pb.Acc[e];
      for (int e = 0; e < 6; ++e) acc[21 + e] += pb.bc[e];
      acc[27] += pb.rho;
      if (!(pb.App[0] > 0.0 && pb.App[3] > 0.0 && pb.App[5] > 0.0)) diag_ok = 0;
    }
    diag_ok = __syncthreads_and(diag_ok);
    block_reduce(acc, NRED, s_red, s_sum);
    if (tid == 0) {
      const double rho = s_sum[27];
      if (a.trace) a.trace[b * (a.n_iter + 1) + it] = (float)rho;
      if (it == 0) {
        s_best_r = rho;
        for (int i = 0; i < 12; ++i) s_best[i] = s_T[i];
      } else if (rho < s_best_r) {          // :160-165
        s_best_r = rho;
        for (int i = 0; i < 12; ++i) s_best[i] = s_T[i];
        s_lambda = s_lambda / 3.5;
      } else {
        s_lambda = s_lambda * 1.5;
      }
      // camera diagonal: 0,6,11,15,18,20 in the packed upper triangle
      const int di[6] = {0, 6, 11, 15, 18, 20};
      int ok = diag_ok;
      for (int k = 0; k < 6; ++k) ok = ok && (s_sum[di[k]] > 0.0);
      s_precond = ok;
    }
    __syncthreads();
    if (it == a.n_iter) break;
    const double lambda = s_lambda;
    const int precond = s_precond;
    double Acc_tot[21], bc_tot[6];
    for (int e = 0; e < 21; ++e) Acc_tot[e] = s_sum[e];
    for (int e = 0; e < 6; ++e) bc_tot[e] = s_sum[21 + e];
    __syncthreads();

    // ---- pass 2: Schur complement of the point blocks ----
    for (int e = 0; e < NRED; ++e) acc[e] = 0.0;
    int sing = 0;
    for (int i = tid; i < N; i += NT) {
      if (!(CF(i) > 0.f)) continue;
      PointBlocks pb;
      const double p[3] = {P[3 * i], P[3 * i + 1], P[3 * i + 2]};
      point_blocks(T, p, k0[2 * i], k0[2 * i + 1], k1[2 * i], k1[2 * i + 1], (double)CF(i) * w_scale, pb);
      double M[6] = {pb.App[0], pb.App[1], pb.App[2], pb.App[3], pb.App[4], pb.App[5]};
      if (precond) {   // (A + lambda * max(diag A, 1e-12)) (:171-184)
        M[0] += lambda * fmax(pb.App[0], 1e-12); M[3] += lambda * fmax(pb.App[3], 1e-12);
        M[5] += lambda * fmax(pb.App[5], 1e-12);
      } else {
        M[0] += lambda; M[3] += lambda; M[5] += lambda;
      }
      double Mi[6];
      if (!inv3_sym(M, Mi)) { sing = 1; continue; }
      // Y = Acp Mi (6x3)
      double Y[18];
      for (int r = 0; r < 6; ++r) {
        const double a0 = pb.Acp[r * 3], a1 = pb.Acp[r * 3 + 1], a2 = pb.Acp[r * 3 + 2];
        Y[r * 3 + 0] = a0 * Mi[0] + a1 * Mi[1] + a2 * Mi[2];
        Y[r * 3 + 1] = a0 * Mi[1] + a1 * Mi[3] + a2 * Mi[4];
        Y[r * 3 + 2] = a0 * Mi[2] + a1 * Mi[4] + a2 * Mi[5];
      }
      int e = 0;
      for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c)
          acc[e++] += Y[r * 3] * pb.Acp[c * 3] + Y[r * 3 + 1] * pb.Acp[c * 3 + 1] + Y[r * 3 + 2] * pb.Acp[c * 3 + 2];
      for (int r = 0; r < 6; ++r)
        acc[21 + r] += Y[r * 3] * pb.bp[0] + Y[r * 3 + 1] * pb.bp[1] + Y[r * 3 + 2] * pb.bp[2];
    }
    sing = __syncthreads_or(sing);
    block_reduce(acc, 27, s_red, s_sum);
    if (tid == 0) {
      double S[36], g[6];
      int e = 0;
      for (int r = 0; r < 6; ++r)
        for (int c = r; c < 6; ++c) {
          double v = Acc_tot[e] - s_sum[e];
          if (r == c) v += precond ? lambda * fmax(Acc_tot[e], 1e-12) : lambda;
          S[r * 6 + c] = v; S[c * 6 + r] = v;
          ++e;
        }
      for (int r = 0; r < 6; ++r) g[r] = bc_tot[r] - s_sum[21 + r];
      int ok = !sing && lu_solve_small<6>(S, g);
      for (int r = 0; r < 6; ++r) ok = ok && isfinite(g[r]);
      s_ok = ok;
      if (ok) {
        for (int r = 0; r < 6; ++r) s_delta[r] = g[r];
        double Tn[12];
        for (int i = 0; i < 12; ++i) Tn[i] = s_T[i];
        apply_se3_update(g, Tn);
        for (int i = 0; i < 12; ++i) s_T[i] = Tn[i];
      }
    }
    __syncthreads();
    if (s_ok) {
      // back-substitution: delta_p = Mi (bp - Acp^T delta_c); uses the OLD pose blocks (T)
      double dc[6];
      for (int r = 0; r < 6; ++r) dc[r] = s_delta[r];
      for (int i = tid; i < N; i += NT) {
        if (!(CF(i) > 0.f)) continue;
        PointBlocks pb;
        const double p[3] = {P[3 * i], P[3 * i + 1], P[3 * i + 2]};
        point_blocks(T, p, k0[2 * i], k0[2 * i + 1], k1[2 * i], k1[2 * i + 1], (double)CF(i) * w_scale, pb);
        double M[6] = {pb.App[0], pb.App[1], pb.App[2], pb.App[3], pb.App[4], pb.App[5]};
        if (precond) {
          M[0] += lambda * fmax(pb.App[0], 1e-12); M[3] += lambda * fmax(pb.App[3], 1e-12);
          M[5] += lambda * fmax(pb.App[5], 1e-12);
        } else {
          M[0] += lambda; M[3] += lambda; M[5] += lambda;
        }
        double Mi[6];
        inv3_sym(M, Mi);
        double rhs[3];
        for (int j = 0; j < 3; ++j) {
          double s = pb.bp[j];
          for (int r = 0; r < 6; ++r) s -= pb.Acp[r * 3 + j] * dc[r];
          rhs[j] = s;
        }
        P[3 * i + 0] = p[0] + Mi[0] * rhs[0] + Mi[1] * rhs[1] + Mi[2] * rhs[2];
        P[3 * i + 1] = p[1] + Mi[1] * rhs[0] + Mi[3] * rhs[1] + Mi[4] * rhs[2];
        P[3 * i + 2] = p[2] + Mi[2] * rhs[0] + Mi[4] * rhs[1] + Mi[5] * rhs[2];
      }
    }
    __syncthreads();
  }
  if (tid < 12) Tout[tid] = (float)s_best[tid];
  if (tid >= 12 && tid < 16) Tout[tid] = tid == 15 ? 1.f : 0.f;
}

}  // namespace

extern "C" int mvm_ba2view(const float* kpts0_norm, const float* kpts1_norm, const float* conf,
                           const float* T_init, int batch, int n, int n_iterations, float* T_out,
                           unsigned char* valid_batch, double* pts_ws, float* trace,
                           const int* n_valid, const unsigned char* mask, void* stream) {
  MvmProfScope prof__(MVM_TAG_BA2, (cudaStream_t)stream);
  MVM_REQUIRE(kpts0_norm && kpts1_norm && conf && T_init && T_out && valid_batch && pts_ws);
  MVM_REQUIRE(batch >= 1 && n >= 1 && n_iterations >= 0);
  Ba2Args a;
  a.k0n = kpts0_norm; a.k1n = kpts1_norm; a.conf = conf; a.T_init = T_init; a.N = n;
  a.n_valid = n_valid; a.mask = mask;
  a.n_iter = n_iterations; a.T_out = T_out; a.valid_batch = valid_batch; a.pts = pts_ws;
  a.trace = trace;
  ba2_kernel<<<batch, NT, 0, (cudaStream_t)stream>>>(a);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}
