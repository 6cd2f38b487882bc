// Ground-truth matches of an image pair for the training consumers: the reference's
// compute_gt_matches_of_image_pair (helpers.py:121-203, transform_kpts :115-119, set_weight :205-213) without the
// [bs, N, N] error matrix.  Three launches:
//   gt_project_kernel   depth look-up at the (truncated) keypoint pixel, reprojection into the other view with
//                       M = K_dst . T . inv(K_src) (fp32 products in the reference's left-to-right order)
//   gt_argmin_kernel    for every keypoint the closest keypoint of the other view under the symmetric reprojection
//                       error  (|p10_j - k0_i| + |p01_i - k1_j|) / 2  -- rows and columns in one launch, the error tile
//                       never leaves registers; ties resolve to the smallest index like torch.argmin
//   gt_assign_kernel    mutual / threshold / depth-consistency tests, the "drop" sets, class-balancing weights
// Outputs follow the reference: indices [bs, 2, N+1] int64 (-1 = no match, last entry = dustbin), weights [bs, 2, N+1].
#include "../../include/mvm_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace {

struct GtArgs {
  const float* kpts[2];     // [bs, N, 2]
  const float* K[2];        // [bs, 16]
  const float* T01;         // [bs, 16]
  const float* depth[2];    // [bs, H, W]
  int bs, n, H, W;
  float max_matched, min_unmatched;
  // workspace
  float* proj[2];           // [bs, N, 2]  keypoints of view s reprojected into the other view
  float* zproj[2];          // [bs, N]     their depth in the other view
  float* d[2];              // [bs, N]     depth at the keypoint
  int* amin[2];             // [bs, N]     closest keypoint of the other view
  float* aerr[2];           // [bs, N]     error at that minimum
  long long* indices;       // [bs, 2, N+1]
  float* weights;           // [bs, 2, N+1]
};

// 4x4 inverse by Gauss-Jordan with partial pivoting in double, rounded to float (torch.linalg.inv works in fp32 LU:
// the two agree to fp32 rounding)
__device__ void inv4(const float* A, float* out) {
  double a[4][8];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) { a[i][j] = A[i * 4 + j]; a[i][4 + j] = i == j ? 1.0 : 0.0; }
  for (int c = 0; c < 4; ++c) {
    int p = c;
    for (int r = c + 1; r < 4; ++r) if (fabs(a[r][c]) > fabs(a[p][c])) p = r;
    if (p != c) for (int j = 0; j < 8; ++j) { const double t = a[c][j]; a[c][j] = a[p][j]; a[p][j] = t; }
    const double inv = 1.0 / a[c][c];
    for (int j = 0; j < 8; ++j) a[c][j] *= inv;
    for (int r = 0; r < 4; ++r) if (r != c) { const double f = a[r][c]; for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j]; }
  }
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[i * 4 + j] = (float)a[i][4 + j];
}
__device__ void mul4(const float* A, const float* B, float* C) {      // fp32, k ascending, no contraction
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = __fmul_rn(A[i * 4], B[j]);
      for (int k = 1; k < 4; ++k) s = __fadd_rn(s, __fmul_rn(A[i * 4 + k], B[k * 4 + j]));
      C[i * 4 + j] = s;
    }
}

// grid (ceil(N / 128), bs, 2): direction s = view s -> view 1 - s
__global__ void __launch_bounds__(128) gt_project_kernel(const __grid_constant__ GtArgs g) {
  const int b = blockIdx.y, s = blockIdx.z;
  __shared__ float M[16];
  if (threadIdx.x == 0) {
    float T[16], Ki[16], A[16];
    if (s == 0) for (int i = 0; i < 16; ++i) T[i] = g.T01[b * 16 + i];
    else inv4(g.T01 + b * 16, T);
    inv4(g.K[s] + b * 16, Ki);
    mul4(g.K[1 - s] + b * 16, T, A);        // (K_dst @ T) @ inv(K_src): left to right like the reference expression
    mul4(A, Ki, M);
  }
  __syncthreads();
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= g.n) return;
  const float* kp = g.kpts[s] + ((long long)b * g.n + i) * 2;
  long long x = (long long)kp[0], y = (long long)kp[1];             // .long(): truncation
  x = x < 0 ? 0 : (x >= g.W ? g.W - 1 : x);                         // (the reference would raise on an out-of-image pixel)
  y = y < 0 ? 0 : (y >= g.H ? g.H - 1 : y);
  const float dd = g.depth[s][((long long)b * g.H + y) * g.W + x];
  const float v[4] = {__fmul_rn((float)x, dd), __fmul_rn((float)y, dd), dd, 1.f};
  float p[3];
  for (int r = 0; r < 3; ++r) {
    float acc = __fmul_rn(M[r * 4], v[0]);
    for (int k = 1; k < 4; ++k) acc = __fadd_rn(acc, __fmul_rn(M[r * 4 + k], v[k]));
    p[r] = acc;
  }
  const long long o = (long long)b * g.n + i;
  g.d[s][o] = dd;
  g.zproj[s][o] = p[2];
  g.proj[s][o * 2] = __fdiv_rn(p[0], p[2]);
  g.proj[s][o * 2 + 1] = __fdiv_rn(p[1], p[2]);
}

// grid (ceil(N / 128), bs, 2): side s = every keypoint of view s finds its closest keypoint of view 1 - s
__global__ void __launch_bounds__(128) gt_argmin_kernel(const __grid_constant__ GtArgs g) {
  const int b = blockIdx.y, s = blockIdx.z, o = 1 - s;
  const int i = blockIdx.x * 128 + threadIdx.x;
  __shared__ float4 tile[128];      // (other keypoint x, y (truncated), other projection x, y)
  float ax = 0.f, ay = 0.f, pax = 0.f, pay = 0.f;
  if (i < g.n) {
    const float* kp = g.kpts[s] + ((long long)b * g.n + i) * 2;
    ax = (float)(long long)kp[0]; ay = (float)(long long)kp[1];
    pax = g.proj[s][((long long)b * g.n + i) * 2]; pay = g.proj[s][((long long)b * g.n + i) * 2 + 1];
  }
  float best = INFINITY;
  int arg = 0;
  bool have = false;
  for (int j0 = 0; j0 < g.n; j0 += 128) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    if (j < g.n) {
      const float* kq = g.kpts[o] + ((long long)b * g.n + j) * 2;
      tile[threadIdx.x] = make_float4((float)(long long)kq[0], (float)(long long)kq[1],
                                      g.proj[o][((long long)b * g.n + j) * 2], g.proj[o][((long long)b * g.n + j) * 2 + 1]);
    }
    __syncthreads();
    const int lim = min(128, g.n - j0);
    for (int jj = 0; jj < lim; ++jj) {
      const float4 t = tile[jj];
      // |proj_other_j - my keypoint| + |my projection - other keypoint_j|, halved; the reference's operation order
      const float dx1 = __fsub_rn(t.z, ax), dy1 = __fsub_rn(t.w, ay);
      const float dx2 = __fsub_rn(pax, t.x), dy2 = __fsub_rn(pay, t.y);
      const float e1 = __fsqrt_rn(__fadd_rn(__fmul_rn(dx1, dx1), __fmul_rn(dy1, dy1)));
      const float e2 = __fsqrt_rn(__fadd_rn(__fmul_rn(dx2, dx2), __fmul_rn(dy2, dy2)));
      const float e = __fdiv_rn(__fadd_rn(e1, e2), 2.0f);
      // first minimum wins; a NaN error counts as smaller than any number (torch.argmin propagates NaN)
      if (!have || e < best || (e != e && best == best)) { best = e; arg = j0 + jj; have = true; }
    }
  }
  if (i < g.n) {
    g.amin[s][(long long)b * g.n + i] = arg;
    g.aerr[s][(long long)b * g.n + i] = best;
  }
}

// one CTA per batch item
__global__ void __launch_bounds__(256) gt_assign_kernel(const __grid_constant__ GtArgs g) {
  const int b = blockIdx.x, n = g.n, nb = n + 1;
  long long* idx0 = g.indices + (long long)b * 2 * nb;
  long long* idx1 = idx0 + nb;
  float* w0 = g.weights + (long long)b * 2 * nb;
  float* w1 = w0 + nb;
  const float* d0 = g.d[0] + (long long)b * n;
  const float* d1 = g.d[1] + (long long)b * n;
  const float* z01 = g.zproj[0] + (long long)b * n;
  const float* z10 = g.zproj[1] + (long long)b * n;
  const int* rmin = g.amin[0] + (long long)b * n;
  const int* cmin = g.amin[1] + (long long)b * n;
  const float* rerr = g.aerr[0] + (long long)b * n;
  const float* cerr = g.aerr[1] + (long long)b * n;
  __shared__ int s_match, s_drop;
  if (threadIdx.x == 0) { s_match = 0; s_drop = 0; }
  for (int i = threadIdx.x; i < nb; i += blockDim.x) { idx0[i] = -1; idx1[i] = -1; w0[i] = 0.f; w1[i] = 0.f; }
  __syncthreads();
  int n_match = 0, n_drop = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int i1 = rmin[i];
    const bool both = cmin[i1] == i;
    const bool small10 = rerr[i] <= g.max_matched;
    const float dd0 = d0[i], md1 = d1[i1];
    const bool v0 = dd0 > 1e-6f, v1 = md1 > 1e-6f;
    bool m = both && small10 && v0 && v1;
    if (m) m = (__fdiv_rn(fabsf(__fsub_rn(z01[i], md1)), md1) < 0.1f) && (__fdiv_rn(fabsf(__fsub_rn(z10[i1], dd0)), dd0) < 0.1f);
    if (m) { idx0[i] = i1; idx1[i1] = i; ++n_match; }                 // mutual minima: i1 is written by one thread only
    else if (!v0 || !v1 || rerr[i] <= g.min_unmatched) { w0[i] = -1.f; ++n_drop; }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    if (idx1[j] != -1) continue;
    const int i0 = cmin[j];
    const bool v1 = d1[j] > 1e-6f, v0 = d0[i0] > 1e-6f;
    if (!v0 || !v1 || cerr[j] <= g.min_unmatched) { w1[j] = -1.f; ++n_drop; }
  }
  atomicAdd(&s_match, n_match);
  atomicAdd(&s_drop, n_drop);
  __syncthreads();
  // class balancing (helpers.py:190-198): float32 like the reference's tensors
  float mw = __fdiv_rn(__fmul_rn(2.f, (float)s_match), __fsub_rn(__fmul_rn(2.f, (float)n), (float)s_drop));
  float uw = __fdiv_rn(0.5f, __fsub_rn(1.f, mw));
  mw = __fdiv_rn(0.5f, mw);
  if (!(isfinite(mw) && isfinite(uw))) { mw = 0.f; uw = 0.f; }
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    w0[i] = w0[i] == -1.f ? 0.f : (idx0[i] == -1 ? uw : mw);
    w1[i] = w1[i] == -1.f ? 0.f : (idx1[i] == -1 ? uw : mw);
  }
}

}  // namespace

extern "C" size_t mvm_gt_matches_workspace_bytes(int bs, int n) {
  // 2 x (proj [n,2] + zproj + d + aerr) floats + 2 x amin ints, per batch item
  return (size_t)bs * n * (2 * (2 + 1 + 1 + 1) * sizeof(float) + 2 * sizeof(int)) + 256;
}

extern "C" int mvm_gt_matches_pair(const float* kpts0, const float* kpts1, const float* K0, const float* K1,
                                   const float* T0to1, const float* depth0, const float* depth1, int bs, int n, int H,
                                   int W, float max_matched_reproj_err, float min_unmatched_reproj_err,
                                   long long* indices, float* weights, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  MVM_REQUIRE(kpts0 && kpts1 && K0 && K1 && T0to1 && depth0 && depth1 && indices && weights && workspace);
  MVM_REQUIRE(bs >= 1 && n >= 1 && H >= 1 && W >= 1);
  if (workspace_bytes < mvm_gt_matches_workspace_bytes(bs, n)) return MVM_ERR_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  GtArgs g;
  g.kpts[0] = kpts0; g.kpts[1] = kpts1; g.K[0] = K0; g.K[1] = K1; g.T01 = T0to1; g.depth[0] = depth0; g.depth[1] = depth1;
  g.bs = bs; g.n = n; g.H = H; g.W = W; g.max_matched = max_matched_reproj_err; g.min_unmatched = min_unmatched_reproj_err;
  float* f = reinterpret_cast<float*>(workspace);
  const size_t bn = (size_t)bs * n;
  for (int v = 0; v < 2; ++v) { g.proj[v] = f; f += 2 * bn; g.zproj[v] = f; f += bn; g.d[v] = f; f += bn; g.aerr[v] = f; f += bn; }
  int* ip = reinterpret_cast<int*>(f);
  for (int v = 0; v < 2; ++v) { g.amin[v] = ip; ip += bn; }
  g.indices = indices; g.weights = weights;
  MvmProfScope prof__(MVM_TAG_MISC, s);
  const dim3 grid(mvm_div_up(n, 128), bs, 2);
  gt_project_kernel<<<grid, 128, 0, s>>>(g);
  MVM_CHECK_LAUNCH();
  gt_argmin_kernel<<<grid, 128, 0, s>>>(g);
  MVM_CHECK_LAUNCH();
  gt_assign_kernel<<<bs, 256, 0, s>>>(g);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}
