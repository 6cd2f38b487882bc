// log_optimal_transport for TRAINING: the forward of superglue.py:143-172 keeping the potentials of
// every iteration, and the EXACT gradient of the unrolled iterations -- what autograd computes for the reference.
//   forward   u^t = log_mu - LSE_j(Z + v^(t-1)),  v^t = log_nu - LSE_i(Z + u^t),  t = 1..T,  out = Z + u^T + v^T - norm
//   backward  gu = rowsum G, gv = colsum G, dZ = G;  for t = T..1:
//               P = exp(Z + u^t + v^t - log_nu)   (columns sum to 1):  W = P gv_j;  dZ -= W;  gu_i -= sum_j W
//               Q = exp(Z + u^t - log_mu + v^(t-1)) (rows sum to 1):   W = Q gu_i;  dZ -= W;  gv_j = -sum_i W;  gu = 0
//             d scores = dZ[:m, :n],  d alpha = sum of dZ over the dustbin row and column
// (Z is the augmented matrix; every exponent above is <= 0, so nothing overflows however far the scores of an untrained
// network spread.)  One CTA per problem; the matrix and dZ stay in L2 (0.64 MB each at 400 keypoints); row-direction
// reductions by a warp per row, column-direction reductions by a thread per column (coalesced across the warp).
#include "../../include/mvm_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace {

struct SkArgs {
  const float* scores;   // [B, m, n]
  float* out;            // forward: couplings [B, m+1, n+1];  backward: dZ [B, m+1, n+1] (holds G on entry)
  float* pot;            // [B, iters, m + n + 2]  (u^t | v^t)
  const float* alpha_p;  // device scalar (bin_score)
  double* d_alpha;       // backward: accumulated over the problems
  int m, n, iters;
};

__device__ __forceinline__ float zin(const float* __restrict__ sc, int i, int j, int m, int n, float alpha) {
  return (i < m && j < n) ? sc[(long long)i * n + j] : alpha;
}

__global__ void __launch_bounds__(1024) sinkhorn_train_fwd_kernel(const SkArgs g) {
  extern __shared__ float sm[];
  const int m = g.m, n = g.n, b = blockIdx.x;
  float* u = sm;               // m + 1
  float* v = u + (m + 1);      // n + 1
  const float* sc = g.scores + (long long)b * m * n;
  const float alpha = *g.alpha_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float norm = -logf((float)(m + n));
  const float log_mu_bin = logf((float)n) + norm, log_nu_bin = logf((float)m) + norm;
  for (int j = tid; j <= n; j += blockDim.x) v[j] = 0.f;
  __syncthreads();
  for (int it = 0; it < g.iters; ++it) {
    float* pot = g.pot + ((long long)b * g.iters + it) * (m + n + 2);
    for (int i = warp; i <= m; i += nwarps) {
      float mx = -INFINITY, s = 0.f;
      for (int j = lane; j <= n; j += 32) {
        const float x = zin(sc, i, j, m, n, alpha) + v[j];
        const float mn = fmaxf(mx, x);
        s = s * expf(mx - mn) + expf(x - mn);
        mx = mn;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, mx, o), so = __shfl_xor_sync(0xffffffffu, s, o);
        const float mn = fmaxf(mx, mo);
        s = (mx == -INFINITY ? 0.f : s * expf(mx - mn)) + (mo == -INFINITY ? 0.f : so * expf(mo - mn));
        mx = mn;
      }
      if (lane == 0) { const float r = (i < m ? norm : log_mu_bin) - (logf(s) + mx); u[i] = r; pot[i] = r; }
    }
    __syncthreads();
    for (int j = tid; j <= n; j += blockDim.x) {
      float mx = -INFINITY, s = 0.f;
      for (int i = 0; i <= m; ++i) {
        const float x = zin(sc, i, j, m, n, alpha) + u[i];
        const float mn = fmaxf(mx, x);
        s = s * expf(mx - mn) + expf(x - mn);
        mx = mn;
      }
      const float r = (j < n ? norm : log_nu_bin) - (logf(s) + mx);
      v[j] = r;
      pot[m + 1 + j] = r;
    }
    __syncthreads();
  }
  float* out = g.out + (long long)b * (m + 1) * (n + 1);
  for (int i = warp; i <= m; i += nwarps)
    for (int j = lane; j <= n; j += 32) out[(long long)i * (n + 1) + j] = zin(sc, i, j, m, n, alpha) + u[i] + v[j] - norm;
}

__global__ void __launch_bounds__(1024) sinkhorn_train_bwd_kernel(const SkArgs g) {
  extern __shared__ float sm[];
  const int m = g.m, n = g.n, b = blockIdx.x, ld = n + 1;
  float* u = sm;                // m + 1   u^t
  float* v = u + (m + 1);       // n + 1   v^t
  float* vp = v + (n + 1);      // n + 1   v^(t-1)
  float* gu = vp + (n + 1);     // m + 1
  float* gv = gu + (m + 1);     // n + 1
  __shared__ double red[32];
  const float* sc = g.scores + (long long)b * m * n;
  float* dZ = g.out + (long long)b * (m + 1) * ld;
  const float alpha = *g.alpha_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float norm = -logf((float)(m + n));
  const float log_mu_bin = logf((float)n) + norm, log_nu_bin = logf((float)m) + norm;
  // gu = row sums of G, gv = column sums of G (dZ already holds G)
  for (int i = warp; i <= m; i += nwarps) {
    float s = 0.f;
    for (int j = lane; j <= n; j += 32) s += dZ[(long long)i * ld + j];
    s = warp_sum(s);
    if (lane == 0) gu[i] = s;
  }
  for (int j = tid; j <= n; j += blockDim.x) {
    float s0 = 0.f, s1 = 0.f;
    int i = 0;
    for (; i + 1 <= m; i += 2) { s0 += dZ[(long long)i * ld + j]; s1 += dZ[(long long)(i + 1) * ld + j]; }
    if (i <= m) s0 += dZ[(long long)i * ld + j];
    gv[j] = s0 + s1;
  }
  __syncthreads();
  for (int t = g.iters; t >= 1; --t) {
    const float* pot = g.pot + ((long long)b * g.iters + (t - 1)) * (m + n + 2);
    for (int i = tid; i <= m; i += blockDim.x) u[i] = pot[i];
    for (int j = tid; j <= n; j += blockDim.x) {
      v[j] = pot[m + 1 + j] - (j < n ? norm : log_nu_bin);          // v^t - log_nu
      vp[j] = t > 1 ? (pot - (m + n + 2))[m + 1 + j] : 0.f;         // v^(t-1)
    }
    __syncthreads();
    // pass A (rows): W = exp(Z + u^t + v^t - log_nu) gv_j
    for (int i = warp; i <= m; i += nwarps) {
      const float ui = u[i];
      float acc = 0.f;
      for (int j = lane; j <= n; j += 32) {
        const float w = expf(zin(sc, i, j, m, n, alpha) + ui + v[j]) * gv[j];
        dZ[(long long)i * ld + j] -= w;
        acc += w;
      }
      acc = warp_sum(acc);
      if (lane == 0) gu[i] -= acc;
    }
    __syncthreads();
    // pass B (columns): W = exp(Z + u^t - log_mu + v^(t-1)) gu_i;  u is shifted by -log_mu in place first
    for (int i = tid; i <= m; i += blockDim.x) u[i] -= (i < m ? norm : log_mu_bin);
    __syncthreads();
    for (int j = tid; j <= n; j += blockDim.x) {
      const float vj = vp[j];
      float a0 = 0.f, a1 = 0.f;
      int i = 0;
      for (; i + 1 <= m; i += 2) {
        const float w0 = expf(zin(sc, i, j, m, n, alpha) + u[i] + vj) * gu[i];
        const float w1 = expf(zin(sc, i + 1, j, m, n, alpha) + u[i + 1] + vj) * gu[i + 1];
        dZ[(long long)i * ld + j] -= w0;
        dZ[(long long)(i + 1) * ld + j] -= w1;
        a0 += w0; a1 += w1;
      }
      if (i <= m) {
        const float w0 = expf(zin(sc, i, j, m, n, alpha) + u[i] + vj) * gu[i];
        dZ[(long long)i * ld + j] -= w0;
        a0 += w0;
      }
      gv[j] = -(a0 + a1);
    }
    __syncthreads();
    for (int i = tid; i <= m; i += blockDim.x) gu[i] = 0.f;
    __syncthreads();
  }
  // d alpha: dustbin row and column
  double s = 0.0;
  for (int j = tid; j <= n; j += blockDim.x) s += (double)dZ[(long long)m * ld + j];
  for (int i = tid; i < m; i += blockDim.x) s += (double)dZ[(long long)i * ld + n];
  s = warp_sum_d(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = lane < nwarps ? red[lane] : 0.0;
    s = warp_sum_d(s);
    if (lane == 0) atomicAdd(g.d_alpha, s);
  }
}

}  // namespace

extern "C" size_t mvm_sinkhorn_train_pot_floats(int batch, int m, int n, int iters) {
  return (size_t)batch * (size_t)iters * (size_t)(m + n + 2);
}

extern "C" int mvm_sinkhorn_train_forward(const float* scores, const float* alpha, int batch, int m, int n, int iters,
                                          float* out, float* pot, void* stream) {
  MVM_REQUIRE(scores && alpha && out && pot && batch >= 1 && m >= 1 && n >= 1 && iters >= 1 && m <= 8192 && n <= 8192);
  SkArgs g;
  g.scores = scores; g.out = out; g.pot = pot; g.alpha_p = alpha; g.d_alpha = nullptr; g.m = m; g.n = n; g.iters = iters;
  const int smem = (m + n + 2) * 4;
  mvm_once_per_device(MVM_ONCE_SINKHORN_TRAIN, [&] {
    cudaFuncSetAttribute(sinkhorn_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8194 * 4);
    cudaFuncSetAttribute(sinkhorn_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 8194 * 4);
  });
  cudaStream_t s = (cudaStream_t)stream;
  MvmProfScope prof__(MVM_TAG_SINKHORN, s);
  sinkhorn_train_fwd_kernel<<<batch, 1024, smem, s>>>(g);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}

// dZ: [B, m+1, n+1], holds the gradient w.r.t. the couplings on entry and the gradient w.r.t. the augmented score matrix
// on return (its inner block is d scores); d_alpha (one double, zeroed by the caller) += the dustbin entries.
extern "C" int mvm_sinkhorn_train_backward(const float* scores, const float* alpha, const float* pot, int batch, int m, int n,
                                           int iters, float* dZ, double* d_alpha, void* stream) {
  MVM_REQUIRE(scores && alpha && pot && dZ && d_alpha && batch >= 1 && m >= 1 && n >= 1 && iters >= 1 && m <= 8192 && n <= 8192);
  SkArgs g;
  g.scores = scores; g.out = dZ; g.pot = const_cast<float*>(pot); g.alpha_p = alpha; g.d_alpha = d_alpha;
  g.m = m; g.n = n; g.iters = iters;
  const int smem = (2 * (m + 1) + 3 * (n + 1)) * 4;
  mvm_once_per_device(MVM_ONCE_SINKHORN_TRAIN, [&] {
    cudaFuncSetAttribute(sinkhorn_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8194 * 4);
    cudaFuncSetAttribute(sinkhorn_train_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 8194 * 4);
  });
  cudaStream_t s = (cudaStream_t)stream;
  MvmProfScope prof__(MVM_TAG_SINKHORN, s);
  sinkhorn_train_bwd_kernel<<<batch, 1024, smem, s>>>(g);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}
