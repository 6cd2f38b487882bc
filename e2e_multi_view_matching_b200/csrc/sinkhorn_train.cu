// log_optimal_transport for TRAINING: the forward of superglue.py:143-172 keeping the potentials of
// every iteration, and the EXACT gradient of the unrolled iterations -- what autograd computes for the reference.
//   forward   u^t = log_mu - LSE_j(Z + v^(t-1)),  v^t = log_nu - LSE_i(Z + u^t),  t = 1..T,  out = Z + u^T + v^T - norm
//   backward  gu = rowsum G, gv = colsum G, dZ = G;  for t = T..1:
//               P = exp(Z + u^t + v^t - log_nu)   (columns sum to 1):  W = P gv_j;  dZ -= W;  gu_i -= sum_j W
//               Q = exp(Z + u^t - log_mu + v^(t-1)) (rows sum to 1):   W = Q gu_i;  dZ -= W;  gv_j = -sum_i W;  gu = 0
//             d scores = dZ[:m, :n],  d alpha = sum of dZ over the dustbin row and column
// (Z is the augmented matrix; every exponent above is <= 0, so nothing overflows however far the scores of an untrained
// network spread.)  One CTA per problem, every pair of a training step in ONE launch; the matrix and dZ stay in L2
// (0.64 MB each at 400 keypoints).  Row-direction reductions: a warp per row.  Column-direction reductions: every warp
// walks its rows with per-lane accumulators for the columns lane + 32 k (coalesced 128-byte row segments, KC independent
// chains per thread), then the 32 per-warp partials of a column are merged through shared memory in fixed order
// (deterministic) -- a thread-per-column loop over all rows was a 400-deep dependent chain on 13 warps (290 ms per
// cfg5 step before).  expf on the conditional-rescale online log-sum-exp: ~1 exponential per element.
#include "../../include/mvm_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace {

struct SkArgs {
  const float* scores;   // problem b: rows of n scores, row stride sld, at scores + b * sstride ([B, m, n] packed, or the
                         // inner block of [B, m+1, n+1] buffers as the score GEMM of the matcher writes them)
  long long sld, sstride;
  float* out;            // forward: couplings [B, m+1, n+1];  backward: dZ [B, m+1, n+1] (holds G on entry)
  float* pot;            // [B, iters, m + n + 2]  (u^t | v^t)
  const float* alpha_p;  // device scalar (bin_score)
  double* d_alpha;       // backward: accumulated over the problems
  int m, n, iters;
};

__device__ __forceinline__ float zin_(const float* __restrict__ sc, long long sld, int i, int j, int m, int n, float alpha) {
  return (i < m && j < n) ? sc[(long long)i * sld + j] : alpha;
}
#define zin(sc, i, j, m, n, alpha) zin_(sc, sld, i, j, m, n, alpha)

constexpr int CH = 416;   // columns merged per shared-memory round (13 x 32)
constexpr int KB = 7;     // columns per thread whose loads are issued together

// online log-sum-exp update with one exponential in the common case.  Every exponent in this file is <= 0 (a difference
// to a running maximum, or a log-probability): __expf (ex2.approx of x log2 e) is accurate to ~1e-7 relative there.
__device__ __forceinline__ void lse_push(float& mx, float& s, float x) {
  if (x > mx) { s = s * expf(mx - x) + 1.f; mx = x; }
  else s += expf(x - mx);
}
__device__ __forceinline__ void lse_merge(float& mx, float& s, float mo, float so) {
  if (mo == -INFINITY) return;
  if (mo > mx) { s = s * expf(mx - mo) + so; mx = mo; }
  else s += so * expf(mo - mx);
}

template <int KC>
__global__ void __launch_bounds__(1024) sinkhorn_train_fwd_kernel(const SkArgs g) {
  extern __shared__ float sm[];
  const int m = g.m, n = g.n, b = blockIdx.x;
  float* u = sm;                 // m + 1
  float* v = u + (m + 1);        // n + 1
  float* pmx = v + (n + 1);      // [32][CH] per-warp column partials
  float* psm = pmx + 32 * CH;
  const float* sc = g.scores + (long long)b * g.sstride;
  const long long sld = g.sld;
  const float alpha = *g.alpha_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float norm = -logf((float)(m + n));
  const float log_mu_bin = logf((float)n) + norm, log_nu_bin = logf((float)m) + norm;
  for (int j = tid; j <= n; j += 1024) v[j] = 0.f;
  __syncthreads();
  for (int it = 0; it < g.iters; ++it) {
    float* pot = g.pot + ((long long)b * g.iters + it) * (m + n + 2);
    // rows: u_i = log_mu_i - LSE_j(Z_ij + v_j)
    for (int i = warp; i <= m; i += 32) {
      // chunks of KB columns in two phases: the loads of a chunk are in flight together before the dependent
      // log-sum-exp chain consumes them (a load -> push -> load loop serialised 13 L2 latencies per row)
      float mx = -INFINITY, s = 0.f;
#pragma unroll
      for (int k0 = 0; k0 < KC; k0 += KB) {
        float z[KB];
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int j = lane + 32 * (k0 + kk);
          z[kk] = (k0 + kk < KC && j <= n) ? zin(sc, i, j, m, n, alpha) : 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int j = lane + 32 * (k0 + kk);
          if (k0 + kk < KC && j <= n) lse_push(mx, s, z[kk] + v[j]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, mx, o), so = __shfl_xor_sync(0xffffffffu, s, o);
        lse_merge(mx, s, mo, so);
      }
      if (lane == 0) { const float r = (i < m ? norm : log_mu_bin) - (logf(s) + mx); u[i] = r; pot[i] = r; }
    }
    __syncthreads();
    // columns: v_j = log_nu_j - LSE_i(Z_ij + u_i)
    float cmx[KC], csm[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) { cmx[k] = -INFINITY; csm[k] = 0.f; }
    for (int i = warp; i <= m; i += 32) {
      const float ui = u[i];
#pragma unroll
      for (int k0 = 0; k0 < KC; k0 += KB) {
        float z[KB];
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int j = lane + 32 * (k0 + kk);
          z[kk] = (k0 + kk < KC && j <= n) ? zin(sc, i, j, m, n, alpha) : 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int j = lane + 32 * (k0 + kk);
          if (k0 + kk < KC && j <= n) lse_push(cmx[k0 + kk], csm[k0 + kk], z[kk] + ui);
        }
      }
    }
    for (int c0 = 0; c0 <= n; c0 += CH) {
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int j = lane + 32 * k;
        if (j >= c0 && j < c0 + CH) { pmx[warp * CH + j - c0] = cmx[k]; psm[warp * CH + j - c0] = csm[k]; }
      }
      __syncthreads();
      for (int jj = tid; jj < CH && c0 + jj <= n; jj += 1024) {
        float mx = -INFINITY, s = 0.f;
        for (int w = 0; w < 32; ++w) lse_merge(mx, s, pmx[w * CH + jj], psm[w * CH + jj]);
        const int j = c0 + jj;
        const float r = (j < n ? norm : log_nu_bin) - (logf(s) + mx);
        v[j] = r;
        pot[m + 1 + j] = r;
      }
      __syncthreads();
    }
  }
  float* out = g.out + (long long)b * (m + 1) * (n + 1);
  for (int i = warp; i <= m; i += 32)
    for (int j = lane; j <= n; j += 32) out[(long long)i * (n + 1) + j] = zin(sc, i, j, m, n, alpha) + u[i] + v[j] - norm;
}

template <int KC>
__global__ void __launch_bounds__(1024) sinkhorn_train_bwd_kernel(const SkArgs g) {
  extern __shared__ float sm[];
  const int m = g.m, n = g.n, b = blockIdx.x, ld = n + 1;
  float* u = sm;                // m + 1   u^t
  float* v = u + (m + 1);       // n + 1   v^t - log_nu
  float* vp = v + (n + 1);      // n + 1   v^(t-1)
  float* gu = vp + (n + 1);     // m + 1
  float* gv = gu + (m + 1);     // n + 1
  float* part = gv + (n + 1);   // [32][CH] per-warp column partial sums
  __shared__ double red[32];
  const float* sc = g.scores + (long long)b * g.sstride;
  const long long sld = g.sld;
  float* dZ = g.out + (long long)b * (m + 1) * ld;
  const float alpha = *g.alpha_p;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float norm = -logf((float)(m + n));
  const float log_mu_bin = logf((float)n) + norm, log_nu_bin = logf((float)m) + norm;
  float acc[KC];
  // merges the per-warp column partials `acc` into gv_j = sign * sum (fixed order)
  auto merge_columns = [&](float sign) {
    for (int c0 = 0; c0 <= n; c0 += CH) {
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        const int j = lane + 32 * k;
        if (j >= c0 && j < c0 + CH) part[warp * CH + j - c0] = acc[k];
      }
      __syncthreads();
      for (int jj = tid; jj < CH && c0 + jj <= n; jj += 1024) {
        float s = 0.f;
        for (int w = 0; w < 32; ++w) s += part[w * CH + jj];
        gv[c0 + jj] = sign * s;
      }
      __syncthreads();
    }
  };
  // gu = row sums of G, gv = column sums of G (dZ already holds G)
#pragma unroll
  for (int k = 0; k < KC; ++k) acc[k] = 0.f;
  for (int i = warp; i <= m; i += 32) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
      const int j = lane + 32 * k;
      if (j <= n) { const float x = dZ[(long long)i * ld + j]; s += x; acc[k] += x; }
    }
    s = warp_sum(s);
    if (lane == 0) gu[i] = s;
  }
  merge_columns(1.f);
  for (int t = g.iters; t >= 1; --t) {
    const float* pot = g.pot + ((long long)b * g.iters + (t - 1)) * (m + n + 2);
    for (int i = tid; i <= m; i += 1024) u[i] = pot[i];
    for (int j = tid; j <= n; j += 1024) {
      v[j] = pot[m + 1 + j] - (j < n ? norm : log_nu_bin);          // v^t - log_nu
      vp[j] = t > 1 ? (pot - (m + n + 2))[m + 1 + j] : 0.f;         // v^(t-1)
    }
    __syncthreads();
    // pass A (rows): W = exp(Z + u^t + v^t - log_nu) gv_j;  dZ -= W;  gu_i -= sum_j W
    for (int i = warp; i <= m; i += 32) {
      const float ui = u[i];
      float a = 0.f;
      // chunks of KB columns: all loads of a chunk first (the stores to dZ would otherwise order every later load behind
      // them: the compiler cannot prove that dZ and the scores do not alias), then the exponentials, then the stores
#pragma unroll
      for (int k0 = 0; k0 < KC; k0 += KB) {
        float z[KB], d[KB];
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int j = lane + 32 * (k0 + kk);
          const bool ok = k0 + kk < KC && j <= n;
          z[kk] = ok ? zin(sc, i, j, m, n, alpha) : 0.f;
          d[kk] = ok ? dZ[(long long)i * ld + j] : 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int j = lane + 32 * (k0 + kk);
          if (k0 + kk < KC && j <= n) {
            const float w = expf(z[kk] + ui + v[j]) * gv[j];
            dZ[(long long)i * ld + j] = d[kk] - w;
            a += w;
          }
        }
      }
      a = warp_sum(a);
      if (lane == 0) gu[i] -= a;
    }
    __syncthreads();
    // pass B (columns): W = exp(Z + u^t - log_mu + v^(t-1)) gu_i;  dZ -= W;  gv_j = -sum_i W;  gu = 0
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[k] = 0.f;
    for (int i = warp; i <= m; i += 32) {
      const float ui = u[i] - (i < m ? norm : log_mu_bin), gi = gu[i];
#pragma unroll
      for (int k0 = 0; k0 < KC; k0 += KB) {
        float z[KB], d[KB];
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int j = lane + 32 * (k0 + kk);
          const bool ok = k0 + kk < KC && j <= n;
          z[kk] = ok ? zin(sc, i, j, m, n, alpha) : 0.f;
          d[kk] = ok ? dZ[(long long)i * ld + j] : 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < KB; ++kk) {
          const int j = lane + 32 * (k0 + kk);
          if (k0 + kk < KC && j <= n) {
            const float w = expf(z[kk] + ui + vp[j]) * gi;
            dZ[(long long)i * ld + j] = d[kk] - w;
            acc[k0 + kk] += w;
          }
        }
      }
    }
    merge_columns(-1.f);          // (its barriers also order pass B's reads of gu before the reset below)
    for (int i = tid; i <= m; i += 1024) gu[i] = 0.f;
    __syncthreads();
  }
  // d alpha: dustbin row and column
  double s = 0.0;
  for (int j = tid; j <= n; j += 1024) s += (double)dZ[(long long)m * ld + j];
  for (int i = tid; i < m; i += 1024) s += (double)dZ[(long long)i * ld + n];
  s = warp_sum_d(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    s = red[lane];
    s = warp_sum_d(s);
    if (lane == 0) atomicAdd(g.d_alpha, s);
  }
}

constexpr int SK_MAX_N = 1055;    // KC = 33 covers n + 1 <= 1056 columns
inline int fwd_smem(int m, int n) { return (m + n + 2 + 2 * 32 * CH) * 4; }
inline int bwd_smem(int m, int n) { return (2 * (m + 1) + 3 * (n + 1) + 32 * CH) * 4; }
void set_attrs() {
  mvm_once_per_device(MVM_ONCE_SINKHORN_TRAIN, [&] {
    cudaFuncSetAttribute(sinkhorn_train_fwd_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(SK_MAX_N, SK_MAX_N));
    cudaFuncSetAttribute(sinkhorn_train_fwd_kernel<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, fwd_smem(SK_MAX_N, SK_MAX_N));
    cudaFuncSetAttribute(sinkhorn_train_bwd_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(SK_MAX_N, SK_MAX_N));
    cudaFuncSetAttribute(sinkhorn_train_bwd_kernel<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem(SK_MAX_N, SK_MAX_N));
  });
}

}  // namespace

extern "C" size_t mvm_sinkhorn_train_pot_floats(int batch, int m, int n, int iters) {
  return (size_t)batch * (size_t)iters * (size_t)(m + n + 2);
}

extern "C" int mvm_sinkhorn_train_forward(const float* scores, long long scores_ld, long long scores_stride, const float* alpha,
                                          int batch, int m, int n, int iters, float* out, float* pot, void* stream) {
  MVM_REQUIRE(scores_ld >= n && scores_stride >= 0);
  MVM_REQUIRE(scores && alpha && out && pot && batch >= 1 && m >= 1 && n >= 1 && iters >= 1 && m <= SK_MAX_N && n <= SK_MAX_N);
  SkArgs g;
  g.scores = scores; g.sld = scores_ld; g.sstride = scores_stride; g.out = out; g.pot = pot; g.alpha_p = alpha; g.d_alpha = nullptr; g.m = m; g.n = n; g.iters = iters;
  const int smem = fwd_smem(m, n);
  set_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  MvmProfScope prof__(MVM_TAG_SINKHORN, s);
  if (n + 1 <= 13 * 32) sinkhorn_train_fwd_kernel<13><<<batch, 1024, smem, s>>>(g);
  else sinkhorn_train_fwd_kernel<33><<<batch, 1024, smem, s>>>(g);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}

// dZ: [B, m+1, n+1], holds the gradient w.r.t. the couplings on entry and the gradient w.r.t. the augmented score matrix
// on return (its inner block is d scores); d_alpha (one double, zeroed by the caller) += the dustbin entries.
extern "C" int mvm_sinkhorn_train_backward(const float* scores, long long scores_ld, long long scores_stride, const float* alpha,
                                           const float* pot, int batch, int m, int n, int iters, float* dZ, double* d_alpha,
                                           void* stream) {
  MVM_REQUIRE(scores_ld >= n && scores_stride >= 0);
  MVM_REQUIRE(scores && alpha && pot && dZ && d_alpha && batch >= 1 && m >= 1 && n >= 1 && iters >= 1 && m <= SK_MAX_N && n <= SK_MAX_N);
  SkArgs g;
  g.scores = scores; g.sld = scores_ld; g.sstride = scores_stride; g.out = dZ; g.pot = const_cast<float*>(pot); g.alpha_p = alpha; g.d_alpha = d_alpha;
  g.m = m; g.n = n; g.iters = iters;
  const int smem = bwd_smem(m, n);
  set_attrs();
  cudaStream_t s = (cudaStream_t)stream;
  MvmProfScope prof__(MVM_TAG_SINKHORN, s);
  if (n + 1 <= 13 * 32) sinkhorn_train_bwd_kernel<13><<<batch, 1024, smem, s>>>(g);
  else sinkhorn_train_bwd_kernel<33><<<batch, 1024, smem, s>>>(g);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}
