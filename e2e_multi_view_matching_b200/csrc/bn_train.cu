// BatchNorm1d in TRAINING mode on point-major activations (the train branch of the matcher, SURVEY.md 8 a8 / f-2:
// multi_view_matcher.py:8-22 MLP = Conv1d + BatchNorm1d + ReLU; statistics over all B*T*N points of the call):
//   y = act(gamma * (x - mean_c) / sqrt(var_c + eps) + beta),  var biased for the normalisation,
//   running_mean <- (1 - momentum) running_mean + momentum mean,  running_var likewise with the UNBIASED variance
// x [rows, C] with row stride ld; only rows whose index inside their n_pad-row view slot is < n_valid count (the
// padding rows of a slot are left untouched), and only the slots s with s % slot_mod == slot_rem (the pairwise train
// path normalises every view on its own, multi_view_matcher.py:169-173 / superglue.py:131-140).  Two launches: per-channel sums in double (shifted by the first valid
// row: no cancellation), then normalise in place.
#include "../../include/mvm_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace {

// grid (row blocks), block = C threads (<= 1024): thread c accumulates channel c over the block's rows
__global__ void bn_sums_kernel(const float* __restrict__ x, int rows, int C, int ld, int n_pad, int n_valid,
                               int slot_mod, int slot_rem, int rows_per_block, double* sums) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const float shift = (float)sums[2 * C + c];   // row 0 of the buffer (always a valid row), stored by bn_shift_kernel
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  double s = 0.0, q = 0.0;
  for (int r = r0; r < r1; ++r) {
    if (r % n_pad >= n_valid || (r / n_pad) % slot_mod != slot_rem) continue;
    const double d = (double)(x[(long long)r * ld + c] - shift);
    s += d; q += d * d;
  }
  atomicAdd(sums + c, s);
  atomicAdd(sums + C + c, q);
}

__global__ void bn_apply_kernel(float* __restrict__ x, int rows, int C, int ld, int n_pad, int n_valid, int slot_mod,
                                int slot_rem,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int relu,
                                float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                const double* __restrict__ sums, long long count) {
  const int c = threadIdx.x;
  if (c >= C) return;
  const double sh = sums[2 * C + c];              // row 0's ORIGINAL value (block 0 rewrites row 0 below)
  const double n = (double)count;
  const double mean_s = sums[c] / n;                                    // mean of (x - shift)
  const double var = fmax(sums[C + c] / n - mean_s * mean_s, 0.0);      // biased
  const double mean = mean_s + sh;
  const float scale = (float)((double)gamma[c] / sqrt(var + (double)eps));
  const float bias = (float)((double)beta[c] - mean * (double)scale);
  if (blockIdx.x == 0 && running_mean != nullptr) {
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
  }
  const int rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (int r = r0; r < r1; ++r) {
    if (r % n_pad >= n_valid || (r / n_pad) % slot_mod != slot_rem) continue;
    float* p = x + (long long)r * ld + c;
    float y = fmaf(*p, scale, bias);
    if (relu) y = fmaxf(y, 0.f);
    *p = y;
  }
}

__global__ void bn_shift_kernel(const float* __restrict__ x, int C, long long first_row_off, double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { sums[c] = 0.0; sums[C + c] = 0.0; sums[2 * C + c] = (double)x[first_row_off + c]; }
}

}  // namespace

extern "C" int mvm_batchnorm_train(float* x, int rows, int C, int ld, int n_pad, int n_valid, int slot_mod,
                                   int slot_rem, const float* gamma,
                                   const float* beta, float eps, int relu, float* running_mean, float* running_var,
                                   float momentum, double* ws /* 3 C doubles */, void* stream) {
  MVM_REQUIRE(x && gamma && beta && ws && rows >= 1 && C >= 1 && C <= 1024 && ld >= C);
  MVM_REQUIRE(n_pad >= 1 && n_valid >= 1 && n_valid <= n_pad && rows % n_pad == 0);
  MVM_REQUIRE(slot_mod >= 1 && slot_rem >= 0 && slot_rem < slot_mod && (rows / n_pad) % slot_mod == 0);
  MVM_REQUIRE((running_mean == nullptr) == (running_var == nullptr));
  cudaStream_t s = (cudaStream_t)stream;
  MvmProfScope prof__(MVM_TAG_MISC, s);
  const int threads = ((C + 31) / 32) * 32;
  const int blocks = rows < 592 ? rows : 592;                 // 4 x 148
  const int rpb = (rows + blocks - 1) / blocks;
  const long long count = (long long)(rows / n_pad / slot_mod) * n_valid;
  bn_shift_kernel<<<(C + 255) / 256, 256, 0, s>>>(x, C, (long long)slot_rem * n_pad * ld, ws);
  MVM_CHECK_LAUNCH();
  bn_sums_kernel<<<blocks, threads, 0, s>>>(x, rows, C, ld, n_pad, n_valid, slot_mod, slot_rem, rpb, ws);
  MVM_CHECK_LAUNCH();
  bn_apply_kernel<<<blocks, threads, 0, s>>>(x, rows, C, ld, n_pad, n_valid, slot_mod, slot_rem, gamma, beta, eps, relu, running_mean,
                                             running_var, momentum, ws, count);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}
