// Keypoint encoder front (3->32->64->128 with folded BN + ReLU) and the channel-first ->
// point-major descriptor transpose.  Reference: normalize_keypoints superglue.py:65-72,
// KeypointEncoder multi_view_matcher.py:24-37.  The 128->256->256 tail runs as GEMMs.
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int PTS = 32;  // points per block

template <int CIN, int COUT>
__device__ __forceinline__ void dense_relu(const float* __restrict__ in, float* __restrict__ out,
                                           const float* __restrict__ w, const float* __restrict__ b) {
  // in [PTS][CIN], out [PTS][COUT]; threads stride over (point, channel) outputs.
  for (int o = threadIdx.x; o < PTS * COUT; o += blockDim.x) {
    const int p = o / COUT, c = o % COUT;
    float acc = __ldg(b + c);
    const float* wr = w + c * CIN;
    const float* ip = in + p * CIN;
#pragma unroll 8
    for (int k = 0; k < CIN; ++k) acc = fmaf(__ldg(wr + k), ip[k], acc);
    out[o] = fmaxf(acc, 0.f);
  }
}

__global__ void __launch_bounds__(256) kenc_front_kernel(const float* __restrict__ kpts,
                                                         const float* __restrict__ kscores,
                                                         const float* w0, const float* b0,
                                                         const float* w1, const float* b1,
                                                         const float* w2, const float* b2,
                                                         float* __restrict__ h3, int n_points,
                                                         float cx, float cy, float scale) {
  __shared__ float s_in[PTS * 3];
  __shared__ float s_h1[PTS * 32];
  __shared__ float s_h2[PTS * 64];
  const int p0 = blockIdx.x * PTS;
  for (int i = threadIdx.x; i < PTS; i += blockDim.x) {
    const int p = p0 + i;
    float x = 0.f, y = 0.f, s = 0.f;
    if (p < n_points) {
      x = (kpts[2 * p] - cx) / scale;
      y = (kpts[2 * p + 1] - cy) / scale;
      s = kscores[p];
    }
    s_in[i * 3 + 0] = x; s_in[i * 3 + 1] = y; s_in[i * 3 + 2] = s;
  }
  __syncthreads();
  dense_relu<3, 32>(s_in, s_h1, w0, b0);
  __syncthreads();
  dense_relu<32, 64>(s_h1, s_h2, w1, b1);
  __syncthreads();
  // last front layer straight to global (coalesced over channels)
  for (int o = threadIdx.x; o < PTS * 128; o += blockDim.x) {
    const int p = o / 128, c = o % 128;
    if (p0 + p >= n_points) continue;
    float acc = __ldg(b2 + c);
    const float* wr = w2 + c * 64;
    const float* ip = s_h2 + p * 64;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) acc = fmaf(__ldg(wr + k), ip[k], acc);
    h3[(long long)(p0 + p) * 128 + c] = fmaxf(acc, 0.f);
  }
}

// [V, C, n_pad] -> [V, n_pad, C]
__global__ void transpose_cn_kernel(const float* __restrict__ in, float* __restrict__ out, int C,
                                    int n_pad) {
  __shared__ float tile[32][33];
  const int v = blockIdx.z;
  const float* src = in + (long long)v * C * n_pad;
  float* dst = out + (long long)v * C * n_pad;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    tile[r][threadIdx.x] = src[(long long)(c0 + r) * n_pad + n0 + threadIdx.x];
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    dst[(long long)(n0 + r) * C + c0 + threadIdx.x] = tile[threadIdx.x][r];
}

}  // namespace

int launch_kenc_front(const float* kpts, const float* kscores, const float* const* w,
                      const float* const* b, float* h3, int n_points, float img_w, float img_h,
                      cudaStream_t stream) {
  MvmProfScope prof__(MVM_TAG_KENC, stream);
  const float scale = 0.7f * fmaxf(img_w, img_h);
  kenc_front_kernel<<<mvm_div_up(n_points, PTS), 256, 0, stream>>>(
      kpts, kscores, w[0], b[0], w[1], b[1], w[2], b[2], h3, n_points, img_w * 0.5f, img_h * 0.5f,
      scale);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}

int launch_transpose_cn(const float* in, float* out, int n_views_total, int C, int n_pad,
                        cudaStream_t stream) {
  MVM_REQUIRE(C % 32 == 0 && n_pad % 32 == 0);
  dim3 grid(n_pad / 32, C / 32, n_views_total), block(32, 8);
  transpose_cn_kernel<<<grid, block, 0, stream>>>(in, out, C, n_pad);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}
