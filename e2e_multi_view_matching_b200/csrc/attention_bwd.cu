// Backward of the multi-head attention with multi-view key/value segments (what autograd computes for
// attention() / MultiHeadedAttention, superglue.py:87-109, with the cross source = concatenation of the other views,
// multi_view_matcher.py:65-86).  Flash-style: the probability tensor is recomputed tile by tile, never stored.
//   S = Q K^T / 8,  P = softmax_j(S),  O = P V
//   D_i = sum_d dO_id O_id,  dP = dO V^T,  dS = P o (dP - D),  dQ = dS K / 8,  dK = dS^T Q / 8,  dV = P^T dO
// Two kernels, each in two variants: 64 x 64 x 64 tile products on the tensor cores (mma.sync TF32, three passes on split
// operands, default) or register-tiled on the fp32 CUDA cores (cross-check, mvm_debug_set_attention_backward_variant):
//   attn_bwd_dq_kernel   CTA = (64 queries, head, view slot): sweep 1 over the key tiles -> row log-sum-exp L_i and D_i
//                        (kept for the second kernel), sweep 2 -> dQ
//   attn_bwd_dkv_kernel  CTA = (64 keys, head, view slot): loops over every query tile that attends to these keys -> dK, dV
// Deterministic (no atomics).  Rows beyond a view's keypoint count are masked on both sides and get zero gradients.
#include "../../include/mvm_b200.h"
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int BT = 64, LD = 68, QKV_LD = 768, O_LD = 256;
constexpr int TILE = BT * LD;
constexpr float SCALE = 0.125f;   // 1 / sqrt(64)

struct BwdArgs {
  const float* qkv; const float* out; const float* dout; float* dqkv; float* lse; float* dsum;
  int n_pad; int is_cross; AttnSegs segs;
};

// 64 x 64 tile of a row-major matrix (row stride ld) -> smem [64][LD]; rows >= nvalid are zero-filled
__device__ __forceinline__ void load_tile(float* s, const float* g, long long ld, int nvalid, int tid) {
  for (int i = tid; i < 64 * 16; i += 256) {
    const int r = i >> 4, c4 = (i & 15) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nvalid) x = *reinterpret_cast<const float4*>(g + (long long)r * ld + c4);
    *reinterpret_cast<float4*>(s + r * LD + c4) = x;
  }
}

// acc[rr][cc] += sum_d A[ty*4+rr][d] * B[tx+16*cc][d]
__device__ __forceinline__ void mm_nt(float (&acc)[4][4], const float* A, const float* B, int ty, int tx) {
#pragma unroll 2
  for (int d = 0; d < 64; d += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(A + (ty * 4 + i) * LD + d);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = *reinterpret_cast<const float4*>(B + (tx + 16 * i) * LD + d);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
}

// acc[rr][dd] += sum_j A[ty*4+rr][j] * B[j][tx*4+dd]
__device__ __forceinline__ void mm_nn(float (&acc)[4][4], const float* A, const float* B, int ty, int tx) {
#pragma unroll 2
  for (int j = 0; j < 64; j += 4) {
    float a[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = *reinterpret_cast<const float4*>(A + (ty * 4 + i) * LD + j);
      a[i][0] = t.x; a[i][1] = t.y; a[i][2] = t.z; a[i][3] = t.w;
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float4 b = *reinterpret_cast<const float4*>(B + (j + jj) * LD + tx * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(a[i][jj], b.x, acc[i][0]);
        acc[i][1] = fmaf(a[i][jj], b.y, acc[i][1]);
        acc[i][2] = fmaf(a[i][jj], b.z, acc[i][2]);
        acc[i][3] = fmaf(a[i][jj], b.w, acc[i][3]);
      }
    }
  }
}

// acc[cc][dd] += sum_r A[r][ty*4+cc] * B[r][tx*4+dd]
__device__ __forceinline__ void mm_tn(float (&acc)[4][4], const float* A, const float* B, int ty, int tx) {
#pragma unroll 4
  for (int r = 0; r < 64; ++r) {
    const float4 a = *reinterpret_cast<const float4*>(A + r * LD + ty * 4);
    const float4 b = *reinterpret_cast<const float4*>(B + r * LD + tx * 4);
    acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
    acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
    acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
    acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
    acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]);
    acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
    acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]);
    acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
  }
}

__device__ __forceinline__ void zero16(float (&a)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// does query view t attend to key view s?
__device__ __forceinline__ bool attends(int is_cross, int t, int s) { return is_cross ? (s != t) : (s == t); }

__global__ void __launch_bounds__(256) attn_bwd_dq_kernel(const __grid_constant__ BwdArgs g) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;
  float* dOs = Qs + TILE;
  float* Ks = dOs + TILE;
  float* Vs = Ks + TILE;
  float* Ss = Vs + TILE;
  __shared__ float Lsm[64], Dsm[64];
  const int q0 = blockIdx.x * BT, h = blockIdx.y, v = blockIdx.z;
  const int T = g.segs.n_views, t = v % T, b = v / T, n_pad = g.n_pad;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int cnt_q = g.segs.counts[t];
  float* dq_out = g.dqkv + ((long long)v * n_pad + q0) * QKV_LD + h * 64;
  float* lse_out = g.lse + ((long long)v * 4 + h) * n_pad + q0;
  float* dsum_out = g.dsum + ((long long)v * 4 + h) * n_pad + q0;
  if (q0 >= cnt_q) {          // padding tile: zero gradient
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(dq_out + (long long)(ty * 4 + i) * QKV_LD + tx * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 64) { lse_out[tid] = 0.f; dsum_out[tid] = 0.f; }
    return;
  }
  const int nvq = min(64, cnt_q - q0);
  load_tile(Qs, g.qkv + ((long long)v * n_pad + q0) * QKV_LD + h * 64, QKV_LD, nvq, tid);
  load_tile(dOs, g.dout + ((long long)v * n_pad + q0) * O_LD + h * 64, O_LD, nvq, tid);
  load_tile(Ks, g.out + ((long long)v * n_pad + q0) * O_LD + h * 64, O_LD, nvq, tid);
  __syncthreads();
  {   // D_i = dO_i . O_i : four threads per row
    const int r = tid >> 2, p = tid & 3;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) s = fmaf(dOs[r * LD + p * 16 + d], Ks[r * LD + p * 16 + d], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (p == 0) Dsm[r] = s;
  }
  // ---- sweep 1: row log-sum-exp of the scaled scores
  float m_run[4], l_run[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m_run[i] = -INFINITY; l_run[i] = 0.f; }
  for (int s = 0; s < T; ++s) {
    if (!attends(g.is_cross, t, s)) continue;
    const int cnt = g.segs.counts[s];
    const long long vs = (long long)b * T + s;
    for (int k0 = 0; k0 < cnt; k0 += BT) {
      __syncthreads();
      load_tile(Ks, g.qkv + (vs * n_pad + k0) * QKV_LD + 256 + h * 64, QKV_LD, min(64, cnt - k0), tid);
      __syncthreads();
      float acc[4][4];
      zero16(acc);
      mm_nt(acc, Qs, Ks, ty, tx);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k0 + tx + 16 * j >= cnt) continue;
          const float sc = acc[i][j] * SCALE;
          const float mn = fmaxf(m_run[i], sc);
          l_run[i] = l_run[i] * __expf(m_run[i] - mn) + __expf(sc - mn);
          m_run[i] = mn;
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const float mo = __shfl_xor_sync(0xffffffffu, m_run[i], o);
      const float lo = __shfl_xor_sync(0xffffffffu, l_run[i], o);
      const float mn = fmaxf(m_run[i], mo);
      if (mn == -INFINITY) { l_run[i] = 0.f; continue; }
      l_run[i] = l_run[i] * __expf(m_run[i] - mn) + lo * __expf(mo - mn);
      m_run[i] = mn;
    }
    if (tx == 0) Lsm[ty * 4 + i] = l_run[i] > 0.f ? m_run[i] + logf(l_run[i]) : INFINITY;   // no keys: P = 0
  }
  __syncthreads();
  if (tid < 64) { lse_out[tid] = Lsm[tid]; dsum_out[tid] = Dsm[tid]; }
  // ---- sweep 2: dQ
  float dq[4][4];
  zero16(dq);
  for (int s = 0; s < T; ++s) {
    if (!attends(g.is_cross, t, s)) continue;
    const int cnt = g.segs.counts[s];
    const long long vs = (long long)b * T + s;
    for (int k0 = 0; k0 < cnt; k0 += BT) {
      __syncthreads();
      load_tile(Ks, g.qkv + (vs * n_pad + k0) * QKV_LD + 256 + h * 64, QKV_LD, min(64, cnt - k0), tid);
      load_tile(Vs, g.qkv + (vs * n_pad + k0) * QKV_LD + 512 + h * 64, QKV_LD, min(64, cnt - k0), tid);
      __syncthreads();
      float sc[4][4], dp[4][4];
      zero16(sc);
      zero16(dp);
      mm_nt(sc, Qs, Ks, ty, tx);
      mm_nt(dp, dOs, Vs, ty, tx);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = ty * 4 + i;
        const float L = Lsm[row], D = Dsm[row];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = tx + 16 * j;
          float ds = 0.f;
          if (k0 + col < cnt && row < nvq) ds = expf(sc[i][j] * SCALE - L) * (dp[i][j] - D) * SCALE;
          Ss[row * LD + col] = ds;
        }
      }
      __syncthreads();
      mm_nn(dq, Ss, Ks, ty, tx);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(dq_out + (long long)(ty * 4 + i) * QKV_LD + tx * 4) =
        make_float4(dq[i][0], dq[i][1], dq[i][2], dq[i][3]);
}

__global__ void __launch_bounds__(256) attn_bwd_dkv_kernel(const __grid_constant__ BwdArgs g) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = Ks + TILE;
  float* Qs = Vs + TILE;
  float* dOs = Qs + TILE;
  float* Ps = dOs + TILE;
  float* dSs = Ps + TILE;
  __shared__ float Lsm[64], Dsm[64];
  const int k0 = blockIdx.x * BT, h = blockIdx.y, v = blockIdx.z;
  const int T = g.segs.n_views, t = v % T, b = v / T, n_pad = g.n_pad;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int cnt_k = g.segs.counts[t];
  float* dk_out = g.dqkv + ((long long)v * n_pad + k0) * QKV_LD + 256 + h * 64;
  float* dv_out = dk_out + 256;
  float dk[4][4], dv[4][4];
  zero16(dk);
  zero16(dv);
  if (k0 < cnt_k) {
    const int nvk = min(64, cnt_k - k0);
    load_tile(Ks, g.qkv + ((long long)v * n_pad + k0) * QKV_LD + 256 + h * 64, QKV_LD, nvk, tid);
    load_tile(Vs, g.qkv + ((long long)v * n_pad + k0) * QKV_LD + 512 + h * 64, QKV_LD, nvk, tid);
    for (int s = 0; s < T; ++s) {
      if (!attends(g.is_cross, s, t)) continue;     // query view s attends to key view t
      const int cnt_q = g.segs.counts[s];
      const long long vq = (long long)b * T + s;
      for (int q0 = 0; q0 < cnt_q; q0 += BT) {
        const int nvq = min(64, cnt_q - q0);
        __syncthreads();
        load_tile(Qs, g.qkv + (vq * n_pad + q0) * QKV_LD + h * 64, QKV_LD, nvq, tid);
        load_tile(dOs, g.dout + (vq * n_pad + q0) * O_LD + h * 64, O_LD, nvq, tid);
        if (tid < 64) {
          Lsm[tid] = g.lse[(vq * 4 + h) * n_pad + q0 + tid];
          Dsm[tid] = g.dsum[(vq * 4 + h) * n_pad + q0 + tid];
        }
        __syncthreads();
        float sc[4][4], dp[4][4];
        zero16(sc);
        zero16(dp);
        mm_nt(sc, Qs, Ks, ty, tx);
        mm_nt(dp, dOs, Vs, ty, tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = ty * 4 + i;
          const float L = Lsm[row], D = Dsm[row];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = tx + 16 * j;
            float p = 0.f;
            if (col < nvk && row < nvq) p = expf(sc[i][j] * SCALE - L);
            Ps[row * LD + col] = p;
            dSs[row * LD + col] = p * (dp[i][j] - D) * SCALE;
          }
        }
        __syncthreads();
        mm_tn(dv, Ps, dOs, ty, tx);
        mm_tn(dk, dSs, Qs, ty, tx);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    *reinterpret_cast<float4*>(dk_out + (long long)(ty * 4 + i) * QKV_LD + tx * 4) = make_float4(dk[i][0], dk[i][1], dk[i][2], dk[i][3]);
    *reinterpret_cast<float4*>(dv_out + (long long)(ty * 4 + i) * QKV_LD + tx * 4) = make_float4(dv[i][0], dv[i][1], dv[i][2], dv[i][3]);
  }
}

// ======================================================================================================================
// Tensor-core variant (default): the same two kernels with the 64 x 64 x 64 products on mma.sync m16n8k8 TF32, every
// product as three MMAs on split operands (x = hi + lo, hi = x truncated to tf32, lo = x - hi; lo.lo dropped) with fp32
// accumulation -- fp32-faithful like the 3xTF32 GEMMs.  8 warps: warp w owns rows (w & 3) * 16 .. + 16 and columns
// (w >> 2) * 32 .. + 32 of a tile product; a thread holds rows gid, gid + 8 and columns 2 tig, 2 tig + 1 of each of its
// four 8-column blocks (gid = lane / 4, tig = lane % 4).
// ======================================================================================================================
constexpr int LDM = 72;             // row pitch of the tensor-core variant's tiles (conflict-free for every fragment pattern below)
constexpr int TILE_M = BT * LDM;

__device__ __forceinline__ void load_tile_m(float* s, const float* g, long long ld, int nvalid, int tid) {
  for (int i = tid; i < 64 * 16; i += 256) {
    const int r = i >> 4, c4 = (i & 15) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nvalid) x = *reinterpret_cast<const float4*>(g + (long long)r * ld + c4);
    *reinterpret_cast<float4*>(s + r * LDM + c4) = x;
  }
}

// x = hi + lo with hi = x truncated to tf32 (one logic op) and lo = x - hi (exact; the tensor core reads its upper 19
// bits).  The ncu capture of the first version (cvt.rna on both parts: 3 conversions / subtractions per fragment element
// in the inner loop) had the ALU pipe as its top pipe at 51 %, the tensor pipe at 36-42 %.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc[nt][.] += A(16 rows m0.. x 64) * B(64 x 8 columns n0 + 8 nt ..), nt = 0..3.
//   A_T == false: A stored [m][k] (row-major),  A_T == true: A stored [k][m]
//   B_T == true:  B stored [n][k] ("NT" product, both operands contraction-contiguous),  B_T == false: B stored [k][n]
// NT products relabel the contraction index inside a k-step (fragment slots k = tig / tig + 4 read the memory columns
// 2 tig / 2 tig + 1 of BOTH operands), so that every fragment pair is one 64-bit shared-memory load.
template <bool A_T, bool B_T>
__device__ __forceinline__ void warp_mma(float (&acc)[4][4], const float* __restrict__ A, const float* __restrict__ B,
                                         int m0, int n0, int gid, int tig) {
  constexpr bool PAIR = !A_T && B_T;
#pragma unroll 2
  for (int k0 = 0; k0 < 64; k0 += 8) {
    float af[4];
    if (PAIR) {
      const float2 x = *reinterpret_cast<const float2*>(A + (m0 + gid) * LDM + k0 + 2 * tig);
      const float2 y = *reinterpret_cast<const float2*>(A + (m0 + gid + 8) * LDM + k0 + 2 * tig);
      af[0] = x.x; af[2] = x.y; af[1] = y.x; af[3] = y.y;
    } else if (!A_T) {
      af[0] = A[(m0 + gid) * LDM + k0 + tig];      af[1] = A[(m0 + gid + 8) * LDM + k0 + tig];
      af[2] = A[(m0 + gid) * LDM + k0 + tig + 4];  af[3] = A[(m0 + gid + 8) * LDM + k0 + tig + 4];
    } else {
      af[0] = A[(k0 + tig) * LDM + m0 + gid];      af[1] = A[(k0 + tig) * LDM + m0 + gid + 8];
      af[2] = A[(k0 + tig + 4) * LDM + m0 + gid];  af[3] = A[(k0 + tig + 4) * LDM + m0 + gid + 8];
    }
    uint32_t ahi[4], alo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(af[i], ahi[i], alo[i]);
    uint32_t bhi[4][2], blo[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int n = n0 + nt * 8 + gid;
      float bf[2];
      if (PAIR) { const float2 x = *reinterpret_cast<const float2*>(B + n * LDM + k0 + 2 * tig); bf[0] = x.x; bf[1] = x.y; }
      else if (B_T) { bf[0] = B[n * LDM + k0 + tig]; bf[1] = B[n * LDM + k0 + tig + 4]; }
      else     { bf[0] = B[(k0 + tig) * LDM + n]; bf[1] = B[(k0 + tig + 4) * LDM + n]; }
      split_tf32(bf[0], bhi[nt][0], blo[nt][0]);
      split_tf32(bf[1], bhi[nt][1], blo[nt][1]);
    }
    // the three passes of a product go to the same accumulator: issue them pass-major, so that four independent MMAs sit
    // between two dependent ones (the ncu capture of the nt-major order had `wait` -- fixed-latency dependencies -- as its
    // top stall reason)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], alo, bhi[nt]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], ahi, blo[nt]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_tf32(acc[nt], ahi, bhi[nt]);
  }
}

__global__ void __launch_bounds__(256) attn_bwd_dq_mma_kernel(const __grid_constant__ BwdArgs g) {
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;
  float* dOs = Qs + TILE_M;
  float* Ks = dOs + TILE_M;
  float* Vs = Ks + TILE_M;
  float* Ss = Vs + TILE_M;
  __shared__ float Lsm[64], Dsm[64], Pm[2][64], Pl[2][64];
  const int q0 = blockIdx.x * BT, h = blockIdx.y, v = blockIdx.z;
  const int T = g.segs.n_views, t = v % T, b = v / T, n_pad = g.n_pad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
  const int cnt_q = g.segs.counts[t];
  float* dq_out = g.dqkv + ((long long)v * n_pad + q0) * QKV_LD + h * 64;
  float* lse_out = g.lse + ((long long)v * 4 + h) * n_pad + q0;
  float* dsum_out = g.dsum + ((long long)v * 4 + h) * n_pad + q0;
  if (q0 >= cnt_q) {          // padding tile: zero gradient
    const int tx = tid & 15, ty = tid >> 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(dq_out + (long long)(ty * 4 + i) * QKV_LD + tx * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 64) { lse_out[tid] = 0.f; dsum_out[tid] = 0.f; }
    return;
  }
  const int nvq = min(64, cnt_q - q0);
  load_tile_m(Qs, g.qkv + ((long long)v * n_pad + q0) * QKV_LD + h * 64, QKV_LD, nvq, tid);
  load_tile_m(dOs, g.dout + ((long long)v * n_pad + q0) * O_LD + h * 64, O_LD, nvq, tid);
  load_tile_m(Ks, g.out + ((long long)v * n_pad + q0) * O_LD + h * 64, O_LD, nvq, tid);
  __syncthreads();
  {   // D_i = dO_i . O_i : four threads per row
    const int r = tid >> 2, p = tid & 3;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 16; ++d) s = fmaf(dOs[r * LDM + p * 16 + d], Ks[r * LDM + p * 16 + d], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (p == 0) Dsm[r] = s;
  }
  // ---- sweep 1: row log-sum-exp (rows m0 + gid and m0 + gid + 8; this thread's 8 columns of every key tile)
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  for (int s = 0; s < T; ++s) {
    if (!attends(g.is_cross, t, s)) continue;
    const int cnt = g.segs.counts[s];
    const long long vs = (long long)b * T + s;
    for (int k0 = 0; k0 < cnt; k0 += BT) {
      __syncthreads();
      load_tile_m(Ks, g.qkv + (vs * n_pad + k0) * QKV_LD + 256 + h * 64, QKV_LD, min(64, cnt - k0), tid);
      __syncthreads();
      float acc[4][4];
      zero16(acc);
      warp_mma<false, true>(acc, Qs, Ks, m0, n0, gid, tig);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = k0 + n0 + nt * 8 + 2 * tig + (e & 1), r = e >> 1;
          if (col >= cnt) continue;
          const float sc = acc[nt][e] * SCALE;
          const float mn = fmaxf(m_run[r], sc);
          l_run[r] = l_run[r] * __expf(m_run[r] - mn) + __expf(sc - mn);
          m_run[r] = mn;
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {       // the four lanes of a quad hold the same rows
      const float mo = __shfl_xor_sync(0xffffffffu, m_run[r], o);
      const float lo = __shfl_xor_sync(0xffffffffu, l_run[r], o);
      const float mn = fmaxf(m_run[r], mo);
      if (mn == -INFINITY) { l_run[r] = 0.f; continue; }
      l_run[r] = l_run[r] * __expf(m_run[r] - mn) + lo * __expf(mo - mn);
      m_run[r] = mn;
    }
    if (tig == 0) { Pm[warp >> 2][m0 + gid + 8 * r] = m_run[r]; Pl[warp >> 2][m0 + gid + 8 * r] = l_run[r]; }
  }
  __syncthreads();
  if (tid < 64) {                            // merge the two column halves of the row
    const float ma = Pm[0][tid], mb = Pm[1][tid], mn = fmaxf(ma, mb);
    float L = INFINITY;                      // no keys: P = 0
    if (mn != -INFINITY) {
      const float l = Pl[0][tid] * __expf(ma - mn) + Pl[1][tid] * __expf(mb - mn);
      if (l > 0.f) L = mn + logf(l);
    }
    Lsm[tid] = L;
    lse_out[tid] = L;
    dsum_out[tid] = Dsm[tid];
  }
  __syncthreads();
  // ---- sweep 2: dQ
  float dq[4][4];
  zero16(dq);
  for (int s = 0; s < T; ++s) {
    if (!attends(g.is_cross, t, s)) continue;
    const int cnt = g.segs.counts[s];
    const long long vs = (long long)b * T + s;
    for (int k0 = 0; k0 < cnt; k0 += BT) {
      __syncthreads();
      load_tile_m(Ks, g.qkv + (vs * n_pad + k0) * QKV_LD + 256 + h * 64, QKV_LD, min(64, cnt - k0), tid);
      load_tile_m(Vs, g.qkv + (vs * n_pad + k0) * QKV_LD + 512 + h * 64, QKV_LD, min(64, cnt - k0), tid);
      __syncthreads();
      float sc[4][4], dp[4][4];
      zero16(sc);
      zero16(dp);
      warp_mma<false, true>(sc, Qs, Ks, m0, n0, gid, tig);
      warp_mma<false, true>(dp, dOs, Vs, m0, n0, gid, tig);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = m0 + gid + 8 * r;
        const float L = Lsm[row], D = Dsm[row];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int col = n0 + nt * 8 + 2 * tig;
          float2 ds = make_float2(0.f, 0.f);
          if (row < nvq) {
            if (k0 + col < cnt) ds.x = expf(sc[nt][2 * r] * SCALE - L) * (dp[nt][2 * r] - D) * SCALE;
            if (k0 + col + 1 < cnt) ds.y = expf(sc[nt][2 * r + 1] * SCALE - L) * (dp[nt][2 * r + 1] - D) * SCALE;
          }
          *reinterpret_cast<float2*>(Ss + row * LDM + col) = ds;
        }
      }
      __syncthreads();
      warp_mma<false, false>(dq, Ss, Ks, m0, n0, gid, tig);
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<float2*>(dq_out + (long long)(m0 + gid + 8 * r) * QKV_LD + n0 + nt * 8 + 2 * tig) =
          make_float2(dq[nt][2 * r], dq[nt][2 * r + 1]);
}

__global__ void __launch_bounds__(256) attn_bwd_dkv_mma_kernel(const __grid_constant__ BwdArgs g) {
  extern __shared__ __align__(16) float smem[];
  float* Ks = smem;
  float* Vs = Ks + TILE_M;
  float* Qs = Vs + TILE_M;
  float* dOs = Qs + TILE_M;
  float* Ps = dOs + TILE_M;
  float* dSs = Ps + TILE_M;
  __shared__ float Lsm[64], Dsm[64];
  const int k0 = blockIdx.x * BT, h = blockIdx.y, v = blockIdx.z;
  const int T = g.segs.n_views, t = v % T, b = v / T, n_pad = g.n_pad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gid = lane >> 2, tig = lane & 3;
  const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 32;
  const int cnt_k = g.segs.counts[t];
  float* dk_out = g.dqkv + ((long long)v * n_pad + k0) * QKV_LD + 256 + h * 64;
  float* dv_out = dk_out + 256;
  float dk[4][4], dv[4][4];
  zero16(dk);
  zero16(dv);
  if (k0 < cnt_k) {
    const int nvk = min(64, cnt_k - k0);
    load_tile_m(Ks, g.qkv + ((long long)v * n_pad + k0) * QKV_LD + 256 + h * 64, QKV_LD, nvk, tid);
    load_tile_m(Vs, g.qkv + ((long long)v * n_pad + k0) * QKV_LD + 512 + h * 64, QKV_LD, nvk, tid);
    for (int s = 0; s < T; ++s) {
      if (!attends(g.is_cross, s, t)) continue;     // query view s attends to key view t
      const int cnt_q = g.segs.counts[s];
      const long long vq = (long long)b * T + s;
      for (int q0 = 0; q0 < cnt_q; q0 += BT) {
        const int nvq = min(64, cnt_q - q0);
        __syncthreads();
        load_tile_m(Qs, g.qkv + (vq * n_pad + q0) * QKV_LD + h * 64, QKV_LD, nvq, tid);
        load_tile_m(dOs, g.dout + (vq * n_pad + q0) * O_LD + h * 64, O_LD, nvq, tid);
        if (tid < 64) {
          Lsm[tid] = g.lse[(vq * 4 + h) * n_pad + q0 + tid];
          Dsm[tid] = g.dsum[(vq * 4 + h) * n_pad + q0 + tid];
        }
        __syncthreads();
        float sc[4][4], dp[4][4];       // [query row][key column]
        zero16(sc);
        zero16(dp);
        warp_mma<false, true>(sc, Qs, Ks, m0, n0, gid, tig);
        warp_mma<false, true>(dp, dOs, Vs, m0, n0, gid, tig);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int row = m0 + gid + 8 * r;
          const float L = Lsm[row], D = Dsm[row];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const int col = n0 + nt * 8 + 2 * tig;
            float2 p = make_float2(0.f, 0.f);
            if (row < nvq) {
              if (col < nvk) p.x = expf(sc[nt][2 * r] * SCALE - L);
              if (col + 1 < nvk) p.y = expf(sc[nt][2 * r + 1] * SCALE - L);
            }
            *reinterpret_cast<float2*>(Ps + row * LDM + col) = p;
            *reinterpret_cast<float2*>(dSs + row * LDM + col) =
                make_float2(p.x * (dp[nt][2 * r] - D) * SCALE, p.y * (dp[nt][2 * r + 1] - D) * SCALE);
          }
        }
        __syncthreads();
        warp_mma<true, false>(dv, Ps, dOs, m0, n0, gid, tig);      // dV[key][d] += sum_q P[q][key] dO[q][d]
        warp_mma<true, false>(dk, dSs, Qs, m0, n0, gid, tig);      // dK[key][d] += sum_q dS[q][key] Q[q][d]
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const long long o = (long long)(m0 + gid + 8 * r) * QKV_LD + n0 + nt * 8 + 2 * tig;
      *reinterpret_cast<float2*>(dk_out + o) = make_float2(dk[nt][2 * r], dk[nt][2 * r + 1]);
      *reinterpret_cast<float2*>(dv_out + o) = make_float2(dv[nt][2 * r], dv[nt][2 * r + 1]);
    }
}

int g_attn_bwd_variant = 1;    // 1 = mma.sync TF32x3 (default), 0 = fp32 CUDA cores (cross-check)

constexpr int SMEM_DQ = 5 * TILE * 4, SMEM_DKV = 6 * TILE * 4;
constexpr int SMEM_DQ_M = 5 * TILE_M * 4, SMEM_DKV_M = 6 * TILE_M * 4;

}  // namespace

extern "C" int mvm_attention_backward(const float* qkv, const float* out, const float* dout, float* dqkv, float* ws,
                                      int batch, int n_views, int n_pad, const int* counts, int is_cross, void* stream) {
  MVM_REQUIRE(qkv && out && dout && dqkv && ws && counts && batch >= 1 && n_views >= 1 && n_views <= 8);
  MVM_REQUIRE(n_pad >= 64 && n_pad % 64 == 0);
  BwdArgs g;
  g.qkv = qkv; g.out = out; g.dout = dout; g.dqkv = dqkv; g.n_pad = n_pad; g.is_cross = is_cross;
  const long long V = (long long)batch * n_views;
  g.lse = ws; g.dsum = ws + V * 4 * n_pad;
  g.segs.n_views = n_views;
  for (int t = 0; t < 8; ++t) {
    g.segs.counts[t] = t < n_views ? counts[t] : 0;
    MVM_REQUIRE(g.segs.counts[t] >= 0 && g.segs.counts[t] <= n_pad);
  }
  mvm_once_per_device(MVM_ONCE_ATTN_BWD, [&] {
    cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ);
    cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV);
    cudaFuncSetAttribute(attn_bwd_dq_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ_M);
    cudaFuncSetAttribute(attn_bwd_dkv_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV_M);
  });
  cudaStream_t s = (cudaStream_t)stream;
  MvmProfScope prof__(MVM_TAG_ATTN, s);
  const dim3 grid(n_pad / 64, 4, (unsigned)V);
  if (g_attn_bwd_variant == 1) attn_bwd_dq_mma_kernel<<<grid, 256, SMEM_DQ_M, s>>>(g);
  else attn_bwd_dq_kernel<<<grid, 256, SMEM_DQ, s>>>(g);
  MVM_CHECK_LAUNCH();
  if (g_attn_bwd_variant == 1) attn_bwd_dkv_mma_kernel<<<grid, 256, SMEM_DKV_M, s>>>(g);
  else attn_bwd_dkv_kernel<<<grid, 256, SMEM_DKV, s>>>(g);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}

extern "C" int mvm_debug_set_attention_backward_variant(int variant) {
  MVM_REQUIRE(variant == 0 || variant == 1);
  g_attn_bwd_variant = variant;
  return MVM_OK;
}
