// Persistent tcgen05 GEMM for the 3xTF32 (fp32-faithful) production path of the matcher's 1x1 convolutions
// (superglue.py:51-62,101-121; multi_view_matcher.py:8-53):
//   C[M,N] = act(alpha * [A | A2][M,K] . W[N,K]^T + bias[N]) + R[M,N]         (fp32, both operands K-major)
//   D += A_hi.W_hi + A_hi.W_lo + A_lo.W_hi   with hi = rn_tf32(x), lo = rn_tf32(x - hi)
//
// Why this shape (measured on the one-tile-per-CTA kernel of gemm_tc.cu, profiles/r01_v16_*): with every
// operand plane in shared memory a 128x256x32 k-block moves 272 KB through SMEM (TMA writes, the splitter's
// read-modify-write of A, 3 passes x (A + W) operand reads) against 1536 tensor-pipe cycles -- 177 B/clk on
// a 128 B/clk port: the GEMM was shared-memory-bandwidth bound at ~55 % tensor-pipe utilisation, and the
// epilogue (TMEM drain + global stores) was not overlapped with anything.  Here:
//   * the A operand lives in TENSOR MEMORY: splitter warps read the raw fp32 tile TMA landed in SMEM once,
//     split it in registers and tcgen05.st the hi / lo planes; the three UMMAs read A from TMEM, so SMEM
//     only carries the TMA writes, one A read and the W operand reads (112 KB per 128x128x32 k-block);
//   * one CTA per SM walks a static tile schedule (persistent), with TWO accumulators in TMEM: the epilogue
//     warps drain tile i while the tensor pipe already works on tile i+1;
//   * C leaves through shared-memory staging + TMA stores (full 128-byte lines) instead of 16-byte
//     row-strided stores.
//
// Warp roles (320 threads):
//   warp 0      TMA producer: per k-block A [128 x 32] (raw fp32) + W_hi, W_lo [128 x 32] into a 4-deep ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (M128 N128 K8, kind::tf32, A from TMEM)
//   warps 2-5   A splitters (thread r owns tile row r = TMEM lane r)
//   warps 6-9   epilogue: tcgen05.ld -> alpha/bias/ReLU/residual -> swizzled staging -> TMA store
//               (V^T / lo planes of the QKV projection as in gemm_tc.cu)
// TMEM (512 columns): accumulators [0,128) [128,256); A stages s at 256 + 64 s: hi [0,32) lo [32,64).
#include <cstring>
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128, BN = 128;
constexpr int NTHREADS = 320;
// Two operand arithmetics (template parameter F16 of the kernel):
//   tf32x3 (F16 = false): k-block 32, W planes fp32 (tf32 values), kind::tf32 K = 8 per MMA, 4 stages of 48 KB
//   fp16x3 (F16 = true):  k-block 64, W planes fp16 (pre-scaled by a power of two, packing.py), kind::f16 K = 16 per
//                         MMA: the same 22-bit operands at half the tensor-pipe time; 3 stages of 64 KB
// A W plane tile is 16 KB in both ([128 x 32] fp32 or [128 x 64] fp16, 128-byte rows), the raw fp32 A tile 16 / 32 KB.
template <bool F16> struct GCfg {
  static constexpr int BK = F16 ? 64 : 32;
  static constexpr int STAGES = F16 ? 3 : 4;
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int W_BYTES = 16384;
  static constexpr int STAGE_BYTES = A_BYTES + 2 * W_BYTES;
};
static_assert(GCfg<true>::STAGES * GCfg<true>::STAGE_BYTES == GCfg<false>::STAGES * GCfg<false>::STAGE_BYTES, "same ring size");
constexpr int STG_BYTES = 32 * 128;               // one staged [32 rows x 32 cols] block per epilogue warp
constexpr int OFF_STG = GCfg<false>::STAGES * GCfg<false>::STAGE_BYTES;     // 4 warps x 2 buffers
constexpr int OFF_BIAS = OFF_STG + 8 * STG_BYTES;
constexpr int OFF_BAR = OFF_BIAS + BN * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr uint32_t TM_ACC = 0, TM_A = 256;

struct PArgs {
  const float* bias;
  const float* R; int ldr;
  float* C; int ldc;
  int M, N, K, K1;
  float alpha;
  int relu;
  float* VT; int vt_col0; int n_pad;   // transposed output for columns >= vt_col0 (see gemm_tc.cu)
  float* KLO; float* VTLO;             // tf32 lo planes of the K / V^T attention operands
  // half-precision operand planes for attention_h3.cu (fp16x3): K hi / lo and V hi / lo, all [rows, 256] (V stays
  // key-major: the attention reads it as an MN-major B operand); when set they replace the fp32 K and V thirds of C,
  // KLO, VT and VTLO
  __half* KH16; __half* KL16; __half* VH16; __half* VL16;
  int tiles_m, tiles_n;
  // split-K (weight-gradient GEMMs: few output tiles, a very long contraction): K is cut into `ksplit` slices, slice s is
  // an extra group of row tiles writing its partial product to rows [s * tiles_m * BM, ...) of C (a slab buffer that a
  // small kernel sums afterwards, in fixed order).  1 = off.
  int ksplit;
};

// SCORE mode: one launch computes every (pair, tuple) score matrix  scores = mdesc_a . mdesc_b^T * alpha  into the
// inner [m, n] block of the [m+1, n+1] coupling buffers (multi_view_matcher.py:278-280).  A = descriptors of view
// a (raw fp32, split on chip), W = the tf32 planes of the descriptors of view b.
struct ScoreTab {
  int n_pairs, batch, n_views, n_pad;
  int a[MVM_MAX_PAIRS], b[MVM_MAX_PAIRS], m[MVM_MAX_PAIRS], n[MVM_MAX_PAIRS];
  float* scores[MVM_MAX_PAIRS];
};

// rn_tf32 of a finite value (ties away, == cvt.rna.tf32.f32) in two integer instructions
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__device__ __forceinline__ void split_pack_h(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {   // D = F32, A = B = F16, both K-major
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool SCORE, bool F16>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                       const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                       const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmKLO,
                       const __grid_constant__ CUtensorMap tmKH, const __grid_constant__ CUtensorMap tmKL,
                       const __grid_constant__ CUtensorMap tmVH, const __grid_constant__ CUtensorMap tmVL, PArgs g,
                       const __grid_constant__ ScoreTab st) {
  using G_ = GCfg<F16>;
  constexpr int BK = G_::BK, STAGES = G_::STAGES, A_BYTES = G_::A_BYTES, W_BYTES = G_::W_BYTES, STAGE_BYTES = G_::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full = bars;                  // [STAGES] TMA landed
  uint64_t* empty = bars + STAGES;        // [STAGES] MMAs that read the stage retired
  uint64_t* a_ready = bars + 2 * STAGES;  // [STAGES] A planes stored to tensor memory (128 arrivals)
  uint64_t* acc_full = bars + 3 * STAGES;       // [2]
  uint64_t* acc_empty = bars + 3 * STAGES + 2;  // [2] (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = SCORE ? g.K / BK : g.K / BK / g.ksplit;       // k-blocks per tile (per K slice)
  const int per_prob = g.tiles_m * g.tiles_n;
  const int n_tiles = SCORE ? per_prob * st.n_pairs * st.batch : per_prob * g.ksplit;
  // tile -> output tile origin (m0, n0) and the rows the A / W boxes start at
  auto decode = [&](int tile, int& m0, int& n0, int& a_row, int& w_row, int& prob) {
    if (!SCORE) {
      const int r = tile % per_prob;
      prob = tile / per_prob;                                   // K slice (0 without split-K)
      m0 = (r / g.tiles_n) * BM; n0 = (r % g.tiles_n) * BN; a_row = m0; w_row = n0;
    } else {
      prob = tile / per_prob;
      const int r = tile % per_prob;
      m0 = (r / g.tiles_n) * BM; n0 = (r % g.tiles_n) * BN;
      const int p = prob / st.batch, bi = prob % st.batch;
      a_row = (bi * st.n_views + st.a[p]) * st.n_pad + m0;
      w_row = (bi * st.n_views + st.b[p]) * st.n_pad + n0;
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
      tc::mbar_init(a_ready + s, 128);
    }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(acc_full + i, 1); tc::mbar_init(acc_empty + i, 128); }
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA); tc::prefetch_tmap(&tmA2); tc::prefetch_tmap(&tmWhi); tc::prefetch_tmap(&tmWlo);
    tc::prefetch_tmap(&tmC); tc::prefetch_tmap(&tmKLO);
    if (g.KH16) { tc::prefetch_tmap(&tmKH); tc::prefetch_tmap(&tmKL); tc::prefetch_tmap(&tmVH); tc::prefetch_tmap(&tmVL); }
  }
  if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int m0, n0, a_row, w_row, prob;
      decode(tile, m0, n0, a_row, w_row, prob);
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(full + s, STAGE_BYTES);
          uint8_t* sp = smem + s * STAGE_BYTES;
          const int k = kt * BK + (SCORE ? 0 : prob * nk * BK);
          if (k < g.K1) tc::tma_load_2d(sp, &tmA, full + s, k, a_row);
          else tc::tma_load_2d(sp, &tmA2, full + s, k - g.K1, a_row);
          if (F16) {   // second [128 x 32] fp32 box of the 64-wide k-block (K1 is a multiple of 64: same side of the concat)
            if (k < g.K1) tc::tma_load_2d(sp + 16384, &tmA, full + s, k + 32, a_row);
            else tc::tma_load_2d(sp + 16384, &tmA2, full + s, k + 32 - g.K1, a_row);
          }
          tc::tma_load_2d(sp + A_BYTES, &tmWhi, full + s, k, w_row);
          tc::tma_load_2d(sp + A_BYTES + W_BYTES, &tmWlo, full + s, k, w_row);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // the whole warp walks the loop (converged); one elected lane issues
    constexpr uint32_t idesc = F16 ? make_idesc_f16(BM, BN) : tc::make_idesc_tf32(BM, BN);
    uint32_t it = 0;
    int i = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
      const int buf = i & 1;
      tc::mbar_wait(acc_empty + buf, ((i >> 1) & 1) ^ 1);     // epilogue drained this accumulator
      tc::tc_fence_after();
      const uint32_t acc = tmem_base + TM_ACC + buf * BN;
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        tc::mbar_wait(full + s, ph);
        tc::mbar_wait(a_ready + s, ph);
        tc::tc_fence_after();
        const uint32_t whi = tc::smem_u32(smem + s * STAGE_BYTES + A_BYTES), wlo = whi + W_BYTES;
        const uint32_t a_hi = tmem_base + TM_A + s * 64, a_lo = a_hi + 32;
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {        // 4 x (K = 8 tf32 | K = 16 halves) = 32 bytes of a 128-byte row each
            const uint64_t dhi = tc::make_kmajor_sw128_desc(whi + kk * 32);
            if (F16) {
              umma_f16_ts(acc, a_hi + kk * 8, dhi, idesc, (kt | kk) != 0);
              umma_f16_ts(acc, a_hi + kk * 8, tc::make_kmajor_sw128_desc(wlo + kk * 32), idesc, 1);
              umma_f16_ts(acc, a_lo + kk * 8, dhi, idesc, 1);
            } else {
              tc::umma_tf32_ts(acc, a_hi + kk * 8, dhi, idesc, (kt | kk) != 0);
              tc::umma_tf32_ts(acc, a_hi + kk * 8, tc::make_kmajor_sw128_desc(wlo + kk * 32), idesc, 1);
              tc::umma_tf32_ts(acc, a_lo + kk * 8, dhi, idesc, 1);
            }
          }
          tc::umma_commit(empty + s);
          if (kt == nk - 1) tc::umma_commit(acc_full + buf);
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ================================ A splitters ================================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may touch
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(full + s, (it / STAGES) & 1);
        // row r of the 128B-swizzled tile: 16-byte chunk c sits at position c ^ (r & 7)
        const float4* a = reinterpret_cast<const float4*>(smem + s * STAGE_BYTES + row * 128);
        float hi[32], lo[32];
        if (F16) {
          // 64 fp32 of the row (two swizzled 128-byte rows, 16 KB apart) -> 32 + 32 packed half2 columns
          uint32_t* hp = reinterpret_cast<uint32_t*>(hi);
          uint32_t* lp = reinterpret_cast<uint32_t*>(lo);
#pragma unroll
          for (int b2 = 0; b2 < 2; ++b2) {
            const float4* ab = reinterpret_cast<const float4*>(smem + s * STAGE_BYTES + b2 * 16384 + row * 128);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 x = ab[c ^ (row & 7)];
              split_pack_h(x.x, x.y, hp[b2 * 16 + 2 * c], lp[b2 * 16 + 2 * c]);
              split_pack_h(x.z, x.w, hp[b2 * 16 + 2 * c + 1], lp[b2 * 16 + 2 * c + 1]);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 x = a[c ^ (row & 7)];
            hi[4 * c] = tf32_hi(x.x); hi[4 * c + 1] = tf32_hi(x.y); hi[4 * c + 2] = tf32_hi(x.z); hi[4 * c + 3] = tf32_hi(x.w);
            lo[4 * c] = tf32_hi(x.x - hi[4 * c]); lo[4 * c + 1] = tf32_hi(x.y - hi[4 * c + 1]);
            lo[4 * c + 2] = tf32_hi(x.z - hi[4 * c + 2]); lo[4 * c + 3] = tf32_hi(x.w - hi[4 * c + 3]);
          }
        }
        const uint32_t ta = tmem_base + TM_A + s * 64 + lane_addr;
        tc::tmem_st32(ta, hi);
        tc::tmem_st32(ta + 32, lo);
        tc::tmem_st_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(a_ready + s);
      }
    }
  } else {
    // ================================ epilogue ================================
    const int q = warp & 3;
    const int ew = warp - 6;                                  // staging buffers of this warp
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint8_t* stg = smem + OFF_STG + ew * 2 * STG_BYTES;
    const int et = threadIdx.x - 192;                         // 0..127
    int sbuf = 0;
    int i = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
      int m0, n0, a_row_, w_row_, prob;
      decode(tile, m0, n0, a_row_, w_row_, prob);
      const int buf = i & 1;
      const int m = m0 + row;
      // bias of this tile's columns, shared by the four epilogue warps
      asm volatile("bar.sync 1, 128;" ::: "memory");           // previous tile's readers are done
      s_bias[et] = (!SCORE && g.bias) ? __ldg(g.bias + n0 + et) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      tc::mbar_wait(acc_full + buf, (i >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + TM_ACC + buf * BN + lane_addr;
      if (SCORE) {
        // rows of the coupling buffer are n+1 floats long (not 16-byte aligned): no TMA store.  Each warp
        // transposes its 32 x 32 block through an XOR-swizzled staging tile so that a store instruction writes
        // 32 consecutive floats of one row.
        const int p = prob / st.batch, bi = prob % st.batch;
        const int pm = st.m[p], pn = st.n[p];
        float* Cp = st.scores[p] + (long long)bi * (pm + 1) * (pn + 1);
        float* sw = reinterpret_cast<float*>(stg);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          tc::tmem_ld32(taddr + c * 32, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) sw[lane * 32 + (j ^ lane)] = g.alpha * v[j];
          __syncwarp();
          const int gn = n0 + c * 32 + lane;
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int gm = m0 + q * 32 + rr;
            const float x = sw[rr * 32 + (lane ^ rr)];
            if (gm < pm && gn < pn) Cp[(long long)gm * (pn + 1) + gn] = x;
          }
          __syncwarp();
        }
        tc::tc_fence_before();
        tc::mbar_arrive(acc_empty + buf);
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int nb = n0 + c * 32;
        float r[32];
        const bool has_r = g.R != nullptr && m < g.M;
        if (has_r) {
          const float4* r4 = reinterpret_cast<const float4*>(g.R + (long long)m * g.ldr + nb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(r4 + j);
            r[4 * j] = t.x; r[4 * j + 1] = t.y; r[4 * j + 2] = t.z; r[4 * j + 3] = t.w;
          }
        }
        float v[32];
        tc::tmem_ld32(taddr + c * 32, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = g.alpha * v[j] + s_bias[c * 32 + j];
          if (g.relu) x = fmaxf(x, 0.f);
          if (has_r) x += r[j];
          v[j] = x;
        }
        if (g.VT && nb >= g.vt_col0) {
          // V^T (and its tf32 lo plane) for the tf32 attention kernel: lanes = consecutive keypoints -> coalesced
          if (m < g.M) {
            const int slab = m / g.n_pad, ii = m % g.n_pad;
            const long long off = ((long long)slab * (g.N - g.vt_col0) + (nb - g.vt_col0)) * g.n_pad + ii;
            if (g.VTLO) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float h = tf32_hi(v[j]);
                g.VT[off + (long long)j * g.n_pad] = h;
                g.VTLO[off + (long long)j * g.n_pad] = tf32_hi(v[j] - h);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) g.VT[off + (long long)j * g.n_pad] = v[j];
            }
          }
          continue;
        }
        if (g.KH16 && nb >= 256) {
          // K (columns 256..511) and V (512..767) hi / lo planes in half precision: the warp stages its 32 rows x 64 B
          // per plane (64-byte swizzle: conflict-free 16-byte stores) and one TMA store per plane writes full lines --
          // 16-byte row-strided global stores cost 32 L1 wavefronts per instruction and made this epilogue as long as
          // the tile's MMAs (QKV GEMM at 34 % tensor pipe, profiles/r02_v10_gemm_tc_persist_kernel_ncu.txt)
          uint32_t hp[16], lp[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
            hp[j] = *reinterpret_cast<const uint32_t*>(&h);
            lp[j] = *reinterpret_cast<const uint32_t*>(&l);
          }
          tc::tma_store_wait_read<1>();                        // the buffer used two stores ago is free
          __syncwarp();
          uint8_t* sb = stg + sbuf * STG_BYTES + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            *reinterpret_cast<uint4*>(sb + ((j ^ sw) << 4)) = make_uint4(hp[4 * j], hp[4 * j + 1], hp[4 * j + 2], hp[4 * j + 3]);
            *reinterpret_cast<uint4*>(sb + 2048 + ((j ^ sw) << 4)) = make_uint4(lp[4 * j], lp[4 * j + 1], lp[4 * j + 2], lp[4 * j + 3]);
          }
          tc::fence_proxy_async();
          __syncwarp();
          if (tc::elect_one()) {
            const bool is_v = nb >= 512;
            const int col = nb - (is_v ? 512 : 256);
            tc::tma_store_2d(is_v ? &tmVH : &tmKH, stg + sbuf * STG_BYTES, col, m0 + q * 32);
            tc::tma_store_2d(is_v ? &tmVL : &tmKL, stg + sbuf * STG_BYTES + 2048, col, m0 + q * 32);
            tc::tma_store_commit();
          }
          sbuf ^= 1;
          continue;
        }
        const bool split_k = g.KLO != nullptr && nb >= 256 && nb < 512;
        // stage [32 rows x 32 cols] in the 128B-swizzled layout the tensor map expects, then TMA store
        tc::tma_store_wait_read<1>();                          // the buffer used two stores ago is free
        __syncwarp();
        float4* so = reinterpret_cast<float4*>(stg + sbuf * STG_BYTES + lane * 128);
        if (split_k) {
          float l[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { const float h = tf32_hi(v[j]); l[j] = tf32_hi(v[j] - h); v[j] = h; }
#pragma unroll
          for (int j = 0; j < 8; ++j) so[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          tc::fence_proxy_async();
          __syncwarp();
          if (tc::elect_one()) { tc::tma_store_2d(&tmC, stg + sbuf * STG_BYTES, nb, m0 + q * 32); tc::tma_store_commit(); }
          sbuf ^= 1;
          tc::tma_store_wait_read<1>();
          __syncwarp();
          float4* sl = reinterpret_cast<float4*>(stg + sbuf * STG_BYTES + lane * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) sl[j ^ (lane & 7)] = make_float4(l[4 * j], l[4 * j + 1], l[4 * j + 2], l[4 * j + 3]);
          tc::fence_proxy_async();
          __syncwarp();
          if (tc::elect_one()) { tc::tma_store_2d(&tmKLO, stg + sbuf * STG_BYTES, nb - 256, m0 + q * 32); tc::tma_store_commit(); }
          sbuf ^= 1;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) so[j ^ (lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          tc::fence_proxy_async();
          __syncwarp();
          const int c_row = m0 + q * 32 + (SCORE ? 0 : prob * g.tiles_m * BM);     // split-K: slab of this K slice
          if (tc::elect_one()) { tc::tma_store_2d(&tmC, stg + sbuf * STG_BYTES, nb, c_row); tc::tma_store_commit(); }
          sbuf ^= 1;
        }
      }
      // accumulator fully read: hand it back to the MMA warp
      tc::tc_fence_before();
      tc::mbar_arrive(acc_empty + buf);
    }
    tc::tma_store_wait_all();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<512>(tmem_base);
}

}  // namespace

// Requirements (checked by the dispatcher in gemm_tc.cu): pre-split W planes, N % 128 == 0, K % 32 == 0,
// K1 % 32 == 0, 16-byte aligned rows.
int launch_gemm_tc_persist(const GemmDesc& d, float* VT, int vt_col0, int n_pad, float* KLO, float* VTLO,
                           cudaStream_t stream, const HalfPlanes* hp, int ksplit, float* slabs) {
  // split-K: partial products of the K slices go to slabs [ksplit, M, N] (M a multiple of the tile height), summed by
  // launch_splitk_reduce afterwards; no bias / residual / activation / concat in that mode
  MVM_REQUIRE(ksplit >= 1 && (ksplit == 1 || (slabs && d.M % BM == 0 && d.K % (32 * ksplit) == 0 && !d.Whi16 && !d.bias && !d.R && !d.A2 &&
                                               !d.relu && !VT && !KLO && !hp)));
  mvm_once_per_device(MVM_ONCE_GEMM_PERSIST, [&] {
    cudaFuncSetAttribute(gemm_tc_persist_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    cudaFuncSetAttribute(gemm_tc_persist_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  const int n_sm = mvm_dev_info().n_sm;
  // fp16x3 when the half-precision weight planes are given and the shape allows 64-wide k-blocks
  const bool f16 = d.Whi16 != nullptr && d.Wlo16 != nullptr && d.K % 64 == 0 && d.K1 % 64 == 0 && d.wscale > 0.f;
  const CUtensorMap* tA = mvm_get_tmap_2d(d.A, d.M, d.K1, d.lda, BM);
  const CUtensorMap* tA2 = d.A2 ? mvm_get_tmap_2d(d.A2, d.M, d.K - d.K1, d.lda2, BM) : tA;
  const CUtensorMap* tWhi = f16 ? mvm_get_tmap_2d_f16(d.Whi16, d.N, d.K, d.ldw, BN) : mvm_get_tmap_2d(d.Whi, d.N, d.K, d.ldw, BN);
  const CUtensorMap* tWlo = f16 ? mvm_get_tmap_2d_f16(d.Wlo16, d.N, d.K, d.ldw, BN) : mvm_get_tmap_2d(d.Wlo, d.N, d.K, d.ldw, BN);
  const CUtensorMap* tC = ksplit > 1 ? mvm_get_tmap_2d(slabs, (long long)ksplit * d.M, d.N, d.N, 32) : mvm_get_tmap_2d(d.C, d.M, d.N, d.ldc, 32);
  const CUtensorMap* tK = KLO ? mvm_get_tmap_2d(KLO, d.M, 256, 256, 32) : tC;
  const CUtensorMap* tKH = hp ? mvm_get_tmap_2d_f16_store(hp->kh, d.M, 256, 256) : tC;
  const CUtensorMap* tKL = hp ? mvm_get_tmap_2d_f16_store(hp->kl, d.M, 256, 256) : tC;
  const CUtensorMap* tVH = hp ? mvm_get_tmap_2d_f16_store(hp->vh, d.M, 256, 256) : tC;
  const CUtensorMap* tVL = hp ? mvm_get_tmap_2d_f16_store(hp->vl, d.M, 256, 256) : tC;
  if (!tA || !tA2 || !tWhi || !tWlo || !tC || !tK || !tKH || !tKL || !tVH || !tVL) return MVM_ERR_LAUNCH;
  PArgs g;
  g.bias = d.bias; g.R = d.R; g.ldr = d.ldr; g.C = ksplit > 1 ? slabs : d.C; g.ldc = ksplit > 1 ? d.N : d.ldc; g.M = d.M; g.N = d.N; g.K = d.K;
  g.K1 = d.K1; g.alpha = f16 ? d.alpha / d.wscale : d.alpha; g.relu = d.relu; g.VT = VT; g.vt_col0 = vt_col0; g.n_pad = n_pad;
  g.KLO = KLO; g.VTLO = VTLO;
  g.KH16 = hp ? (__half*)hp->kh : nullptr; g.KL16 = hp ? (__half*)hp->kl : nullptr;
  g.VH16 = hp ? (__half*)hp->vh : nullptr; g.VL16 = hp ? (__half*)hp->vl : nullptr;
  g.tiles_m = mvm_div_up(d.M, BM); g.tiles_n = d.N / BN;
  g.ksplit = ksplit;
  const int n_tiles = g.tiles_m * g.tiles_n * ksplit;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  ScoreTab none;
  none.n_pairs = 0; none.batch = 0; none.n_views = 0; none.n_pad = 0;
  if (f16) gemm_tc_persist_kernel<false, true><<<grid, NTHREADS, SMEM_BYTES, stream>>>(*tA, *tA2, *tWhi, *tWlo, *tC, *tK, *tKH, *tKL, *tVH, *tVL, g, none);
  else gemm_tc_persist_kernel<false, false><<<grid, NTHREADS, SMEM_BYTES, stream>>>(*tA, *tA2, *tWhi, *tWlo, *tC, *tK, *tKH, *tKL, *tVH, *tVL, g, none);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}

namespace {
// C[m, n] = sum_s slabs[s][m][n] (fixed order: deterministic)
__global__ void splitk_reduce_kernel(const float4* __restrict__ slabs, float* __restrict__ C, int M, int N4, int ldc, int S) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)M * N4) return;
  const int m = (int)(i / N4), n4 = (int)(i % N4);
  float4 a = slabs[i];
  for (int s = 1; s < S; ++s) {
    const float4 b = slabs[(long long)s * M * N4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  *reinterpret_cast<float4*>(C + (long long)m * ldc + 4 * n4) = a;
}
}  // namespace

int launch_splitk_reduce(const float* slabs, float* C, int M, int N, int ldc, int ksplit, cudaStream_t stream) {
  MVM_REQUIRE(N % 4 == 0 && ldc % 4 == 0);
  const long long n = (long long)M * (N / 4);
  splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(slabs), C, M, N / 4, ldc, ksplit);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}

namespace {
// hi = rn_tf32(x), lo = rn_tf32(x - hi) of a whole buffer (the W-operand planes of the score GEMM)
__global__ void split_planes_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo,
                                    long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = tf32_hi(v.x - h.x); l.y = tf32_hi(v.y - h.y); l.z = tf32_hi(v.z - h.z); l.w = tf32_hi(v.w - h.w);
    hi[i] = h;
    lo[i] = l;
  }
}
}  // namespace

// All (pair, tuple) score matrices on the tensor cores (3xTF32).  mdesc [rows_total, 256] point-major; hi / lo:
// scratch planes of the same size (filled here).
int launch_score_gemm_tc(const float* mdesc, float* hi, float* lo, int n_pad, const PairTable& tab, int batch,
                         float alpha, cudaStream_t stream) {
  MvmProfScope prof__(MVM_TAG_SCORE, stream);
  mvm_once_per_device(MVM_ONCE_GEMM_SCORE, [&] {
    cudaFuncSetAttribute(gemm_tc_persist_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  const int n_sm = mvm_dev_info().n_sm;
  const long long rows = (long long)batch * tab.n_views * n_pad;
  split_planes_kernel<<<n_sm * 4, 256, 0, stream>>>(reinterpret_cast<const float4*>(mdesc), reinterpret_cast<float4*>(hi),
                                                    reinterpret_cast<float4*>(lo), rows * 256 / 4);
  MVM_CHECK_LAUNCH();
  const CUtensorMap* tA = mvm_get_tmap_2d(mdesc, rows, 256, 256, BM);
  const CUtensorMap* tWhi = mvm_get_tmap_2d(hi, rows, 256, 256, BN);
  const CUtensorMap* tWlo = mvm_get_tmap_2d(lo, rows, 256, 256, BN);
  if (!tA || !tWhi || !tWlo) return MVM_ERR_LAUNCH;
  ScoreTab st;
  st.n_pairs = tab.n_pairs; st.batch = batch; st.n_views = tab.n_views; st.n_pad = n_pad;
  int max_m = 0, max_n = 0;
  for (int p = 0; p < tab.n_pairs; ++p) {
    st.a[p] = tab.a[p]; st.b[p] = tab.b[p]; st.m[p] = tab.m[p]; st.n[p] = tab.n[p]; st.scores[p] = tab.scores[p];
    max_m = tab.m[p] > max_m ? tab.m[p] : max_m;
    max_n = tab.n[p] > max_n ? tab.n[p] : max_n;
  }
  PArgs g;
  memset(&g, 0, sizeof(g));
  g.K = 256; g.K1 = 256; g.alpha = alpha; g.ksplit = 1;
  g.tiles_m = mvm_div_up(max_m, BM); g.tiles_n = mvm_div_up(max_n, BN);
  const long long n_tiles = (long long)g.tiles_m * g.tiles_n * tab.n_pairs * batch;
  const int grid = n_tiles < n_sm ? (int)n_tiles : n_sm;
  gemm_tc_persist_kernel<true, false><<<grid, NTHREADS, SMEM_BYTES, stream>>>(*tA, *tA, *tWhi, *tWlo, *tA, *tA, *tA, *tA, *tA, *tA, g, st);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}
