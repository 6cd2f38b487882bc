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
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 4;
constexpr int NTHREADS = 320;
constexpr int A_BYTES = BM * BK * 4;              // 16 KB
constexpr int W_BYTES = BN * BK * 4;              // 16 KB per plane
constexpr int STAGE_BYTES = A_BYTES + 2 * W_BYTES;
constexpr int STG_BYTES = 32 * 128;               // one staged [32 rows x 32 cols] block per epilogue warp
constexpr int OFF_STG = STAGES * STAGE_BYTES;     // 4 warps x 2 buffers
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
  int tiles_m, tiles_n;
};

// rn_tf32 of a finite value (ties away, == cvt.rna.tf32.f32) in two integer instructions
__device__ __forceinline__ float tf32_hi(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                       const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                       const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmKLO, PArgs g) {
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
  const int nk = g.K / BK;
  const int n_tiles = g.tiles_m * g.tiles_n;

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
      const int m0 = (tile / g.tiles_n) * BM, n0 = (tile % g.tiles_n) * BN;
      for (int kt = 0; kt < nk; ++kt, ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(empty + s, ((it / STAGES) & 1) ^ 1);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(full + s, STAGE_BYTES);
          uint8_t* st = smem + s * STAGE_BYTES;
          const int k = kt * BK;
          if (k < g.K1) tc::tma_load_2d(st, &tmA, full + s, k, m0);
          else tc::tma_load_2d(st, &tmA2, full + s, k - g.K1, m0);
          tc::tma_load_2d(st + A_BYTES, &tmWhi, full + s, k, n0);
          tc::tma_load_2d(st + A_BYTES + W_BYTES, &tmWlo, full + s, k, n0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // the whole warp walks the loop (converged); one elected lane issues
    constexpr uint32_t idesc = tc::make_idesc_tf32(BM, BN);
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
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t dhi = tc::make_kmajor_sw128_desc(whi + kk * 32);
            tc::umma_tf32_ts(acc, a_hi + kk * 8, dhi, idesc, (kt | kk) != 0);
            tc::umma_tf32_ts(acc, a_hi + kk * 8, tc::make_kmajor_sw128_desc(wlo + kk * 32), idesc, 1);
            tc::umma_tf32_ts(acc, a_lo + kk * 8, dhi, idesc, 1);
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
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 x = a[c ^ (row & 7)];
          hi[4 * c] = tf32_hi(x.x); hi[4 * c + 1] = tf32_hi(x.y); hi[4 * c + 2] = tf32_hi(x.z); hi[4 * c + 3] = tf32_hi(x.w);
          lo[4 * c] = tf32_hi(x.x - hi[4 * c]); lo[4 * c + 1] = tf32_hi(x.y - hi[4 * c + 1]);
          lo[4 * c + 2] = tf32_hi(x.z - hi[4 * c + 2]); lo[4 * c + 3] = tf32_hi(x.w - hi[4 * c + 3]);
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
      const int m0 = (tile / g.tiles_n) * BM, n0 = (tile % g.tiles_n) * BN;
      const int buf = i & 1;
      const int m = m0 + row;
      // bias of this tile's columns, shared by the four epilogue warps
      asm volatile("bar.sync 1, 128;" ::: "memory");           // previous tile's readers are done
      s_bias[et] = g.bias ? __ldg(g.bias + n0 + et) : 0.f;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      tc::mbar_wait(acc_full + buf, (i >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + TM_ACC + buf * BN + lane_addr;
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
          // V^T (and its lo plane) for the attention kernel: lanes = consecutive keypoints -> coalesced
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
          if (tc::elect_one()) { tc::tma_store_2d(&tmC, stg + sbuf * STG_BYTES, nb, m0 + q * 32); tc::tma_store_commit(); }
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
                           cudaStream_t stream) {
  static bool attr = false;
  static int n_sm = 0;
  if (!attr) {
    cudaFuncSetAttribute(gemm_tc_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  const CUtensorMap* tA = mvm_get_tmap_2d(d.A, d.M, d.K1, d.lda, BM);
  const CUtensorMap* tA2 = d.A2 ? mvm_get_tmap_2d(d.A2, d.M, d.K - d.K1, d.lda2, BM) : tA;
  const CUtensorMap* tWhi = mvm_get_tmap_2d(d.Whi, d.N, d.K, d.ldw, BN);
  const CUtensorMap* tWlo = mvm_get_tmap_2d(d.Wlo, d.N, d.K, d.ldw, BN);
  const CUtensorMap* tC = mvm_get_tmap_2d(d.C, d.M, d.N, d.ldc, 32);
  const CUtensorMap* tK = KLO ? mvm_get_tmap_2d(KLO, d.M, 256, 256, 32) : tC;
  if (!tA || !tA2 || !tWhi || !tWlo || !tC || !tK) return MVM_ERR_LAUNCH;
  PArgs g;
  g.bias = d.bias; g.R = d.R; g.ldr = d.ldr; g.C = d.C; g.ldc = d.ldc; g.M = d.M; g.N = d.N; g.K = d.K;
  g.K1 = d.K1; g.alpha = d.alpha; g.relu = d.relu; g.VT = VT; g.vt_col0 = vt_col0; g.n_pad = n_pad;
  g.KLO = KLO; g.VTLO = VTLO;
  g.tiles_m = mvm_div_up(d.M, BM); g.tiles_n = d.N / BN;
  const int n_tiles = g.tiles_m * g.tiles_n;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  gemm_tc_persist_kernel<<<grid, NTHREADS, SMEM_BYTES, stream>>>(*tA, *tA2, *tWhi, *tWlo, *tC, *tK, g);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}
