// tcgen05 GEMM with fused epilogue for the 1x1 convolutions of the matcher
// (superglue.py:51-62,101-121; multi_view_matcher.py:8-53):
//   C[M,N] = act(alpha * [A | A2][M,K] . W[N,K]^T + bias[N]) + R[M,N]      (all fp32, K-major)
//
// Warp roles (192 threads, one 128 x BN output tile per CTA):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled [128 x 32] A and [BN x 32] W tiles
//               into a STAGES-deep ring (mbarrier expect_tx / complete_tx)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, M = 128, N = BN,
//               K = 8 per instruction, accumulator in TMEM), tcgen05.commit frees the ring slots
//   warps 2-5   (NPASS == 3) operand splitters during the main loop: lo = x - tf32(x) of every
//               landed tile into a second smem buffer, so that D += A.W + A.W_lo + A_lo.W
//               reproduces fp32 products to ~2^-21 (the "3xTF32" scheme) with no extra HBM traffic;
//               then the epilogue: tcgen05.ld 32x32b -> registers -> bias/ReLU/residual -> global
// NPASS == 1 is the single-pass TF32 mode (what torch 1.10 did by default on Ampere+).
#include "common.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace {

constexpr int BM = 128, BK = 32;
constexpr int NTHREADS = 192;

struct GemmTcArgs {
  const float* bias;
  const float* R; int ldr;
  float* C; int ldc;
  int M, N, K, K1;
  float alpha;
  int relu;
  // optional transposed output for columns >= vt_col0: VT[(m / n_pad), n - vt_col0, m % n_pad]
  float* VT; int vt_col0; int n_pad;
  // optional tf32 hi/lo planes for the attention operands (3xTF32 mode): columns [256,512) (= K) are
  // stored as rn_tf32 in C with the remainder in KLO [M,256]; V^T likewise in VT / VTLO
  float* KLO; float* VTLO;
};

template <int BN, int NPASS>
struct Cfg {
  static constexpr int STAGES = (BN == 256) ? (NPASS == 3 ? 2 : 4) : (NPASS == 3 ? 3 : 4);
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int W_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = (A_BYTES + W_BYTES) * (NPASS == 3 ? 2 : 1);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int NPASS, bool PRE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmWlo, GemmTcArgs g) {
  using C_ = Cfg<BN, NPASS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C_::STAGES * C_::STAGE_BYTES);
  uint64_t* full = bars;                       // TMA landed
  uint64_t* empty = bars + C_::STAGES;         // MMAs that read the slot retired
  uint64_t* split = bars + 2 * C_::STAGES;     // lo planes written (NPASS == 3)
  uint64_t* tmem_full = bars + 3 * C_::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C_::STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nk = g.K / BK;

  auto stage_A = [&](int s) { return smem + s * C_::STAGE_BYTES; };
  auto stage_W = [&](int s) { return smem + s * C_::STAGE_BYTES + C_::A_BYTES; };
  auto stage_Alo = [&](int s) { return smem + s * C_::STAGE_BYTES + C_::A_BYTES + C_::W_BYTES; };
  auto stage_Wlo = [&](int s) { return smem + s * C_::STAGE_BYTES + 2 * C_::A_BYTES + C_::W_BYTES; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < C_::STAGES; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
      tc::mbar_init(split + s, 128);
    }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmA2);
    tc::prefetch_tmap(&tmW);
  }
  if (warp == 1) tc::tmem_alloc<BN>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % C_::STAGES;
      const uint32_t ph = (kt / C_::STAGES) & 1;
      tc::mbar_wait(empty + s, ph ^ 1);
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx(full + s, C_::A_BYTES + C_::W_BYTES * (PRE ? 2 : 1));
        const int k = kt * BK;
        if (k < g.K1) tc::tma_load_2d(stage_A(s), &tmA, full + s, k, m0);
        else tc::tma_load_2d(stage_A(s), &tmA2, full + s, k - g.K1, m0);
        tc::tma_load_2d(stage_W(s), &tmW, full + s, k, n0);            // PRE: the rn_tf32 plane of W
        if (PRE) tc::tma_load_2d(stage_Wlo(s), &tmWlo, full + s, k, n0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // converged warp, one elected lane issues (see tc::elect_one)
    constexpr uint32_t idesc = tc::make_idesc_tf32(BM, BN);
    for (int kt = 0; kt < nk; ++kt) {
      const int s = kt % C_::STAGES;
      const uint32_t ph = (kt / C_::STAGES) & 1;
      tc::mbar_wait(full + s, ph);
      if (NPASS == 3) tc::mbar_wait(split + s, ph);
      tc::tc_fence_after();
      const uint32_t a = tc::smem_u32(stage_A(s)), w = tc::smem_u32(stage_W(s));
      const uint32_t alo = tc::smem_u32(stage_Alo(s)), wlo = tc::smem_u32(stage_Wlo(s));
      if (tc::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk) {
          const uint32_t off = kk * 32;   // 8 tf32 = 32 bytes along K inside the 128B swizzle span
          const uint64_t da = tc::make_kmajor_sw128_desc(a + off), dw = tc::make_kmajor_sw128_desc(w + off);
          tc::umma_tf32(tmem_base, da, dw, idesc, (kt | kk) != 0);
          if (NPASS == 3) {
            tc::umma_tf32(tmem_base, da, tc::make_kmajor_sw128_desc(wlo + off), idesc, 1);
            tc::umma_tf32(tmem_base, tc::make_kmajor_sw128_desc(alo + off), dw, idesc, 1);
          }
        }
        tc::umma_commit(empty + s);
        if (kt == nk - 1) tc::umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else {
    // ================================ splitters, then epilogue ================================
    const int et = threadIdx.x - 64;   // 0..127
    if (NPASS == 3) {
      for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % C_::STAGES;
        const uint32_t ph = (kt / C_::STAGES) & 1;
        tc::mbar_wait(full + s, ph);
        // lo planes are element-wise images of the landed tiles: same (swizzled) offsets
        // hi = rn_tf32(x) replaces the landed tile in place, lo = rn_tf32(x - hi) goes to the second
        // buffer; both are element-wise images of the tile, so the (swizzled) offsets carry over
        float4* a = reinterpret_cast<float4*>(stage_A(s));
        float4* alo = reinterpret_cast<float4*>(stage_Alo(s));
#pragma unroll
        for (int i = 0; i < (BM * BK / 4) / 128; ++i) {
          const float4 x = a[et + i * 128];
          float4 h;
          h.x = tc::tf32_rn(x.x); h.y = tc::tf32_rn(x.y); h.z = tc::tf32_rn(x.z); h.w = tc::tf32_rn(x.w);
          a[et + i * 128] = h;
          alo[et + i * 128] = make_float4(tc::tf32_rn(x.x - h.x), tc::tf32_rn(x.y - h.y), tc::tf32_rn(x.z - h.z),
                                          tc::tf32_rn(x.w - h.w));
        }
        float4* w = reinterpret_cast<float4*>(stage_W(s));
        float4* wlo = reinterpret_cast<float4*>(stage_Wlo(s));
#pragma unroll
        for (int i = 0; i < (PRE ? 0 : (BN * BK / 4) / 128); ++i) {
          const float4 x = w[et + i * 128];
          float4 h;
          h.x = tc::tf32_rn(x.x); h.y = tc::tf32_rn(x.y); h.z = tc::tf32_rn(x.z); h.w = tc::tf32_rn(x.w);
          w[et + i * 128] = h;
          wlo[et + i * 128] = make_float4(tc::tf32_rn(x.x - h.x), tc::tf32_rn(x.y - h.y), tc::tf32_rn(x.z - h.z),
                                          tc::tf32_rn(x.w - h.w));
        }
        tc::fence_proxy_async();      // generic-proxy writes -> visible to the tensor core (async proxy)
        tc::mbar_arrive(split + s);
      }
    }
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const int q = warp & 3;                     // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int m = m0 + row;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    for (int c = 0; c < BN / 32; ++c) {
      float v[32];
      tc::tmem_ld32(taddr + c * 32, v);
      tc::tmem_ld_wait();
      if (m < g.M) {
        const int nb = n0 + c * 32;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = g.alpha * v[j];
          if (g.bias) x += __ldg(g.bias + nb + j);
          if (g.relu) x = fmaxf(x, 0.f);
          v[j] = x;
        }
        if (g.R) {
          const float4* r4 = reinterpret_cast<const float4*>(g.R + (long long)m * g.ldr + nb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r = r4[j];
            v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
          }
        }
        if (g.VT && nb >= g.vt_col0) {
          const int slab = m / g.n_pad, i = m % g.n_pad;
          const long long off = ((long long)slab * (g.N - g.vt_col0) + (nb - g.vt_col0)) * g.n_pad + i;
          if (g.VTLO) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float hi = tc::tf32_rn(v[j]);
              g.VT[off + (long long)j * g.n_pad] = hi;
              g.VTLO[off + (long long)j * g.n_pad] = tc::tf32_rn(v[j] - hi);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) g.VT[off + (long long)j * g.n_pad] = v[j];
          }
        } else if (g.KLO && nb >= 256 && nb < 512) {
          float4* o = reinterpret_cast<float4*>(g.C + (long long)m * g.ldc + nb);
          float4* ol = reinterpret_cast<float4*>(g.KLO + (long long)m * 256 + (nb - 256));
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 h;
            h.x = tc::tf32_rn(v[4 * j]); h.y = tc::tf32_rn(v[4 * j + 1]); h.z = tc::tf32_rn(v[4 * j + 2]); h.w = tc::tf32_rn(v[4 * j + 3]);
            o[j] = h;
            ol[j] = make_float4(tc::tf32_rn(v[4 * j] - h.x), tc::tf32_rn(v[4 * j + 1] - h.y), tc::tf32_rn(v[4 * j + 2] - h.z),
                                tc::tf32_rn(v[4 * j + 3] - h.w));
          }
        } else {
          float4* o = reinterpret_cast<float4*>(g.C + (long long)m * g.ldc + nb);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<BN>(tmem_base);
}

// ---- host: tensor-map cache -------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}
typedef std::tuple<const void*, long long, long long, long long, long long, long long, int> TmKey;
std::map<TmKey, CUtensorMap*> g_tmaps;
std::mutex g_tmap_mu;

template <int BN, int NPASS, bool PRE>
int launch_cfg(const GemmDesc& d, float* VT, int vt_col0, int n_pad, float* KLO, float* VTLO, cudaStream_t stream) {
  using C_ = Cfg<BN, NPASS>;
  // cheap and idempotent: set on every launch (per-device attribute; 12 template instances)
  cudaFuncSetAttribute(gemm_tc_kernel<BN, NPASS, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM_BYTES);
  const CUtensorMap* tA = mvm_get_tmap_2d(d.A, d.M, d.K1, d.lda, BM);
  const CUtensorMap* tA2 = d.A2 ? mvm_get_tmap_2d(d.A2, d.M, d.K - d.K1, d.lda2, BM) : tA;
  const CUtensorMap* tW = mvm_get_tmap_2d(PRE ? d.Whi : d.W, d.N, d.K, d.ldw, BN);
  const CUtensorMap* tWlo = PRE ? mvm_get_tmap_2d(d.Wlo, d.N, d.K, d.ldw, BN) : tW;
  if (!tA || !tA2 || !tW || !tWlo) return MVM_ERR_LAUNCH;
  GemmTcArgs g;
  g.bias = d.bias; g.R = d.R; g.ldr = d.ldr; g.C = d.C; g.ldc = d.ldc; g.M = d.M; g.N = d.N; g.K = d.K;
  g.K1 = d.K1; g.alpha = d.alpha; g.relu = d.relu; g.VT = VT; g.vt_col0 = vt_col0; g.n_pad = n_pad; g.KLO = KLO; g.VTLO = VTLO;
  dim3 grid(d.N / BN, mvm_div_up(d.M, BM));
  gemm_tc_kernel<BN, NPASS, PRE><<<grid, NTHREADS, C_::SMEM_BYTES, stream>>>(*tA, *tA2, *tW, *tWlo, g);
  MVM_CHECK_LAUNCH();
  return MVM_OK;
}

}  // namespace

const CUtensorMap* mvm_get_tmap_3d(const float* base, long long slabs, long long rows, long long cols,
                                   long long ld_row, long long ld_slab, int box_rows) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  TmKey key(base, slabs, rows, cols, ld_row, ld_slab, box_rows);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return it->second;
  EncodeFn enc = get_encode();
  if (!enc) return nullptr;
  CUtensorMap* tm = new CUtensorMap;
  const int rank = slabs > 0 ? 3 : 2;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)(slabs > 0 ? slabs : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld_row * 4, (cuuint64_t)ld_slab * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[mvm_b200] cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld\n", (int)r, rows, cols, ld_row);
    delete tm;
    return nullptr;
  }
  g_tmaps[key] = tm;
  return tm;
}

const CUtensorMap* mvm_get_tmap_2d(const float* base, long long rows, long long cols, long long ld, int box_rows) {
  return mvm_get_tmap_3d(base, 0, rows, cols, ld, 0, box_rows);
}

// 2-D fp16 row-major [rows, cols] with row stride ld (elements), box = [box_rows, 64 cols = 128 B], 128B swizzle
const CUtensorMap* mvm_get_tmap_2d_f16(const void* base, long long rows, long long cols, long long ld, int box_rows) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  TmKey key(base, -16, rows, cols, ld, 0, box_rows);      // slabs = -16 marks the half-precision maps
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return it->second;
  EncodeFn enc = get_encode();
  if (!enc) return nullptr;
  CUtensorMap* tm = new CUtensorMap;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[mvm_b200] cuTensorMapEncodeTiled (f16) failed (%d) rows=%lld cols=%lld ld=%lld\n", (int)r, rows, cols, ld);
    delete tm;
    return nullptr;
  }
  g_tmaps[key] = tm;
  return tm;
}

// 2-D fp16 row-major [rows, cols] STORE map of the persistent GEMM's plane epilogue: box = [32 rows, 32 cols = 64 B],
// 64-byte swizzle (a warp stages 32 rows x 64 B conflict-free and one TMA store writes them as full lines)
const CUtensorMap* mvm_get_tmap_2d_f16_store(const void* base, long long rows, long long cols, long long ld) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  TmKey key(base, -17, rows, cols, ld, 0, 32);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return it->second;
  EncodeFn enc = get_encode();
  if (!enc) return nullptr;
  CUtensorMap* tm = new CUtensorMap;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[mvm_b200] cuTensorMapEncodeTiled (f16 store) failed (%d) rows=%lld cols=%lld ld=%lld\n", (int)r, rows, cols, ld);
    delete tm;
    return nullptr;
  }
  g_tmaps[key] = tm;
  return tm;
}

int g_gemm_bn = 256;   // output tile width of the one-tile-per-CTA tcgen05 GEMM (128 or 256), see mvm_debug_set_gemm_tile
extern "C" void mvm_debug_set_gemm_tile(int bn) { g_gemm_bn = bn == 256 ? 256 : 128; }
int g_gemm_persist = 1;   // 1: the persistent kernel of gemm_tc_persist.cu serves the 3xTF32 path (default)
extern "C" void mvm_debug_set_gemm_kernel(int persistent) { g_gemm_persist = persistent ? 1 : 0; }

// GEMM on the tensor cores.  Requirements: K, K1 multiples of 32, N multiple of 128, 16-byte aligned
// rows (lda/ldw/ldc/ldr multiples of 4).  n_pass: 3 = fp32-faithful 3xTF32, 1 = single-pass TF32.
int mvm_default_gemm_tile() { return g_gemm_bn; }
int mvm_default_gemm_persistent() { return g_gemm_persist; }

int launch_gemm_tc(const GemmDesc& d, int n_pass, float* VT, int vt_col0, int n_pad, cudaStream_t stream,
                   float* KLO, float* VTLO, int gemm_tile, int gemm_persist) {
  if (gemm_tile < 0) gemm_tile = g_gemm_bn;            // stage-level callers: the process defaults
  if (gemm_persist < 0) gemm_persist = g_gemm_persist;
  MVM_REQUIRE(d.batch == 1 && d.K % BK == 0 && d.K1 % BK == 0 && d.N % 128 == 0);
  MVM_REQUIRE(d.lda % 4 == 0 && d.ldw % 4 == 0 && d.ldc % 4 == 0 && (d.R == nullptr || d.ldr % 4 == 0));
  MVM_REQUIRE(d.A2 == nullptr || d.lda2 % 4 == 0);
  MvmProfScope prof__(MVM_TAG_GEMM, stream);
  if (gemm_persist && n_pass == 3 && d.Whi && d.Wlo) return launch_gemm_tc_persist(d, VT, vt_col0, n_pad, KLO, VTLO, stream);
  if (gemm_tile == 256 && d.N % 256 == 0) {
    if (n_pass == 3 && d.Whi && d.Wlo) return launch_cfg<256, 3, true>(d, VT, vt_col0, n_pad, KLO, VTLO, stream);
    if (n_pass == 1) return launch_cfg<256, 1, false>(d, VT, vt_col0, n_pad, nullptr, nullptr, stream);
  }
  if (n_pass == 3 && d.Whi && d.Wlo) return launch_cfg<128, 3, true>(d, VT, vt_col0, n_pad, KLO, VTLO, stream);
  if (n_pass == 3) return launch_cfg<128, 3, false>(d, VT, vt_col0, n_pad, KLO, VTLO, stream);
  return launch_cfg<128, 1, false>(d, VT, vt_col0, n_pad, nullptr, nullptr, stream);
}
