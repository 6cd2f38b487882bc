// Internal launch prototypes (host side) of the matcher kernels.
#pragma once
#include "common.cuh"

constexpr int MVM_MAX_PAIRS = 28;  // C(8,2)

// Per-pair table passed BY VALUE to the pair-stage kernels (score GEMM, Sinkhorn, match
// extraction, confidence head): one launch covers every (pair, batch) problem and the
// whole forward stays CUDA-graph capturable (no host->device table copies).
struct PairTable {
  int n_pairs;
  int n_views;                 // views per tuple (slot of view t in tuple b = b*n_views + t)
  int a[MVM_MAX_PAIRS], b[MVM_MAX_PAIRS];   // view ids, a < b
  int m[MVM_MAX_PAIRS], n[MVM_MAX_PAIRS];   // true keypoint counts of a and b
  float* scores[MVM_MAX_PAIRS];             // [batch, m+1, n+1]
  int64_t* matches_a[MVM_MAX_PAIRS];        // [batch, m]
  int64_t* matches_b[MVM_MAX_PAIRS];        // [batch, n]
  float* ms_a[MVM_MAX_PAIRS];
  float* ms_b[MVM_MAX_PAIRS];
  float* conf[MVM_MAX_PAIRS];               // [batch, m] (or nullptr)
  long long ws_off[MVM_MAX_PAIRS];          // float offset of this pair's (u,v) scratch
};
typedef PairTable SinkhornTable;

int launch_kenc_front(const float* kpts, const float* kscores, const float* const* w,
                      const float* const* b, float* h3, int n_points, float img_w, float img_h,
                      cudaStream_t stream);
int launch_transpose_cn(const float* in, float* out, int n_views_total, int C, int n_pad,
                        cudaStream_t stream);

// tcgen05 GEMM (gemm_tc.cu): n_pass 3 = fp32-faithful 3xTF32, 1 = single-pass TF32
// gemm_tile (128 / 256) and gemm_persist (0 / 1) select the kernel; -1 = the process defaults
int launch_gemm_tc(const GemmDesc& d, int n_pass, float* VT, int vt_col0, int n_pad, cudaStream_t stream,
                   float* KLO = nullptr, float* VTLO = nullptr, int gemm_tile = -1, int gemm_persist = -1);
int mvm_default_gemm_tile();
int mvm_default_gemm_persistent();
// persistent 3xTF32 kernel (A operand in tensor memory, double-buffered accumulators, TMA-store epilogue)
// hp != nullptr (QKV projection, vt_col0 = 512): the K third and V^T leave as half-precision hi / lo planes for
// launch_attention_h3 instead of the fp32 / tf32 buffers (kh, kl, vh, vl: all [rows, 256], as __half; V stays key-major)
struct HalfPlanes { void* kh; void* kl; void* vh; void* vl; };
int launch_gemm_tc_persist(const GemmDesc& d, float* VT, int vt_col0, int n_pad, float* KLO, float* VTLO,
                           cudaStream_t stream, const HalfPlanes* hp = nullptr, int ksplit = 1, float* slabs = nullptr);
int launch_splitk_reduce(const float* slabs, float* C, int M, int N, int ldc, int ksplit, cudaStream_t stream);
// every (pair, tuple) score matrix in one launch of the persistent kernel (3xTF32); hi / lo: scratch [rows, 256]
struct PairTable;
int launch_score_gemm_tc(const float* mdesc, float* hi, float* lo, int n_pad, const PairTable& tab, int batch,
                         float alpha, cudaStream_t stream);

struct AttnSegs {
  int n_views;
  int counts[8];
};
// klo / vtlo: tf32 remainder planes of K [rows,256] and V^T (n_pass == 3; K and V^T then hold the rn_tf32 parts)
int launch_attention_tc(const float* qkv, const float* vt, float* out, int batch, int n_pad, AttnSegs segs,
                        int is_cross, int n_pass, cudaStream_t stream, const float* klo = nullptr,
                        const float* vtlo = nullptr);
int launch_attention_simt(const float* qkv, float* out, int batch, int n_pad, AttnSegs segs,
                          int is_cross, cudaStream_t stream);
// fp32-faithful attention with half-precision operand planes (fp16x3, attention_h3.cu); planes as in HalfPlanes
struct __half;
int launch_attention_h3(const float* qkv, const __half* kh, const __half* kl, const __half* vh, const __half* vl,
                        float* out, int batch, int n_pad, AttnSegs segs, int is_cross, cudaStream_t stream);

// scores[p][bi] inner block = mdesc[a] . mdesc[b]^T * alpha   (mdesc: [views, n_pad, 256])
int launch_score_gemm_simt(const float* mdesc, int n_pad, const PairTable& tab, int batch,
                           float alpha, cudaStream_t stream);

int launch_sinkhorn_ref(const SinkhornTable& tab, int batch, float bin_score, int iters,
                        float* ws, cudaStream_t stream);
// production dispatch (sinkhorn_exp.cu): cluster kernel up to 1024 x 1024, multi-CTA kernel beyond;
// variant 0 = automatic, 1 = multi-CTA kernel, 2 = cluster kernel
int launch_sinkhorn(const SinkhornTable& tab, int batch, float bin_score, int iters, float* ws,
                    cudaStream_t stream, int variant = 0);
int launch_sinkhorn_multicta(const SinkhornTable& tab, int batch, float bin_score, int iters, float* ws,
                             cudaStream_t stream);
// one thread-block cluster per problem (sinkhorn_cl.cu)
int sinkhorn_cluster_size(int max_m, int max_n);
int sinkhorn_cluster_max_active(int C, int n);
int launch_sinkhorn_cluster(const SinkhornTable& tab, int batch, float bin_score, int iters, int C,
                            cudaStream_t stream, int rr = 0);
int launch_sinkhorn_log(const SinkhornTable& tab, int batch, float bin_score, int iters, float* ws,
                        cudaStream_t stream);
size_t sinkhorn_ws_floats(int n_pairs, int batch, int n_pad);

// idx_ws: ints [n_pairs*batch*2*n_pad] + floats; see match.cu
int launch_extract_matches(const PairTable& tab, int batch, int n_pad, float thresh, int* idx_ws,
                           cudaStream_t stream);

// feat [n_pairs*batch*n_pad, 512] = cat(mdesc_a[i], mdesc_b[match(i)]); sc [rows] = Z[i, match(i)]
int launch_conf_gather(const float* mdesc, const PairTable& tab, int batch, int n_pad, float* feat,
                       float* sc, cudaStream_t stream);
int launch_conf_c0(const float* sc, const float* w, const float* b, float* out, long long rows,
                   cudaStream_t stream);
int launch_conf_final(const float* h, const float* wl, float bl, const PairTable& tab, int batch,
                      int n_pad, cudaStream_t stream);
