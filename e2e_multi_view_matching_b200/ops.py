"""Stage-level entry points (torch tensors in/out) over the C ABI.  Names follow the reference
functions they replace (superglue.py:87-172, multi_view_matcher.py:288-300)."""
import ctypes as C

import torch

from . import _lib


def rn_tf32(x):
    """cvt.rna.tf32.f32 on a tensor: round to nearest (ties away) on the 13 dropped mantissa bits."""
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def linear(a, w, bias=None, a2=None, residual=None, relu=False, alpha=1.0, tc_passes=0, presplit=False):
    """act(alpha * [a|a2] @ w.T + bias) + residual on point-major activations (Conv1d k=1).
    tc_passes: 0 = fp32 CUDA cores, 3 = tcgen05 3xTF32, 1 = tcgen05 single-pass TF32.
    presplit (with tc_passes=3): hand W over as its tf32 hi/lo planes, like the packed matcher weights do --
    this is the production path (persistent kernel)."""
    lib = _lib.lib()
    if tc_passes == 'h16':      # fp16x3 on the persistent kernel (what the packed matcher weights use by default)
        M, K1 = a.shape
        K = K1 + (a2.shape[1] if a2 is not None else 0)
        N = w.shape[0]
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
        scale = 64.0
        ws = w.double() * scale
        w_hi = ws.to(torch.float16)
        w_lo = (ws - w_hi.double()).to(torch.float16).contiguous()
        w_hi = w_hi.contiguous()
        rc = lib.mvm_linear_tc_h16(_lib.ptr(a), a.stride(0), _lib.ptr(a2), a2.stride(0) if a2 is not None else 0, K1,
                                   _lib.ptr(w_hi), _lib.ptr(w_lo), scale, w_hi.stride(0), _lib.ptr(bias), _lib.ptr(residual),
                                   residual.stride(0) if residual is not None else 0, _lib.ptr(out), N, M, N, K,
                                   float(alpha), int(relu), _lib.stream_ptr())
        _lib.check(rc, 'mvm_linear_tc_h16')
        return out
    if tc_passes:
        M, K1 = a.shape
        K = K1 + (a2.shape[1] if a2 is not None else 0)
        N = w.shape[0]
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
        if presplit:
            assert tc_passes == 3
            w_hi = rn_tf32(w)
            w_lo = rn_tf32(w - w_hi)
            rc = lib.mvm_linear_tc_presplit(_lib.ptr(a), a.stride(0), _lib.ptr(a2), a2.stride(0) if a2 is not None else 0,
                                            K1, _lib.ptr(w_hi), _lib.ptr(w_lo), w_hi.stride(0), _lib.ptr(bias),
                                            _lib.ptr(residual), residual.stride(0) if residual is not None else 0,
                                            _lib.ptr(out), N, M, N, K, float(alpha), int(relu), _lib.stream_ptr())
            _lib.check(rc, 'mvm_linear_tc_presplit')
            return out
        rc = lib.mvm_linear_tc(_lib.ptr(a), a.stride(0), _lib.ptr(a2), a2.stride(0) if a2 is not None else 0,
                               K1, _lib.ptr(w), w.stride(0), _lib.ptr(bias), _lib.ptr(residual),
                               residual.stride(0) if residual is not None else 0, _lib.ptr(out), N, M, N, K,
                               float(alpha), int(relu), int(tc_passes), _lib.stream_ptr())
        _lib.check(rc, 'mvm_linear_tc')
        return out
    M, K1 = a.shape
    K = K1 + (a2.shape[1] if a2 is not None else 0)
    N = w.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    rc = lib.mvm_linear(_lib.ptr(a), a.stride(0), _lib.ptr(a2), a2.stride(0) if a2 is not None else 0,
                        K1, _lib.ptr(w), w.stride(0), _lib.ptr(bias), _lib.ptr(residual),
                        residual.stride(0) if residual is not None else 0, _lib.ptr(out), N, M, N, K,
                        float(alpha), int(relu), _lib.stream_ptr())
    _lib.check(rc, 'mvm_linear')
    return out


def attention(qkv, batch, n_views, counts, is_cross, tc_passes=0):
    """qkv [batch*n_views, n_pad, 768] (q|k|v, head-contiguous) -> [batch*n_views, n_pad, 256]."""
    lib = _lib.lib()
    V, n_pad, _ = qkv.shape
    out = torch.zeros(V, n_pad, 256, dtype=torch.float32, device=qkv.device)
    cnt = (C.c_int * n_views)(*counts)
    if tc_passes == 'h3':     # fp16x3: the planes the QKV GEMM epilogue writes in the matcher
        k = qkv[:, :, 256:512].contiguous()
        kh = k.half()
        kl = (k - kh.float()).half().reshape(-1, 256).contiguous()
        kh = kh.reshape(-1, 256).contiguous()
        v = qkv[:, :, 512:].contiguous()                      # V stays key-major [rows, 256] (MN-major B operand)
        vh = v.half()
        vl = (v - vh.float()).half().reshape(-1, 256).contiguous()
        vh = vh.reshape(-1, 256).contiguous()
        rc = lib.mvm_attention_h3(_lib.ptr(qkv), _lib.ptr(kh), _lib.ptr(kl), _lib.ptr(vh), _lib.ptr(vl),
                                  _lib.ptr(out), batch, n_views, n_pad, cnt, int(is_cross), _lib.stream_ptr())
        _lib.check(rc, 'mvm_attention_h3')
        return out
    if tc_passes:
        vt = qkv[:, :, 512:].transpose(1, 2).contiguous()      # [V, 256, n_pad]
        klo = vtlo = None
        if tc_passes == 3:      # what the QKV GEMM epilogue does in 3xTF32 mode: rn_tf32 planes + remainders
            def rn(x):          # round to nearest (ties away) on the 13 dropped mantissa bits
                return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
            qkv = qkv.clone()
            k = qkv[:, :, 256:512].contiguous()
            khi = rn(k)
            klo = rn(k - khi).reshape(-1, 256).contiguous()
            qkv[:, :, 256:512] = khi
            vhi = rn(vt)
            vtlo = rn(vt - vhi).contiguous()
            vt = vhi
        rc = lib.mvm_attention_tc(_lib.ptr(qkv), _lib.ptr(vt), _lib.ptr(out), batch, n_views, n_pad, cnt,
                                  int(is_cross), int(tc_passes), _lib.ptr(klo), _lib.ptr(vtlo), _lib.stream_ptr())
        _lib.check(rc, 'mvm_attention_tc')
        return out
    rc = lib.mvm_attention(_lib.ptr(qkv), _lib.ptr(out), batch, n_views, n_pad, cnt, int(is_cross),
                           _lib.stream_ptr())
    _lib.check(rc, 'mvm_attention')
    return out


def log_optimal_transport(scores, alpha, iters, ref_kernel=False, kernel=None):
    """superglue.py:152-172: scores [B,m,n] -> couplings [B,m+1,n+1]."""
    lib = _lib.lib()
    B, m, n = scores.shape
    Z = torch.empty(B, m + 1, n + 1, dtype=torch.float32, device=scores.device)
    Z[:, :m, :n] = scores
    nws = lib.mvm_sinkhorn_workspace_floats(1, B, max(m, n))
    ws = torch.empty(nws, dtype=torch.float32, device=scores.device)
    variants = {'multicta': 1, 'cluster': 2, 'cluster6': 3, 'cluster2': 4}
    if kernel in variants:
        rc = lib.mvm_log_optimal_transport_ex(_lib.ptr(Z), B, m, n, float(alpha), int(iters), _lib.ptr(ws),
                                              variants[kernel], _lib.stream_ptr())
        _lib.check(rc, 'mvm_log_optimal_transport_ex')
        return Z
    fn = {None: lib.mvm_log_optimal_transport, 'ref': lib.mvm_log_optimal_transport_ref,
          'log': lib.mvm_log_optimal_transport_logdomain}['ref' if ref_kernel else kernel]
    rc = fn(_lib.ptr(Z), B, m, n, float(alpha), int(iters), _lib.ptr(ws), _lib.stream_ptr())
    _lib.check(rc, 'mvm_log_optimal_transport')
    return Z


def extract_matches(Z, match_threshold=0.0):
    """multi_view_matcher.py:288-300 on couplings [B,m+1,n+1]."""
    lib = _lib.lib()
    B, m1, n1 = Z.shape
    m, n = m1 - 1, n1 - 1
    dev = Z.device
    m0 = torch.empty(B, m, dtype=torch.int64, device=dev)
    m1_ = torch.empty(B, n, dtype=torch.int64, device=dev)
    s0 = torch.empty(B, m, dtype=torch.float32, device=dev)
    s1 = torch.empty(B, n, dtype=torch.float32, device=dev)
    n_pad = (max(m, n) + 63) // 64 * 64
    ws = torch.empty(3 * B * n_pad, dtype=torch.int32, device=dev)
    rc = lib.mvm_extract_matches(_lib.ptr(Z.contiguous()), B, m, n, float(match_threshold), _lib.ptr(m0),
                                 _lib.ptr(m1_), _lib.ptr(s0), _lib.ptr(s1), _lib.ptr(ws), _lib.stream_ptr())
    _lib.check(rc, 'mvm_extract_matches')
    return m0, m1_, s0, s1
