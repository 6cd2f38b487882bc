"""GPU stage parity: each CUDA stage against the oracle (or fp64 torch math) on seeded inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_linear_epilogues():
    from e2e_multi_view_matching_b200 import ops
    g = torch.Generator().manual_seed(0)
    for (M, N, K1, K2) in [(300, 256, 128, 0), (1024, 768, 256, 0), (130, 512, 256, 256), (64, 256, 512, 0)]:
        a = torch.randn(M, K1, generator=g).cuda()
        a2 = torch.randn(M, K2, generator=g).cuda() if K2 else None
        w = torch.randn(N, K1 + K2, generator=g).cuda() / 16
        b = torch.randn(N, generator=g).cuda()
        r = torch.randn(M, N, generator=g).cuda()
        out = ops.linear(a, w, bias=b, a2=a2, residual=r, relu=True)
        A = torch.cat([a, a2], 1) if a2 is not None else a
        ref = torch.relu(A.double() @ w.double().T + b.double()) + r.double()
        assert (out.double() - ref).abs().max().item() < 2e-4


def test_attention_self_and_cross_ragged():
    from e2e_multi_view_matching_b200 import ops
    from oracle.matcher import attention as oracle_attention
    rng = np.random.default_rng(1)
    B, T, n_pad = 2, 3, 128
    counts = [100, 128, 77]
    qkv = rng.standard_normal((B * T, n_pad, 768)).astype(np.float32)
    for is_cross in (0, 1):
        out = ops.attention(torch.from_numpy(qkv).cuda(), B, T, counts, is_cross).cpu().numpy()
        for b in range(B):
            for t in range(T):
                v = b * T + t
                segs = [s for s in range(T) if (s != t if is_cross else s == t)]
                q = qkv[v, :counts[t], 0:256].reshape(counts[t], 4, 64)
                k = np.concatenate([qkv[b * T + s, :counts[s], 256:512] for s in segs], 0).reshape(-1, 4, 64)
                vv = np.concatenate([qkv[b * T + s, :counts[s], 512:768] for s in segs], 0).reshape(-1, 4, 64)
                # oracle layout [B, d, h, N]
                o = oracle_attention(q.transpose(2, 1, 0)[None], k.transpose(2, 1, 0)[None],
                                     vv.transpose(2, 1, 0)[None])[0]      # [d, h, n]
                ref = o.transpose(2, 1, 0).reshape(counts[t], 256)
                np.testing.assert_allclose(out[v, :counts[t]], ref, atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize('shape', [(1, 60, 50), (3, 128, 128), (2, 33, 200), (1, 300, 257)])
@pytest.mark.parametrize('spread', [1.0, 12.0, 40.0])
def test_sinkhorn_kernels_vs_oracle(shape, spread):
    from e2e_multi_view_matching_b200 import ops
    from oracle.matcher import log_optimal_transport
    B, m, n = shape
    rng = np.random.default_rng(m * 1000 + n)
    s = (rng.standard_normal((B, m, n)) * spread).astype(np.float32)
    ref = log_optimal_transport(s, 1.0, 100)
    for kernel in ('ref', 'log', None, 'cluster', 'cluster6', 'cluster2', 'multicta'):
        Z = ops.log_optimal_transport(torch.from_numpy(s).cuda(), 1.0, 100, kernel=kernel).cpu().numpy()
        err = np.abs(Z - ref)
        lim = (1e-4 if spread < 20 else 6e-4) + 1e-5 * np.abs(ref)
        assert (err <= lim).all(), (kernel, float(err.max()), float((err - lim).max()))


def test_sinkhorn_full_size_marginals():
    """1024 x 1024 (BASELINE size): size-independent property -- row/col marginals of exp(Z)."""
    from e2e_multi_view_matching_b200 import ops
    g = torch.Generator().manual_seed(3)
    m = n = 1024
    s = (torch.randn(2, m, n, generator=g) * 4).cuda()
    Z = ops.log_optimal_transport(s, 1.0, 100)
    P = torch.exp(Z.double()) / (m + n)
    assert (P[:, :m, :].sum(2) * (m + n) - 1).abs().max().item() < 5e-3
    assert (P[:, :, :n].sum(1) * (m + n) - 1).abs().max().item() < 5e-3
    Zr = ops.log_optimal_transport(s, 1.0, 100, ref_kernel=True)
    assert (Z - Zr).abs().max().item() < 2e-4


def test_extract_matches_vs_oracle():
    from e2e_multi_view_matching_b200 import ops
    from oracle.matcher import extract_matches
    rng = np.random.default_rng(5)
    for (B, m, n) in [(2, 70, 90), (1, 128, 128), (1, 257, 64)]:
        Z = rng.standard_normal((B, m + 1, n + 1)).astype(np.float32)
        for thr in (0.0, 0.2):
            r = extract_matches(Z, thr)
            g = ops.extract_matches(torch.from_numpy(Z).cuda(), thr)
            assert np.array_equal(r[0], g[0].cpu().numpy())
            assert np.array_equal(r[1], g[1].cpu().numpy())
            np.testing.assert_allclose(g[2].cpu().numpy(), r[2], rtol=1e-5)
            np.testing.assert_allclose(g[3].cpu().numpy(), r[3], rtol=1e-5)


def test_pack_views_matches_slice_copies():
    """mvm_pack_views == the per-view zero-padded slice copies (multi_view_matcher.py:229-262 layout), ragged views."""
    import ctypes as C
    from e2e_multi_view_matching_b200 import _lib
    lib = _lib.lib()
    B, counts, n_pad = 3, [100, 192, 1, 77], 192
    T = len(counts)
    g = torch.Generator().manual_seed(5)
    views = [(torch.randn(B, n, 2, generator=g).cuda(), torch.rand(B, n, generator=g).cuda(),
              torch.randn(B, 256, n, generator=g).cuda()) for n in counts]
    kp = torch.full((B, T, n_pad, 2), 7.0, device='cuda')           # stale contents must be overwritten by the padding
    sc = torch.full((B, T, n_pad), 7.0, device='cuda')
    de = torch.full((B, T, 256, n_pad), 7.0, device='cuda')
    ptrs = [(C.c_void_p * T)(*[v[i].data_ptr() for v in views]) for i in range(3)]
    rc = lib.mvm_pack_views(ptrs[0], ptrs[1], ptrs[2], (C.c_int * T)(*counts), B, T, n_pad, _lib.ptr(kp), _lib.ptr(sc),
                            _lib.ptr(de), _lib.stream_ptr())
    assert rc == 0
    for t, (k, s, d) in enumerate(views):
        n = counts[t]
        assert torch.equal(kp[:, t, :n], k) and torch.equal(sc[:, t, :n], s) and torch.equal(de[:, t, :, :n], d)
        assert kp[:, t, n:].abs().max().item() == 0 if n < n_pad else True
        assert sc[:, t, n:].abs().sum().item() == 0 and de[:, t, :, n:].abs().sum().item() == 0
    assert lib.mvm_pack_views(ptrs[0], ptrs[1], ptrs[2], (C.c_int * T)(*[n_pad + 1] * T), B, T, n_pad, _lib.ptr(kp),
                              _lib.ptr(sc), _lib.ptr(de), _lib.stream_ptr()) != 0      # a view longer than n_pad is refused
