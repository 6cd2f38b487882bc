"""GPU: tcgen05 tensor-core kernels against fp64 math and the fp32 CUDA-core kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape', [(128, 128, 32, 0), (256, 256, 128, 0), (1024, 768, 256, 0), (320, 512, 256, 256),
                                   (5120, 256, 512, 0), (192, 256, 256, 0)])
@pytest.mark.parametrize('passes', [3, 1])
def test_gemm_tc_vs_fp64(shape, passes):
    from e2e_multi_view_matching_b200 import ops
    M, N, K1, K2 = shape
    g = torch.Generator().manual_seed(M + N + K1)
    a = torch.randn(M, K1, generator=g).cuda()
    a2 = torch.randn(M, K2, generator=g).cuda() if K2 else None
    w = (torch.randn(N, K1 + K2, generator=g) / 16).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).cuda()
    out = ops.linear(a, w, bias=b, a2=a2, residual=r, relu=True, tc_passes=passes)
    torch.cuda.synchronize()
    A = torch.cat([a, a2], 1) if a2 is not None else a
    ref = torch.relu(A.double() @ w.double().T + b.double()) + r.double()
    err = (out.double() - ref).abs().max().item()
    tol = 3e-5 if passes == 3 else 2e-2
    assert err < tol, (shape, passes, err)
    if passes == 3:
        simt = ops.linear(a, w, bias=b, a2=a2, residual=r, relu=True)
        assert (simt - out).abs().max().item() < 3e-5
