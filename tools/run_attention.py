import sys, torch
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200 import ops
mode = sys.argv[1] if len(sys.argv) > 1 else 'h3'
mode = mode if mode == 'h3' else int(mode)
g = torch.Generator().manual_seed(0)
B, T, N = 4, 5, 1024
qkv = torch.randn(B * T, N, 768, generator=g).cuda()
for cross in (0, 1):
    for _ in range(2):
        o = ops.attention(qkv, B, T, [N] * T, cross, tc_passes=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        o = ops.attention(qkv, B, T, [N] * T, cross, tc_passes=mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 4 * N * (N * (T - 1) if cross else N) * 256 * B * T
    print('mode %s cross %d: %.3f ms  %.1f TFLOP/s algorithmic' % (mode, cross, ms, fl / ms / 1e9))
