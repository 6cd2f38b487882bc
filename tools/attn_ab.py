"""A/B of the two fp16-plane attention kernels on the bench shape (14 tuples x 5 views x 1024 keypoints):
variant 1 = one softmax group, two CTAs per SM; variant 0 = two softmax groups, one CTA per SM.
CUDA-event time per launch (L2 flushed between launches) + CTA-level clock trace of the two-CTA kernel."""
import sys, ctypes, torch
import numpy as np
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200 import ops, _lib
lib = _lib.lib()
B, T, N = 14, 5, 1024
g = torch.Generator().manual_seed(0)
qkv = torch.randn(B * T, N, 768, generator=g).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
k = qkv[:, :, 256:512].contiguous()
kh = k.half()
kl = (k - kh.float()).half().reshape(-1, 256).contiguous()
kh = kh.reshape(-1, 256).contiguous()
v = qkv[:, :, 512:].contiguous()
vth = v.half().reshape(-1, 256).contiguous()
vtl = (v - vth.reshape(v.shape).float()).half().reshape(-1, 256).contiguous()
cnt = (ctypes.c_int * T)(*([N] * T))
out_buf = torch.zeros(B * T, N, 256, device='cuda')


def run(is_cross):
    rc = lib.mvm_attention_h3(_lib.ptr(qkv), _lib.ptr(kh), _lib.ptr(kl), _lib.ptr(vth), _lib.ptr(vtl), _lib.ptr(out_buf),
                              B, T, N, cnt, int(is_cross), _lib.stream_ptr())
    assert rc == 0
    return out_buf


outs = {}
for variant in (1, 0):
    lib.mvm_debug_set_attention_h3_variant(variant)
    for is_cross in (0, 1):
        for _ in range(3):
            out = run(is_cross)
        outs[(variant, is_cross)] = out.clone()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(is_cross)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        flops = 4.0 * B * T * 4 * N * (N if not is_cross else (T - 1) * N) * 64
        print('variant %d %s: median %.1f us  min %.1f us  (%.0f TFLOP/s algorithmic)'
              % (variant, 'cross' if is_cross else 'self ', np.median(ts), min(ts), flops / np.median(ts) / 1e6))
for is_cross in (0, 1):
    d = (outs[(1, is_cross)] - outs[(0, is_cross)]).abs().max().item()
    print('max |variant 1 - variant 0| %s: %.3g' % ('cross' if is_cross else 'self', d))
lib.mvm_debug_set_attention_h3_variant(1)
buf = torch.zeros(64 * 16 + 2048 * 8 + 64 * 4, dtype=torch.int64, device='cuda')
lib.mvm_debug_set_attention_timing.argtypes = [ctypes.c_void_p]
for is_cross in (0, 1):
    buf.zero_()
    lib.mvm_debug_set_attention_timing(ctypes.c_void_p(buf.data_ptr()))
    run(is_cross)
    torch.cuda.synchronize()
    lib.mvm_debug_set_attention_timing(ctypes.c_void_p(0))
    c = buf[64 * 16:64 * 16 + 2048 * 8].view(2048, 8).cpu().numpy()
    c = c[c[:, 1] > 0]
    print('two-CTA kernel, %s layer, %d CTAs traced' % ('cross' if is_cross else 'self', len(c)))
    lab = ['setup (barriers, TMEM alloc, sync)', 'Q -> TMEM', 'first S read', 'key-tile loop', 'normalise + store + sync', 'dealloc']
    for k, l in enumerate(lab):
        d = c[:, k + 2] - c[:, k + 1]
        print('  %-36s avg %7.0f clk  (min %d max %d)' % (l, d.mean(), d.min(), d.max()))
    print('  CTA lifetime                         avg %7.0f clk' % (c[:, 7] - c[:, 1]).mean())
