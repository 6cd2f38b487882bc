"""clock64 trace of one attention CTA (debug hook mvm_debug_set_attention_timing): where a key tile's time goes."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200 import ops, _lib
lib = _lib.lib()
mode = sys.argv[1] if len(sys.argv) > 1 else 'h3'        # 'h3' (fp16x3), 3 (tf32x3) or 1 (single-pass tf32)
mode = mode if mode == 'h3' else int(mode)
buf = torch.zeros(64 * 16 + 2048 * 8 + 64 * 4, dtype=torch.int64, device='cuda')
lib.mvm_debug_set_attention_timing.argtypes = [ctypes.c_void_p]
g = torch.Generator().manual_seed(0)
B, T, N = 8, 5, 1024
qkv = torch.randn(B * T, N, 768, generator=g).cuda()
for _ in range(2):
    ops.attention(qkv, B, T, [N] * T, 1, tc_passes=mode)
lib.mvm_debug_set_attention_timing(ctypes.c_void_p(buf.data_ptr()))
ops.attention(qkv, B, T, [N] * T, 1, tc_passes=mode)
torch.cuda.synchronize()
lib.mvm_debug_set_attention_timing(ctypes.c_void_p(0))
t = buf[:64 * 16].view(64, 16).cpu().numpy()
c = buf[64 * 16:64 * 16 + 2048 * 8].view(2048, 8).cpu().numpy()
pr = buf[64 * 16 + 2048 * 8:].view(64, 4).cpu().numpy()
t0 = t[0, 8]
names = ['sm:wait S', 'sm:S ready', 'sm:ld done', 'sm:max done', 'sm:P half a', 'sm:P half b', 'S:issued', 'S:committed',
         'S:iter start', 'S:K+buffer ready', 'PV:P a ready', 'PV:a issued', 'PV:P b ready', 'PV:b issued', 'PV:commit1', 'PV:commit2']
print('tile ' + ' '.join('%16s' % n for n in names))
for j in range(20, 28):
    print('%4d ' % j + ' '.join('%16d' % (t[j, k] - t0) for k, n in enumerate(names)))
print('period per tile: %.0f clk' % ((t[40, 8] - t[20, 8]) / 20))
if mode == 'h3' and pr.any():
    print('producer (relative to the same origin): tile, K slot free, K requested, V slot free, V requested | S warp: iter start, K landed')
    for j in list(range(0, 10)) + list(range(20, 28)):
        print('%4d ' % j + ' '.join('%9d' % (pr[j, k] - t0) for k in range(4)) + ' | %9d %9d' % (t[j, 8] - t0, t[j, 9] - t0))
for a, b, label in ((1, 2, 'S ready -> ld done'), (2, 3, 'ld -> max/vote'), (3, 4, 'exp half a + st + arrive'), (4, 5, 'exp half b + st + arrive'),
                    (0, 1, 'softmax warp waits for S'), (8, 9, 'S warp: waits (K, buffer)'), (9, 6, 'S warp: 24 UMMAs issue'), (6, 7, 'S warp: two commits'),
                    (10, 11, 'PV warp: 12 UMMAs a'), (11, 12, 'PV warp: wait P b'), (12, 13, 'PV warp: 12 UMMAs b'), (13, 15, 'PV warp: two commits')):
    d = [t[j, b] - t[j, a] for j in range(16, 48)]
    print('%-30s avg %.0f clk' % (label, sum(d) / len(d)))
d = [t[j + 1, 10] - t[j, 15] for j in range(16, 47)]
print('%-30s avg %.0f clk' % ('PV warp: wait P a (next tile)', sum(d) / len(d)))

# ---- CTA level: per-phase durations and the gap between consecutive CTAs on one SM
import numpy as np
c = c[:1280]
lab = ['setup (barriers, TMEM alloc, sync)', 'Q -> TMEM', 'first S read', 'key-tile loop', 'merge + store + sync', 'dealloc']
for k, l in enumerate(lab):
    d = c[:, k + 2] - c[:, k + 1]
    print('%-36s avg %7.0f clk  (min %d max %d)' % (l, d.mean(), d.min(), d.max()))
print('CTA lifetime                         avg %7.0f clk' % (c[:, 7] - c[:, 1]).mean())
gaps = []
for sm in np.unique(c[:, 0]):
    rows = c[c[:, 0] == sm]
    rows = rows[np.argsort(rows[:, 1])]
    gaps += list(rows[1:, 1] - rows[:-1, 7])
print('gap between CTAs on one SM           avg %7.0f clk (n=%d)' % (np.mean(gaps), len(gaps)))
