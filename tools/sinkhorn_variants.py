"""Parity + timing of the Sinkhorn kernel variants on the GPU (multi-CTA / cluster with 8 or 6 register rows).
python tools/sinkhorn_variants.py [n_problems]  -> prints max |Z - oracle| per shape and ms per launch."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from e2e_multi_view_matching_b200 import ops, _lib  # noqa: E402
from oracle.matcher import log_optimal_transport  # noqa: E402

lib = _lib.lib()
print('max active clusters 1024x1024:', lib.mvm_sinkhorn_max_active_clusters(1024, 1024),
      ' 512x512:', lib.mvm_sinkhorn_max_active_clusters(512, 512),
      ' 128x128:', lib.mvm_sinkhorn_max_active_clusters(128, 128))

for (B, m, n, spread) in [(2, 60, 50, 1.0), (1, 300, 257, 12.0), (2, 128, 128, 40.0), (1, 513, 1000, 4.0),
                          (1, 1024, 1024, 4.0), (1, 1000, 1024, 30.0), (1, 777, 650, 12.0)]:
    rng = np.random.default_rng(m * 1000 + n)
    s = (rng.standard_normal((B, m, n)) * spread).astype(np.float32)
    t0 = time.time()
    ref = log_optimal_transport(s, 1.0, 100)
    t_or = time.time() - t0
    line = '%dx%dx%d spread %g (oracle %.1fs):' % (B, m, n, spread, t_or)
    for kernel in ('multicta', 'cluster', 'cluster6', 'cluster2'):
        Z = ops.log_optimal_transport(torch.from_numpy(s).cuda(), 1.0, 100, kernel=kernel).cpu().numpy()
        err = np.abs(Z - ref)
        line += '  %s %.2e (rel-excess %.2e)' % (kernel, err.max(), (err - 1e-5 * np.abs(ref)).max())
    print(line, flush=True)

NP = int(sys.argv[1]) if len(sys.argv) > 1 else 140
g = torch.Generator().manual_seed(0)
s = (torch.randn(NP, 1024, 1024, generator=g) * 4).cuda()
for kernel in ('multicta', 'cluster', 'cluster6', 'cluster2'):
    for _ in range(2):
        ops.log_optimal_transport(s, 1.0, 100, kernel=kernel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.log_optimal_transport(s, 1.0, 100, kernel=kernel)
    e1.record()
    torch.cuda.synchronize()
    # the wrapper also copies the scores into the [m+1, n+1] buffer: time that alone and subtract
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(3):
        Z = torch.empty(NP, 1025, 1025, dtype=torch.float32, device='cuda')
        Z[:, :1024, :1024] = s
    f1.record()
    torch.cuda.synchronize()
    print('%s: %d problems 1024x1024, 100 iterations: %.3f ms per launch (copy-in %.3f ms subtracted)'
          % (kernel, NP, (e0.elapsed_time(e1) - f0.elapsed_time(f1)) / 3, f0.elapsed_time(f1) / 3), flush=True)

# phase breakdown of the cluster kernel (TIMING instance, thread 0 of CTA 0)
import ctypes
t = torch.zeros(8, dtype=torch.int64, device='cuda')
lib.mvm_debug_set_sinkhorn_timing.argtypes = [ctypes.c_void_p]
lib.mvm_debug_set_sinkhorn_timing(ctypes.c_void_p(t.data_ptr()))
s7 = s[:7].contiguous()
for _ in range(2):
    ops.log_optimal_transport(s7, 1.0, 100, kernel='cluster2')
torch.cuda.synchronize()
names = ['row pass', 'barrier 1 (+col absorb)', 'a + col pass', 'barrier 2 + push', 'cluster sync 1 + merge', 'cluster sync 2']
print('cluster2 phases, cycles per iteration:', {n: round(v / 100) for n, v in zip(names, t[:6].tolist())}, flush=True)
lib.mvm_debug_set_sinkhorn_timing(ctypes.c_void_p(0))
