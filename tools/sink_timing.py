"""Phase trace of the cluster Sinkhorn (thread 0 of CTA 0, clock() deltas summed over the 100 iterations; the timing
instance exists for the 6-register-row variant): where an iteration's cycles go."""
import sys, ctypes, torch
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200 import ops, _lib
lib = _lib.lib()
t = torch.zeros(8, dtype=torch.int64, device='cuda')
lib.mvm_debug_set_sinkhorn_timing.argtypes = [ctypes.c_void_p]
lib.mvm_debug_set_sinkhorn_timing(ctypes.c_void_p(t.data_ptr()))
g = torch.Generator().manual_seed(0)
names = ['row pass', 'absorb check + barrier', 'a_i + column pass', 'barrier + push partials', 'wait partials + merge + broadcast', 'wait b + barrier']
for B in (7, 1):
    s = (torch.randn(B, 1024, 1024, generator=g) * 4).cuda()
    for _ in range(2):
        Z = ops.log_optimal_transport(s, 1.0, 100, kernel='cluster2')
    torch.cuda.synchronize()
    v = t[:6].tolist()
    print('B=%d  total %.0f clk/iter' % (B, sum(v) / 100))
    for n, x in zip(names, v):
        print('    %-36s %6.0f clk/iter' % (n, x / 100))
lib.mvm_debug_set_sinkhorn_timing(ctypes.c_void_p(0))
