import sys, ctypes, torch
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200 import ops, _lib
lib = _lib.lib()
t = torch.zeros(8, dtype=torch.int64, device='cuda')
lib.mvm_debug_set_sinkhorn_timing.argtypes = [ctypes.c_void_p]
lib.mvm_debug_set_sinkhorn_timing(ctypes.c_void_p(t.data_ptr()))
g = torch.Generator().manual_seed(0)
for B in (7, 1):
    s = (torch.randn(B, 1024, 1024, generator=g) * 4).cuda()
    for _ in range(2):
        Z = ops.log_optimal_transport(s, 1.0, 100)
    torch.cuda.synchronize()
    names = ['row pass', 'col pass', 'barrier 1', 'merge', 'barrier 2']
    print('B=%d' % B, {n: '%.2f us/iter' % (v / 100 / 1965.0) for n, v in zip(names, t[:5].tolist())})
