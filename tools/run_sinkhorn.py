import sys, torch
sys.path.insert(0, '.')
from e2e_multi_view_matching_b200 import ops
g = torch.Generator().manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 7
s = (torch.randn(B, 1024, 1024, generator=g) * 4).cuda()
for _ in range(3):
    Z = ops.log_optimal_transport(s, 1.0, 100)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    Z = ops.log_optimal_transport(s, 1.0, 100)
e1.record(); torch.cuda.synchronize()
print('B=%d problems: %.3f ms per call' % (B, e0.elapsed_time(e1) / 5))
