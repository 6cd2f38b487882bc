"""Phase cycle counters of one mvba CTA on the bench workload (debug hook mvm_debug_set_mvba_timing)."""
import sys, ctypes
sys.path.insert(0, '.')
import numpy as np, torch
from e2e_multi_view_matching_b200 import _lib
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
from e2e_multi_view_matching_b200.pipeline import MultiViewPipeline
from e2e_multi_view_matching_b200.synthetic import make_state_dict, make_scene_tuple_inputs
import bench
lib = _lib.lib()
sd = make_state_dict(len(bench.LAYERS), seed=0, final_proj_gain=bench.GAIN)
model = MultiViewMatcher({'GNN_layers': bench.LAYERS}).eval()
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
model = model.cuda()
pipe = MultiViewPipeline(model)
d = make_scene_tuple_inputs(1000, bench.T_VIEWS, bench.N_KPTS, batch=8)
data = {k: torch.from_numpy(v).cuda() for k, v in d.items() if isinstance(v, np.ndarray) and not k.startswith('image') and not k.startswith('landmark')}
data.update({k: torch.empty(v.shape, device='meta') for k, v in d.items() if k.startswith('image')})
data['ids'] = d['ids']
t = torch.zeros(8, dtype=torch.int64, device='cuda')
lib.mvm_debug_set_mvba_timing.argtypes = [ctypes.c_void_p]
for _ in range(2):
    res, pose = pipe(data)
lib.mvm_debug_set_mvba_timing(ctypes.c_void_p(t.data_ptr()))
res, pose = pipe(data)
torch.cuda.synchronize()
lib.mvm_debug_set_mvba_timing(ctypes.c_void_p(0))
v = t.tolist()
names = ['pass A (points)', 'exchange 116 (barrier)', 'assemble + diag', 'cholesky 24x24', 'pass B (points)', 'exchange 4 (barrier)', 'decision']
it = max(v[7], 1)
print('iterations', v[7], 'n_matches per pair', pose['n_matches'][0].tolist(), 'ba iterations', pose['ba_iterations'].tolist())
for n, c in zip(names, v[:7]):
    print('%-26s %8.2f us/iter' % (n, c / it / 1965.0))
print('%-26s %8.2f us/iter' % ('total', sum(v[:7]) / it / 1965.0))
