"""Does the pose stage of one step fit under the Sinkhorn launch of the next?  The production Sinkhorn keeps 7 x 16-CTA
clusters resident (112 of 148 SMs); this probe times (a) one 140-problem Sinkhorn launch, (b) one pose stage, (c) both
on two streams, with CUDA events."""
import sys, ctypes as C
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from e2e_multi_view_matching_b200 import _lib
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
from e2e_multi_view_matching_b200.pipeline import MultiViewPipeline

lib = _lib.lib()
cfg = bench.CONFIGS['cfg3']
dev = torch.device('cuda:0')
B = cfg['batch']
sd = bench.make_weights(cfg)
model = MultiViewMatcher({'GNN_layers': cfg['layers'], 'multi_frame_matching': True}).eval()
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
model = model.to(dev)
pipe = MultiViewPipeline(model)
data_np = bench.make_inputs(cfg, cfg['seed_base'], B)
data = {k: torch.from_numpy(v).to(dev) for k, v in data_np.items() if isinstance(v, np.ndarray) and not k.startswith(('image', 'landmark'))}
data.update({k: torch.empty(v.shape, device='meta') for k, v in data_np.items() if k.startswith('image')})
data['ids'] = data_np['ids']
res, pose = pipe(data)
state = pipe.matcher._engine.last
intr = [data['intr%d' % i] for i in state['view_ids']]
torch.cuda.synchronize()

# a 140-problem Sinkhorn launch on scratch copies of the score matrices (timing does not depend on the values)
P = len(state['pair_ids'])
Z = torch.randn(B * P, 1025, 1025, device=dev) * 3
nws = lib.mvm_sinkhorn_workspace_floats(1, B * P, 1024)
ws = torch.empty(nws, dtype=torch.float32, device=dev)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()


def sink():
    rc = lib.mvm_log_optimal_transport(_lib.ptr(Z), B * P, 1024, 1024, 1.0, 100, _lib.ptr(ws), C.c_void_p(sA.cuda_stream))
    assert rc == 0


def pose_stage():
    with torch.cuda.stream(sB):
        return pipe.pose.run(state, intr, global_ba=True)


def timed(fa, fb, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sA)
        sB.wait_event(e0)
        if fa: fa()
        if fb: fb()
        eb = torch.cuda.Event(); eb.record(sB)
        sA.wait_event(eb)
        e1.record(sA)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return np.median(ts)


for _ in range(2):
    sink(); pose_stage()
torch.cuda.synchronize()
a, b, c = timed(sink, None), timed(None, pose_stage), timed(sink, pose_stage)
print('Sinkhorn alone %.2f ms | pose stage alone %.2f ms | both on two streams %.2f ms (sum %.2f)' % (a, b, c, a + b))
