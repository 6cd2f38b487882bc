"""Capture one matcher forward into a CUDA graph, replay it, compare with the eager outputs."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from tests.util import load_case, case_inputs
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher

meta, ref = load_case('mv4_ragged_sharp')
sd, data = case_inputs(meta)
model = MultiViewMatcher({'multi_frame_matching': True, 'GNN_layers': meta['layers'], 'conf_mlp': True}).eval()
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
model = model.cuda()
tdata = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data.items()}
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        eager = model(tdata)
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
eager = {k: v.clone() for k, v in eager.items() if v is not None}
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = model(tdata)
for _ in range(2):
    for v in out.values():
        if v is not None:
            v.zero_()
    g.replay()
torch.cuda.synchronize()
bad = [k for k in eager if not torch.equal(eager[k], out[k])]
print('graph capture ok' if not bad else 'MISMATCH %s' % bad, len(eager), 'outputs')
