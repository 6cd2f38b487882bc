"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): ragged multi-view matcher in the default math
mode + the pose stage.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from tests.util import load_case, case_inputs
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
from e2e_multi_view_matching_b200.pipeline import MultiViewPipeline

meta, ref = load_case('mv4_ragged_sharp')
sd, data = case_inputs(meta)
model = MultiViewMatcher({'multi_frame_matching': True, 'GNN_layers': meta['layers'][:4], 'conf_mlp': True}).eval()
sd = {k: v for k, v in sd.items() if not any(k.startswith('gnn.layers.%d.' % l) for l in range(4, 64))}
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
model = model.cuda()
tdata = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data.items()}
T = len(meta['counts'])
K = torch.tensor([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1.0]])[None]
for i in range(T):
    tdata['intr%d' % i] = K
pipe = MultiViewPipeline(model)
res, pose = pipe(tdata)
torch.cuda.synchronize()
print('sanitize smoke ok', {k: tuple(v.shape) for k, v in pose.items() if hasattr(v, 'shape')}.get('extrinsics'))
# training-side kernels: ground-truth matches on a tiny scene
from e2e_multi_view_matching_b200.training import compute_gt_matches_of_image_pair
g = torch.Generator().manual_seed(0)
kp = torch.stack([torch.randint(0, 64, (2, 50), generator=g), torch.randint(0, 48, (2, 50), generator=g)], -1).float().cuda()
Kk = torch.eye(4)[None].repeat(2, 1, 1); Kk[:, 0, 0] = Kk[:, 1, 1] = 60.0; Kk[:, 0, 2] = 31.5; Kk[:, 1, 2] = 23.5
Tt = torch.eye(4)[None].repeat(2, 1, 1); Tt[:, 0, 3] = 0.05
dep = (torch.rand(2, 48, 64, generator=g) * 2 + 1).cuda()
idx, w = compute_gt_matches_of_image_pair(kp, kp.flip(1), Kk.cuda(), Kk.cuda(), Tt.cuda(), dep, dep, 5.0, 15.0)
torch.cuda.synchronize()
print('gt matches ok', tuple(idx.shape), int((idx[:, 0, :-1] >= 0).sum()))
