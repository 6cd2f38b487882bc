"""Where do the engine's and the oracle's pose errors differ?  Per tuple: max |err_engine - err_oracle| over the 10 pairs,
global-BA iterations / termination of both, match differences."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench
from oracle import pipeline as OP, mvba as M
from oracle.matcher_torch import matcher_forward
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
from e2e_multi_view_matching_b200.pipeline import MultiViewPipeline

cfg = bench.CONFIGS['cfg3']
sd = bench.make_weights(cfg)
model = MultiViewMatcher({'GNN_layers': cfg['layers'], 'multi_frame_matching': True}).eval()
model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
model = model.cuda()
n_units, N = 32, cfg['parity_kpts']
data = bench.make_inputs(cfg, cfg['seed_base'] + 500, n_units, kpts=N)
tdata = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) and not k.startswith('image') else
             (torch.empty(v.shape, device='meta') if isinstance(v, np.ndarray) else v)) for k, v in data.items()}
res, pose = MultiViewPipeline(model)(tdata)
eng = np.array([e[0] for e in MultiViewPipeline.pair_errors(tdata, pose, 5)]).reshape(n_units, 10)
its = pose['ba_iterations'].cpu().numpy()
cost = pose['ba_cost'].cpu().numpy()
torch.set_num_threads(bench.cpu_threads())
for b in range(n_units):
    one = OP._one(data, b)
    r = matcher_forward(sd, {'GNN_layers': cfg['layers'], 'multi_frame_matching': True}, one)
    scene = {'kpts': [one['keypoints%d' % i][0] for i in range(5)], 'K': one['intr0'][0][:3, :3], 'matches': {}, 'conf': {}}
    flips = 0
    for i1 in range(5):
        for i0 in range(i1):
            scene['matches'][(i0, i1)] = r['matches%d_%d_%d' % (i0, i0, i1)][0]
            scene['conf'][(i0, i1)] = r['conf_scores_%d_%d' % (i0, i1)][0, :, 0]
            flips += int((res['matches%d_%d_%d' % (i0, i0, i1)][b].cpu().numpy() != scene['matches'][(i0, i1)]).sum())
    out = M.multi_view_pipeline(scene)
    ora = np.array([e[0] for e in OP.tuple_errors(sd, cfg['layers'], data, b)])
    d = np.abs(eng[b] - ora)
    tree_d = np.abs(pose['extrinsics_tree'][b].cpu().numpy() - out['extr_tree']).max()
    init_d = np.abs(pose['extrinsics_init'][b].cpu().numpy() - out['extr_init']).max()
    print('tuple %2d: max |d err| %.4f deg  match flips %3d  BA its engine %2d oracle %2d (%s)  cost %.3e / %.3e  tree diff %.1e init diff %.1e'
          % (b, d.max(), flips, its[b], out['info']['iterations'], out['info']['termination'], cost[b, 1], out['info']['final_cost'],
             tree_d, init_d), flush=True)
