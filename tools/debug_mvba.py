import sys
sys.path.insert(0, '.')
import numpy as np, torch
from oracle import mvba as M
from tests.test_mv_gpu import _state_from_scene
from e2e_multi_view_matching_b200.pose_optimization.multi_view.pose_engine import MultiViewPoseEngine
for (T, n, outl) in [(3, 60, 0.0), (5, 100, 0.1)]:
    scenes = [M.make_multi_view_scene(s, T, n, outlier_frac=outl) for s in (1, 2)]
    state = _state_from_scene(scenes)
    K = torch.from_numpy(scenes[0]['K'])[None].repeat(len(scenes), 1, 1)
    for mi in (1, 2, 3, 5, 10, 50):
        out = MultiViewPoseEngine(max_iterations_ba=mi).run(state, [K] * T)
        torch.cuda.synchronize()
        for b, sc in enumerate(scenes[:1]):
            ref = M.multi_view_pipeline(sc, max_iterations=mi)
            d = np.abs(out['extrinsics'][b].cpu().numpy() - ref['extr']).max()
            print(T, 'max_it', mi, 'ours it', int(out['ba_iterations'][b]), 'cost', out['ba_cost'][b].cpu().numpy(),
                  '| ref', ref['info']['iterations'], ref['info']['initial_cost'], ref['info']['final_cost'], ref['info']['termination'], '| max|dE| %.2e' % d)
