"""Golden scalars for the validation pass: the reference's OWN helpers.run_matcher
(helpers.py:243-260) -- reference MultiViewMatcher in eval mode, compute_match_loss, run_weighted_8_point with
choose_closest, rotation / translation losses -- imported unmodified through oracle/ref_shim.py and run on CPU on
seeded synthetic tuples -> tests/golden/validation_*.npz.  TEST INFRASTRUCTURE ONLY (needs /root/reference).

The ground-truth assignments come from the scene's landmark ids (two keypoints match iff they observe the same
landmark), weighted like helpers.py:190-213; the GPU test rebuilds the same inputs from the seeds."""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
CASES = [dict(name='mv3_128', views=3, kpts=128, batch=2, layers=(['self'] + ['cross'] * 2) * 2, wseed=21, iseed=31, gain=12.0),
         dict(name='pair_192', views=2, kpts=192, batch=3, layers=['self', 'cross'] * 3, wseed=22, iseed=32, gain=12.0)]


def gt_from_landmarks(la, lb):
    """indices [2, n+1] int64, weights [2, n+1] float32 for one pair of one batch item."""
    n = len(la)
    pos = {int(l): j for j, l in enumerate(lb)}
    i0 = np.array([pos.get(int(l), -1) for l in la] + [-1], np.int64)
    i1 = np.full(n + 1, -1, np.int64)
    for i, j in enumerate(i0[:n]):
        if j >= 0:
            i1[j] = i
    m = int((i0 >= 0).sum())
    mw = np.float32(2.0 * m) / np.float32(2.0 * n)
    uw = np.float32(0.5) / (np.float32(1.0) - mw)
    mw = np.float32(0.5) / mw
    w0 = np.where(i0 >= 0, mw, uw).astype(np.float32)
    w1 = np.where(i1 >= 0, mw, uw).astype(np.float32)
    return np.stack([i0, i1]), np.stack([w0, w1])


def build(case):
    from e2e_multi_view_matching_b200.synthetic import make_scene_tuple_inputs, make_state_dict
    data = make_scene_tuple_inputs(case['iseed'], n_views=case['views'], n_kpts=case['kpts'], batch=case['batch'])
    for b_ in range(case['views']):
        for a_ in range(b_):
            pairs = [gt_from_landmarks(data['landmark%d' % a_][b], data['landmark%d' % b_][b]) for b in range(case['batch'])]
            data['gt_indices_%d_%d' % (a_, b_)] = np.stack([p[0] for p in pairs])
            data['gt_weights_%d_%d' % (a_, b_)] = np.stack([p[1] for p in pairs])
    sd = make_state_dict(len(case['layers']), seed=case['wseed'], final_proj_gain=case['gain'], conf_head='score')
    return data, sd


def main():
    ref_shim.load()
    import helpers
    from models.models.multi_view_matcher import MultiViewMatcher
    assert helpers.__file__.startswith('/root/reference')
    torch.set_num_threads(8)
    report = {}
    for case in CASES:
        data_np, sd = build(case)
        model = MultiViewMatcher({'multi_frame_matching': case['views'] > 2, 'GNN_layers': case['layers'], 'conf_mlp': True}).eval()
        model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        matcher = type('W', (), {'module': model, '__call__': lambda self, d: model(d)})()
        out = {}
        for dtype, tag in ((torch.float32, 'f32'), (torch.float64, 'f64')):
            model.to(dtype)
            torch.set_default_dtype(dtype)          # the reference creates identity matrices etc. in the default dtype
            data = {k: (torch.from_numpy(v).to(dtype) if isinstance(v, np.ndarray) and v.dtype == np.float32 else
                        (torch.from_numpy(v) if isinstance(v, np.ndarray) else v)) for k, v in data_np.items()}
            opt = types.SimpleNamespace(pose_loss=True)
            with torch.no_grad():
                losses, result = helpers.run_matcher(opt, data, matcher)
            out[tag] = {k: float(v) for k, v in losses.items()}
        model.float()
        torch.set_default_dtype(torch.float32)
        print(case['name'], out)
        report[case['name']] = out
        np.savez(os.path.join(OUT, 'validation_%s.npz' % case['name']), meta=json.dumps(case),
                 **{'%s_%s' % (k, tag): np.float64(v) for tag, d in out.items() for k, v in d.items()})
    json.dump(report, open(os.path.join(OUT, 'validation_report.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
