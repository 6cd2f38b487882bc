"""Multi-view benchmark driver with the reference's flow (eval_multi_view.py:89-165): matcher on
fixed 5-tuples, then eval_bundle_adjust, then pose / translation / rotation AUC@5/10/20 written as
JSON (eval_multi_view.py:70-87).  The reference reads ScanNet/Matterport/MegaDepth tuples and a
trained checkpoint; neither exists offline, so tuples are synthetic scenes
(synthetic.make_scene_tuple_inputs) and weights are either a reference checkpoint given with --ckpt
(helpers.load_ckpt format: {'model': state_dict with 'module.' prefix}) or seeded random weights.

    python -m e2e_multi_view_matching_b200.eval_multi_view --n_tuples 16 --out result.json
"""
import argparse
import json

import numpy as np
import torch

from .models.multi_view_matcher import MultiViewMatcher
from .pipeline import MultiViewPipeline, pose_auc
from .synthetic import make_state_dict, make_scene_tuple_inputs


def write_result(pose_errors, file):
    thresholds = [5, 10, 20]
    metrics = dict()
    for name, errs in (('pose', pose_errors[0]), ('transl', pose_errors[1]), ('rot', pose_errors[2])):
        for thresh, auc in zip(thresholds, pose_auc(errs, thresholds)):
            metrics["{}_AUC@{}deg".format(name, thresh)] = auc * 100.0
    if file:
        with open(file, 'w') as tf:
            json.dump(metrics, tf, indent=4)
    return metrics


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--n_tuples', type=int, default=8)
    ap.add_argument('--tuple_size', type=int, default=5)
    ap.add_argument('--max_keypoints', type=int, default=1024)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--dataset', default='scannet', choices=['scannet', 'matterport', 'megadepth'])
    ap.add_argument('--ckpt', default=None)
    ap.add_argument('--seed', type=int, default=0)
    ap.add_argument('--math_mode', type=int, default=3)
    ap.add_argument('--out', default=None)
    opt = ap.parse_args(argv)
    import e2e_multi_view_matching_b200 as pkg
    pkg.set_math_mode(opt.math_mode)
    # GNN depth per dataset (train.py:262-268)
    layers = ['self', 'cross'] * 9 if opt.dataset == 'megadepth' else (['self'] + ['cross'] * 3) * 7
    matcher = MultiViewMatcher({'multi_frame_matching': True, 'GNN_layers': layers}).eval()
    if opt.ckpt:
        sd = torch.load(opt.ckpt, map_location='cpu')
        sd = sd.get('model', sd)
        # the reference loads with strict=False (helpers.py:48), which hides key mismatches: load the same way,
        # but say what did not line up
        missing, unexpected = matcher.load_state_dict({k[7:] if k.startswith('module.') else k: v for k, v in sd.items()},
                                                      strict=False)
        if missing or unexpected:
            import logging
            logging.warning('checkpoint keys: %d missing (%s...), %d unexpected (%s...)', len(missing),
                            ', '.join(missing[:3]), len(unexpected), ', '.join(unexpected[:3]))
    else:
        sd = make_state_dict(len(layers), seed=opt.seed, final_proj_gain=12.0, conf_head='score')
        matcher.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    matcher = matcher.cuda()
    pipe = MultiViewPipeline(matcher)
    pose_errors = [[], [], []]
    with torch.no_grad():
        for start in range(0, opt.n_tuples, opt.batch):
            b = min(opt.batch, opt.n_tuples - start)
            data = make_scene_tuple_inputs(1000 + start, opt.tuple_size, opt.max_keypoints, batch=b)
            data = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) and not k.startswith('image')
                        else (torch.empty(v.shape, device='meta') if isinstance(v, np.ndarray) else v))
                    for k, v in data.items()}
            _, pose = pipe(data)
            for e in MultiViewPipeline.pair_errors(data, pose, opt.tuple_size):
                for i in range(3):
                    pose_errors[i].append(e[i])
    metrics = write_result(pose_errors, opt.out)
    print(json.dumps(metrics))
    return metrics


if __name__ == '__main__':
    main()
