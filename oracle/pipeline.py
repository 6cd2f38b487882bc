"""CPU restatement of the reference's two eval loops, end to end.  TEST INFRASTRUCTURE ONLY.

  tuple_errors  ==  eval_multi_view.py:152-162 + eval_bundle_adjust (:21-68): matcher on a tuple, per pair
                    w8pt + two-view BA, spanning tree, rotation averaging + LUD, global BA, pair errors
  pair_error    ==  eval_pairs.py:208-267 for the `w8pt` / `w8pt_ba` modes

Built from the pinned pieces: oracle/matcher_torch.py (== the reference matcher, tests/test_oracle_matcher.py),
oracle/pose.py (== the reference's two-view files, oracle/make_pose_golden.py), oracle/mvba.py + ba_init.py
(gtest known answers).  Used by the AUC-parity test and by bench.py's `pose_auc` block."""
import numpy as np

from . import mvba as M
from . import pose as P
from .matcher_torch import matcher_forward


def _one(data, b):
    return {k: (v[b:b + 1] if isinstance(v, np.ndarray) and v.ndim >= 1 and not k.startswith('image') else v)
            for k, v in data.items()}


def tuple_errors(sd, layers, data, b, conf_thresh=0.0):
    """-> list of (pose_error, err_t, err_R) for every pair id0 < id1 of tuple b (eval_multi_view.py:53-66)."""
    T = len(data['ids'])
    one = _one(data, b)
    res = matcher_forward(sd, {'GNN_layers': layers, 'multi_frame_matching': True}, one)
    scene = {'kpts': [one['keypoints%d' % i][0] for i in range(T)], 'K': one['intr0'][0][:3, :3],
             'matches': {}, 'conf': {}}
    for i1 in range(T):
        for i0 in range(i1):
            scene['matches'][(i0, i1)] = res['matches%d_%d_%d' % (i0, i0, i1)][0]
            scene['conf'][(i0, i1)] = res['conf_scores_%d_%d' % (i0, i1)][0, :, 0]
    out = M.multi_view_pipeline(scene, conf_thresh=conf_thresh)
    extr = out['extr']
    errs = []
    for id1 in range(T):
        for id0 in range(id1):
            pose0, pose1 = one['pose%d' % id0][0].astype(np.float64), one['pose%d' % id1][0].astype(np.float64)
            T_021 = np.linalg.inv(pose1) @ pose0                     # cam->world poses (eval_multi_view.py:59)
            Tp = extr[id1] @ np.linalg.inv(extr[id0])
            et, er = P.compute_pose_error(T_021, Tp[:3, :3], Tp[:3, 3])
            errs.append((max(et, er), et, er))
    return errs


def pair_error(sd, layers, data, b, eval_mode='w8pt_ba', match_threshold=0.0):
    """-> pose error of pair b (eval_pairs.py:212-267); np.inf when the pose cannot be computed."""
    one = _one(data, b)
    res = matcher_forward(sd, {'GNN_layers': layers, 'multi_frame_matching': False}, one)
    kpts0, kpts1 = one['keypoints0'][0], one['keypoints1'][0]
    matches = res['matches0_0_1'][0]
    conf = res['conf_scores_0_1'][0, :, 0]
    valid = matches > -1
    mk0, mk1, mconf = kpts0[valid], kpts1[matches[valid]], conf[valid]
    cm = mconf > match_threshold
    K0 = one['intr0'][0][:3, :3].astype(np.float64)[None]
    K1 = one['intr1'][0][:3, :3].astype(np.float64)[None]
    T_0to1 = np.linalg.inv(one['pose1'][0].astype(np.float64)) @ one['pose0'][0].astype(np.float64)
    Tw, info = P.estimate_relative_pose_w8pt(mk0[cm].astype(np.float64)[None], mk1[cm].astype(np.float64)[None], K0, K1,
                                             mconf[cm].astype(np.float64)[None, :, None], determine_inliers=True)
    if Tw is None:
        return np.inf
    if 'ba' in eval_mode:
        c = info['confidence'].copy()
        c[~info['pos_depth_mask']] = 0.0
        ext, vb = P.run_bundle_adjust_2_view(info['kpts0_norm'], info['kpts1_norm'], c, Tw, 10)
        if vb[0]:
            Tw = ext
    et, er = P.compute_pose_error(T_0to1, Tw[0, :3, :3], Tw[0, :3, 3])
    return max(et, er)
