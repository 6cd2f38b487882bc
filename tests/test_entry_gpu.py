"""GPU: the eval entry points at BASELINE.json's full sizes, checked through size-independent
properties (Sinkhorn marginals, match symmetry, rigid-pose validity) and, for a pair, against the
CPU port of the whole path (matcher + w8pt + BA) on identical synthetic inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _to_cuda(data):
    return {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) and not k.startswith('image')
                else (torch.empty(v.shape, device='meta') if isinstance(v, np.ndarray) else v)) for k, v in data.items()}


def _check_matches(res, a, b, n):
    m0 = res['matches%d_%d_%d' % (a, a, b)]
    m1 = res['matches%d_%d_%d' % (b, a, b)]
    assert m0.dtype == torch.int64 and m0.shape[1] == n
    for bi in range(m0.shape[0]):
        i = torch.nonzero(m0[bi] >= 0)[:, 0]
        assert (m1[bi][m0[bi][i]] == i).all()          # mutual consistency
    Z = res['scores_%d_%d' % (a, b)].double()
    P = torch.exp(Z) / (2 * n)
    assert (P[:, :n, :].sum(2) * 2 * n - 1).abs().max().item() < 2e-2     # 100 iterations: near marginals
    assert torch.isfinite(res['conf_scores_%d_%d' % (a, b)]).all()


@pytest.mark.parametrize('mode', [3, 1])
def test_cfg2_pairs_batch32_w8pt_ba(mode):
    """BASELINE configs[1]: ScanNet-shape 2-view 1024 kpts w8pt_ba, batch 32."""
    import e2e_multi_view_matching_b200 as pkg
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    from e2e_multi_view_matching_b200.pipeline import PairPipeline
    from e2e_multi_view_matching_b200.synthetic import make_state_dict, make_scene_tuple_inputs
    layers = ['self', 'cross'] * 9
    sd = make_state_dict(len(layers), seed=0, final_proj_gain=12.0)
    m = MultiViewMatcher({'multi_frame_matching': False, 'GNN_layers': layers}).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    pipe = PairPipeline(m.cuda(), 'w8pt_ba')
    data = _to_cuda(make_scene_tuple_inputs(31, 2, 1024, batch=32))
    pkg.set_math_mode(mode)
    res, pose = pipe(data)
    _check_matches(res, 0, 1, 1024)
    T = pose['T_021'].double()
    R = T[:, :3, :3]
    assert (R @ R.transpose(1, 2) - torch.eye(3, device=R.device, dtype=R.dtype)).abs().max().item() < 1e-5
    assert (torch.linalg.det(R) - 1).abs().max().item() < 1e-5
    assert pose['success'].all()


def test_cfg4_megadepth_shape_2048():
    """BASELINE configs[3]: MegaDepth-shape 2-view 2048 kpts w8pt_ba."""
    from e2e_multi_view_matching_b200 import eval_pairs
    r = eval_pairs.main(['--dataset', 'megadepth', '--n_pairs', '2', '--batch', '2', '--eval_mode', 'w8pt_ba'])
    assert r['n_pairs'] == 2 and r['cannot_compute_pose'] == 0


def test_cfg3_multi_view_entry_point():
    """BASELINE configs[2]: 5-tuple multi-view GN-BA, 1024 kpts."""
    from e2e_multi_view_matching_b200 import eval_multi_view
    m = eval_multi_view.main(['--n_tuples', '2', '--batch', '2'])
    assert set(m.keys()) >= {'pose_AUC@5deg', 'transl_AUC@10deg', 'rot_AUC@20deg'}
    assert all(np.isfinite(v) for v in m.values())


def test_pair_pipeline_vs_cpu_port_end_to_end():
    """End-to-end parity track (SURVEY.md §8d ii): matcher output fed to the pose stage, engine vs the
    CPU port on identical inputs (256 kpts so the dense reference BA stays small)."""
    import e2e_multi_view_matching_b200 as pkg
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    from e2e_multi_view_matching_b200.pipeline import PairPipeline
    from e2e_multi_view_matching_b200.synthetic import make_state_dict, make_scene_tuple_inputs
    from oracle.matcher import matcher_forward
    from oracle import pose as P
    layers = ['self', 'cross'] * 3
    sd = make_state_dict(len(layers), seed=3, final_proj_gain=14.0)
    data_np = make_scene_tuple_inputs(77, 2, 256, batch=1)
    m = MultiViewMatcher({'multi_frame_matching': False, 'GNN_layers': layers}).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    res, pose = PairPipeline(m.cuda(), 'w8pt_ba')(_to_cuda(data_np))
    ref = matcher_forward(sd, {'multi_frame_matching': False, 'GNN_layers': layers}, data_np)
    m0 = ref['matches0_0_1'][0]
    c = ref['conf_scores_0_1'][0, :, 0]
    assert (res['matches0_0_1'][0].cpu().numpy() == m0).mean() > 0.99
    valid = (m0 >= 0) & (c > 0)
    K = data_np['intr0'].astype(np.float64)
    k0 = data_np['keypoints0'][0][valid].astype(np.float64)[None]
    k1 = data_np['keypoints1'][0][m0[valid]].astype(np.float64)[None]
    Tw, info = P.estimate_relative_pose_w8pt(k0, k1, K, K, c[valid].astype(np.float64)[None, :, None], determine_inliers=True)
    cn = info['confidence'].copy()
    cn[~info['pos_depth_mask']] = 0
    ext, vb = P.run_bundle_adjust_2_view(info['kpts0_norm'], info['kpts1_norm'], cn, Tw, 10)
    T_ref = ext[0] if vb[0] else Tw[0]
    T = pose['T_021'][0].double().cpu().numpy()
    if (res['matches0_0_1'][0].cpu().numpy() == m0).all():
        np.testing.assert_allclose(T, T_ref, atol=1e-3)
    et, er = P.compute_pose_error(T_ref, T[:3, :3], T[:3, 3])
    assert er < 0.5 and et < 2.0
