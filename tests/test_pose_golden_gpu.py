"""CUDA two-view pose kernels against fixtures produced by the REFERENCE's own code
(oracle/make_pose_golden.py -> tests/golden/pose_*.npz): weighted eight-point incl. cheirality vote,
positive-depth mask and inlier test (masks compared EXACTLY), choose-closest branch, and the 10-iteration LM
bundle adjustment incl. items with <= 6 matches.  The kernels compute in fp64, so they are held to the
reference's double-precision run (tight) and to its shipped fp32 run at that run's own measured noise
(tests/golden/pose_report.json: the dense fp32 LU moves rotation by up to 8e-4 and the scale gauge by 3e-2)."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from tests.util import GOLDEN

pytestmark = pytest.mark.gpu

W8PT = sorted(glob.glob(os.path.join(GOLDEN, 'pose_w8pt_n*.npz')))
BA = sorted(glob.glob(os.path.join(GOLDEN, 'pose_ba_*.npz')))


def tdir(T):
    t = T[..., :3, 3]
    return t / np.linalg.norm(t, axis=-1, keepdims=True)


def _cuda(z, *keys):
    return [torch.from_numpy(z[k]).cuda() for k in keys]


@pytest.mark.parametrize('path', W8PT, ids=[os.path.basename(p)[5:-4] for p in W8PT])
def test_w8pt_vs_reference_golden(path):
    from e2e_multi_view_matching_b200.pose_optimization.two_view.estimate_relative_pose import estimate_relative_pose_w8pt
    z = np.load(path)
    k0, k1, K, c = _cuda(z, 'kpts0', 'kpts1', 'intr', 'conf')
    T, info = estimate_relative_pose_w8pt(k0, k1, K, K, c, determine_inliers=True)
    T = T.cpu().numpy()
    assert np.abs(T - z['T64']).max() < 2e-6, np.abs(T - z['T64']).max()       # (R,t) contract: 1e-4 rel
    assert np.abs(T - z['T32']).max() < 2e-5, np.abs(T - z['T32']).max()
    # masks: exact
    assert np.array_equal(info['pos_depth_mask'].cpu().numpy(), z['pos64'])
    assert np.array_equal(info['inliers'].cpu().numpy(), z['inl64'])
    assert np.array_equal(info['pos_depth_mask'].cpu().numpy(), z['pos32'])
    assert np.array_equal(info['inliers'].cpu().numpy(), z['inl32'])
    np.testing.assert_allclose(info['confidence'].cpu().numpy(), z['conf32'], rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(info['kpts0_norm'].cpu().numpy(), z['k0n32'], atol=1e-6)
    np.testing.assert_allclose(info['kpts1_norm'].cpu().numpy(), z['k1n32'], atol=1e-6)


def test_w8pt_choose_closest_vs_reference_golden():
    from e2e_multi_view_matching_b200.pose_optimization.two_view.estimate_relative_pose import estimate_relative_pose_w8pt
    z = np.load(os.path.join(GOLDEN, 'pose_w8pt_closest_b4_n200.npz'))
    k0, k1, K, c, Tg = _cuda(z, 'kpts0', 'kpts1', 'intr', 'conf', 'T_gt')
    T, info = estimate_relative_pose_w8pt(k0, k1, K, K, c, choose_closest=True, T_021=Tg)
    assert np.abs(T.cpu().numpy() - z['T64']).max() < 2e-6
    assert np.abs(T.cpu().numpy() - z['T32']).max() < 2e-5
    assert np.array_equal(info['pos_depth_mask'].cpu().numpy(), z['pos64'])


@pytest.mark.parametrize('path', BA, ids=[os.path.basename(p)[5:-4] for p in BA])
def test_ba2view_vs_reference_golden(path):
    from e2e_multi_view_matching_b200.pose_optimization.two_view.estimate_relative_pose import run_bundle_adjust_2_view
    z = np.load(path)
    name = json.loads(str(z['meta']))['name']
    rep = json.load(open(os.path.join(GOLDEN, 'pose_report.json')))
    k0, k1, c, Ti = _cuda(z, 'kpts0_norm', 'kpts1_norm', 'conf', 'T_init')
    ext, valid = run_bundle_adjust_2_view(k0, k1, c, Ti, n_iterations=10)
    assert np.array_equal(valid.cpu().numpy(), z['valid64']) and np.array_equal(valid.cpu().numpy(), z['valid32'])
    ext = ext.cpu().numpy()
    assert ext.shape == z['ext64'].shape
    if ext.size == 0:
        return
    # the reference's algorithm in double precision: 1e-4 rel contract, met with margin
    assert np.abs(ext - z['ext64']).max() < 1e-5, np.abs(ext - z['ext64']).max()
    # the reference's shipped fp32 arithmetic: within its own distance to its double-precision run
    rot32 = np.abs(ext[:, :3, :3] - z['ext32'][:, :3, :3]).max()
    dir32 = np.abs(tdir(ext) - tdir(z['ext32'])).max()
    assert rot32 <= 1.05 * rep['ba_%s_ref32_vs_ref64_rot' % name] + 2e-5, rot32
    assert dir32 <= 1.05 * rep['ba_%s_ref32_vs_ref64_tdir' % name] + 2e-5, dir32
