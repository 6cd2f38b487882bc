"""SuperPoint front-end (csrc/superpoint.cu) against fixtures produced by the UNMODIFIED reference SuperPoint with seeded
weights (oracle/make_superpoint_golden.py): keypoint coordinates exact, scores and descriptors within fp32 noise."""
import json
import os

import numpy as np
import pytest
import torch

from tests.util import GOLDEN

pytestmark = pytest.mark.gpu

CASES = ['120x160_all', '240x320_top200_b2', '480x640_top1024']


@pytest.mark.parametrize('name', CASES)
def test_superpoint_vs_reference_golden(name):
    from e2e_multi_view_matching_b200.models.superpoint import SuperPoint
    from e2e_multi_view_matching_b200.synthetic import make_superpoint_state_dict, make_image
    z = np.load(os.path.join(GOLDEN, 'superpoint_%s.npz' % name))
    meta = json.loads(str(z['meta']))
    sp = SuperPoint({'max_keypoints': meta['max_keypoints']}).eval()
    sp.load_state_dict({k: torch.from_numpy(v) for k, v in make_superpoint_state_dict(meta['wseed']).items()}, strict=True)
    sp = sp.cuda()
    img = torch.from_numpy(make_image(meta['seed'], meta['height'], meta['width'], meta['batch'])).cuda()
    out = sp({'image': [img]})
    assert len(out['keypoints']) == meta['batch']
    for b in range(meta['batch']):
        kp_ref = z['keypoints%d' % b].astype(np.int64)
        kp = out['keypoints'][b].cpu().numpy()
        assert kp.dtype == np.float32 and kp.shape == kp_ref.shape, (kp.shape, kp_ref.shape)
        if meta['max_keypoints'] < 0:
            assert np.array_equal(kp.astype(np.int64), kp_ref)             # nonzero order = row-major, exact
            order = order_ref = np.arange(kp.shape[0])
        else:
            # torch.topk does not promise an order among (near-)equal scores: compare as sets, then align
            key = lambda a: a[:, 1] * 100000 + a[:, 0]
            order, order_ref = np.argsort(key(kp.astype(np.int64))), np.argsort(key(kp_ref))
            assert np.array_equal(kp.astype(np.int64)[order], kp_ref[order_ref])
        sc, sc_ref = out['scores'][b].cpu().numpy()[order], z['scores%d' % b][order_ref]
        np.testing.assert_allclose(sc, sc_ref, rtol=2e-5, atol=1e-7)
        d, d_ref = out['descriptors'][b].cpu().numpy()[:, order], z['descriptors%d' % b][:, order_ref]
        assert d.shape == d_ref.shape == (256, kp.shape[0])
        assert np.abs(d - d_ref).max() < 1e-4, np.abs(d - d_ref).max()      # 3xTF32 1x1 head + fp32 convolutions
        np.testing.assert_allclose(np.linalg.norm(d, axis=0), 1.0, atol=1e-5)


def test_superpoint_feeds_the_matcher():
    """image -> keypoints/descriptors -> MultiViewMatcher on the device (run_super_point, helpers.py:83-96)."""
    from e2e_multi_view_matching_b200.models.superpoint import SuperPoint
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    from e2e_multi_view_matching_b200.synthetic import make_superpoint_state_dict, make_state_dict, make_image
    sp = SuperPoint({'max_keypoints': 256}).eval()
    sp.load_state_dict({k: torch.from_numpy(v) for k, v in make_superpoint_state_dict(0).items()})
    sp = sp.cuda()
    layers = ['self', 'cross'] * 2
    m = MultiViewMatcher({'GNN_layers': layers, 'multi_frame_matching': False}).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in make_state_dict(len(layers), seed=1, final_proj_gain=8.0).items()})
    m = m.cuda()
    imgs = [torch.from_numpy(make_image(5 + i, 240, 320)).cuda() for i in range(2)]
    pred = sp({'image': imgs})
    data = {'ids': [0, 1]}
    for i in range(2):
        data['keypoints%d' % i] = pred['keypoints'][i][None]
        data['scores%d' % i] = pred['scores'][i][None]
        data['descriptors%d' % i] = pred['descriptors'][i][None]
        data['image%d' % i] = imgs[i]
    out = m(data)
    assert out['matches0_0_1'].shape == (1, 256) and out['scores_0_1'].shape == (1, 257, 257)
    assert torch.isfinite(out['scores_0_1']).all()
