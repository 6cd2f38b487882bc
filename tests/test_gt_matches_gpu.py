"""Ground-truth match computation (csrc/gt_matches.cu) against the reference's own
compute_gt_matches_of_image_pair (helpers.py:121-203): fixtures written by oracle/make_gt_matches_golden.py."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'gt_matches_*.npz')))


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[11:-4] for p in CASES])
def test_gt_matches_vs_reference_golden(path):
    from e2e_multi_view_matching_b200.training import compute_gt_matches_of_image_pair
    z = np.load(path)
    t = {k: torch.from_numpy(z[k]).cuda() for k in ('kpts0', 'kpts1', 'K0', 'K1', 'T', 'depth0', 'depth1')}
    e_match, e_unmatch = [float(x) for x in z['thresholds']]
    idx, w = compute_gt_matches_of_image_pair(t['kpts0'], t['kpts1'], t['K0'], t['K1'], t['T'], t['depth0'], t['depth1'],
                                              e_match, e_unmatch)
    idx, w = idx.cpu().numpy(), w.cpu().numpy()
    ref_i, ref_w = z['indices'], z['weights']
    assert idx.shape == ref_i.shape and idx.dtype == np.int64 and w.shape == ref_w.shape and w.dtype == np.float32
    # a decision is stable when the arg-min margin and the distance of the minimum to both thresholds exceed the
    # float32 noise of the reprojection (1e-3 px is ~100 ulp at 10 px)
    tau = 1e-3
    stable = []
    for margin, emin in ((z['row_margin'], z['row_min']), (z['col_margin'], z['col_min'])):
        stable.append((margin > tau) & (np.abs(emin - e_match) > tau) & (np.abs(emin - e_unmatch) > tau))
    stable = np.stack(stable, 1)                                  # [bs, 2, n]
    mism = idx[:, :, :-1] != ref_i[:, :, :-1]
    assert not (mism & stable).any(), ('index mismatch on a stable keypoint', int((mism & stable).sum()))
    assert (idx[:, :, -1] == -1).all()                            # the dustbin entry never matches
    print(os.path.basename(path), 'index mismatches (all on rounding-level ties):', int(mism.sum()), 'of', mism.size)
    if not mism.any():
        # same decisions => same counts => the class-balancing weights are the same float32 numbers
        assert np.array_equal(w == 0, ref_w == 0)
        np.testing.assert_allclose(w, ref_w, rtol=1e-6, atol=0)
    else:
        assert mism.sum() <= 0.01 * mism.size
        np.testing.assert_allclose(w, ref_w, rtol=0.05, atol=0)   # the counts move by the few flipped ties


def test_gt_matches_rejects_small_workspace():
    from e2e_multi_view_matching_b200 import _lib
    lib = _lib.lib()
    x = torch.zeros(1, 8, 2, device='cuda'); K = torch.eye(4, device='cuda')[None].contiguous(); d = torch.ones(1, 4, 4, device='cuda')
    idx = torch.empty(1, 2, 9, dtype=torch.int64, device='cuda'); w = torch.empty(1, 2, 9, device='cuda')
    ws = torch.empty(16, dtype=torch.uint8, device='cuda')
    rc = lib.mvm_gt_matches_pair(_lib.ptr(x), _lib.ptr(x), _lib.ptr(K), _lib.ptr(K), _lib.ptr(K), _lib.ptr(d), _lib.ptr(d),
                                 1, 8, 4, 4, 5.0, 15.0, _lib.ptr(idx), _lib.ptr(w), _lib.ptr(ws), 16, _lib.stream_ptr())
    assert rc != 0
