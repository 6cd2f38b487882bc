"""CPU: the ground-truth-match fixtures (reference outputs, oracle/make_gt_matches_golden.py) are self-consistent --
mutual assignments, dustbin entries, class-balancing weights as helpers.py:190-213 defines them."""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
CASES = sorted(glob.glob(os.path.join(GOLDEN, 'gt_matches_*.npz')))


def test_fixtures_exist():
    assert len(CASES) >= 4


@pytest.mark.parametrize('path', CASES, ids=[os.path.basename(p)[11:-4] for p in CASES])
def test_reference_outputs_are_consistent(path):
    z = np.load(path)
    idx, w = z['indices'], z['weights']
    bs, _, nb = idx.shape
    n = nb - 1
    assert z['kpts0'].shape == (bs, n, 2) and w.shape == idx.shape
    for b in range(bs):
        i0, i1 = idx[b, 0, :n], idx[b, 1, :n]
        m = i0 >= 0
        assert (i1[i0[m]] == np.nonzero(m)[0]).all()                 # mutual
        assert idx[b, 0, n] == -1 and idx[b, 1, n] == -1              # dustbin entries
        n_match = int(m.sum())
        assert n_match == int((i1 >= 0).sum())
        # weights take three values: 0 (dropped), the match weight, the unmatch weight (helpers.py:205-213)
        vals = np.unique(w[b])
        assert len(vals) <= 3 and (vals >= 0).all()
        if n_match:
            mw = w[b, 0, :n][m][0]
            n_drop = int((w[b, :, :n] == 0).sum())
            expect = np.float32(0.5) / (np.float32(2.0 * n_match) / np.float32(2.0 * n - n_drop))
            assert abs(mw - expect) <= 1e-6 * expect
