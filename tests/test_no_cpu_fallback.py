"""CPU: the product path fails loudly instead of computing on the CPU -- every Python mirror raises MvmError on
non-CUDA tensors, stage wrappers refuse CPU tensors, and the reference's value-type error returns that do not need
a GPU ((None, None) for fewer than eight keypoints) still work."""
import numpy as np
import pytest
import torch

from e2e_multi_view_matching_b200 import _lib, ops
from e2e_multi_view_matching_b200.synthetic import make_state_dict, make_view_inputs


def _matcher(cls_name='MultiViewMatcher'):
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    layers = ['self', 'cross']
    sd = make_state_dict(len(layers), seed=1)
    m = MultiViewMatcher({'GNN_layers': layers}).eval()
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return m


def test_matcher_refuses_cpu_tensors():
    m = _matcher()
    data = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in make_view_inputs(3, [40, 40]).items()}
    with pytest.raises(_lib.MvmError, match='no CPU fallback'):
        m(data)


def test_two_view_pose_refuses_cpu_tensors_but_keeps_value_errors():
    from e2e_multi_view_matching_b200.pose_optimization.two_view.estimate_relative_pose import (
        estimate_relative_pose_w8pt, run_bundle_adjust_2_view)
    K = torch.eye(3)[None]
    k = torch.rand(1, 20, 2)
    with pytest.raises(_lib.MvmError, match='no CPU fallback'):
        estimate_relative_pose_w8pt(k, k, K, K, torch.ones(1, 20, 1))
    assert estimate_relative_pose_w8pt(k[:, :5], k[:, :5], K, K, torch.ones(1, 5, 1)) == (None, None)   # :85-86
    with pytest.raises(_lib.MvmError, match='no CPU fallback'):
        run_bundle_adjust_2_view(k, k, torch.ones(1, 20, 1), torch.eye(4)[None], n_iterations=2)


def test_stage_wrappers_refuse_cpu_tensors():
    with pytest.raises(AssertionError, match='CUDA'):
        ops.linear(torch.zeros(128, 32), torch.zeros(128, 32), tc_passes=3)
    with pytest.raises(AssertionError, match='CUDA'):
        ops.linear(torch.zeros(128, 32), torch.zeros(128, 32))


def test_ransac_modes_are_not_silently_emulated():
    from e2e_multi_view_matching_b200.pose_optimization.multi_view import bundle_adjust_io as IO
    with pytest.raises(NotImplementedError):
        IO.initialize_bundle_adjust(2, {}, {}, None, rel_pose_method='ransac_ba')
