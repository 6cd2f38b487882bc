"""CPU oracle for the matcher + pose hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``e2e_multi_view_matching_b200`` may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / CPU baseline.

Parity pinning status:
  * matcher (oracle/matcher.py, matcher_torch.py): PINNED.  ``oracle/make_golden.py`` and ``make_golden_full.py`` run the
    unmodified reference ``models/models/multi_view_matcher.py``, assert oracle == reference and write
    ``tests/golden/matcher_*.npz`` (incl. full-size cfg2 / cfg3 / cfg4 fixtures).
  * two-view pose (oracle/pose.py): PINNED since round 2.  ``oracle/ref_shim.py`` imports the reference's OWN
    ``estimate_relative_pose.py`` and ``bundle_adjust_gauss_newton_2_view.py`` unmodified, with stub modules for the nine
    kornia 0.7.0 / pytorch3d 0.7.5 leaf functions that are absent from this image (those leaves remain restatements of the
    published functions, cross-checked against OpenCV / SciPy); ``oracle/make_pose_golden.py`` asserts
    oracle/pose.py == reference and writes ``tests/golden/pose_*.npz``.
  * multi-view BA (oracle/mvba.py): pinned on the reference's own gtest known-answer scenes
    (test_ba_problem.cpp:165-184); Ceres itself is absent ("parity unpinned" against Ceres' exact iterates).
  * ba_initializer (oracle/ba_init.py): pinned on the reference's gtest scene (test_ba_init.cpp:93-274); "parity
    unpinned" against Theia's exact iterates.
  * validation pass, ground-truth matches, SuperPoint, training forward / backward: the reference itself is the
    generator (``make_validation_golden.py``, ``make_gt_matches_golden.py``, ``make_superpoint_golden.py``,
    ``make_train_forward_golden.py``, ``make_train_backward_golden.py`` -> ``tests/golden/``).
  * training stage ops (oracle/train_ops.py: float64 torch restatement of every stage op of the training path): PINNED by
    ``tests/test_train_host_logic.py`` -- the product's host-side orchestration on these functions reproduces the
    reference's loss, parameter gradients, couplings / matches / confidences and BatchNorm statistics.
"""
