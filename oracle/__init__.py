"""CPU oracle for the matcher + pose hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``e2e_multi_view_matching_b200`` may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / CPU baseline.

Parity pinning status:
  * matcher (oracle/matcher.py): PINNED.  Checked against the importable
    reference ``models/models/multi_view_matcher.py`` by
    ``oracle/make_golden.py``; fixtures in ``tests/golden/``.
  * two-view pose (oracle/pose.py): "parity unpinned" -- kornia 0.7.0 and
    pytorch3d 0.7.5 are pip-pinned third-party dependencies that are absent
    from /root/reference and from this image; their published algorithms are
    restated, the reference's own call sites are followed line by line;
    cross-checked against OpenCV 4.13 (tests/test_pose_oracle_opencv.py).
  * multi-view BA (oracle/mvba.py): pinned on the reference's own gtest
    known-answer scenes (test_ba_problem.cpp:165-184); Ceres itself is absent
    ("parity unpinned" against Ceres' exact iterates).
  * ba_initializer (oracle/ba_init.py): pinned on the reference's gtest scene
    (test_ba_init.cpp:93-274); "parity unpinned" against Theia's exact iterates.
"""
