"""CPU: the CLI-compatible binaries exist, keep the reference's usage contract (bundle_adjuster.cpp,
ba_initializer.cpp) and refuse problems they cannot solve instead of falling back to anything."""
import os
import subprocess

import numpy as np


def _bin(name):
    from e2e_multi_view_matching_b200 import build
    build.build()
    p = os.path.join(build.BIN, name)
    assert os.access(p, os.X_OK), p
    return p


def test_usage_and_missing_file(tmp_path):
    for name, f in (('bundle_adjuster', 'ba_in.csv'), ('ba_initializer', 'ba_init_in.csv')):
        r = subprocess.run([_bin(name)], capture_output=True, text=True)
        assert r.returncode == 1 and 'Usage: %s <path to read and write>' % name in r.stderr
        r = subprocess.run([_bin(name), str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 2 and f in r.stderr


def test_unsupported_problems_are_refused(tmp_path):
    cam = '1,0,0,0,1,0,0,0,1,0,0,0\n'
    # a point seen by three cameras: not a pairwise problem
    (tmp_path / 'ba_in.csv').write_text('3,0,1,3,1.0,1.0,0.0,0.0\n0,0,0.1,0.1,1.0\n1,0,0.1,0.1,1.0\n2,0,0.1,0.1,1.0\n' +
                                        cam * 3 + '0.2,0.2,2.0\n')
    r = subprocess.run([_bin('bundle_adjuster'), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 2 and 'more than two observations' in r.stderr
    assert not (tmp_path / 'ba_out.csv').exists()
    # header / body mismatch
    (tmp_path / 'ba_in.csv').write_text('2,0,2,4,1.0,1.0,0.0,0.0\n0,0,0.1,0.1,1.0\n1,0,0.1,0.1,1.0\n' + cam * 2 + '0.2,0.2,2.0\n')
    r = subprocess.run([_bin('bundle_adjuster'), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 2 and 'do not match the header' in r.stderr
    # view ids with a hole
    (tmp_path / 'ba_init_in.csv').write_text('0,1,0,0,0,1,0,0,0,1\n2,1,0,0,0,1,0,0,0,1\n')
    r = subprocess.run([_bin('ba_initializer'), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 2 and 'view ids' in r.stderr


def test_result_reader_and_weights(tmp_path):
    from e2e_multi_view_matching_b200.pose_optimization.multi_view import bundle_adjust_io as IO
    (tmp_path / 'ba_out.csv').write_text('1,0,0,0,1,0,0,0,1,0,0,0\n0,1,0,-1,0,0,0,0,1,0.5,-0.25,2\n')
    E = IO.read_bundle_adjust_result(str(tmp_path / 'ba_out.csv'))
    assert len(E) == 2 and np.array_equal(E[0], np.eye(4))
    np.testing.assert_array_equal(E[1][:3, :3], [[0, -1, 0], [1, 0, 0], [0, 0, 1]])     # file is column-major
    np.testing.assert_array_equal(E[1][:3, 3], [0.5, -0.25, 2])
    obs = np.array([[0.1, 0.2, 0.5], [0.3, 0.4, 1.5]])
    out = IO.normalize_confidences(obs.copy())
    np.testing.assert_allclose(out[:, 2], obs[:, 2] / (0.5 * (2.0 + 1e-3)))
    np.testing.assert_array_equal(out[:, :2], obs[:, :2])
