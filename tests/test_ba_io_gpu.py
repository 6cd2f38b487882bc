"""GPU: the file-protocol route of the multi-view stage (SURVEY.md §8 b "Multi-view BA", f-4): the three
Python functions of bundle_adjust_io.py and the two CLI-compatible binaries, against (i) the CPU oracle solving
the very CSV files that were written and (ii) the device-resident MultiViewPoseEngine."""
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _data_result(scene):
    T = len(scene['kpts'])
    data, result = {}, {}
    for v in range(T):
        data['keypoints%d' % v] = torch.from_numpy(scene['kpts'][v])[None].cuda()
        data['intr%d' % v] = torch.from_numpy(scene['K'])[None]
    for (a, b), m in scene['matches'].items():
        result['matches%d_%d_%d' % (a, a, b)] = torch.from_numpy(m)[None].cuda()
        result['conf_scores_%d_%d' % (a, b)] = torch.from_numpy(scene['conf'][(a, b)])[None, :, None].cuda()
    return data, result


def _parse_ba_in(path):
    from oracle import mvba as M
    cams, pts, oc, op, oxy, ow = [], [], [], [], [], []
    hdr = None
    for line in open(path):
        e = line.strip().split(',')
        if len(e) == 8:
            hdr = [float(x) for x in e]
        elif len(e) == 3:
            pts.append([float(x) for x in e])
        elif len(e) == 5:
            oc.append(int(e[0])); op.append(int(e[1])); oxy.append([float(e[2]), float(e[3])]); ow.append(float(e[4]))
        elif len(e) == 12:
            v = [float(x) for x in e]
            R = np.array(v[:9]).reshape(3, 3).T
            cams.append(np.concatenate([M.R_to_angle_axis(R), v[9:]]))
    assert hdr is not None and int(hdr[0]) == len(cams) and int(hdr[2]) == len(pts) and int(hdr[3]) == len(oc)
    return M.BaProblem(cams, pts, oc, op, oxy, ow, fixed_cam=int(hdr[1]), intr=tuple(hdr[4:]))


def _parse_ba_init_in(path):
    from oracle import mvba as M
    rot, prot, ppos = {}, {}, {}
    for line in open(path):
        e = line.strip().split(',')
        if len(e) == 10:
            rot[int(e[0])] = M.R_to_angle_axis(np.array([float(x) for x in e[1:]]).reshape(3, 3).T)
        elif len(e) == 14:
            k = (int(e[0]), int(e[1]))
            prot[k] = M.R_to_angle_axis(np.array([float(x) for x in e[2:11]]).reshape(3, 3).T)
            ppos[k] = np.array([float(x) for x in e[11:]])
    return np.array([rot[v] for v in range(len(rot))]), prot, ppos


def _run(binary, tmp):
    from e2e_multi_view_matching_b200.pose_optimization.multi_view import bundle_adjust_io as IO
    r = subprocess.run([os.path.join(IO.BUNDLE_ADJUSTMENT_BUILD_DIR, binary), str(tmp)], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, (binary, r.stdout, r.stderr)
    return r.stdout


@pytest.mark.parametrize('T,n,seed', [(3, 80, 3), (5, 120, 4)])
def test_file_protocol_against_oracle_and_engine(tmp_path, T, n, seed):
    from oracle import mvba as M, ba_init as BI
    from oracle.pose import compute_pose_error
    from e2e_multi_view_matching_b200.pose_optimization.multi_view import bundle_adjust_io as IO
    scene = M.make_multi_view_scene(seed, T, n, outlier_frac=0.1)
    data, result = _data_result(scene)

    # --- eval_bundle_adjust, step by step, through the files (eval_multi_view.py:21-51) ---
    pw = IO.initialize_bundle_adjust(T, data, result, str(tmp_path / 'ba_init_in.csv'))
    _run('ba_initializer', tmp_path)
    extr_init = np.array(IO.read_bundle_adjust_result(str(tmp_path / 'ba_init_out.csv')))
    IO.write_bundle_adjust_problem(T, pw, extr_init, str(tmp_path / 'ba_in.csv'))
    out = _run('bundle_adjuster', tmp_path)
    extr = np.array(IO.read_bundle_adjust_result(str(tmp_path / 'ba_out.csv')))
    assert extr.shape == (T, 4, 4) and 'iterations' in out

    # --- (i) the oracle on the same files ---
    rot0, prot, ppos = _parse_ba_init_in(str(tmp_path / 'ba_init_in.csv'))
    rot = BI.robust_rotation_averaging(T, prot, rot0)
    pos = BI.lud_positions(T, ppos, rot)
    for v in range(T):
        R = M.angle_axis_to_R(rot[v])
        np.testing.assert_allclose(extr_init[v, :3, :3], R, atol=2e-4)
        np.testing.assert_allclose(extr_init[v, :3, 3], -R @ pos[v], atol=2e-4 * max(1.0, np.abs(pos).max()))
    pb = _parse_ba_in(str(tmp_path / 'ba_in.csv'))
    # triangulated points written to the file = DLT with the file's cameras
    cams_o, _, info = M.solve(pb)
    ref = M.cams_to_extrinsics(cams_o)
    np.testing.assert_allclose(extr[0], ref[0], atol=1e-9)          # fixed camera untouched
    if info['termination'] != 'max_iterations':
        for v in range(1, T):
            et, er = compute_pose_error(ref[v], extr[v][:3, :3], extr[v][:3, 3])
            assert er < 0.05 and et < 0.5, (v, et, er, info)

    # --- (ii) in-process routes give the same answer as the file route ---
    extr_init2 = IO.ba_initialize(T, pw)
    np.testing.assert_allclose(extr_init2, extr_init, atol=1e-6)
    extr2 = np.array(IO.solve_bundle_adjust(T, pw, extr_init))
    for v in range(1, T):
        et, er = compute_pose_error(extr[v], extr2[v][:3, :3], extr2[v][:3, 3])
        assert er < 0.05 and et < 0.5, (v, et, er)      # two LM runs stop within the function tolerance of each other
    extr3 = np.array(IO.solve(T, data, result))
    for v in range(1, T):       # free scale gauge + LM stopping tolerance: compare what the reference evaluates
        et, er = compute_pose_error(extr2[v], extr3[v][:3, :3], extr3[v][:3, 3])
        # the two routes start from initial poses that differ in the 12th digit (CSV precision); on the 5-view
        # scene with outliers LM runs into the 50-iteration limit and the end points drift apart
        assert (er < 0.05 and et < 0.5) if T == 3 else (er < 0.5 and et < 2.0), (v, et, er)


def test_pair_wise_data_keys_and_counts(tmp_path):
    """pair_wise_data carries the reference's keys (bundle_adjust_io.py:62-133) and the w8pt_ba mode keeps every
    valid match while counting inliers separately (:114-117)."""
    from oracle import mvba as M
    from e2e_multi_view_matching_b200.pose_optimization.multi_view import bundle_adjust_io as IO
    scene = M.make_multi_view_scene(7, 3, 64, outlier_frac=0.2)
    data, result = _data_result(scene)
    pw = IO.initialize_bundle_adjust(3, data, result, str(tmp_path / 'ba_init_in.csv'), conf_thresh=0.1)
    for (a, b), m in scene['matches'].items():
        valid = (m >= 0) & (scene['conf'][(a, b)] > 0.1)
        assert pw['mkpts%d_%d_%d' % (a, a, b)].shape == (valid.sum(), 2)
        np.testing.assert_array_equal(pw['mkpts%d_%d_%d' % (b, a, b)], scene['kpts'][b][m[valid]])
        assert pw['conf%d_%d_%d' % (a, a, b)].shape == (valid.sum(), 1)
        assert 0 < pw['inlier_count%d_%d' % (a, b)] <= valid.sum()
        assert pw['rel_pose%d_%d' % (a, b)].shape == (4, 4)
    assert all('abs_init_pose%d' % v in pw for v in range(3))
    lines = open(tmp_path / 'ba_init_in.csv').read().strip().split('\n')
    assert [len(l.split(',')) for l in lines[:3]] == [10, 10, 10] and all(len(l.split(',')) == 14 for l in lines[3:])
    with pytest.raises(NotImplementedError):
        IO.initialize_bundle_adjust(3, data, result, None, rel_pose_method='ransac')
