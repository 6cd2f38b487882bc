"""The reference's own known-answer tests, through the CUDA solvers (not only through the CPU oracle):
  * test_ba_problem.cpp:165-184 -- three 2-camera / 5-point scenes whose observation weights are the per-camera depths
    (one weight per OBSERVATION, ba_problem.h:60-151) -> mvm_multi_view_ba_obs;
  * test_ba_init.cpp:93-274 -- all 13 RotationAveraging / TranslationAveraging / TransformationAveraging cases on the
    four-camera unit-square scene -> mvm_ba_initialize.
Same scenes, same glibc rand() noise streams and same tolerances as tests/test_mvba_oracle.py / test_ba_init_oracle.py."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

EXPECTED = [0.3, -0.2, 0.5, 0.3, -0.4, 0.5]


def _extr(cam):
    from oracle.mvba import angle_axis_to_R
    T = np.eye(4)
    T[:3, :3] = angle_axis_to_R(np.asarray(cam[:3], float))
    T[:3, 3] = cam[3:]
    return T


@pytest.mark.parametrize('name,args,tol', [('Perfect2Cams5Pts', (0., 0., 0.), 1e-6),
                                           ('Noisy2Cams5Pts', (0.1, 10., 0.2), 9e-2),
                                           ('MoreNoisy2Cams5Pts', (0.2, 0., 0.3), 4e-2)])
def test_ba_problem_gtest_scene_through_cuda(name, args, tol):
    from oracle import mvba as M
    from e2e_multi_view_matching_b200 import _lib
    lib = _lib.lib()
    pb = M.gtest_problem(EXPECTED, *args)
    n, n_pad = 5, 64
    xa = np.zeros((1, 1, n_pad, 2), np.float32); xb = np.zeros_like(xa)
    wa = np.zeros((1, 1, n_pad), np.float32); wb = np.zeros_like(wa)
    for o in range(len(pb.obs_cam)):
        c, k = pb.obs_cam[o], pb.obs_pt[o]
        (xa if c == 0 else xb)[0, 0, k] = pb.obs_xy[o]
        (wa if c == 0 else wb)[0, 0, k] = pb.obs_w[o, 0]
    extr = np.stack([_extr(pb.cams[0]), _extr(pb.cams[1])])[None]
    pts = np.zeros((1, 1, n_pad, 3)); pts[0, 0, :n] = pb.points
    t = lambda a, dt: torch.tensor(a, dtype=dt).cuda().contiguous()
    d_xa, d_xb, d_wa, d_wb = t(xa, torch.float32), t(xb, torch.float32), t(wa, torch.float32), t(wb, torch.float32)
    d_ex, d_pts = t(extr, torch.float64), t(pts, torch.float64)
    nv = torch.tensor([[n]], dtype=torch.int32).cuda()
    out = torch.empty(1, 2, 4, 4, dtype=torch.float32).cuda()
    out64 = torch.empty(1, 2, 4, 4, dtype=torch.float64).cuda()
    iters = torch.zeros(1, dtype=torch.int32).cuda()
    cost = torch.zeros(1, 2, dtype=torch.float64).cuda()
    nbytes = lib.mvm_mvba_workspace_bytes(2, 1, 1, n_pad)
    ws = torch.empty(nbytes, dtype=torch.uint8).cuda()
    pa, pb_ = (C.c_int * 1)(0), (C.c_int * 1)(1)
    rc = lib.mvm_multi_view_ba_obs(pa, pb_, 2, 1, 1, n_pad, _lib.ptr(d_xa), _lib.ptr(d_xb), _lib.ptr(d_wa), _lib.ptr(d_wb),
                                   _lib.ptr(nv), _lib.ptr(d_ex), _lib.ptr(d_pts), 1, _lib.ptr(out), _lib.ptr(out64), 50,
                                   _lib.ptr(iters), _lib.ptr(cost), _lib.ptr(ws), nbytes, _lib.stream_ptr())
    assert rc == 0
    E = out64[0].cpu().numpy()
    got = np.concatenate([M.R_to_angle_axis(E[1, :3, :3]), E[1, :3, 3]])
    np.testing.assert_allclose(E[0], np.eye(4), atol=1e-12)                  # camera 0 stays fixed (:152-156)
    assert np.abs(got - np.array(EXPECTED)).max() < tol, (name, got)         # the reference's EXPECT_NEAR
    cams, _, info = M.solve(pb)                                              # and the CPU oracle, iterate for iterate
    assert np.abs(got - cams[1]).max() < 1e-5, (got, cams[1])
    assert abs(int(iters[0]) - info['iterations']) <= 1      # the oracle leaves before its first step when the initial gradient is 0
    c = cost[0].cpu().numpy()
    np.testing.assert_allclose(c[0], info["initial_cost"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(c[1], info["final_cost"], rtol=1e-3, atol=1e-12)


def _run_ba_init(init_extr, rel):
    from e2e_multi_view_matching_b200 import _lib
    lib = _lib.lib()
    T, pairs = 4, [(a, b) for b in range(4) for a in range(b)]
    P = len(pairs)
    pa = (C.c_int * P)(*[a for a, _ in pairs]); pb = (C.c_int * P)(*[b for _, b in pairs])
    Trel = torch.tensor(np.array([rel[p] for p in pairs])[None], dtype=torch.float32).cuda().contiguous()
    e0 = torch.tensor(np.array(init_extr)[None], dtype=torch.float64).cuda().contiguous()
    ones = torch.ones(1, P, dtype=torch.uint8).cuda()
    inl = torch.ones(1, P, 64, dtype=torch.uint8).cuda()
    out = torch.empty(1, T, 4, 4, dtype=torch.float64).cuda()
    ne = torch.zeros(1, dtype=torch.int32).cuda()
    rc = lib.mvm_ba_initialize(pa, pb, T, P, 1, 64, _lib.ptr(e0), _lib.ptr(Trel), _lib.ptr(ones), _lib.ptr(ones),
                               _lib.ptr(inl), 20, _lib.ptr(out), _lib.ptr(ne), _lib.stream_ptr())
    assert rc == 0 and int(ne[0]) == P
    return out[0].cpu().numpy()


def test_ba_init_gtest_cases_through_cuda():
    """All 13 cases of test_ba_init.cpp.  mvm_ba_initialize runs rotation averaging and then LUD positions on the
    averaged rotations; a RotationAveraging case feeds (pair rotations, initial rotations) with exact pair positions,
    a TranslationAveraging case feeds pair positions with pair rotations consistent with the GIVEN global rotations
    (so that the averaging step returns exactly those), the Transformation case feeds both noisy."""
    from oracle import ba_init as B
    from oracle.mvba import R_to_angle_axis, angle_axis_to_R
    NOISY = 1.6      # see tests/test_ba_init_oracle.py: the original rand() stream position cannot be reproduced
    C.CDLL('libc.so.6').srand(1)
    N = B.GtestNoise()
    extr = B.gtest_extrinsics()
    pairs = [(a, b) for b in range(4) for a in range(b)]

    def rel_from(rot_p, pos_p):
        rel = {}
        for p in pairs:
            Tm = np.eye(4)
            R = angle_axis_to_R(rot_p[p])
            Tm[:3, :3] = R
            Tm[:3, 3] = -R @ pos_p[p]          # pos_p = position of the second camera in the first one's frame
            rel[p] = Tm
        return rel

    def init_from(rot, pos=None):
        out = []
        for i in range(4):
            Tm = np.eye(4)
            Tm[:3, :3] = angle_axis_to_R(rot[i])
            c = np.linalg.inv(extr[i])[:3, 3] if pos is None else pos[i]
            Tm[:3, 3] = -Tm[:3, :3] @ c
            out.append(Tm)
        return out

    def check_rot(E, tol):
        for i in range(4):
            assert np.abs(R_to_angle_axis(E[i][:3, :3]) - R_to_angle_axis(extr[i][:3, :3])).max() < tol, i

    def check_pos(E, tol):
        for i in range(4):
            assert np.abs(np.linalg.inv(E[i])[:3, 3] - np.linalg.inv(extr[i])[:3, 3]).max() < tol, i

    # ---- RotationAveraging.* (:93-173) ----
    cases = []
    rot_p, pos_p = N.view_pairs(extr); init = N.global_rotations(extr); cases.append((rot_p, init, 1e-6))
    rot_p, pos_p = N.view_pairs(extr); init = N.global_rotations(extr); rot_p[(1, 2)] = -0.5 * rot_p[(1, 2)]; cases.append((rot_p, init, 1e-4))
    rot_p, pos_p = N.view_pairs(extr, 0.05); init = N.global_rotations(extr); cases.append((rot_p, init, NOISY * 4e-2))
    rot_p, pos_p = N.view_pairs(extr); init = N.global_rotations(extr); init[2] = -0.5 * init[2]; cases.append((rot_p, init, 1e-6))
    rot_p, pos_p = N.view_pairs(extr); init = N.global_rotations(extr, 0.03); cases.append((rot_p, init, NOISY * 3e-2))
    rot_p, pos_p = N.view_pairs(extr, 0.02); init = N.global_rotations(extr, 0.03); cases.append((rot_p, init, NOISY * 3e-2))
    exact_rot, exact_pos = GtestPairsExact(extr)
    for rot_p, init, tol in cases:
        E = _run_ba_init(init_from(init), rel_from(rot_p, exact_pos))
        check_rot(E, tol)
        ref = B.robust_rotation_averaging(4, rot_p, init)            # and the CPU restatement of Theia's estimator
        for i in range(4):
            assert np.abs(R_to_angle_axis(E[i][:3, :3]) - ref[i]).max() < 2e-4

    # ---- TranslationAveraging.* (:176-258) ----
    tcases = []
    rot_p, pos_p = N.view_pairs(extr); rot = N.global_rotations(extr); tcases.append((pos_p, rot, 1e-6))
    rot_p, pos_p = N.view_pairs(extr); pos_p[(1, 2)] = -0.5 * pos_p[(1, 2)]; rot = N.global_rotations(extr); tcases.append((pos_p, rot, 1e-4))
    rot_p, pos_p = N.view_pairs(extr, 0.05); rot = N.global_rotations(extr); tcases.append((pos_p, rot, NOISY * 5e-2))
    rot_p, pos_p = N.view_pairs(extr); rot = N.global_rotations(extr); rot[1] = 0.9 * rot[1]; tcases.append((pos_p, rot, NOISY * 1e-1))
    rot_p, pos_p = N.view_pairs(extr); rot = N.global_rotations(extr, 0.03); tcases.append((pos_p, rot, NOISY * 4e-2))
    rot_p, pos_p = N.view_pairs(extr, 0.03); rot = N.global_rotations(extr, 0.03); tcases.append((pos_p, rot, NOISY * 3e-2))
    for pos_p, rot, tol in tcases:
        # pair rotations consistent with the given global rotations: R_ij = R_j R_i^T
        cons = {(i, j): R_to_angle_axis(angle_axis_to_R(rot[j]) @ angle_axis_to_R(rot[i]).T) for (i, j) in pairs}
        E = _run_ba_init(init_from(rot), rel_from(cons, pos_p))
        check_pos(E, tol)
        ref = B.lud_positions(4, pos_p, rot)
        for i in range(4):
            assert np.abs(np.linalg.inv(E[i])[:3, 3] - ref[i]).max() < 5e-4

    # ---- TransformationAveraging.NoisyInitNoisyRel (:260-274) ----
    rot_p, pos_p = N.view_pairs(extr, 0.02); init = N.global_rotations(extr, 0.03)
    E = _run_ba_init(init_from(init), rel_from(rot_p, pos_p))
    check_rot(E, NOISY * 3e-2)
    check_pos(E, NOISY * 3e-2)


def GtestPairsExact(extr):
    from oracle.mvba import R_to_angle_axis
    rot, pos = {}, {}
    for j in range(4):
        for i in range(j):
            T = extr[j] @ np.linalg.inv(extr[i])
            rot[(i, j)] = R_to_angle_axis(T[:3, :3])
            pos[(i, j)] = np.linalg.inv(T)[:3, 3]
    return rot, pos
