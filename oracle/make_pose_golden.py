"""Generate tests/golden/pose_*.npz by running the REFERENCE's own two-view pose files.

Run in the authoring container only (needs /root/reference):
    python -m oracle.make_pose_golden
``oracle/ref_shim.py`` imports ``pose_optimization/two_view/{estimate_relative_pose,
bundle_adjust_gauss_newton_2_view}.py`` unmodified (the nine absent kornia / pytorch3d leaf functions are
stubbed there).  This script (1) builds seeded two-view scenes, (2) runs the reference functions on them in
fp32 -- the arithmetic the reference ships -- and once more with ``torch.set_default_dtype(float64)``, i.e.
the reference's own code in double precision (its fp32 run is noisy: the dense fp32 LU of the LM step moves
the free scale gauge by up to 1e-2), (3) asserts that the numpy restatement ``oracle/pose.py`` reproduces the
reference (this is what pins the oracle) and (4) stores inputs and outputs as fixtures.  The fixtures travel
to the GPU box; the reference does not.
"""
import json
import os
import warnings

import numpy as np
import torch

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

W8PT_CASES = [
    # name, seed, n matches, outlier fraction, 4x4 intrinsics
    dict(name='n8', seed=101, n=8, outl=0.0, k4=False),
    dict(name='n12', seed=102, n=12, outl=0.0, k4=False),
    dict(name='n40', seed=103, n=40, outl=0.2, k4=True),
    dict(name='n150', seed=104, n=150, outl=0.3, k4=False),
    dict(name='n500', seed=105, n=500, outl=0.3, k4=False),
    dict(name='n1024', seed=106, n=1024, outl=0.3, k4=True),
    dict(name='n1024_clean', seed=107, n=1024, outl=0.0, k4=False),
    dict(name='n2048', seed=108, n=2048, outl=0.3, k4=False),
]
CLOSEST_CASES = [dict(name='closest_b4_n200', seeds=[111, 112, 113, 114], n=200, outl=0.1)]
BA_CASES = [
    # name, seeds (batch), n, outlier fraction, extra zero-confidence fraction, items forced to <= 6 matches
    dict(name='ba_n8', seeds=[121], n=8, outl=0.0, zero=0.0, starve=[]),
    dict(name='ba_n40', seeds=[122], n=40, outl=0.2, zero=0.1, starve=[]),
    dict(name='ba_n150_b3', seeds=[123, 124, 125], n=150, outl=0.3, zero=0.2, starve=[1]),
    dict(name='ba_n500', seeds=[126], n=500, outl=0.3, zero=0.0, starve=[]),
    dict(name='ba_n1024', seeds=[127], n=1024, outl=0.3, zero=0.05, starve=[]),
    dict(name='ba_b4_mixed', seeds=[128, 129, 130, 131], n=64, outl=0.1, zero=0.5, starve=[0, 3]),
    # exactly 7 valid matches: the smallest problem the reference solves (n_matches > 6, :134)
    dict(name='ba_min7', seeds=[132, 133], n=30, outl=0.0, zero=0.0, starve=[], keep=7),
]


def _scene(seed, n, outl):
    from oracle import pose as P
    return P.make_two_view_scene(seed, n, outlier_frac=outl)


def _k4(K):
    K4 = np.tile(np.eye(4, dtype=K.dtype), (K.shape[0], 1, 1))
    K4[:, :3, :3] = K
    return K4


def _np(x):
    return None if x is None else x.detach().cpu().numpy()


def _run_ref(fn, tdt):
    """Run `fn` with the reference's default dtype set to tdt (the reference creates its work tensors with
    torch.eye / torch.zeros, i.e. in the default dtype)."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(tdt)
    try:
        with torch.no_grad():
            return fn()
    finally:
        torch.set_default_dtype(old)


def tdir(T):
    t = T[..., :3, 3]
    return t / np.linalg.norm(t, axis=-1, keepdims=True)


def main():
    warnings.filterwarnings('ignore')
    from oracle import ref_shim
    from oracle import pose as P
    erp, _ = ref_shim.load()
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    report = {}

    # ---- weighted eight-point, eval branch (B = 1, the only batch size the reference uses it with) ----
    for case in W8PT_CASES:
        sc = _scene(case['seed'], case['n'], case['outl'])
        K = _k4(sc['intr']) if case['k4'] else sc['intr']
        out = {}
        for tag, dt, tdt in (('32', np.float32, torch.float32), ('64', np.float64, torch.float64)):
            a = [torch.from_numpy(x.astype(dt)) for x in (sc['kpts0'], sc['kpts1'], K, K, sc['conf'])]
            T, info = _run_ref(lambda: erp.estimate_relative_pose_w8pt(*a, determine_inliers=True), tdt)
            To, io = P.estimate_relative_pose_w8pt(sc['kpts0'].astype(dt), sc['kpts1'].astype(dt), K.astype(dt),
                                                   K.astype(dt), sc['conf'].astype(dt), determine_inliers=True)
            tol = 1e-9 if dt == np.float64 else 2e-5
            err = float(np.abs(_np(T) - To).max())
            assert err < tol, (case['name'], tag, err)
            assert np.array_equal(_np(info['pos_depth_mask']), io['pos_depth_mask']) or dt == np.float32
            assert np.array_equal(_np(info['inliers']), io['inliers']) or dt == np.float32
            out['T' + tag] = _np(T)
            out['pos' + tag] = _np(info['pos_depth_mask'])
            out['inl' + tag] = _np(info['inliers'])
            out['conf' + tag] = _np(info['confidence'])
            out['k0n' + tag] = _np(info['kpts0_norm'])
            out['k1n' + tag] = _np(info['kpts1_norm'])
            report['w8pt_%s_%s_oracle_err' % (case['name'], tag)] = err
        report['w8pt_%s_ref32_vs_ref64' % case['name']] = float(np.abs(out['T32'] - out['T64']).max())
        report['w8pt_%s_mask_flips_32_vs_64' % case['name']] = int((out['pos32'] != out['pos64']).sum() + (out['inl32'] != out['inl64']).sum())
        np.savez_compressed(os.path.join(OUT, 'pose_w8pt_%s.npz' % case['name']), meta=json.dumps(case),
                            kpts0=sc['kpts0'], kpts1=sc['kpts1'], intr=K, conf=sc['conf'], **out)
        print('w8pt', case['name'], 'ok; ref32 vs ref64', report['w8pt_%s_ref32_vs_ref64' % case['name']])

    # ---- weighted eight-point, training branch (choose_closest, batched) ----
    for case in CLOSEST_CASES:
        scs = [_scene(s, case['n'], case['outl']) for s in case['seeds']]
        sc = {k: np.concatenate([s[k] for s in scs], 0) for k in scs[0]}
        out = {}
        for tag, dt, tdt in (('32', np.float32, torch.float32), ('64', np.float64, torch.float64)):
            a = [torch.from_numpy(x.astype(dt)) for x in (sc['kpts0'], sc['kpts1'], sc['intr'], sc['intr'], sc['conf'])]
            Tg = torch.from_numpy(sc['T_021'].astype(dt))
            T, info = _run_ref(lambda: erp.estimate_relative_pose_w8pt(*a, choose_closest=True, T_021=Tg), tdt)
            To, _ = P.estimate_relative_pose_w8pt(sc['kpts0'].astype(dt), sc['kpts1'].astype(dt), sc['intr'].astype(dt),
                                                  sc['intr'].astype(dt), sc['conf'].astype(dt), choose_closest=True,
                                                  T_021=sc['T_021'].astype(dt))
            err = float(np.abs(_np(T) - To).max())
            assert err < (1e-9 if dt == np.float64 else 2e-5), (case['name'], tag, err)
            out['T' + tag] = _np(T)
            out['pos' + tag] = _np(info['pos_depth_mask'])
        np.savez_compressed(os.path.join(OUT, 'pose_w8pt_%s.npz' % case['name']), meta=json.dumps(case),
                            kpts0=sc['kpts0'], kpts1=sc['kpts1'], intr=sc['intr'], conf=sc['conf'], T_gt=sc['T_021'], **out)
        print('w8pt', case['name'], 'ok')

    # ---- two-view LM bundle adjustment (run_bundle_adjust_2_view) ----
    for case in BA_CASES:
        rng = np.random.default_rng(case['seeds'][0] + 7)
        k0n, k1n, cf, Ti = [], [], [], []
        for bi, s in enumerate(case['seeds']):
            sc = _scene(s, case['n'], case['outl'])
            # inputs as eval_pairs.py:239-255 builds them: w8pt (fp32 reference), confidences of matches
            # with negative depth zeroed; here additionally a seeded fraction of zero confidences
            a = [torch.from_numpy(x) for x in (sc['kpts0'], sc['kpts1'], sc['intr'], sc['intr'], sc['conf'])]
            T, info = _run_ref(lambda: erp.estimate_relative_pose_w8pt(*a, determine_inliers=True), torch.float32)
            c = _np(info['confidence'])[0, :, 0].copy()
            c[~_np(info['pos_depth_mask'])[0]] = 0.0
            c[rng.uniform(size=c.shape) < case['zero']] = 0.0
            if case.get('keep'):
                keep = np.flatnonzero(c > 0)[:case['keep']]
                c2 = np.zeros_like(c)
                c2[keep] = c[keep]
                c = c2
            if bi in case['starve']:
                keep = np.flatnonzero(c > 0)[:rng.integers(0, 7)]
                c2 = np.zeros_like(c)
                c2[keep] = c[keep]
                c = c2
            k0n.append(_np(info['kpts0_norm'])[0]); k1n.append(_np(info['kpts1_norm'])[0]); cf.append(c); Ti.append(_np(T)[0])
        k0n, k1n, cf, Ti = np.stack(k0n), np.stack(k1n), np.stack(cf)[..., None], np.stack(Ti)
        out = {}
        for tag, dt, tdt in (('32', np.float32, torch.float32), ('64', np.float64, torch.float64)):
            a = [torch.from_numpy(x.astype(dt)) for x in (k0n, k1n, cf, Ti)]
            ext, valid = _run_ref(lambda: erp.run_bundle_adjust_2_view(*a, 10), tdt)
            exto, valido = P.run_bundle_adjust_2_view(k0n.astype(dt), k1n.astype(dt), cf.astype(dt), Ti.astype(dt), 10)
            assert np.array_equal(_np(valid), valido), (case['name'], tag)
            e = _np(ext)
            if dt == np.float64:
                err = float(np.abs(e - exto).max()) if e.size else 0.0
                assert err < 1e-7, (case['name'], err)
                report['ba_%s_64_oracle_err' % case['name']] = err
            elif e.size:
                # fp32 dense-LU noise moves the free scale gauge: compare rotation and translation direction
                rerr = float(np.abs(e[:, :3, :3] - exto[:, :3, :3]).max())
                derr = float(np.abs(tdir(e) - tdir(exto)).max())
                report['ba_%s_32_oracle_rot_err' % case['name']] = rerr
                report['ba_%s_32_oracle_tdir_err' % case['name']] = derr
            out['ext' + tag] = e
            out['valid' + tag] = _np(valid)
        if out['ext32'].size:
            report['ba_%s_ref32_vs_ref64_abs' % case['name']] = float(np.abs(out['ext32'] - out['ext64']).max())
            report['ba_%s_ref32_vs_ref64_rot' % case['name']] = float(np.abs(out['ext32'][:, :3, :3] - out['ext64'][:, :3, :3]).max())
            report['ba_%s_ref32_vs_ref64_tdir' % case['name']] = float(np.abs(tdir(out['ext32']) - tdir(out['ext64'])).max())
        np.savez_compressed(os.path.join(OUT, 'pose_%s.npz' % case['name']), meta=json.dumps(case),
                            kpts0_norm=k0n, kpts1_norm=k1n, conf=cf, T_init=Ti, **out)
        print('ba', case['name'], 'ok; valid', out['valid64'].tolist(),
              {k.split(case['name'] + '_')[1]: v for k, v in report.items() if k.startswith('ba_%s_' % case['name'])})

    with open(os.path.join(OUT, 'pose_report.json'), 'w') as f:
        json.dump(report, f, indent=1)


if __name__ == '__main__':
    main()
