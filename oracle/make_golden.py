"""Generate tests/golden/matcher_*.npz by running the REAL reference matcher.

Run in the authoring container only (needs /root/reference):
    python -m oracle.make_golden
It (1) builds seeded weights/inputs with oracle/weights.py, (2) loads them into
the unmodified reference ``MultiViewMatcher`` (imported from /root/reference by
path), (3) stores the reference outputs as fixtures and (4) asserts that the
numpy oracle (oracle/matcher.py) agrees with the reference on every case, which
is what pins the oracle.  The fixtures travel to the GPU box; the reference
does not.
"""
import os
import sys
import json

import numpy as np
import torch

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

CASES = [
    # name, multi_frame, layer names, per-view keypoint counts, weight seed, input seed, correlated
    # every case carries a final_proj gain that gives the assignment realistic contrast (a trained model's
    # couplings are sharp; with gain 1 the random-weight couplings are flat and >90 % of the rows have a
    # top-2 margin below fp32 noise, which would make the match comparison vacuous)
    dict(name='pair_small_ragged', multi=False, layers=['self', 'cross'] * 2, counts=[60, 50], wseed=1, iseed=11, corr=False, gain=16.0),
    dict(name='pair_18l_128', multi=False, layers=['self', 'cross'] * 9, counts=[128, 128], wseed=2, iseed=12, corr=True, gain=6.0),
    dict(name='mv3_ragged', multi=True, layers=['self', 'cross', 'cross'] * 2, counts=[70, 64, 50], wseed=3, iseed=13, corr=False, gain=16.0),
    dict(name='mv5_28l_96', multi=True, layers=(['self'] + ['cross'] * 3) * 7, counts=[96] * 5, wseed=4, iseed=14, corr=True, gain=6.0),
    dict(name='pair3_mv_false', multi=False, layers=['self', 'cross'], counts=[40, 33, 47], wseed=5, iseed=15, corr=False, gain=16.0),
    # one deliberately flat case (gain 1): exercises the score tolerance where every row is a near-tie
    dict(name='pair_flat', multi=False, layers=['self', 'cross'] * 2, counts=[64, 64], wseed=7, iseed=17, corr=False, gain=1.0),
    dict(name='pair_18l_128_sharp', multi=False, layers=['self', 'cross'] * 9, counts=[128, 128], wseed=2, iseed=12, corr=True, gain=12.0),
    dict(name='mv5_28l_96_sharp', multi=True, layers=(['self'] + ['cross'] * 3) * 7, counts=[96] * 5, wseed=4, iseed=14, corr=True, gain=12.0),
    dict(name='mv4_ragged_sharp', multi=True, layers=(['self'] + ['cross'] * 2) * 3, counts=[90, 77, 64, 81], wseed=6, iseed=16, corr=False, gain=16.0),
]


def build_inputs(case):
    from oracle.weights import make_view_inputs, make_correlated_view_inputs
    if case['corr']:
        return make_correlated_view_inputs(case['iseed'], len(case['counts']), case['counts'][0])
    return make_view_inputs(case['iseed'], case['counts'])


def run_reference(case, sd_np, data_np, double=False):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from models.models.multi_view_matcher import MultiViewMatcher  # the unmodified reference
    torch.manual_seed(0)
    model = MultiViewMatcher({'multi_frame_matching': case['multi'], 'GNN_layers': case['layers'],
                              'conf_mlp': True}).eval()
    sd_t = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
    model.load_state_dict(sd_t, strict=True)
    if double:      # the reference's own code in double precision: yardstick for its fp32 arithmetic noise
        model = model.double()
    data = {k: ((torch.from_numpy(v).double() if double else torch.from_numpy(v)) if isinstance(v, np.ndarray) else v)
            for k, v in data_np.items()}
    with torch.no_grad():
        out = model(data)
    return {k: v.numpy() for k, v in out.items() if v is not None}


def main():
    from oracle.weights import make_state_dict
    from oracle.matcher import matcher_forward
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    report = {}
    for case in CASES:
        sd = make_state_dict(len(case['layers']), seed=case['wseed'], final_proj_gain=case.get('gain', 1.0))
        data = build_inputs(case)
        ref = run_reference(case, sd, data)
        ora = matcher_forward(sd, {'multi_frame_matching': case['multi'], 'GNN_layers': case['layers']}, data)
        stats = {}
        for k, v in ref.items():
            o = ora[k]
            if k.startswith('matches'):
                assert np.array_equal(v, o), (case['name'], k, int((v != o).sum()))
                stats[k] = int((v >= 0).sum())
            else:
                err = float(np.abs(v - o).max())
                assert err < 2e-4, (case['name'], k, err)
                stats[k] = err
        # top-2 margin of the assignment rows (tie sensitivity of the argmax); fp32 noise of the reference
        from tests.util import stable_rows
        ref64 = run_reference(case, sd, data, double=True)
        margins, st, tot, noise = [], 0, 0, 0.0
        for k, v in ref.items():
            if k.startswith('scores_'):
                inner = np.sort(v[:, :-1, :-1], axis=2)
                margins.append(float((inner[..., -1] - inner[..., -2]).min()))
                s0, s1 = stable_rows(v, 2e-3)
                st += int(s0.sum() + s1.sum()); tot += int(s0.size + s1.size)
                noise = max(noise, float(np.abs(v.astype(np.float64) - ref64[k]).max()))
        stats['min_top2_margin'] = min(margins)
        stats['stable_frac_tau_2e-3'] = st / tot
        stats['max_abs_ref32_vs_ref64'] = noise
        assert case['name'] == 'pair_flat' or st / tot >= 0.9, (case['name'], st / tot)
        report[case['name']] = stats
        np.savez_compressed(os.path.join(OUT, 'matcher_%s.npz' % case['name']),
                            meta=json.dumps({k: case[k] for k in case}), **ref)
        print(case['name'], 'ok; max |oracle-ref| =',
              max(v for k, v in stats.items() if isinstance(v, float) and k.startswith(('scores', 'conf', 'matching'))),
              'min top2 margin', stats['min_top2_margin'], 'stable', round(st / tot, 4), 'ref32-vs-ref64', noise)
    with open(os.path.join(OUT, 'matcher_report.json'), 'w') as f:
        json.dump(report, f, indent=1)


if __name__ == '__main__':
    main()
