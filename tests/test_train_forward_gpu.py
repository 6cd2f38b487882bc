"""TRAIN-MODE forward (models/train_forward.py, csrc/bn_train.cu) against the unmodified
reference MultiViewMatcher in .train() with full_output: batch-statistics BatchNorm, stacked views / per-view pairwise
path, running-statistics updates.  Fixtures: oracle/make_train_forward_golden.py (fp32 and fp64 runs of the reference)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _model(case, sd, full_output=True):
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    model = MultiViewMatcher({'multi_frame_matching': case['multi'], 'GNN_layers': case['layers'], 'conf_mlp': True,
                              'full_output': full_output})
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return model.cuda().train()


@pytest.mark.parametrize('name', ['mv3_64', 'pair_128'])
def test_train_forward_vs_reference_golden(name):
    from oracle.make_train_forward_golden import build, STATS
    z = np.load(os.path.join(GOLDEN, 'train_forward_%s.npz' % name))
    case = json.loads(str(z['meta']))
    sd, data_np = build(case)
    model = _model(case, sd)
    data = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    with torch.no_grad():
        res = model(data)
    keys = [k[5:] for k in z.files if k.startswith('f64__') and '__stat__' not in k and '__inter__' not in k]
    assert sorted(k for k, v in res.items() if v is not None) == sorted(keys)
    worst, fails = {}, []
    for k in sorted(keys):
        r32, r64 = z['f32__' + k], z['f64__' + k]
        got = res[k].cpu().numpy()
        assert got.shape == r64.shape, (k, got.shape, r64.shape)
        if k.startswith('matches'):
            # a keypoint's decision is compared where it is stable: the top-2 margins of its row AND of the matched
            # column (mutual check) and the distance of its score from the match threshold all exceed 10x the
            # reference's own fp32-vs-fp64 deviation of that coupling matrix
            from tests.util import stable_rows
            pair = k.split('_', 1)[1]
            view = k[len('matches'):].split('_')[0]
            Z64, Z32 = z['f64__scores_' + pair], z['f32__scores_' + pair]
            tau = 10.0 * float(np.abs(Z32.astype(np.float64) - Z64).max())
            first = view == pair.split('_')[0]
            st0, st1 = stable_rows(Z64, tau)
            inner = Z64[:, :-1, :-1]
            best = inner.max(2) if first else inner.max(1)
            stable = st0 if first else st1
            if model.match_threshold > 0:
                stable = stable & (np.abs(best - np.log(model.match_threshold)) > tau)
            bad = int(((got != r64) & stable).sum())
            print(name, k, 'mismatch %d of %d, on stable rows %d (stable %d, ref fp32-vs-fp64 mismatches %d)'
                  % (int((got != r64).sum()), got.size, bad, int(stable.sum()), int((r32 != r64).sum())))
            if bad > max(1, int(0.02 * stable.sum())):
                fails.append((k, bad))
            continue
        noise = float(np.abs(r32.astype(np.float64) - r64).max())
        err = float(np.abs(got.astype(np.float64) - r64).max())
        if k.startswith('matching_scores') or k.startswith('conf_scores'):
            # a flipped match zeroes its score / changes the inputs of the confidence head: compare where the match agrees
            mk = 'matches' + k.split('scores', 1)[1].lstrip('_') if k.startswith('matching') else 'matches%s_%s' % (k.split('_')[2], k.split('_', 2)[2])
            same = res[mk].cpu().numpy() == z['f64__' + mk]
            d = np.abs(got.astype(np.float64) - r64).reshape(same.shape)
            err = float(d[same].max()) if same.any() else 0.0
        kind = k.split('_')[0]
        worst[kind] = max(worst.get(kind, 0.0), err / max(noise, 1e-6))
        print(name, k, 'max err %.3g  reference fp32-vs-fp64 %.3g' % (err, noise))
        if err > max(6.0 * noise, 2e-4):
            fails.append((k, err, noise))
    assert not fails, fails
    for k in STATS:
        st = dict(model.state_dict())[k].cpu().numpy()
        r32, r64 = z['f32__stat__' + k], z['f64__stat__' + k]
        if k.endswith('num_batches_tracked'):
            assert int(st) == int(r64)
            continue
        noise = float(np.abs(r32.astype(np.float64) - r64).max())
        assert np.abs(st.astype(np.float64) - r64).max() <= max(4.0 * noise, 1e-5 * max(1.0, np.abs(r64).max())), k
    print(name, 'error / reference fp32 noise per output kind:', {k: round(v, 2) for k, v in worst.items()})


def test_train_forward_output_gating_and_eval_unchanged():
    """Without full_output the train branch returns only the couplings (multi_view_matcher.py:316-319); switching back to
    eval() runs the fused forward with the UPDATED running statistics folded in."""
    from oracle.make_train_forward_golden import build, CASES
    case = CASES[0]
    sd, data_np = build(case)
    model = _model(case, sd, full_output=False)
    data = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    res = model(data)
    assert sorted(res) == ['scores_0_1', 'scores_0_2', 'scores_1_2'] and res['scores_0_1'].requires_grad   # MatcherTrainFn graph
    with torch.no_grad():
        assert not model(data)['scores_0_1'].requires_grad
    rm = model.kenc.encoder[1].running_mean.clone()
    before = model.eval()(data)['scores_0_1'].clone()
    with torch.no_grad():
        model.train()(data)                                          # another train-mode call moves the statistics again
    assert not torch.equal(rm, model.kenc.encoder[1].running_mean)
    after = model.eval()(data)['scores_0_1']
    assert not torch.equal(before, after)                            # the packed weights were rebuilt from the new buffers


def test_full_output_with_autograd_on_the_kernels():
    """`full_output` + autograd: couplings with MatcherTrainFn's graph, match / confidence outputs beside them; backward
    fills every matcher parameter's gradient with finite values."""
    from oracle.make_train_forward_golden import build, CASES
    case = CASES[0]
    sd, data_np = build(case)
    model = _model(case, sd, full_output=True)
    data = {k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    res = model(data)
    assert res['scores_0_1'].requires_grad and not res['conf_scores_0_1'].requires_grad and res['matches0_0_1'].dtype == torch.int64
    sum(res['scores_%s' % k].square().mean() for k in ('0_1', '0_2', '1_2')).backward()
    named = dict(model.named_parameters())
    assert all(named[k].grad is not None and torch.isfinite(named[k].grad).all() for k in named if not k.startswith('conf_mlp'))
