"""Host-side orchestration of the TRAINING step (models/train_forward.py: MatcherTrainFn forward + backward) against the
gradients of the unmodified reference (tests/golden/train_backward_*.npz, oracle/make_train_backward_golden.py), with the
stage kernels replaced by the float64 torch stand-ins of oracle/train_ops.py -- runs without a GPU.  What it pins: which
tensors are saved, the head permutation of q/k/v/merge and its inverse on the gradients, the concat / residual routing,
the per-view BatchNorm groups of the pairwise path, the pair loops and the accumulation over pairs.  The kernels
themselves are checked against the same stand-ins on the GPU (tests/test_train_backward_gpu.py)."""
import json
import os

import numpy as np
import pytest
import torch

from tests import emul_ops

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
PATCHED = ['pack_views', 'pair_scores', 'linear', 'attention', 'attention_backward', 'transpose_split', 'linear_presplit', 'linear_presplit_splitk', 'colsum',
           'batchnorm_train', 'batchnorm_train_backward', 'sinkhorn_train_forward', 'sinkhorn_train_backward']


def match_loss(log_p, idx, w):
    """helpers.py:228-241 restated (the product's compute_match_loss is a CUDA kernel)."""
    bs, ft, _ = log_p.shape
    l0 = -log_p.reshape(bs * ft, ft)[range(bs * ft), idx[:, 0].reshape(-1)]
    l1 = -log_p.transpose(1, 2).reshape(bs * ft, ft)[range(bs * ft), idx[:, 1].reshape(-1)]
    return (torch.dot(l0, w[:, 0].reshape(-1)) + torch.dot(l1, w[:, 1].reshape(-1))) / bs


def check_gradients(model, z, tol_noise, tol_rel, what, return_ratios=False):
    from oracle.make_train_backward_golden import sample_index
    names = [k[len('grad__'):] for k in z.files if k.startswith('grad__')]
    params = dict(model.named_parameters())
    assert sorted(n for n, p in params.items() if p.grad is not None) == sorted(names)
    scale = max(float(np.abs(z['grad__' + k]).max()) for k in names)
    worst = (0.0, None)
    ratios = []
    for k in names:
        g = params[k].grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
        ref = z['grad__' + k].astype(np.float64)
        got = g[sample_index(k, g.size)]
        tol = tol_noise * float(z['noise__' + k]) + tol_rel * max(float(np.abs(ref).max()), 1e-3 * scale)
        err = float(np.abs(got - ref).max())
        if err / tol > worst[0]:
            worst = (err / tol, k)
        ratios.append(err / max(float(z['noise__' + k]), 1e-30))
        assert err <= tol, (what, k, err, tol, float(z['noise__' + k]), float(np.abs(ref).max()))
        norm = float(np.linalg.norm(g))
        assert abs(norm - float(z['norm__' + k])) <= tol_noise * float(z['noise__' + k]) * np.sqrt(g.size) + 10 * tol_rel * max(float(z['norm__' + k]), 1e-3 * scale), k
    print(what, 'gradient error / reference fp32 noise over the parameters: median %.2f, 90 %% %.2f, max %.2f'
          % (np.median(ratios), np.percentile(ratios, 90), max(ratios)))
    if return_ratios:
        return worst, ratios
    return worst


@pytest.mark.parametrize('name', ['mv3_64', 'mv4_100', 'pair_96'])
def test_train_step_orchestration_vs_reference(name, monkeypatch):
    from oracle.make_train_backward_golden import build
    from e2e_multi_view_matching_b200 import ops, _lib
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    for f in PATCHED:
        monkeypatch.setattr(ops, f, getattr(emul_ops, f))
    monkeypatch.setattr(_lib, 'require_cuda', lambda device, what: None)
    z = np.load(os.path.join(GOLDEN, 'train_backward_%s.npz' % name))
    case = json.loads(str(z['meta']))
    data_np, sd = build(case)
    model = MultiViewMatcher({'multi_frame_matching': case['multi'], 'GNN_layers': case['layers'], 'conf_mlp': True,
                              'full_output': False})
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.train()
    data = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    result = model(data)
    assert sorted(result) == sorted('scores_%d_%d' % (a, b) for b in range(case['views']) for a in range(b))
    loss = 0.0
    for b in range(case['views']):
        for a in range(b):
            key = '%d_%d' % (a, b)
            assert result['scores_' + key].requires_grad
            loss = loss + match_loss(result['scores_' + key], data['gt_indices_' + key], data['gt_weights_' + key])
    assert abs(float(loss) - float(z['loss_f64'])) <= 4 * abs(float(z['loss_f32']) - float(z['loss_f64'])) + 1e-6 * abs(float(z['loss_f64']))
    loss.backward()
    # the stand-ins compute in float64 and hand float32 tensors on: the error budget is the reference's own fp32 noise
    worst = check_gradients(model, z, tol_noise=3.0, tol_rel=2e-5, what=name)
    print(name, 'worst gradient error / tolerance: %.3f at %s' % worst)


def test_no_grad_train_forward_has_no_graph(monkeypatch):
    from oracle.make_train_backward_golden import build, CASES
    from e2e_multi_view_matching_b200 import ops, _lib
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    for f in PATCHED:
        monkeypatch.setattr(ops, f, getattr(emul_ops, f))
    monkeypatch.setattr(_lib, 'require_cuda', lambda device, what: None)
    case = CASES[0]
    data_np, sd = build(case)
    model = MultiViewMatcher({'multi_frame_matching': True, 'GNN_layers': case['layers'], 'conf_mlp': True, 'full_output': False})
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.train()
    data = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    with torch.no_grad():
        res = model(data)
    assert not res['scores_0_1'].requires_grad


@pytest.mark.parametrize('name', ['mv3_64', 'pair_128'])
def test_train_forward_full_output_orchestration_vs_reference(name, monkeypatch):
    """The train-mode forward with `full_output` (match extraction + ConfidenceMLP with batch-statistics BatchNorm, the
    per-view BatchNorm groups of the pairwise path, running statistics) on the stand-ins against the unmodified reference
    in .train() (tests/golden/train_forward_*.npz)."""
    from oracle.make_train_forward_golden import build, STATS
    from tests.util import stable_rows
    from e2e_multi_view_matching_b200 import ops, _lib
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    for f in PATCHED + ['extract_matches']:
        monkeypatch.setattr(ops, f, getattr(emul_ops, f))
    monkeypatch.setattr(_lib, 'require_cuda', lambda device, what: None)
    z = np.load(os.path.join(GOLDEN, 'train_forward_%s.npz' % name))
    case = json.loads(str(z['meta']))
    sd, data_np = build(case)
    model = MultiViewMatcher({'multi_frame_matching': case['multi'], 'GNN_layers': case['layers'], 'conf_mlp': True, 'full_output': True})
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.train()
    data = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    with torch.no_grad():
        res = model(data)
    keys = [k[5:] for k in z.files if k.startswith('f64__') and '__stat__' not in k and '__inter__' not in k]
    assert sorted(k for k, v in res.items() if v is not None) == sorted(keys)
    for k in sorted(keys):
        r32, r64 = z['f32__' + k], z['f64__' + k]
        got = res[k].numpy()
        assert got.shape == r64.shape, k
        if k.startswith('scores_'):
            noise = float(np.abs(r32.astype(np.float64) - r64).max())
            assert np.abs(got - r64).max() <= 3 * noise + 1e-5, (k, float(np.abs(got - r64).max()), noise)
        elif k.startswith('matches'):
            pair = k.split('_', 1)[1]
            first = k[len('matches'):].split('_')[0] == pair.split('_')[0]
            Z64, Z32 = z['f64__scores_' + pair], z['f32__scores_' + pair]
            st0, st1 = stable_rows(Z64, 10.0 * float(np.abs(Z32.astype(np.float64) - Z64).max()))
            st = st0 if first else st1
            assert np.array_equal(got[st], r64[st]), (k, int((got[st] != r64[st]).sum()))
        elif k.startswith('conf_scores'):
            pair = k[len('conf_scores_'):]
            same = res['matches%s_%s' % (pair.split('_')[0], pair)].numpy() == z['f64__matches%s_%s' % (pair.split('_')[0], pair)]
            noise = float(np.abs(r32.astype(np.float64) - r64)[..., 0][same].max())
            assert np.abs(got.astype(np.float64) - r64)[..., 0][same].max() <= 3 * noise + 1e-5, k
    state = dict(model.state_dict())
    for k in STATS:
        st, r32, r64 = state[k].numpy(), z['f32__stat__' + k], z['f64__stat__' + k]
        if k.endswith('num_batches_tracked'):
            assert int(st) == int(r64), k
        else:
            assert np.abs(st.astype(np.float64) - r64).max() <= 3 * float(np.abs(r32.astype(np.float64) - r64).max()) + 1e-6, k


def test_full_output_with_autograd_keeps_the_extra_outputs(monkeypatch):
    """`full_output` with autograd enabled: the couplings carry MatcherTrainFn's graph, the match / confidence outputs come
    back beside them without one, and the backward still reaches every matcher parameter."""
    from oracle.make_train_backward_golden import build, CASES
    from e2e_multi_view_matching_b200 import ops, _lib
    from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher
    for f in PATCHED + ['extract_matches']:
        monkeypatch.setattr(ops, f, getattr(emul_ops, f))
    monkeypatch.setattr(_lib, 'require_cuda', lambda device, what: None)
    case = CASES[0]
    data_np, sd = build(case)
    model = MultiViewMatcher({'multi_frame_matching': True, 'GNN_layers': case['layers'], 'conf_mlp': True, 'full_output': True})
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model.train()
    data = {k: (torch.from_numpy(v) if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    res = model(data)
    assert res['scores_0_1'].requires_grad and not res['conf_scores_0_1'].requires_grad
    assert res['matches0_0_1'].dtype == torch.int64 and res['matching_scores2_0_2'].shape == res['matches2_0_2'].shape
    assert sorted(k.split('_')[0] for k in res) == sorted(['scores'] * 3 + ['matches0', 'matches1', 'matches0', 'matches2', 'matches1', 'matches2'] +
                                                           ['matching'] * 6 + ['conf'] * 3)
    sum(res['scores_%s' % k].sum() for k in ('0_1', '0_2', '1_2')).backward()
    named = dict(model.named_parameters())
    assert all(named[k].grad is not None for k in named if not k.startswith('conf_mlp'))
    assert all(named[k].grad is None for k in named if k.startswith('conf_mlp'))        # no graph through the confidences


def test_landmark_gt_matches_equals_the_golden_generator():
    """synthetic.landmark_gt_matches (ground truth of the training bench) == the per-item construction the reference goldens
    were generated with (oracle/make_validation_golden.gt_from_landmarks: helpers.py:190-213 weights)."""
    from e2e_multi_view_matching_b200.synthetic import landmark_gt_matches, make_scene_tuple_inputs
    from oracle.make_validation_golden import gt_from_landmarks
    d = make_scene_tuple_inputs(3, n_views=3, n_kpts=100, batch=2)
    for a, b in ((0, 1), (0, 2), (1, 2)):
        idx, w = landmark_gt_matches(d['landmark%d' % a], d['landmark%d' % b])
        for i in range(2):
            ri, rw = gt_from_landmarks(d['landmark%d' % a][i], d['landmark%d' % b][i])
            assert np.array_equal(ri, idx[i]) and np.allclose(rw, w[i], rtol=1e-6)
        assert idx.dtype == np.int64 and w.dtype == np.float32 and (idx[:, :, -1] == -1).all()
