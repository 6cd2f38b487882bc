"""Stage-by-stage error of the training step on the GPU: the kernels against the float64 stand-ins (oracle/train_ops.py,
run on the CPU of the same box) on one of the reference golden cases -- forward activations per layer, then every
gradient of the backward per layer.  Usage: python tools/train_diag.py [case]   (default mv3_64)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import emul_ops  # noqa: E402
from tests.test_train_host_logic import PATCHED, match_loss  # noqa: E402
from oracle.make_train_backward_golden import build  # noqa: E402
from e2e_multi_view_matching_b200 import ops, _lib  # noqa: E402
from e2e_multi_view_matching_b200.models.multi_view_matcher import MultiViewMatcher  # noqa: E402
from e2e_multi_view_matching_b200.models import train_forward as TF  # noqa: E402


def run(case, sd, data_np, device):
    model = MultiViewMatcher({'multi_frame_matching': case['multi'], 'GNN_layers': case['layers'], 'conf_mlp': True, 'full_output': False})
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    model = model.to(device).train()
    data = {k: (torch.from_numpy(v).to(device) if isinstance(v, np.ndarray) else v) for k, v in data_np.items()}
    fdbg = {}
    with torch.no_grad():
        import copy
        TF._forward(copy.deepcopy(model), data, None if case['multi'] else [0, 1], save=False, debug=fdbg)
    model._train_debug = {}
    res = model(data)
    loss = 0.0
    for b in range(case['views']):
        for a in range(b):
            key = '%d_%d' % (a, b)
            loss = loss + match_loss(res['scores_' + key], data['gt_indices_' + key], data['gt_weights_' + key])
    loss.backward()
    return fdbg, model._train_debug, {k: v.detach() for k, v in res.items()}, float(loss)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)), float(b.abs().max())


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'mv3_64'
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'train_backward_%s.npz' % name))
    case = json.loads(str(z['meta']))
    data_np, sd = build(case)
    gf, gb, gres, gl = run(case, sd, data_np, 'cuda')
    torch.cuda.synchronize()
    saved = {f: getattr(ops, f) for f in PATCHED}
    rc = _lib.require_cuda
    for f in PATCHED:
        setattr(ops, f, getattr(emul_ops, f))
    _lib.require_cuda = lambda device, what: None
    try:
        cf, cb, cres, cl = run(case, sd, data_np, 'cpu')
    finally:
        for f, v in saved.items():
            setattr(ops, f, v)
        _lib.require_cuda = rc
    print('loss gpu %.6f  stand-ins %.6f  reference fp64 %.6f fp32 %.6f' % (gl, cl, float(z['loss_f64']), float(z['loss_f32'])))
    for i, (a, b) in enumerate(zip(gf['x_layers'], cf['x_layers'])):
        print('forward x after layer %d (%s): rel err %.2e (|x| max %.3g)' % ((i, case['layers'][i]) + rel(a, b)))
    for k in sorted(gres):
        print('forward %s: abs err %.3g' % (k, float((gres[k].double().cpu() - cres[k].double()).abs().max())))
    for k in ('g_mdesc', 'g_gnn'):
        print('backward %s: rel err %.2e (max %.3g)' % ((k,) + rel(gb[k], cb[k])))
    L = len(case['layers'])
    for j, (a, b) in enumerate(zip(gb['layers'], cb['layers'])):
        i = L - 1 - j
        print('backward layer %d (%s): ' % (i, case['layers'][i]) + '  '.join('%s %.2e' % (k, rel(a[k], b[k])[0]) for k in ('g_hid', 'g_merged', 'g_msg', 'g_qkv', 'gx')))
    print('backward g_kenc: rel err %.2e (max %.3g)' % rel(gb['g_kenc'], cb['g_kenc']))


if __name__ == '__main__':
    main()
