"""TRAIN-MODE forward of the matcher: the `self.training` branches of
models/models/multi_view_matcher.py -- stacked views (:219-226), KeypointEncoder / AttentionalPropagation /
ConfidenceMLP with BatchNorm1d in training mode, i.e. BATCH statistics over all B*T*N points of the call and
running-statistics updates (:8-22), combined cross attention over the other views (:65-86), `full_output` gating
(:187,287,316-319).

This is the functional first slice, not the fast path: the eval forward runs as one fused C call on packed weights
with the BatchNorms folded into the convolutions; batch statistics cannot be folded, so this forward sequences the
stage-level kernels (tcgen05 GEMMs, the fp16x3 attention, the cluster Sinkhorn, match extraction) from Python and
inserts `mvm_batchnorm_train` between them.  It has NO backward: the result tensors do not require grad (the backward
kernels of attention / GEMMs are not built), so it serves loss evaluation and BatchNorm statistics
collection, not optimisation."""
import ctypes as C

import torch

from .. import _lib
from .. import ops


def _round_up(x, m):
    return (x + m - 1) // m * m


def _head_perm(device):
    """new channel h*64 + d  <-  reference channel d*4 + h (superglue.py:106: .view(B, 64, 4, N))."""
    return torch.arange(256, device=device).view(64, 4).t().reshape(-1)


def _lin(x, w, b, relu=False, a2=None, residual=None, alpha=1.0):
    """Conv1d(k=1) on point-major rows: tcgen05 fp16x3 when the shape allows, fp32 CUDA cores otherwise."""
    w = w.detach().float().contiguous()
    b = None if b is None else b.detach().float().contiguous()
    K = x.shape[1] + (a2.shape[1] if a2 is not None else 0)
    N = w.shape[0]
    if K % 64 == 0 and N % 128 == 0 and x.shape[1] % 64 == 0:
        return ops.linear(x, w, bias=b, a2=a2, residual=residual, relu=relu, alpha=alpha, tc_passes='h16')
    if K % 16:                                   # first encoder layer (3 inputs), confidence score branch (1 input)
        assert a2 is None
        pad = _round_up(K, 16) - K
        x = torch.nn.functional.pad(x, (0, pad)).contiguous()
        w = torch.nn.functional.pad(w, (0, pad)).contiguous()
    return ops.linear(x, w, bias=b, a2=a2, residual=residual, relu=relu, alpha=alpha, tc_passes=0)


def _bn(x, bn, n_pad, n_valid, relu=True, groups=1):
    """nn.BatchNorm1d in training mode, in place on x [rows, C]; updates the module's running statistics.  groups > 1:
    the view slots s = g (mod groups) are normalised one group after the other (one BatchNorm call per view, as the
    pairwise train path makes them)."""
    lib = _lib.lib()
    rows, Cc = x.shape
    assert x.is_contiguous() and bn.weight is not None and bn.track_running_stats
    momentum = 0.1 if bn.momentum is None else bn.momentum
    ws = torch.empty(3 * Cc, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        for g in range(groups):
            _lib.check(lib.mvm_batchnorm_train(_lib.ptr(x), rows, Cc, x.stride(0), n_pad, n_valid, groups, g,
                                               _lib.ptr(bn.weight.detach().float().contiguous()),
                                               _lib.ptr(bn.bias.detach().float().contiguous()), float(bn.eps), int(relu),
                                               _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), float(momentum),
                                               _lib.ptr(ws), _lib.stream_ptr()), 'mvm_batchnorm_train')
            bn.num_batches_tracked += 1
    return x


def _conv(seq, i):
    return seq[i].weight[:, :, 0], seq[i].bias


def train_forward(model, data, view_ids=None, debug=None):
    """-> result dict of MultiViewMatcher.multi_match in training mode for the views `view_ids` (default: all); with
    `view_ids` = a pair, of MultiViewMatcher.match (pairwise mode: keypoint encoder and every GNN layer are applied to
    one view at a time, so their BatchNorms see one view per call)."""
    cfg = model.config
    g_bn = 1 if view_ids is None else len(view_ids)
    lib = _lib.lib()
    ids = list(range(len(data['ids']))) if view_ids is None else list(view_ids)
    T = len(ids)
    views = [model._view(data, i) for i in ids]
    dev = views[0][0].device
    if dev.type != 'cuda':
        raise _lib.MvmError('MultiViewMatcher needs CUDA tensors (no CPU fallback)')
    B, N = views[0][0].shape[:2]
    assert N > 0 and all(v[0].shape[:2] == (B, N) for v in views), 'training uses a fixed number of keypoints per view'
    n_pad = max(64, _round_up(N, 64))
    rows = B * T * n_pad
    h_img, w_img = data['image0'].shape[-2:]
    # ---- gather the views (one launch) -> point-major rows, slot = b * T + t
    kp = torch.empty(B, T, n_pad, 2, dtype=torch.float32, device=dev)
    sc = torch.empty(B, T, n_pad, dtype=torch.float32, device=dev)
    de = torch.empty(B, T, 256, n_pad, dtype=torch.float32, device=dev)
    views = [tuple(x.detach().contiguous() for x in v) for v in views]
    ptrs = [(C.c_void_p * T)(*[v[i].data_ptr() for v in views]) for i in range(3)]
    with torch.cuda.device(dev):
        _lib.check(lib.mvm_pack_views(ptrs[0], ptrs[1], ptrs[2], (C.c_int * T)(*([N] * T)), B, T, n_pad, _lib.ptr(kp),
                                      _lib.ptr(sc), _lib.ptr(de), _lib.stream_ptr()), 'mvm_pack_views')
    x_desc = de.permute(0, 1, 3, 2).reshape(rows, 256).contiguous()
    # ---- normalize_keypoints (superglue.py:65-72) + KeypointEncoder (multi_view_matcher.py:24-37)
    center = torch.tensor([w_img / 2.0, h_img / 2.0], dtype=torch.float32, device=dev)
    scaling = 0.7 * float(max(w_img, h_img))
    inp = torch.zeros(rows, 16, dtype=torch.float32, device=dev)
    inp[:, 0:2] = ((kp - center) / scaling).reshape(rows, 2)
    inp[:, 2] = sc.reshape(rows)
    enc = model.kenc.encoder
    h = inp
    for i in (0, 3, 6, 9):
        w, b = _conv(enc, i)
        if i == 0:
            w = torch.nn.functional.pad(w, (0, 13))
        h = _bn(_lin(h, w, b), enc[i + 1], n_pad, N, groups=g_bn)
    kenc_out = _lin(h, *_conv(enc, 12))
    x = kenc_out + x_desc
    if debug is not None:
        debug['kenc'] = kenc_out.view(B, T, n_pad, 256)[:, :, :N].clone()
    # ---- MultiFrameAttentionalGNN, train branch (multi_view_matcher.py:65-86)
    perm = _head_perm(dev)
    counts = [N] * T
    for layer, name in zip(model.gnn.layers, model.gnn.names):
        attn = layer.attn
        wqkv = torch.cat([attn.proj[i].weight[:, :, 0][perm] for i in range(3)], 0)
        bqkv = torch.cat([attn.proj[i].bias[perm] for i in range(3)], 0)
        qkv = _lin(x, wqkv, bqkv)
        msg = ops.attention(qkv.view(B * T, n_pad, 768), B, T, counts, 1 if name == 'cross' else 0, tc_passes='h3')
        merged = _lin(msg.view(rows, 256), attn.merge.weight[:, :, 0][:, perm], attn.merge.bias)
        hid = _lin(x, *_conv(layer.mlp, 0), a2=merged)
        _bn(hid, layer.mlp[1], n_pad, N, groups=g_bn)
        delta = _lin(hid, *_conv(layer.mlp, 3))
        if debug is not None and 'layer0_delta' not in debug:
            debug['layer0_delta'] = delta.view(B, T, n_pad, 256)[:, :, :N].clone()
        x = x + delta
    # ---- final projection, scores, optimal transport (multi_view_matcher.py:275-285)
    if debug is not None:
        debug['gnn'] = x.view(B, T, n_pad, 256)[:, :, :N].clone()
    md = _lin(x, model.final_proj.weight[:, :, 0], model.final_proj.bias).view(B, T, n_pad, 256)
    result = {}
    full = bool(cfg['full_output'])
    slot = {v: s for s, v in enumerate(ids)}
    for id1 in ids:
        for id0 in ids:
            if id0 >= id1:
                continue
            a, b_ = slot[id0], slot[id1]
            m0 = md[:, a, :N].contiguous()                                    # [B, N, 256]
            m1 = md[:, b_, :N].contiguous()
            raw = torch.stack([_lin(m0[i], m1[i], None, alpha=1.0 / 16.0) for i in range(B)], 0)      # [B, N, N]
            # log-domain kernel: couplings of an untrained / early-training network span thousands of nats, beyond the
            # range of the scaling-domain production kernels (which serve the eval path)
            Z = ops.log_optimal_transport(raw, float(model.bin_score), int(cfg['sinkhorn_iterations']), kernel='log')
            key = '{}_{}'.format(id0, id1)
            result['scores_' + key] = Z
            if not full:
                continue
            i0, i1, s0, s1 = ops.extract_matches(Z, model.match_threshold)
            conf = None
            if cfg['conf_mlp']:
                # inputs of ConfidenceMLP (multi_view_matcher.py:302-306): -1 wraps to the last keypoint / dustbin column
                bi = torch.arange(B, device=dev).unsqueeze(-1).expand(B, N)
                m1g = m1[bi, i0]                                               # [B, N, 256]
                add = Z[bi, torch.arange(N, device=dev).unsqueeze(0).expand(B, N), i0].reshape(B * N, 1)
                cm = model.conf_mlp
                f = _lin(m0.reshape(B * N, 256), *_conv(cm.layers_f, 0), a2=m1g.reshape(B * N, 256).contiguous())
                _bn(f, cm.layers_f[1], N, N)
                f = _bn(_lin(f, *_conv(cm.layers_f, 3)), cm.layers_f[4], N, N)
                c = _bn(_lin(add.contiguous(), *_conv(cm.layers_c, 0)), cm.layers_c[1], N, N)
                c = _bn(_lin(c, *_conv(cm.layers_c, 3)), cm.layers_c[4], N, N)
                w_last, b_last = _conv(cm.layers, 0)
                logit = _lin((f + c).contiguous(), torch.nn.functional.pad(w_last, (0, 0, 0, 15)), torch.nn.functional.pad(b_last, (0, 15)))
                conf = torch.sigmoid(logit[:, :1]).reshape(B, N, 1)
            result['matches{}_{}'.format(id0, key)] = i0
            result['matches{}_{}'.format(id1, key)] = i1
            result['matching_scores{}_{}'.format(id0, key)] = s0
            result['matching_scores{}_{}'.format(id1, key)] = s1
            result['conf_scores_' + key] = conf
    return result
