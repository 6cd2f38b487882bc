"""numpy fp32 restatement of the reference matcher forward (eval mode).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Checked against the real
reference by oracle/make_golden.py (fixtures: tests/golden/matcher_*.npz).

Every function cites the reference lines it follows (paths relative to
/root/reference).  Op order mirrors the reference (conv, then BatchNorm in eval
mode, then ReLU) -- nothing is folded here; folding is the product's business.
Layout is the reference's channel-first [B, C, N].
"""
import numpy as np

from .weights import BN_EPS

F32 = np.float32


def _conv1d(sd, key, x):
    """nn.Conv1d(kernel_size=1, bias=True) on [B,C,N]  (superglue.py:56-57)."""
    w = sd[key + '.weight'][:, :, 0]
    b = sd[key + '.bias']
    return (np.matmul(w[None], x) + b[None, :, None]).astype(F32)


def _bn_eval(sd, key, x):
    """nn.BatchNorm1d in eval mode (superglue.py:59-60)."""
    mean = sd[key + '.running_mean'][None, :, None]
    var = sd[key + '.running_var'][None, :, None]
    g = sd[key + '.weight'][None, :, None]
    b = sd[key + '.bias'][None, :, None]
    return ((x - mean) / np.sqrt(var + F32(BN_EPS)) * g + b).astype(F32)


def mlp(sd, prefix, channels, x, last_layer=True):
    """MLP factory semantics: Conv1d [+BN +ReLU] stacks; with last_layer=False the
    BN+ReLU also follow the final conv (multi_view_matcher.py:8-22,
    superglue.py:51-62).  Sequential indices: conv at 3*(i-1), BN at 3*(i-1)+1."""
    n = len(channels)
    for i in range(1, n):
        x = _conv1d(sd, '%s.%d' % (prefix, 3 * (i - 1)), x)
        if i < (n - 1 if last_layer else n):
            x = _bn_eval(sd, '%s.%d' % (prefix, 3 * (i - 1) + 1), x)
            x = np.maximum(x, F32(0))
    return x


def normalize_keypoints(kpts, image_shape):
    """superglue.py:65-72."""
    _, _, height, width = image_shape
    size = np.array([[width, height]], dtype=F32)
    center = size / F32(2)
    scaling = size.max(1, keepdims=True) * F32(0.7)
    return ((kpts - center[:, None, :]) / scaling[:, None, :]).astype(F32)


def keypoint_encoder(sd, kpts, scores, kenc_layers=(32, 64, 128, 256), desc_dim=256):
    """multi_view_matcher.py:24-37 (cat(x,y,score) -> MLP)."""
    inputs = np.concatenate([kpts.transpose(0, 2, 1), scores[:, None, :]], axis=1).astype(F32)
    return mlp(sd, 'kenc.encoder', [3] + list(kenc_layers) + [desc_dim], inputs)


def attention(query, key, value):
    """superglue.py:87-91; q [B,64,4,N], k/v [B,64,4,M]."""
    dim = query.shape[1]
    scores = np.einsum('bdhn,bdhm->bhnm', query, key).astype(F32) / F32(dim ** .5)
    scores = scores - scores.max(-1, keepdims=True)
    e = np.exp(scores)
    prob = (e / e.sum(-1, keepdims=True)).astype(F32)
    return np.einsum('bhnm,bdhm->bdhn', prob, value).astype(F32)


def multi_headed_attention(sd, prefix, x, source, num_heads=4):
    """superglue.py:94-109.  .view(B, dim, heads, N): channel c = d*heads + h."""
    b, d_model, _ = x.shape
    dim = d_model // num_heads
    q = _conv1d(sd, prefix + '.proj.0', x).reshape(b, dim, num_heads, -1)
    k = _conv1d(sd, prefix + '.proj.1', source).reshape(b, dim, num_heads, -1)
    v = _conv1d(sd, prefix + '.proj.2', source).reshape(b, dim, num_heads, -1)
    out = attention(q, k, v)
    return _conv1d(sd, prefix + '.merge', out.reshape(b, d_model, -1))


def attentional_propagation(sd, prefix, x, source):
    """superglue.py:112-121."""
    d = x.shape[1]
    message = multi_headed_attention(sd, prefix + '.attn', x, source)
    return mlp(sd, prefix + '.mlp', [2 * d, 2 * d, d], np.concatenate([x, message], axis=1))


def attentional_gnn(sd, names, desc0, desc1):
    """Pair GNN, superglue.py:124-140."""
    for l, name in enumerate(names):
        p = 'gnn.layers.%d' % l
        if name == 'cross':
            src0, src1 = desc1, desc0
        else:
            src0, src1 = desc0, desc1
        delta0 = attentional_propagation(sd, p, desc0, src0)
        delta1 = attentional_propagation(sd, p, desc1, src1)
        desc0, desc1 = desc0 + delta0, desc1 + delta1
    return desc0, desc1


def multi_frame_gnn(sd, names, desc, ids):
    """Multi-view GNN, eval branch (multi_view_matcher.py:87-100): cross source =
    concatenation of the other views in ascending id; deltas applied after all
    views of the layer are computed."""
    desc = list(desc)
    ids = sorted(ids)
    for l, name in enumerate(names):
        p = 'gnn.layers.%d' % l
        if name == 'cross':
            delta = {}
            for i in ids:
                src = np.concatenate([desc[j] for j in ids if j != i], axis=-1)
                delta[i] = attentional_propagation(sd, p, desc[i], src)
            for i in ids:
                desc[i] = desc[i] + delta[i]
        else:
            for i in ids:
                desc[i] = desc[i] + attentional_propagation(sd, p, desc[i], desc[i])
    return desc


def _logsumexp(x, axis):
    m = x.max(axis=axis, keepdims=True)
    return (np.log(np.exp(x - m).sum(axis=axis, keepdims=True)) + m).squeeze(axis).astype(F32)


def log_optimal_transport(scores, alpha, iters):
    """superglue.py:143-172 (dustbin augmentation, 100x {row LSE, col LSE},
    Z + u + v - norm)."""
    b, m, n = scores.shape
    alpha = F32(alpha)
    Z = np.empty((b, m + 1, n + 1), F32)
    Z[:, :m, :n] = scores
    Z[:, :m, n] = alpha
    Z[:, m, :] = alpha
    norm = F32(-np.log(F32(m + n)))
    log_mu = np.concatenate([np.full(m, norm, F32), [np.log(F32(n)) + norm]]).astype(F32)
    log_nu = np.concatenate([np.full(n, norm, F32), [np.log(F32(m)) + norm]]).astype(F32)
    u = np.zeros((b, m + 1), F32)
    v = np.zeros((b, n + 1), F32)
    for _ in range(iters):
        u = log_mu[None] - _logsumexp(Z + v[:, None, :], axis=2)
        v = log_nu[None] - _logsumexp(Z + u[:, :, None], axis=1)
    return (Z + u[:, :, None] + v[:, None, :] - norm).astype(F32)


def extract_matches(scores, match_threshold=0.0):
    """multi_view_matcher.py:288-300 (threshold 0.; SuperGlue class uses 0.2,
    superglue.py:268-278).  Returns int64 indices with -1, fp32 scores."""
    inner = scores[:, :-1, :-1]
    idx0 = inner.argmax(2)
    idx1 = inner.argmax(1)
    max0 = inner.max(2)
    b, m = idx0.shape
    n = idx1.shape[1]
    ar0 = np.arange(m)[None]
    ar1 = np.arange(n)[None]
    mutual0 = ar0 == np.take_along_axis(idx1, idx0, 1)
    mutual1 = ar1 == np.take_along_axis(idx0, idx1, 1)
    ms0 = np.where(mutual0, np.exp(max0), F32(0)).astype(F32)
    ms1 = np.where(mutual1, np.take_along_axis(ms0, idx1, 1), F32(0)).astype(F32)
    valid0 = mutual0 & (ms0 > F32(match_threshold))
    valid1 = mutual1 & np.take_along_axis(valid0, idx1, 1)
    i0 = np.where(valid0, idx0, -1).astype(np.int64)
    i1 = np.where(valid1, idx1, -1).astype(np.int64)
    return i0, i1, ms0, ms1


def confidence_mlp(sd, mdesc0, mdesc1, scores, indices0):
    """multi_view_matcher.py:39-53 and call site :302-306.  -1 wraps to the last
    keypoint of view 1 / selects the dustbin column of scores."""
    b, d, n0 = mdesc0.shape
    bi = np.arange(b)[:, None]
    add = scores[bi, np.arange(n0)[None], indices0][:, None, :].astype(F32)  # [B,1,N0]
    m1 = mdesc1.transpose(0, 2, 1)[bi, indices0].transpose(0, 2, 1)          # [B,256,N0]
    inputs = np.concatenate([mdesc0, m1], axis=1)
    out_f = mlp(sd, 'conf_mlp.layers_f', [2 * d, 2 * d, d], inputs, last_layer=False)
    out_c = mlp(sd, 'conf_mlp.layers_c', [1, d, d], add, last_layer=False)
    z = mlp(sd, 'conf_mlp.layers', [d, 1], out_f + out_c)
    return (F32(1) / (F32(1) + np.exp(-z))).astype(F32).transpose(0, 2, 1)   # [B,N0,1]


def _pair_head(sd, cfg, desc0, desc1, id0, id1, result):
    """final_proj + score einsum + OT + matches + conf (multi_view_matcher.py:275-315)."""
    d = cfg['descriptor_dim']
    mdesc0 = _conv1d(sd, 'final_proj', desc0)
    mdesc1 = _conv1d(sd, 'final_proj', desc1)
    scores = np.einsum('bdn,bdm->bnm', mdesc0, mdesc1).astype(F32) / F32(d ** .5)
    scores = log_optimal_transport(scores, sd['bin_score'], cfg['sinkhorn_iterations'])
    i0, i1, ms0, ms1 = extract_matches(scores, cfg.get('match_threshold', 0.0))
    result['matches%d_%d_%d' % (id0, id0, id1)] = i0
    result['matches%d_%d_%d' % (id1, id0, id1)] = i1
    result['matching_scores%d_%d_%d' % (id0, id0, id1)] = ms0
    result['matching_scores%d_%d_%d' % (id1, id0, id1)] = ms1
    result['scores_%d_%d' % (id0, id1)] = scores
    if cfg.get('conf_mlp', True):
        result['conf_scores_%d_%d' % (id0, id1)] = confidence_mlp(sd, mdesc0, mdesc1, scores, i0)
    return result


DEFAULT_CONFIG = {
    'descriptor_dim': 256,
    'keypoint_encoder': [32, 64, 128, 256],
    'GNN_layers': ['self', 'cross'] * 9,
    'sinkhorn_iterations': 100,
    'multi_frame_matching': True,
    'conf_mlp': True,
}


def matcher_forward(sd, config, data):
    """MultiViewMatcher.forward in eval mode (multi_view_matcher.py:322-332):
    pairwise ``match`` for every id0<id1 when multi_frame_matching is False
    (:150-215), else ``multi_match`` (:217-320)."""
    cfg = {**DEFAULT_CONFIG, **config}
    t = len(data['ids'])
    names = cfg['GNN_layers']
    result = {}
    if not cfg['multi_frame_matching']:
        for id1 in range(t):
            for id0 in range(id1):
                k0 = normalize_keypoints(data['keypoints%d' % id0], data['image%d' % id0].shape)
                k1 = normalize_keypoints(data['keypoints%d' % id1], data['image%d' % id1].shape)
                d0 = data['descriptors%d' % id0] + keypoint_encoder(sd, k0, data['scores%d' % id0])
                d1 = data['descriptors%d' % id1] + keypoint_encoder(sd, k1, data['scores%d' % id1])
                d0, d1 = attentional_gnn(sd, names, d0, d1)
                _pair_head(sd, cfg, d0, d1, id0, id1, result)
        return result
    desc = []
    for i in range(t):
        k = normalize_keypoints(data['keypoints%d' % i], data['image0'].shape)
        desc.append(data['descriptors%d' % i] + keypoint_encoder(sd, k, data['scores%d' % i]))
    desc = multi_frame_gnn(sd, names, desc, list(range(t)))
    for id1 in range(t):
        for id0 in range(id1):
            _pair_head(sd, cfg, desc[id0], desc[id1], id0, id1, result)
    return result
