"""TEST INFRASTRUCTURE: torch (CPU, float64 inside) stand-ins for the stage ops of e2e_multi_view_matching_b200.ops, with
the same signatures.  tests/test_train_host_logic.py patches them in to check the HOST-SIDE orchestration of the training
forward / backward (models/train_forward.py: what is saved, the head permutations, the concat / residual routing, the
BatchNorm groups, the pair loops) against the reference goldens without a GPU; the GPU tests check every kernel against
the same functions.  Never imported by the package."""
import torch

D = torch.float64


def pack_views(views, n_pad):
    T, B = len(views), views[0][0].shape[0]
    kp = torch.zeros(B, T, n_pad, 2)
    sc = torch.zeros(B, T, n_pad)
    de = torch.zeros(B, T, 256, n_pad)
    for t, (k, s, d) in enumerate(views):
        n = k.shape[1]
        kp[:, t, :n], sc[:, t, :n], de[:, t, :, :n] = k, s, d
    return kp, sc, de


def linear(a, w, bias=None, a2=None, residual=None, relu=False, alpha=1.0, tc_passes=0, presplit=False):
    A = a if a2 is None else torch.cat([a, a2], 1)
    y = alpha * (A.to(D) @ w.to(D).t())
    if bias is not None:
        y = y + bias.to(D)
    if relu:
        y = y.clamp_min(0)
    if residual is not None:
        y = y + residual.to(D)
    return y.float()


def _attend(qkv, batch, n_views, counts, is_cross):
    """qkv [V, n_pad, 768] double -> [V, n_pad, 256]; keys of a view beyond its count are masked, query rows beyond the
    count produce zeros."""
    V, n_pad, _ = qkv.shape
    out = torch.zeros(V, n_pad, 256, dtype=qkv.dtype)
    for b in range(batch):
        for t in range(n_views):
            v = b * n_views + t
            src = [s for s in range(n_views) if (s != t if is_cross else s == t)]
            if counts[t] == 0 or not src:
                continue
            for h in range(4):
                q = qkv[v, :counts[t], h * 64:(h + 1) * 64]
                k = torch.cat([qkv[b * n_views + s, :counts[s], 256 + h * 64:256 + (h + 1) * 64] for s in src], 0)
                val = torch.cat([qkv[b * n_views + s, :counts[s], 512 + h * 64:512 + (h + 1) * 64] for s in src], 0)
                p = torch.softmax(q @ k.t() / 8.0, dim=1)
                out[v, :counts[t], h * 64:(h + 1) * 64] = p @ val
    return out


def attention(qkv, batch, n_views, counts, is_cross, tc_passes=0):
    return _attend(qkv.to(D), batch, n_views, list(counts), is_cross).float()


def attention_backward(qkv, out, dout, batch, n_views, counts, is_cross):
    with torch.enable_grad():
        x = qkv.to(D).clone().requires_grad_(True)
        o = _attend(x, batch, n_views, list(counts), is_cross)
        (g,) = torch.autograd.grad(o, x, dout.to(D))
    return g.float()


def transpose_split(x, raw=False, planes=True, out=None):
    xt = x.t().contiguous()
    if out is None:
        return (xt.clone() if raw else None), (xt.clone() if planes else None), (torch.zeros_like(xt) if planes else None)
    r, h, l = out
    if r is not None:
        r.copy_(xt)
    if h is not None:
        h.copy_(xt)
        l.zero_()
    return r, h, l


def linear_presplit(a, w_hi, w_lo, residual=None, alpha=1.0):
    y = alpha * (a.to(D) @ (w_hi.to(D) + w_lo.to(D)).t())
    if residual is not None:
        y = y + residual.to(D)
    return y.float()


def linear_presplit_splitk(a, w_hi, w_lo, ksplit, alpha=1.0):
    return linear_presplit(a, w_hi, w_lo, alpha=alpha)


def colsum(x):
    return x.to(D).sum(0).float()


def _mask(rows, n_pad, n_valid, groups, g):
    r = torch.arange(rows)
    return (r % n_pad < n_valid) & ((r // n_pad) % groups == g)


def batchnorm_train(x, weight, bias, running_mean, running_var, momentum, eps, n_pad, n_valid, relu=True, groups=1,
                    out=None, save=False):
    rows, C = x.shape
    y = x if out is None else out
    stats = torch.empty(groups, 2 * C) if save else None
    for g in range(groups):
        m = _mask(rows, n_pad, n_valid, groups, g)
        xs = x[m].to(D)
        n = xs.shape[0]
        mean, var = xs.mean(0), xs.var(0, unbiased=False)
        invstd = 1.0 / torch.sqrt(var + eps)
        o = (xs - mean) * invstd * weight.to(D) + bias.to(D)
        if relu:
            o = o.clamp_min(0)
        y[m] = o.float()
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean.float())
            running_var.mul_(1 - momentum).add_(momentum * (var * n / max(n - 1, 1)).float())
        if save:
            stats[g, :C], stats[g, C:] = mean.float(), invstd.float()
    return y, stats


def batchnorm_train_backward(x, y, dy, weight, stats, n_pad, n_valid, relu=True):
    rows, C = x.shape
    groups = stats.shape[0]
    dg, db = torch.zeros(C, dtype=D), torch.zeros(C, dtype=D)
    for g in range(groups):
        m = _mask(rows, n_pad, n_valid, groups, g)
        mean, invstd = stats[g, :C].to(D), stats[g, C:].to(D)
        gq = dy[m].to(D)
        if relu:
            gq = gq * (y[m] > 0)
        xh = (x[m].to(D) - mean) * invstd
        sg, sgx = gq.sum(0), (gq * xh).sum(0)
        n = gq.shape[0]
        dy[m] = (weight.to(D) * invstd * (gq - sg / n - xh * sgx / n)).float()
        dg += sgx
        db += sg
    return dg.float(), db.float()


def _ot(scores, alpha, iters):
    """log_optimal_transport (superglue.py:143-172) restated in float64."""
    b, m, n = scores.shape
    Z = torch.cat([torch.cat([scores, alpha.expand(b, m, 1)], 2), alpha.expand(b, 1, n + 1)], 1)
    norm = -torch.log(torch.tensor(float(m + n), dtype=D))
    log_mu = torch.cat([norm.expand(m), (torch.log(torch.tensor(float(n), dtype=D)) + norm)[None]])[None]
    log_nu = torch.cat([norm.expand(n), (torch.log(torch.tensor(float(m), dtype=D)) + norm)[None]])[None]
    u, v = torch.zeros(b, m + 1, dtype=D), torch.zeros(b, n + 1, dtype=D)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    return Z + u.unsqueeze(2) + v.unsqueeze(1) - norm


def pair_scores(md, pairs, N, alpha=1.0 / 16.0):
    B = md.shape[0]
    out = torch.zeros(len(pairs) * B, N + 1, N + 1)
    for p, (a, b) in enumerate(pairs):
        out[p * B:(p + 1) * B, :N, :N] = (alpha * (md[:, a, :N].to(D) @ md[:, b, :N].to(D).transpose(1, 2))).float()
    return out


def sinkhorn_train_forward(scores, alpha, iters, augmented=False):
    if augmented:
        scores = scores[:, :-1, :-1]
    return _ot(scores.to(D), alpha.to(D).reshape(()), iters).float(), None


def sinkhorn_train_backward(scores, alpha, pot, iters, grad_out, augmented=False):
    if augmented:
        scores = scores[:, :-1, :-1]
    """-> (dZ [B, m+1, n+1] with d scores in its inner block (the dustbin entries are folded into d_alpha), d_alpha [1])."""
    with torch.enable_grad():
        s = scores.to(D).clone().requires_grad_(True)
        a = alpha.to(D).reshape(()).clone().requires_grad_(True)
        gs, ga = torch.autograd.grad(_ot(s, a, iters), (s, a), grad_out.to(D))
    b, m, n = scores.shape
    dZ = torch.zeros(b, m + 1, n + 1)
    dZ[:, :m, :n] = gs.float()
    return dZ, ga.reshape(1)


def extract_matches(Z, match_threshold=0.0):
    """multi_view_matcher.py:288-300 restated."""
    inner = Z[:, :-1, :-1]
    max0, max1 = inner.max(2), inner.max(1)
    i0, i1 = max0.indices, max1.indices
    ar0 = torch.arange(i0.shape[1])[None]
    ar1 = torch.arange(i1.shape[1])[None]
    mutual0 = ar0 == i1.gather(1, i0)
    mutual1 = ar1 == i0.gather(1, i1)
    zero = inner.new_tensor(0)
    s0 = torch.where(mutual0, max0.values.exp(), zero)
    s1 = torch.where(mutual1, s0.gather(1, i1), zero)
    v0 = mutual0 & (s0 > match_threshold)
    v1 = mutual1 & v0.gather(1, i1)
    return torch.where(v0, i0, i0.new_tensor(-1)), torch.where(v1, i1, i1.new_tensor(-1)), s0, s1
