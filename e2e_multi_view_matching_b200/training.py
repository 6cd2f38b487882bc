"""Training-side consumers of the matcher output -- FIRST SLICE of SURVEY.md §8 f-2 / a20 (BASELINE cfg5).

Built:
  * compute_match_loss (helpers.py:228-241) as an autograd.Function on CUDA kernels (csrc/train_loss.cu), forward
    and backward;
  * combine_losses (train.py:36-40);
  * compute_gt_matches_of_image_pair / compute_gt_matches (helpers.py:121-226): ground-truth assignments from depth
    maps and poses on CUDA kernels (csrc/gt_matches.cu; the [bs, N, N] error matrix is never materialised);
  * run_matcher (helpers.py:243-260) and validation_step (the loop body of Trainer.validate, train.py:89-106) for the
    matcher in EVAL mode -- exactly what the reference's validation pass runs: matcher forward, match loss, weighted
    eight-point with choose_closest against the ground-truth pose, rotation / translation losses, combine_losses and
    the validation-loss all-reduce;
  * LogOptimalTransport: autograd.Function around the production Sinkhorn kernel whose backward is the EXACT gradient
    of the 100 unrolled iterations (what autograd computes for the reference, superglue.py:143-172).  The backward
    below is stated in torch operations -- it is the executable specification (checked against autograd through the
    reference's function in tests/test_training_gpu.py) of the fused kernel that is the next step: since Z is
    constant over the iterations, d loss / d Z = G - exp(Z) * (A B^T) with A = [exp(u^t) | r^t], B = [c^t | exp(v^(t-1))]
    stacked over the iterations (a rank-2T correction), and the recursion for (gu, gv) needs exactly the two
    matrix-vector passes per iteration the forward kernel already makes over its on-chip copy of the matrix.
Not built (stated, not hidden): the train branch of the matcher forward (batch-statistics BatchNorm over B*T*N,
`full_output`, multi_view_matcher.py:65-86,219-226) and the backward kernels of attention / GEMMs / pose, so
MultiViewMatcher.forward still raises in training mode and cfg5 cannot run end to end yet.
"""
import torch

from . import _lib
from . import ops


class _MatchLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_p, gt_indices, gt_weights):
        if log_p.device.type != 'cuda':
            raise _lib.MvmError('compute_match_loss needs CUDA tensors (no CPU fallback)')
        lib = _lib.lib()
        bs, ft, ft2 = log_p.shape
        assert ft == ft2 and gt_indices.shape == (bs, 2, ft) and gt_weights.shape == (bs, 2, ft)
        lp = log_p.detach().float().contiguous()
        idx = gt_indices.long().contiguous()
        w = gt_weights.float().contiguous()
        part = torch.empty(bs, dtype=torch.float64, device=lp.device)
        loss = torch.empty(1, dtype=torch.float32, device=lp.device)
        with torch.cuda.device(lp.device):
            _lib.check(lib.mvm_match_loss_forward(_lib.ptr(lp), _lib.ptr(idx), _lib.ptr(w), bs, ft, _lib.ptr(part),
                                                  _lib.ptr(loss), _lib.stream_ptr()), 'mvm_match_loss_forward')
        ctx.save_for_backward(idx, w)
        ctx.shape = (bs, ft)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad):
        lib = _lib.lib()
        idx, w = ctx.saved_tensors
        bs, ft = ctx.shape
        g = grad.detach().float().reshape(1).contiguous()
        out = torch.empty(bs, ft, ft, dtype=torch.float32, device=idx.device)
        with torch.cuda.device(idx.device):
            _lib.check(lib.mvm_match_loss_backward(_lib.ptr(idx), _lib.ptr(w), _lib.ptr(g), bs, ft, _lib.ptr(out),
                                                   _lib.stream_ptr()), 'mvm_match_loss_backward')
        return out, None, None


def compute_match_loss(log_p, gt_indices_0_1, gt_weights_0_1):
    """helpers.py:228-241."""
    return _MatchLoss.apply(log_p, gt_indices_0_1, gt_weights_0_1)


def combine_losses(losses, n_pairs, pose_match_ratio, rot_weight, trans_weight):
    """train.py:36-40."""
    losses = {k: v / float(n_pairs) for k, v in losses.items()}
    pose_loss = rot_weight * losses["rot_loss"] + trans_weight * losses["transl_loss"]
    total_loss = (1. - pose_match_ratio) * losses["match_loss"] + pose_match_ratio * pose_loss
    return total_loss, losses


class LogOptimalTransport(torch.autograd.Function):
    """log_optimal_transport(scores, alpha, iters) (superglue.py:152-172) with gradients w.r.t. scores and alpha."""

    @staticmethod
    def forward(ctx, scores, alpha, iters):
        if scores.device.type != 'cuda':
            raise _lib.MvmError('LogOptimalTransport needs CUDA tensors (no CPU fallback)')
        Z = ops.log_optimal_transport(scores.detach().float().contiguous(), float(alpha), int(iters))
        ctx.save_for_backward(scores.detach(), alpha.detach() if torch.is_tensor(alpha) else torch.tensor(float(alpha)))
        ctx.iters = int(iters)
        return Z

    @staticmethod
    def backward(ctx, G):
        scores, alpha = ctx.saved_tensors
        T = ctx.iters
        dt = torch.float64                      # the recursion is short and cheap next to the forward: do it in double
        b, m, n = scores.shape
        dev = scores.device
        Z = torch.empty(b, m + 1, n + 1, dtype=dt, device=dev)
        Z[:, :m, :n] = scores.to(dt)
        Z[:, m, :] = alpha.to(dt).to(dev)
        Z[:, :m, n] = alpha.to(dt).to(dev)
        norm = -torch.log(torch.tensor(float(m + n), dtype=dt, device=dev))
        log_mu = torch.cat([norm.expand(m), (torch.log(torch.tensor(float(n), dtype=dt, device=dev)) + norm)[None]])[None]
        log_nu = torch.cat([norm.expand(n), (torch.log(torch.tensor(float(m), dtype=dt, device=dev)) + norm)[None]])[None]
        us, vs = [], [torch.zeros(b, n + 1, dtype=dt, device=dev)]
        u, v = torch.zeros(b, m + 1, dtype=dt, device=dev), vs[0]
        for _ in range(T):                      # the reference's iteration (superglue.py:143-149), potentials kept
            u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
            v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
            us.append(u)
            vs.append(v)
        G = G.to(dt)
        gu, gv, dZ = G.sum(2), G.sum(1), G.clone()
        for t in range(T, 0, -1):
            u_t, v_t, v_p = us[t - 1], vs[t], vs[t - 1]
            # v^t = log_nu - LSE_i(Z + u^t):  P_ij = exp(Z_ij + u^t_i + v^t_j - log_nu_j)  (columns sum to one)
            P = torch.exp(Z + u_t.unsqueeze(2) + (v_t - log_nu).unsqueeze(1))
            W = P * gv.unsqueeze(1)
            dZ -= W
            gu = gu - W.sum(2)
            # u^t = log_mu - LSE_j(Z + v^(t-1)):  Q_ij = exp(Z_ij + u^t_i + v^(t-1)_j - log_mu_i)  (rows sum to one)
            Q = torch.exp(Z + (u_t - log_mu).unsqueeze(2) + v_p.unsqueeze(1))
            W = Q * gu.unsqueeze(2)
            dZ -= W
            gv = -W.sum(1)
            gu = torch.zeros_like(gu)
        d_scores = dZ[:, :m, :n].to(scores.dtype)
        d_alpha = (dZ[:, m, :].sum() + dZ[:, :m, n].sum()).to(torch.float32)
        return d_scores, d_alpha, None


def log_optimal_transport(scores, alpha, iters):
    return LogOptimalTransport.apply(scores, alpha if torch.is_tensor(alpha) else torch.tensor(float(alpha)), iters)


def compute_gt_matches_of_image_pair(kpts0, kpts1, K0, K1, T0to1, depth0, depth1, max_matched_reproj_err,
                                     min_unmatched_reproj_err):
    """helpers.py:121-203 -> (indices [bs,2,N+1] int64, weights [bs,2,N+1] float32)."""
    if kpts0.device.type != 'cuda':
        raise _lib.MvmError('compute_gt_matches_of_image_pair needs CUDA tensors (no CPU fallback)')
    lib = _lib.lib()
    bs, n, _ = kpts0.shape
    assert kpts1.shape == (bs, n, 2) and depth0.shape == depth1.shape and depth0.dim() == 3
    H, W = depth0.shape[1:]
    dev = kpts0.device
    f = lambda t: t.detach().float().contiguous()
    k0, k1, K0_, K1_, T_, d0, d1 = f(kpts0), f(kpts1), f(K0), f(K1), f(T0to1), f(depth0), f(depth1)
    assert K0_.shape == (bs, 4, 4) and K1_.shape == (bs, 4, 4) and T_.shape == (bs, 4, 4)
    indices = torch.empty(bs, 2, n + 1, dtype=torch.int64, device=dev)
    weights = torch.empty(bs, 2, n + 1, dtype=torch.float32, device=dev)
    nbytes = lib.mvm_gt_matches_workspace_bytes(bs, n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mvm_gt_matches_pair(_lib.ptr(k0), _lib.ptr(k1), _lib.ptr(K0_), _lib.ptr(K1_), _lib.ptr(T_),
                                           _lib.ptr(d0), _lib.ptr(d1), bs, n, H, W, float(max_matched_reproj_err),
                                           float(min_unmatched_reproj_err), _lib.ptr(indices), _lib.ptr(weights),
                                           _lib.ptr(ws), nbytes, _lib.stream_ptr()), 'mvm_gt_matches_pair')
    return indices, weights


def compute_gt_matches(opt, data):
    """helpers.py:215-226: gt_indices_k_m / gt_weights_k_m for every pair k < m of the tuple; pops the depth maps."""
    curr_tuple_size = len(data["ids"])
    for m in range(curr_tuple_size):
        for k in range(m):
            T_k2m = torch.linalg.inv(data["pose" + str(m)]) @ data["pose" + str(k)]
            data["gt_indices_{}_{}".format(k, m)], data["gt_weights_{}_{}".format(k, m)] = \
                compute_gt_matches_of_image_pair(data["keypoints" + str(k)], data["keypoints" + str(m)],
                                                 data["intr" + str(k)], data["intr" + str(m)], T_k2m,
                                                 data["depth" + str(k)], data["depth" + str(m)],
                                                 opt.match_reproj_err, opt.unmatch_reproj_err)
    for m in range(curr_tuple_size):
        data.pop("depth" + str(m))


def run_matcher(opt, data, matcher):
    """helpers.py:243-260.  The matcher must be in eval mode (the validation pass, train.py:66-68): the training
    branch of the forward is not built (module docstring)."""
    from .pose_optimization.two_view.estimate_relative_pose import run_weighted_8_point
    from .pose_optimization.two_view.compute_pose_error import compute_rotation_error, compute_translation_error_as_angle
    curr_tuple_size = len(data["ids"])
    getattr(matcher, 'module', matcher).config["full_output"] = opt.pose_loss        # DataParallel / DDP wrapped or bare
    result = matcher(data)
    match_loss = torch.zeros(1, device=data["pose0"].device)
    rot_loss = torch.zeros(1, device=match_loss.device)
    transl_loss = torch.zeros(1, device=match_loss.device)
    for id1 in range(curr_tuple_size):
        for id0 in range(id1):
            match_loss = match_loss + compute_match_loss(result["scores_{}_{}".format(id0, id1)],
                                                         data["gt_indices_{}_{}".format(id0, id1)],
                                                         data["gt_weights_{}_{}".format(id0, id1)])
            if opt.pose_loss:
                target = torch.linalg.inv(data["pose{}".format(id1)]) @ data["pose{}".format(id0)]
                pred, _ = run_weighted_8_point(data, result, id0, id1, choose_closest=True, target_T_021=target)
                rot_loss = rot_loss + compute_rotation_error(pred, target)
                transl_loss = transl_loss + compute_translation_error_as_angle(pred, target)
    losses = {"match_loss": match_loss, "rot_loss": rot_loss, "transl_loss": transl_loss}
    return losses, result


def validation_step(opt, data, matcher, n_pairs, pose_match_ratio, process_group=None):
    """One batch of Trainer.validate (train.py:89-106) after SuperPoint: ground-truth matches when the batch still
    carries depth maps, run_matcher, combine_losses; the scalar validation loss is all-reduced over the ranks
    (mean) when torch.distributed is initialised.  -> (val_loss tensor [1], losses dict)."""
    with torch.no_grad():
        if "depth0" in data:
            compute_gt_matches(opt, data)
        losses, _ = run_matcher(opt, data, matcher)
        val_loss, losses = combine_losses(losses, n_pairs, pose_match_ratio, opt.rot_weight, opt.trans_weight)
        val_loss = val_loss.reshape(1).clone()
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(val_loss, group=process_group)
            val_loss /= torch.distributed.get_world_size(process_group)
    return val_loss, losses
