"""BundleAdjustGaussNewton2View with the reference's constructor/run contract
(pose_optimization/two_view/bundle_adjust_gauss_newton_2_view.py:101-201) on mvm_ba2view: one CTA
per batch item, fp64, Schur-complement LM step instead of the reference's dense (6+3n)^2 LU.
The strict LU/precondition flags of the reference only change when a failed item stops updating;
singular steps are skipped (non-strict behaviour, the only one the entry scripts use)."""
import logging

import torch

from ... import _lib


class BundleAdjustGaussNewton2View(object):
    def __init__(self, batch_size, n_iterations, jacobi_precond=True, check_lu_info_strict=False, check_precond_strict=False,
                 vary_lm_fact=True, lm_increase=1.5, lm_decrease=3.5):
        assert jacobi_precond and vary_lm_fact and lm_increase == 1.5 and lm_decrease == 3.5, \
            'mvm_ba2view implements the reference defaults (the only configuration its callers use)'
        self.n_imgs = 2
        self.bs = batch_size
        self.n_it = n_iterations
        self.last_trace = None

    def run(self, n_kpts0, n_kpts1, conf, extr1, return_trace=False, n_valid=None, mask=None):
        lib = _lib.lib()
        dev = n_kpts0.device
        if dev.type != 'cuda':
            raise _lib.MvmError('BundleAdjustGaussNewton2View needs CUDA tensors (no CPU fallback)')
        B, N, _ = n_kpts0.shape
        k0 = n_kpts0.float().contiguous()
        k1 = n_kpts1.float().contiguous()
        c = conf.reshape(B, N).float().contiguous()
        T0 = extr1.float().contiguous()
        Tout = torch.empty(B, 4, 4, dtype=torch.float32, device=dev)
        valid = torch.empty(B, dtype=torch.uint8, device=dev)
        pts = torch.empty(B * N * 3, dtype=torch.float64, device=dev)
        trace = torch.zeros(B, self.n_it + 1, dtype=torch.float32, device=dev) if return_trace else None
        with torch.cuda.device(dev):
            rc = lib.mvm_ba2view(_lib.ptr(k0), _lib.ptr(k1), _lib.ptr(c), _lib.ptr(T0), B, N, int(self.n_it),
                                 _lib.ptr(Tout), _lib.ptr(valid), _lib.ptr(pts), _lib.ptr(trace), _lib.ptr(n_valid), _lib.ptr(mask),
                                 _lib.stream_ptr())
        _lib.check(rc, 'mvm_ba2view')
        valid_batch = valid.bool()
        n_valid = int(valid_batch.sum())
        if n_valid != B:
            logging.warning("{} batches are excluded, not enough matches".format(B - n_valid))
        best = torch.eye(4, device=dev).unsqueeze(0).unsqueeze(0).repeat(n_valid, self.n_imgs, 1, 1)
        best[:, 1] = Tout[valid_batch]
        self.last_trace = trace
        return best, valid_batch
