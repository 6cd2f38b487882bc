"""CPU: the two convex programs behind the `ba_initializer` restatement (oracle/ba_init.py), solved by independent
SciPy solvers: least absolute deviations (the L1 step of Chatterjee & Govindu) as a linear program, and the
least-unsquared-deviations position problem (Ozyesil & Singer) by a constrained quasi-Newton method."""
import numpy as np
import pytest
from scipy.optimize import linprog, minimize

from oracle import ba_init as BI, mvba as M


def test_l1_admm_converges_to_the_linear_programming_optimum():
    rng = np.random.default_rng(0)
    n_views, edges = 5, [(i, j) for j in range(5) for i in range(j)]
    A = BI._build_A(n_views, edges)
    b = rng.standard_normal(A.shape[0]) * 0.1
    b[::7] += 1.0                                            # a few gross outliers
    m, n = A.shape
    # min sum t  s.t.  -t <= A x - b <= t
    c = np.concatenate([np.zeros(n), np.ones(m)])
    G = np.block([[A, -np.eye(m)], [-A, -np.eye(m)]])
    h = np.concatenate([b, -b])
    lp = linprog(c, A_ub=G, b_ub=h, bounds=[(None, None)] * n + [(0, None)] * m, method='highs')
    assert lp.status == 0
    x = BI.l1_admm(A, b, max_iter=20000, abs_tol=1e-9, rel_tol=1e-9)
    assert abs(np.abs(A @ x - b).sum() - lp.fun) <= 1e-4 * max(1.0, lp.fun)


def _lud_objective(z, n_views, edges, dirs):
    c = np.vstack([np.zeros(3), z[:3 * (n_views - 1)].reshape(-1, 3)])
    s = z[3 * (n_views - 1):]
    return sum(np.linalg.norm(c[j] - c[i] - s[e] * dirs[e]) for e, (i, j) in enumerate(edges))


@pytest.mark.parametrize('seed,outliers', [(0, 0), (1, 2)])
def test_lud_positions_are_the_constrained_minimiser(seed, outliers):
    rng = np.random.default_rng(seed)
    n_views = 5
    centres = np.vstack([np.zeros(3), rng.standard_normal((n_views - 1, 3))])
    edges = [(i, j) for j in range(n_views) for i in range(j)]
    rot = np.zeros((n_views, 3))                              # identity rotations: directions are world directions
    pair_pos = {}
    for k, (i, j) in enumerate(edges):
        d = centres[j] - centres[i]
        d = d / np.linalg.norm(d) + 0.01 * rng.standard_normal(3)
        if k < outliers:
            d = rng.standard_normal(3)
        pair_pos[(i, j)] = d
    pos = BI.lud_positions(n_views, pair_pos, rot)
    dirs = [pair_pos[e] for e in edges]
    # scales implied by the oracle's positions (projection clipped at the bound)
    s = np.array([max(1.0, dirs[e] @ (pos[j] - pos[i]) / (dirs[e] @ dirs[e])) for e, (i, j) in enumerate(edges)])
    f_oracle = _lud_objective(np.concatenate([pos[1:].reshape(-1), s]), n_views, edges, dirs)
    z0 = np.concatenate([pos[1:].reshape(-1) * 1.1 + 0.05, s + 0.1])
    best = np.inf
    for start in (z0, np.concatenate([rng.standard_normal(3 * (n_views - 1)), np.full(len(edges), 1.5)])):
        res = minimize(_lud_objective, start, args=(n_views, edges, dirs), method='SLSQP',
                       bounds=[(None, None)] * (3 * (n_views - 1)) + [(1.0, None)] * len(edges),
                       options={'maxiter': 2000, 'ftol': 1e-14})
        best = min(best, res.fun)
    # convex problem: nobody gets below the minimum; the oracle's IRLS + active set is at it
    assert f_oracle <= best + 1e-5 * max(1.0, best), (f_oracle, best)
    assert f_oracle >= best - 1e-3 * max(1.0, best), (f_oracle, best)
