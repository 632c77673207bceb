"""Value-iteration / Whittle-index restatement (oracle, test-only).

Follows the vendored, locally modified pymdptoolbox in the reference:
  MDP._bellmanOperator        mdptoolbox/mdp.py:217-246   Q[a,s] = R[a][s] + g * P[a][s,:].V ; max / first argmax
  MDP._computeMatrixReward    mdptoolbox/mdp.py:292-302   R[a][s] = sum_s' P[a][s,s'] R[a,s,s']
  ValueIteration.__init__     mdptoolbox/mdp.py:1293-1318 thresh = eps(1-g)/g
  ValueIteration._boundIter   mdptoolbox/mdp.py:1320-1364 max_iter = ceil(log(thresh/span)/log(g k))
  ValueIteration.run          mdptoolbox/mdp.py:1366-1410 stop 'full': max(V-Vprev) < thresh, or iter == max_iter
and the call site robust_rmab/simulator.py:504-521 (R_i[a,s,s'] = R[i][s], i.e. reward on the
CURRENT state; indexes[i,a,s] = R[i,s] + g V_i . T[i,s,a]).

lambda-probes: the per-(arm, lambda) MDP has reward R[i,s] - lambda*C[a]  (lp_methods.py:63,181).
The Whittle index of (arm, state) is the lambda at which Q(s,1;lambda) = Q(s,0;lambda)
(lp_methods.py:181).  The reference solves that with a Gurobi LP, which cannot run here
(gurobipy absent, unpinned): the index / Hawkins parts below are PARITY UNPINNED against the
reference's LP; they are pinned only against mdptoolbox.ValueIteration per probe.
"""
import math

import numpy as np


def bound_iter(P, R, gamma, eps):
    """_boundIter for one MDP.  P `[A,S,S]`, R `[A,S]` (already collapsed).  Returns max_iter (int).

    Raises ZeroDivisionError / ValueError exactly where the reference's math.log / division does.
    """
    A, S, _ = P.shape
    h = P.min(axis=(0, 1))            # h[ss] = min over (a, s) of P[a][s, ss]   (:1344-1352)
    k = 1 - h.sum()
    V0 = np.zeros(S)
    Q = R + gamma * (P @ V0)          # _bellmanOperator on V = 0
    value = Q.max(axis=0)
    diff = value - V0
    span = diff.max() - diff.min()    # util.getSpan
    max_iter = math.log((eps * (1 - gamma) / gamma) / span) / math.log(gamma * k)
    return int(math.ceil(max_iter))


def value_iteration(P, R, gamma, eps=0.01, stop="full", max_iter=None):
    """ValueIteration(...).run() for one MDP; float64.  Returns (V[S], policy[S], iters, max_iter)."""
    P = np.asarray(P, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    A, S, _ = P.shape
    if max_iter is None:
        try:
            max_iter = bound_iter(P, R, gamma, eps)
        except (ZeroDivisionError, ValueError, OverflowError):
            if stop != "abs":
                raise
            max_iter = 10 ** 6
    thresh = eps * (1 - gamma) / gamma
    V = np.zeros(S)
    it = 0
    while True:
        it += 1
        Vprev = V
        Q = R + gamma * (P @ V)
        policy = Q.argmax(axis=0)
        V = Q.max(axis=0)
        d = V - Vprev
        variation = (d.max() - d.min()) if stop == "fast" else d.max()
        if stop == "abs":                      # not in mdptoolbox: sup-norm change, no iteration bound (true convergence)
            variation = np.abs(d).max()
        if variation < thresh:
            break
        if it == max_iter and stop != "abs":
            break
    return V, policy, it, max_iter


def probe_mdp(T_arm, R_arm, C, lam):
    """(P[A,S,S], R[A,S]) of one arm at one lambda: simulator.py:506-510 with R - lambda*C[a]."""
    P = np.swapaxes(np.asarray(T_arm, np.float64), 0, 1)
    R = np.asarray(R_arm, np.float64)[None, :] - lam * np.asarray(C, np.float64)[:, None]
    return P, R


def batched_vi(T, R, C, lambdas, gamma, eps, stop="full"):
    """All arms x lambda-probes.  T `[N,S,A,S]`, R `[N,S]`, C `[A]`, lambdas `[L]` or `[N,L]`.

    Returns V `[N,L,S]` f64, policy `[N,L,S]` u8, iters `[N,L]` i32, max_iter `[N,L]` i32.  Probes whose
    `_boundIter` raises in the reference (span == 0 or gamma*k in {0,1}) get iters = -1 and V = NaN.
    """
    T = np.asarray(T, np.float64)
    N, S, A, _ = T.shape
    lambdas = np.asarray(lambdas, np.float64)
    if lambdas.ndim == 1:
        lambdas = np.broadcast_to(lambdas[None], (N, lambdas.shape[0]))
    L = lambdas.shape[1]
    V = np.full((N, L, S), np.nan)
    pol = np.zeros((N, L, S), np.uint8)
    iters = np.full((N, L), -1, np.int32)
    mi = np.full((N, L), -1, np.int32)
    for n in range(N):
        for l in range(L):
            P, Rl = probe_mdp(T[n], R[n], C, lambdas[n, l])
            try:
                v, p, it, m = value_iteration(P, Rl, gamma, eps, stop)
            except (ZeroDivisionError, ValueError, OverflowError):
                continue
            V[n, l], pol[n, l], iters[n, l], mi[n, l] = v, p, it, m
    return V, pol, iters, mi


def q_values(T, R, C, V, lambdas, gamma):
    """Q[n,l,s,a] = R[n,s] - lambda*C[a] + gamma * T[n,s,a,:].V[n,l,:]  (simulator.py:518-521)."""
    T = np.asarray(T, np.float64)
    lambdas = np.asarray(lambdas, np.float64)
    if lambdas.ndim == 1:
        lambdas = np.broadcast_to(lambdas[None], (T.shape[0], lambdas.shape[0]))
    ev = np.einsum('nsat,nlt->nlsa', T, V)
    return (np.asarray(R, np.float64)[:, None, :, None]
            - lambdas[:, :, None, None] * np.asarray(C, np.float64)[None, None, None, :] + gamma * ev)


def whittle_bisect(T, R, C, gamma, eps, lam_lo, lam_hi, n_rounds, stop="full"):
    """Binary search over lambda for the indifference point Q(s,1;l) = Q(s,0;l) per (arm, state).

    PARITY UNPINNED vs lp_methods.py:111-212 (Gurobi).  Deterministic schedule shared with the CUDA
    path: per (arm, state) interval [lo, hi]; each round probes mid = (lo+hi)/2 with value_iteration,
    and moves lo = mid where Q(s,1) - Q(s,0) > 0 (acting still preferred), else hi = mid.
    Returns index `[N,S]` = (lo+hi)/2 after n_rounds.
    """
    T = np.asarray(T, np.float64)
    N, S, A, _ = T.shape
    lo = np.full((N, S), float(lam_lo))
    hi = np.full((N, S), float(lam_hi))
    for _ in range(n_rounds):
        mid = 0.5 * (lo + hi)
        V, _, _, _ = batched_vi(T, R, C, mid, gamma, eps, stop)   # probes = one per state
        Q = q_values(T, R, C, V, mid, gamma)                      # [N,S(probe),S,A]
        idx = np.arange(S)
        gap = Q[:, idx, idx, 1] - Q[:, idx, idx, 0]
        act = gap > 0
        lo = np.where(act, mid, lo)
        hi = np.where(act, hi, mid)
    return 0.5 * (lo + hi)


def hawkins_objective(V, start_state, lambdas, B, gamma):
    """lp_methods.py:63 objective (without the 1e-6 tie term): sum_i V_i(s_i; l) + l*B/(1-gamma); `[L]`."""
    N = V.shape[0]
    vs = V[np.arange(N), :, np.asarray(start_state, np.int64)]     # [N,L]
    return vs.sum(axis=0) + np.asarray(lambdas, np.float64) * B / (1 - gamma)


def hawkins_lp_highs(T, R, C, B, start_state, gamma, lambda_lim):
    """lp_methods.hawkins (:11-105) restated for scipy's HiGHS (Gurobi is absent): variables (lambda, L[N,S]),
    minimise sum_i L_i(s_i) + lambda B/(1-gamma) + 1e-6 sum L  s.t.  L_i(s) >= R_i(s) - lambda C[a] + gamma T_i[s,a].L_i.
    Small N only (test infrastructure).  Returns (L_vals [N,S], lambda, objective)."""
    from scipy.optimize import linprog
    T = np.asarray(T, np.float64)
    N, S, A, _ = T.shape
    nv = 1 + N * S
    c = np.full(nv, 1e-6)
    c[0] = B / (1 - gamma)
    for i in range(N):
        c[1 + i * S + int(start_state[i])] += 1.0
    A_ub, b_ub = [], []
    for i in range(N):
        for s in range(S):
            for a in range(A):
                row = np.zeros(nv)
                row[0] = -C[a]                                   # -lambda*C[a] moved to the left: -L + g T.L - lambda C <= -R
                row[1 + i * S:1 + (i + 1) * S] = gamma * T[i, s, a]
                row[1 + i * S + s] -= 1.0
                A_ub.append(row)
                b_ub.append(-R[i, s])
    res = linprog(c, A_ub=np.array(A_ub), b_ub=np.array(b_ub), bounds=[(0, lambda_lim)] + [(None, None)] * (N * S),
                  method="highs")
    assert res.status == 0, res.message
    return res.x[1:].reshape(N, S), res.x[0], res.fun


def hawkins_action_binary(T, R, C, B, state, lam, gamma, eps=1e-10):
    """Hawkins action for binary costs: Q from the converged lambda-charged value function, then exactly B/C[1] arms
    with the largest Q(s,1) - Q(s,0) (the equality knapsack of lp_methods.py:243-250), ties -> higher arm index."""
    N = T.shape[0]
    V, _, _, _ = batched_vi(T, R, C, np.array([lam]), gamma, eps, stop="abs")
    Q = q_values(T, R, C, V, np.array([lam]), gamma)[:, 0]                       # [N,S,A]
    gap = (Q[np.arange(N), state, 1] - Q[np.arange(N), state, 0]).astype(np.float32)
    a = np.zeros(N, dtype=np.int64)
    a[np.argsort(gap, kind="stable")[::-1][:int(B / C[1])]] = 1
    return a


def hawkins_action_knapsack(T, R, C, B, state, lam, gamma, eps=1e-10):
    """Hawkins action for general integer costs (agent_baselines.py:69-91): Q(s_i, a; lambda) from the converged
    lambda-charged value function, then the exact equality knapsack of lp_methods.py:221-271."""
    from .select import knapsack_multichoice
    N = T.shape[0]
    V, _, _, _ = batched_vi(T, R, C, np.array([lam]), gamma, eps, stop="abs")
    Q = q_values(T, R, C, V, np.array([lam]), gamma)[:, 0]                       # [N,S,A]
    return knapsack_multichoice(Q[np.arange(N), state], C, B)


def whittle_lp_highs(T, R, C, B, start_state, a_index, gamma, lambda_lim=None, index_range=None):
    """lp_methods.lp_to_compute_index (:111-212) restated for scipy's HiGHS (Gurobi is absent; small N only, test
    infrastructure).  The reference builds ONE model over all arms, but no constraint couples two arms, so it is
    solved here arm by arm.  Per arm p: variables (index_p, L_p[S]), all with Gurobi's default lower bound 0
    (index_p <= lambda_lim);
        minimise  L_p[start_p]  (+ index_p B/(1-gamma) for ONE arm only: the objective at :163 reads
                  `index_variables[i]` with `i` left over from the loop at :144-146, i.e. arm S-1)
        s.t.      L_p[s] >= R_p[s] - index_p C[a] + gamma T_p[s,a].L_p                      for all s, a   (:166-170)
                  Q(start_p, a_index) == Q(start_p, a_index - 1),  Q(s,a) = R[s] - index C[a] + gamma T[s,a].L  (:176-181)
    Returns (L_vals [N,S], index [N]); NaN where the LP is infeasible (Gurobi would raise on `.x`).
    index_range: optional [N,2] array that receives the smallest and largest index on the LP's optimal face."""
    from scipy.optimize import linprog
    T = np.asarray(T, np.float64)
    R = np.asarray(R, np.float64)
    C = np.asarray(C, np.float64)
    N, S, A, _ = T.shape
    L_vals, index = np.full((N, S), np.nan), np.full(N, np.nan)
    for p in range(N):
        s0 = int(start_state[p])
        c = np.zeros(1 + S)
        c[1 + s0] = 1.0
        if p == S - 1:
            c[0] = B / (1 - gamma)
        A_ub, b_ub = [], []
        for s in range(S):
            for a in range(A):
                row = np.zeros(1 + S)
                row[0] = -C[a]
                row[1:] = gamma * T[p, s, a]
                row[1 + s] -= 1.0
                A_ub.append(row)
                b_ub.append(-R[p, s])
        eq = np.zeros(1 + S)
        eq[0] = -(C[a_index] - C[a_index - 1])
        eq[1:] = gamma * (T[p, s0, a_index] - T[p, s0, a_index - 1])
        bounds = [(0, lambda_lim)] + [(0, None)] * S
        res = linprog(c, A_ub=np.array(A_ub), b_ub=np.array(b_ub), A_eq=eq[None], b_eq=[0.0], bounds=bounds, method="highs")
        if res.status == 0:
            index[p], L_vals[p] = res.x[0], res.x[1:]
            if index_range is not None:
                # the optimal FACE: the LP is often degenerate (a continuum of indices attains the optimum, e.g. every
                # index above the indifference point when the start state's value no longer depends on it), and which
                # vertex a solver returns is then its own business -- so also report the smallest / largest optimal index
                cap = np.concatenate([c, [res.fun + 1e-9 * max(1.0, abs(res.fun))]])
                for k, sign in enumerate((1.0, -1.0)):
                    obj = np.zeros(1 + S)
                    obj[0] = sign
                    r2 = linprog(obj, A_ub=np.array(A_ub + [cap[:-1]]), b_ub=np.array(b_ub + [cap[-1]]), A_eq=eq[None],
                                 b_eq=[0.0], bounds=bounds, method="highs")
                    index_range[p, k] = r2.x[0] if r2.status == 0 else np.nan
    return L_vals, index


def whittle_pin_instances(seed, N=10):
    """Small instances for pinning the Whittle index to the LP restatement: counterexample, ARMMAN and dense random MDPs.
    Yields (name, T [N,S,2,S], R [N,S], C, lambda_lim)."""
    from . import envs as oenv
    rs = np.random.RandomState(seed)
    C = np.array([0.0, 1.0])
    rgc = oenv.ce_param_ranges(N)
    p = rgc[:, 0] + rs.rand(N) * (rgc[:, 1] - rgc[:, 0])
    T = np.zeros((N, 2, 2, 2))
    T[:, 0] = 0.5
    T[:, 1, 0, 0] = 1.0
    T[:, 1, 1, 1], T[:, 1, 1, 0] = p, 1 - p
    yield "ce", T, np.tile([0.0, 1.0], (N, 1)), C, 10.0
    pa = oenv.sample_sub_ranges(oenv.armman_param_ranges(N), rs).mean(-1)
    T = np.zeros((N, 3, 2, 3))
    for s in range(3):
        for a in range(2):
            T[:, s, a] = oenv.armman_probs(np.full(N, s), np.full(N, a), pa[:, s, a]).astype(np.float64)
    yield "armman", T, np.tile([1.0, 0.5, 0.0], (N, 1)), C, 10.0
    S = 4
    R = np.sort(rs.rand(N, S), axis=1)
    yield "random", rs.dirichlet(np.ones(S), size=(N, S, 2)), R, C, float(R.max() / 0.1)


def whittle_lp_pin(index, T, R, C, gamma, lambda_lim, B=3.0, tol=1e-6):
    """Compare indifference-point indices `index [N,S]` (bisection on Q(s,1;l) = Q(s,0;l) under the optimal value
    function) with the reference's index LP (whittle_lp_highs, lp_methods.py:111-212) for start state s on every arm.

    The two definitions coincide where the LP's minimal solution is the Bellman solution; the LP is a relaxation of it
    (L only has to DOMINATE the Bellman backup, so it may buy a lower L[start] with an inflated L elsewhere) and is
    frequently degenerate, so an entry counts as
      'agree'      index == the smallest index on the LP's optimal face;
      'lp_relaxed' the LP optimum is strictly below V^index(start): the LP left the Bellman solution (or the true
                   indifference point is <= 0 and the index variable's lower bound 0 forces that);
      'nonpositive' acting is not preferred even at lambda = 0 (indifference point <= 0): the bisection returns its lower
                   end 0, while the LP (index >= 0 and the equality enforced) has no Bellman solution and inflates L;
      'quirk'      arm S-1, whose index carries the stray B/(1-gamma) objective term (lp_methods.py:163);
      'mismatch'   anything else -- must not happen.
    Returns a dict of counts and the list of mismatches."""
    N, S = index.shape
    out = dict(agree=0, lp_relaxed=0, nonpositive=0, quirk=0, mismatch=0)
    bad = []
    for s in range(S):
        rng_ = np.zeros((N, 2))
        L, _ = whittle_lp_highs(T, R, C, B, np.full(N, s), 1, gamma, lambda_lim, index_range=rng_)
        Vw, _, _, _ = batched_vi(T, R, C, index[:, s:s + 1], gamma, 1e-13, stop="abs")      # per-arm lambda = its index
        for p in range(N):
            if p == S - 1:
                out["quirk"] += 1
            elif abs(rng_[p, 0] - index[p, s]) <= tol:
                out["agree"] += 1
            elif L[p, s] < Vw[p, 0, s] - 1e-9:
                out["lp_relaxed"] += 1
            elif index[p, s] <= tol:
                out["nonpositive"] += 1
            else:
                out["mismatch"] += 1
                bad.append((s, p, float(index[p, s]), rng_[p].tolist(), float(L[p, s]), float(Vw[p, 0, s])))
    return out, bad
