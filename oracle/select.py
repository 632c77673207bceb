"""Budget-constrained action selection restatement (oracle, test-only).

Follows robust_rmab/algos/rmabppo/rmabppo_core.py:
  act_test_deterministic              :289-334  (binary: action 1 for the top-floor(B) arms by p(a=1))
  act_test_deterministic_multiaction  :338-445  (greedy with hiding and sweep restarts)
and the index-policy selection of robust_rmab/simulator.py:309-325 (top-B by index).

Tie rule (the reference's `np.argsort` default is an unstable introsort, so its tie order is
numpy-build dependent): descending by value, ties -> HIGHER arm index first, i.e.
`np.argsort(kind='stable')[::-1]`.  For N <= 16 numpy's introsort is an insertion sort (stable), so
this equals the reference there; for tie-free inputs it equals the reference at any N.
"""
import numpy as np


def _order_desc(keys):
    return np.argsort(keys, kind="stable")[::-1]


def topk_binary(p1, C, B, positive_only=False):
    """rmabppo_core.py:289-334 for one env.  p1 `[N]` -> actions `[N]` int64.
    positive_only: the binary case of the multi-action greedy (:338-445), which never assigns a zero-probability
    action: the selected arms with p1 <= 0 are dropped (their budget stays unused)."""
    p1 = np.asarray(p1)
    N = p1.shape[0]
    a = np.zeros(N, dtype=np.int64)
    order = _order_desc(p1)
    i, spent = 0, 0
    while spent < B and i < N:
        if spent + C[1] > B:
            break
        a[order[i]] = 1
        spent += C[1]
        i += 1
    if positive_only:
        a[~(p1 > 0)] = 0
    return a


def greedy_multiaction(probs, C, B):
    """rmabppo_core.py:371-445, literal restatement for one env.  probs `[N,A]` -> actions `[N]`."""
    pi = np.array(probs, dtype=np.float64)
    C = np.asarray(C)
    N, A = pi.shape
    actions = np.zeros(N, dtype=np.int64)

    def resort():
        return _order_desc(pi.max(axis=1)), np.argsort(pi, axis=1, kind="stable")

    row_order, arg = resort()
    spent = 0
    done = False
    while spent < B and not done:
        i = 0
        while i < N:
            arm = row_order[i]
            arm_a = arg[arm][-1]
            if spent + C[arm_a] - C[actions[arm]] > B:
                done = True
                break
            i += 1
            actions[arm] = arm_a
            pi[arm, :arm_a + 1] = 0
            spent = C[actions].sum()
            over = (C[None, :] - C[actions][:, None]) > (B - spent)
            if over.any():
                i = 0
                pi[over] = 0
                row_order, arg = resort()
            if not pi.sum() > 0:
                done = True
                break
        row_order, arg = resort()
    return actions


def greedy_multiaction_stepwise(probs, C, B):
    """Argmax-per-step formulation used by the CUDA kernel (no sort): same result as
    `greedy_multiaction`.  A sweep visits arms in descending order of the row max SNAPSHOT taken at
    the last (re)start, which is the same as repeatedly taking the max snapshot key among the arms
    not yet visited in this sweep.  Kept here so the equivalence is property-tested on the CPU.
    """
    pi = np.array(probs, dtype=np.float32)
    C = np.asarray(C, dtype=np.float64)
    N, A = pi.shape
    actions = np.zeros(N, dtype=np.int64)
    spent = 0.0
    done = False

    def snapshot():
        return pi.max(axis=1).copy(), np.zeros(N, dtype=bool)

    while spent < B and not done:
        key, visited = snapshot()
        n_vis = 0
        while n_vis < N:
            k = np.where(visited, -1.0, key)
            m = k.max()
            arm = np.flatnonzero(k == m)[-1]              # ties -> higher index
            row = pi[arm]
            arm_a = np.flatnonzero(row == row.max())[-1]  # stable argsort -> last max
            if spent + C[arm_a] - C[actions[arm]] > B:
                done = True
                break
            visited[arm] = True
            n_vis += 1
            actions[arm] = arm_a
            pi[arm, :arm_a + 1] = 0
            spent = C[actions].sum()
            over = (C[None, :] - C[actions][:, None]) > (B - spent)
            if over.any():
                pi[over] = 0
                key, visited = snapshot()
                n_vis = 0
            if not pi.sum() > 0:
                done = True
                break
    return actions


def topk_by_index(index, B):
    """simulator.py:309-325 without the random tie group: top-B arms by index, ties -> higher arm."""
    a = np.zeros(index.shape[0], dtype=np.int64)
    a[_order_desc(index)[:int(B)]] = 1
    return a


def knapsack_multichoice(values, C, B):
    """lp_methods.action_knapsack (:221-271) for one env as an exact dynamic programme (Gurobi is absent; pinned to
    scipy's HiGHS MILP and to brute force in tests): maximise sum_i values[i, a_i] s.t. sum_i C[a_i] == B, one action
    per arm.  Integer costs.  f_i[b] = max_a f_{i-1}[b - C[a]] + values[i, a], ties -> lowest action (the CUDA kernel
    follows the same recurrence and tie rule, so the two agree bit for bit).  Returns (actions `[N]` int64, feasible)."""
    values = np.asarray(values, dtype=np.float64)
    N, A = values.shape
    c = [int(round(float(x))) for x in C]
    B = int(round(float(B)))
    NEG = -1.0e300
    f = np.full(B + 1, NEG)
    f[0] = 0.0
    choice = np.full((N, B + 1), 255, dtype=np.uint8)
    for i in range(N):
        best = np.full(B + 1, NEG)
        arg = np.full(B + 1, 255, dtype=np.uint8)
        for a in range(A):
            if c[a] > B:
                continue
            prev = f[:B + 1 - c[a]]
            cand = np.where(prev > NEG, prev + values[i, a], NEG)
            tgt = slice(c[a], B + 1)
            better = cand > best[tgt]
            best[tgt] = np.where(better, cand, best[tgt])
            arg[tgt] = np.where(better, a, arg[tgt])
        f, choice[i] = best, arg
    feasible = bool(f[B] > NEG) if N else B == 0
    actions = np.zeros(N, dtype=np.int64)
    if feasible:
        b = B
        for i in range(N - 1, -1, -1):
            actions[i] = choice[i, b]
            b -= c[actions[i]]
    return actions, feasible
