"""Hypothesis property tests on the oracle's budget-constrained selection (the definition the CUDA kernels are held
to): feasibility, equivalence of the stepwise formulation with the literal restatement of rmabppo_core.py:338-445, and
binary multi-action == top-B."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import select as osel


@st.composite
def instances(draw):
    N = draw(st.integers(1, 14))
    A = draw(st.integers(2, 4))
    quant = draw(st.sampled_from([0, 4, 8]))                       # coarse grids force exact ties
    rows = []
    for _ in range(N):
        r = np.array(draw(st.lists(st.floats(0.01, 1.0), min_size=A, max_size=A)))
        r = r / r.sum()
        if quant:
            r = np.round(r * quant) / quant
            if r.sum() == 0:
                r[0] = 1.0
        rows.append(r)
    B = draw(st.integers(0, 2 * N + 1))
    return np.array(rows, dtype=np.float32), np.arange(A), B


@settings(max_examples=120, deadline=None)
@given(instances())
def test_greedy_multiaction_properties(inst):
    p, C, B = inst
    a = osel.greedy_multiaction(p, C, B)
    assert C[a].sum() <= B                                           # budget feasibility (simulator.py:255-267)
    assert np.array_equal(a, osel.greedy_multiaction_stepwise(p, C, B))
    if p.shape[1] == 2:
        # binary costs: the greedy is "action 1 for the top-B arms by p1" among the arms with p1 > 0 -- it never
        # assigns a zero-probability action (rmabppo_core.py:434), unlike act_test_deterministic (:289-334)
        assert np.array_equal(a, osel.topk_binary(p[:, 1], C, B, positive_only=True))
        if (p[:, 1] > 0).all():
            assert np.array_equal(a, osel.topk_binary(p[:, 1], C, B))


@st.composite
def knapsack_instances(draw):
    N = draw(st.integers(1, 7))
    A = draw(st.integers(2, 3))
    vals = np.array([[draw(st.integers(-6, 6)) / 2.0 for _ in range(A)] for _ in range(N)])   # half-integers: exact ties
    B = draw(st.integers(0, (A - 1) * N + 1))
    return vals, np.arange(A), B


@settings(max_examples=150, deadline=None)
@given(knapsack_instances())
def test_knapsack_multichoice_is_the_ilp_optimum(inst):
    """The DP the CUDA kernel follows against brute force over all action vectors and against scipy's HiGHS MILP of
    lp_methods.py:221-271 (equality budget, one action per arm): same optimum value, same feasibility."""
    import itertools
    from scipy.optimize import Bounds, LinearConstraint, milp
    vals, C, B = inst
    N, A = vals.shape
    a, feas = osel.knapsack_multichoice(vals, C, B)
    best = None
    for combo in itertools.product(range(A), repeat=N):
        if sum(C[c] for c in combo) == B:
            v = sum(vals[i, c] for i, c in enumerate(combo))
            best = v if best is None else max(best, v)
    assert feas == (best is not None)
    if feas:
        assert C[a].sum() == B
        assert abs(vals[np.arange(N), a].sum() - best) < 1e-12
        cons = [LinearConstraint(np.tile(C.astype(float), N)[None, :], B, B)]
        for i in range(N):
            row = np.zeros(N * A)
            row[i * A:(i + 1) * A] = 1
            cons.append(LinearConstraint(row[None, :], 1, 1))
        res = milp(-vals.reshape(-1), constraints=cons, integrality=np.ones(N * A), bounds=Bounds(0, 1))
        assert res.success and abs(-res.fun - best) < 1e-9
