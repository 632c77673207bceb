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
