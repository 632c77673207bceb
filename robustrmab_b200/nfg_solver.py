"""Normal-form game solver behind the double oracle: the functions of robust_rmab/nfg_solver.py (:14-107) with the
`nashpy` dependency (Lemke-Howson / vertex / support enumeration on a zero-sum game, :20-39) replaced by the two
linear programmes of the zero-sum game, solved with scipy's HiGHS.  Matrices here are a handful of strategies per
side, so this is host code ("negligible compute, pure plumbing").

`nash.Game(payoffs)` with one matrix is the zero-sum game (A, -A): the row player (agent) maximises x^T A y, the column
player (nature) minimises it.  Any pair (maximin x, minimax y) is a Nash equilibrium, and the game value is unique --
WHICH equilibrium is returned when several exist is solver-dependent in the reference too (it takes the first one
Lemke-Howson enumerates).  PARITY UNPINNED vs nashpy (absent from the image): pinned instead to the closed-form
equilibria of the reference's own demo games (:109-137) and to the equilibrium conditions on random games (tests).
"""
import numpy as np
from scipy.optimize import linprog


def _maximin(A):
    """x, v maximising v s.t. (A^T x)_j >= v for all j, x in the simplex."""
    m, n = A.shape
    c = np.zeros(m + 1)
    c[-1] = -1.0
    A_ub = np.hstack([-A.T, np.ones((n, 1))])
    res = linprog(c, A_ub=A_ub, b_ub=np.zeros(n), A_eq=np.hstack([np.ones((1, m)), np.zeros((1, 1))]), b_eq=[1.0],
                  bounds=[(0, None)] * m + [(None, None)], method="highs")
    if res.status != 0:
        raise RuntimeError(f"zero-sum LP failed: {res.message}")
    x = np.clip(res.x[:m], 0.0, None)
    return x / x.sum(), float(res.x[-1])


def solve_game(payoffs):
    """Mixed equilibrium (row strategy, column strategy) of the zero-sum game with row payoffs `payoffs` (:14-44)."""
    A = np.asarray(payoffs, dtype=np.float64)
    if A.ndim != 2 or A.size == 0:
        raise ValueError("payoffs must be a non-empty matrix")
    x, v = _maximin(A)
    y, w = _maximin(-A.T)
    assert abs(v + w) <= 1e-7 * max(1.0, abs(v)), "LP duality gap"
    return x, y


def game_value(payoffs):
    return _maximin(np.asarray(payoffs, dtype=np.float64))[1]


def solve_minimax_regret(payoffs):
    """Minimax-regret optimal mixture (:47-69): subtract each column's maximum, then solve the game."""
    if len(payoffs) == 1:
        eq2 = np.zeros(len(payoffs[0]))
        eq2[0] = 1.0
        return [1.0], eq2
    mod = np.array(payoffs, dtype=np.float64)
    return solve_game(mod - mod.max(axis=0))


def solve_minimax_regret_with_regret_array(regret_array):
    """As above for a matrix that already holds regrets (:71-95)."""
    if len(regret_array) == 1:
        eq2 = np.zeros(len(regret_array[0]))
        eq2[0] = 1.0
        return [1.0], eq2
    return solve_game(np.array(regret_array, dtype=np.float64))


def get_payoff(payoffs, agent_eq, nature_eq):
    """Expected row payoff under the mixed strategies (:97-107)."""
    return float(np.asarray(agent_eq, dtype=np.float64) @ np.asarray(payoffs, dtype=np.float64)
                 @ np.asarray(nature_eq, dtype=np.float64))
