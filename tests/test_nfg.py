"""NFG solver (robust_rmab/nfg_solver.py): the zero-sum LPs against the closed-form equilibria of the reference's own demo
games (nfg_solver.py:109-137) and against the equilibrium conditions on random games.  CPU only (host code)."""
import numpy as np
import pytest

from robustrmab_b200 import nfg_solver as nfg


def test_reference_demo_games():
    # A = [[2,3],[1,4]]: saddle point (row 0, col 0), value 2
    x, y = nfg.solve_game(np.array([[2, 3], [1, 4]]))
    np.testing.assert_allclose(x, [1, 0], atol=1e-9)
    np.testing.assert_allclose(y, [1, 0], atol=1e-9)
    # B = [[3/2,3],[1,4]]: saddle point (row 0, col 0), value 3/2
    x, y = nfg.solve_game(np.array([[1.5, 3], [1, 4]]))
    np.testing.assert_allclose(x, [1, 0], atol=1e-9)
    np.testing.assert_allclose(y, [1, 0], atol=1e-9)
    # C = [[-2,3],[3,-4]]: fully mixed, x = y = (7/12, 5/12), value 1/12
    C = np.array([[-2, 3], [3, -4]])
    x, y = nfg.solve_game(C)
    np.testing.assert_allclose(x, [7 / 12, 5 / 12], atol=1e-9)
    np.testing.assert_allclose(y, [7 / 12, 5 / 12], atol=1e-9)
    assert abs(nfg.get_payoff(C, x, y) - 1 / 12) < 1e-9


@pytest.mark.parametrize("seed", range(8))
def test_equilibrium_conditions_random(seed):
    rs = np.random.RandomState(seed)
    m, n = rs.randint(1, 7), rs.randint(1, 7)
    A = rs.randn(m, n) * 3
    x, y = nfg.solve_game(A)
    assert abs(x.sum() - 1) < 1e-9 and abs(y.sum() - 1) < 1e-9 and (x >= 0).all() and (y >= 0).all()
    v = nfg.get_payoff(A, x, y)
    assert (A @ y).max() <= v + 1e-7          # no profitable row deviation
    assert (x @ A).min() >= v - 1e-7          # no profitable column deviation
    assert abs(v - nfg.game_value(A)) < 1e-7


def test_minimax_regret():
    P = np.array([[5.0, 1.0], [3.0, 2.0], [4.0, 0.5]])
    x, y = nfg.solve_minimax_regret(P)
    reg = P - P.max(axis=0)
    v = nfg.get_payoff(reg, x, y)
    assert (reg @ np.asarray(y)).max() <= v + 1e-7 and (np.asarray(x) @ reg).min() >= v - 1e-7
    x2, y2 = nfg.solve_minimax_regret_with_regret_array(reg)
    np.testing.assert_allclose(nfg.get_payoff(reg, x2, y2), v, atol=1e-9)
    # degenerate single-row case (:57-61)
    x, y = nfg.solve_minimax_regret([[1.0, 2.0, 3.0]])
    assert list(x) == [1.0] and list(y) == [1.0, 0.0, 0.0]
