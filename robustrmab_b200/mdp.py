"""Drop-in for the one class the reference uses from its vendored pymdptoolbox: `mdptoolbox.mdp.ValueIteration`
(mdptoolbox/mdp.py:1180-1410, called at robust_rmab/simulator.py:504-521), solving on the device.

    vi = ValueIteration(transitions, reward, discount, epsilon=0.01, max_iter=1000, initial_value=0,
                        stop_criterion='full')
    vi.run();  vi.V, vi.policy, vi.iter, vi.max_iter, vi.thresh

Kept: argument meaning (transitions `[A,S,S]`; reward `[S]`, `[S,A]` or `[A,S,S]`, collapsed like
MDP._computeReward :292-302), the locally added `stop_criterion` ('full': max(V - Vprev), 'fast': span), the
iteration bound of `_boundIter` (:1320-1364) that always overrides the user's `max_iter` (:1309-1312), `V` / `policy`
as tuples, and the exceptions `_boundIter` raises in the constructor (same expression on numpy scalars: OverflowError
for a zero span, ValueError for a zero contraction factor).
`ValueIterationBatch` solves N MDPs of the same shape in one launch (what the per-arm loop of simulator.py:504-515 is).
"""
import math
import time

import numpy as np
import torch

from . import ops


def _collapse_reward(P, reward):
    """MDP._computeReward: -> R[A][S]."""
    A, S, _ = P.shape
    r = np.asarray(reward, dtype=np.float64)
    if r.ndim == 1:
        return np.tile(r.reshape(1, S), (A, 1))
    if r.ndim == 2:
        return r.T.copy()                                   # [S,A] -> [A,S]
    return np.einsum('ast,ast->as', P, r)                   # sum_s' P[a][s,s'] R[a][s,s']


def _bound_iter(P, R, discount, epsilon):
    """ValueIteration._boundIter with the reference's exceptions."""
    h = P.min(axis=(0, 1))
    k = 1 - h.sum()
    value = R.max(axis=0)                                   # Bellman backup of V = 0
    span = value.max() - value.min()                        # numpy scalar: x / 0 -> inf (warning), like the reference
    with np.errstate(divide="ignore"):
        max_iter = math.log((epsilon * (1 - discount) / discount) / span) / math.log(discount * k)
    return int(math.ceil(max_iter))


class ValueIterationBatch:
    """N MDPs at once: transitions `[N,A,S,S]`, reward `[N,S]`, `[N,S,A]` or `[N,A,S,S]`."""

    def __init__(self, transitions, reward, discount, epsilon=0.01, max_iter=1000, initial_value=0,
                 stop_criterion='full', device="cuda"):
        P = np.asarray(transitions, dtype=np.float64)
        if P.ndim != 4 or P.shape[2] != P.shape[3]:
            raise ValueError("transitions must be [N,A,S,S]")
        if not (0 < discount < 1):
            raise ValueError("the device solver needs 0 < discount < 1")
        if np.any(np.asarray(initial_value) != 0):
            raise NotImplementedError("only initial_value = 0 (the reference's call sites) is supported")
        if np.abs(P.sum(-1) - 1).max() > 1e-8 or P.min() < 0:
            raise ValueError("transitions must be row-stochastic (mdptoolbox.util.check)")
        self.P, self.discount, self.epsilon = P, float(discount), float(epsilon)
        N, A, S, _ = P.shape
        r = np.asarray(reward, dtype=np.float64)
        self.R = np.stack([_collapse_reward(P[n], r[n]) for n in range(N)])      # [N,A,S]
        self.stop_criterion = stop_criterion
        self.thresh = epsilon * (1 - discount) / discount
        self.device = torch.device(device)
        self.V = self.policy = self.iter = self.max_iter = None

    def run(self):
        t0 = time.time()
        T = torch.as_tensor(np.ascontiguousarray(self.P.transpose(0, 2, 1, 3)), device=self.device)     # [N,S,A,S]
        Rsa = torch.as_tensor(np.ascontiguousarray(self.R.transpose(0, 2, 1)), device=self.device)     # [N,S,A]
        V, pol, it, mi = ops.vi_general(T, Rsa, self.discount, self.epsilon, stop=self.stop_criterion)
        self.V, self.policy = V.cpu().numpy(), pol.cpu().numpy().astype(np.int64)
        self.iter, self.max_iter = it.cpu().numpy(), mi.cpu().numpy()
        self.time = time.time() - t0
        return self


class ValueIteration:

    def __init__(self, transitions, reward, discount, epsilon=0.01, max_iter=1000, initial_value=0,
                 stop_criterion='full', device="cuda"):
        P = np.asarray([np.asarray(p, dtype=np.float64) for p in transitions])
        self._batch = ValueIterationBatch(P[None], np.asarray(reward, dtype=np.float64)[None], discount, epsilon, max_iter,
                                          initial_value, stop_criterion, device)
        self.P, self.R = P, self._batch.R[0]
        self.S, self.A = P.shape[1], P.shape[0]
        self.discount, self.epsilon, self.stop_criterion = float(discount), float(epsilon), stop_criterion
        self.thresh = self._batch.thresh
        self.max_iter = _bound_iter(P, self.R, self.discount, self.epsilon)    # raises like the reference's constructor
        self.V = tuple(np.zeros(self.S))
        self.policy, self.iter, self.time = None, 0, None

    def run(self):
        b = self._batch.run()
        if int(b.iter[0]) < 0:
            raise ZeroDivisionError("float division by zero")     # unreachable after the constructor's check
        self.V = tuple(b.V[0].tolist())
        self.policy = tuple(int(x) for x in b.policy[0])
        self.iter, self.time = int(b.iter[0]), b.time
        assert int(b.max_iter[0]) == self.max_iter
