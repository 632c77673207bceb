"""Hawkins (Lagrangian-relaxation) agent policy on the device: `HawkinsAgentPolicy` of
robust_rmab/baselines/agent_baselines.py:35-116, whose reference implementation calls two Gurobi programs per step
(`lp_methods.hawkins` :11-105 and `action_knapsack` :221-271).

Reformulation: for a fixed lambda the LP's minimal L_i is the optimal value function of
arm i's lambda-charged MDP, so   obj(lambda) = sum_i V_i(s_i; lambda) + lambda * B / (1 - gamma)   (:63) is evaluated
for a whole grid of lambda-probes by ONE batched value iteration (rmab_vi_batched in its absolutely-converged
stop mode -- mdptoolbox's own iteration bound leaves V off by a per-MDP constant), minimised over
the grid and refined by re-gridding around the minimiser (obj is convex piecewise linear in lambda).  The action is
the knapsack over Q(s_i, a; lambda*) = R - lambda* C[a] + gamma T.L (:69-91) with the equality budget
`sum x C == B` (lp_methods.py:250): for binary costs C = [0, c] that is action 1 for exactly B / c arms with the largest
Q(s,1) - Q(s,0) (rmab_budget_select_binary); for general integer costs (SIS: C = [0,1,2]) the exact multiple-choice
knapsack dynamic programme rmab_knapsack_multichoice.  Q for all envs comes from ONE batched VI at the envs' lambdas.
PARITY UNPINNED against Gurobi (absent); pinned against HiGHS restatements of the LP and of the ILP at small N (tests).
"""
import numpy as np
import torch

from . import ops

_F64 = torch.float64


class HawkinsAgentPolicy:
    def __init__(self, N, T, R, C, B, gamma, ind, n_probes=64, n_refine=3, eps=1e-8, device="cuda"):
        self.N, self.B, self.gamma, self.ind = int(N), B, float(gamma), ind
        self.T, self.R, self.C = np.asarray(T, np.float64), np.asarray(R, np.float64), np.asarray(C, np.float64)
        self.name = "Hawkins"
        self.n_probes, self.n_refine, self.eps = int(n_probes), int(n_refine), float(eps)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HawkinsAgentPolicy solves on CUDA only (no CPU fallback)")
        self.binary = self.C.shape[0] == 2 and self.C[0] == 0
        if not self.binary and not (np.abs(self.C - np.round(self.C)) < 1e-12).all():
            raise ValueError("the multiple-choice knapsack needs integer action costs")
        self._T = torch.as_tensor(self.T, dtype=_F64, device=self.device).contiguous()
        self._R = torch.as_tensor(self.R, dtype=_F64, device=self.device).contiguous()
        self._C = torch.as_tensor(self.C, dtype=_F64, device=self.device).contiguous()
        self.lambda_lim = float(self.R.max() / (self.C[self.C > 0].min() * (1 - self.gamma)))   # agent_baselines.py:61
        self.max_ws_bytes = 1 << 30            # value-function scratch per refinement chunk

    def __repr__(self):
        return "%s_%s" % (self.name, self.ind)

    def objective(self, s, lambdas):
        """obj[e, k] = sum_i V_i(s_i(e); lambda_k) + lambda_k B / (1 - gamma) for states s `[N,E]`; also returns V, Q."""
        V, _, _, _, Q = ops.vi_batched(self._T, self._R, self._C, lambdas, self.gamma, self.eps, stop="abs", want_q=True)
        idx = s.long()[:, None, :].expand(-1, lambdas.shape[0], -1)                  # [N,L,E]
        vs = torch.gather(V, 2, idx)                                                 # V[n,l,s[n,e]]
        obj = vs.sum(dim=0).t() + lambdas[None, :] * (self.B / (1.0 - self.gamma))   # [E,L]
        return obj, V, Q

    def _objective_per_env(self, s, grids):
        """obj[e, k] = sum_i V_i(s_i(e); grids[e, k]) + grids[e, k] B / (1 - gamma): every env has its OWN probes.  One
        batched VI per chunk of envs over the chunk's flattened grids (the kernel shares a probe list between arms, so
        env e's probes are also solved for the other envs' states and discarded: the cost of batching E refinements)."""
        N, E = s.shape
        P = grids.shape[1]
        S = self._T.shape[1]
        per_env = N * P * S * 8 * 2                                   # V (fp64) + the gathered copy
        chunk = max(1, min(E, int(self.max_ws_bytes // max(per_env, 1))))
        obj = torch.empty((E, P), dtype=_F64, device=self.device)
        for e0 in range(0, E, chunk):
            e1 = min(E, e0 + chunk)
            lam = grids[e0:e1].reshape(-1).contiguous()
            V, _, _, _, _ = ops.vi_batched(self._T, self._R, self._C, lam, self.gamma, self.eps, stop="abs")
            V = V.view(N, e1 - e0, P, S)
            idx = s[:, e0:e1].long()[:, :, None, None].expand(-1, -1, P, 1)
            obj[e0:e1] = torch.gather(V, 3, idx)[..., 0].sum(dim=0) + grids[e0:e1] * (self.B / (1.0 - self.gamma))
        return obj

    def hawkins_lambda_device(self, s):
        """lambda* per env `[E]`: a shared grid of n_probes over [0, lambda_lim], then n_refine rounds that re-grid
        around EACH env's own minimiser (obj is convex piecewise linear in lambda, so the minimiser stays bracketed).
        Final resolution: lambda_lim * (2 / (n_probes - 1)) ** n_refine / (n_probes - 1) for every env."""
        E = s.shape[1]
        unit = torch.linspace(0, 1, self.n_probes, dtype=_F64, device=self.device)
        grid = (self.lambda_lim * unit).contiguous()
        obj, _, _ = self.objective(s, grid)
        best = grid[obj.argmin(dim=1)]                                                # [E]
        step = self.lambda_lim / (self.n_probes - 1)
        for _ in range(self.n_refine):
            lo = torch.clamp(best - step, min=0.0)
            hi = torch.clamp(best + step, max=self.lambda_lim)
            grids = (lo[:, None] + (hi - lo)[:, None] * unit[None, :]).contiguous()   # [E,P]
            obj = self._objective_per_env(s, grids)
            best = torch.gather(grids, 1, obj.argmin(dim=1, keepdim=True))[:, 0]
            step = 2.0 * step / (self.n_probes - 1)
        return best

    def act_test_device(self, s):
        lam = self.hawkins_lambda_device(s)                                          # [E]
        N, E = s.shape
        A = self.C.shape[0]
        # Q at each env's own lambda*: one batched VI with the envs' lambdas as the probes
        _, _, _, _, Q = ops.vi_batched(self._T, self._R, self._C, lam.contiguous(), self.gamma, self.eps, stop="abs", want_q=True)
        idx = s.long()[:, :, None, None].expand(-1, -1, 1, A)                        # Q[n, e, s[n,e], :]
        q = torch.gather(Q, 2, idx)[:, :, 0]                                         # [N,E,A]
        if self.binary:
            gap = (q[..., 1] - q[..., 0]).to(torch.float32).contiguous()             # [N,E]
            return ops.budget_select_binary(gap, 1.0, float(self.B) / float(self.C[1]))
        if abs(float(self.B) - round(float(self.B))) > 1e-9:
            raise ValueError("budget %r cannot be met with equality by integer costs (lp_methods.py:250 is infeasible)" % (self.B,))
        a, status = ops.knapsack_multichoice(q.permute(0, 2, 1).contiguous(), self.C, self.B)
        if bool(status.any()):
            raise ValueError("knapsack infeasible: no action vector spends exactly B = %r" % (self.B,))
        return a

    def act_test(self, o, just_hawkins_lambda=False):
        cur = o.numpy() if isinstance(o, torch.Tensor) else np.asarray(o)
        s = torch.as_tensor(cur.reshape(self.N, -1).astype(np.uint8), device=self.device)
        if just_hawkins_lambda:
            return self.hawkins_lambda_device(s).cpu().numpy()
        a = self.act_test_device(s).cpu().numpy().astype(int)
        return a[:, 0] if a.shape[1] == 1 else a
