"""Drop-in actor-critic: `MLPActorCriticRMAB` of robust_rmab/algos/rmabppo/rmabppo_core.py:131-445 with the N
per-arm actor / critic MLPs packed into two `[N,P]` device tensors and every per-arm Python loop replaced by
one kernel launch.

Kept from the reference: constructor signature, `step(obs, lamb) -> (a, v, logp, q, p1)` (:198-244),
`act(obs, lamb)`, `act_test(obs)` (budget-constrained greedy, :275-279 / :338-445), `get_lambda(obs)` (:281-287),
`get_probs_for_all(obs, lamb)`, `lambda_net` (a torch module, used with SGD at agent_oracle.py:391),
`pi_list[i]` / `v_list[i]` (per-arm module views sharing storage with the packed rows, so the reference's own loss
closures and `Adam(ac.pi_list[i].parameters())` keep working), attributes (`obs_dim, act_dim, act_type,
one_hot_encode, N, C, B, name, ind`), `__repr__ == "RMABPPO_%i"`, picklable with `torch.save(ac)`.

New: `n_envs=E` replicas per call (obs `[N,E]`, lamb `[E]`; with E = 1 shapes are the reference's), and
`step_device` / `act_test_device` that stay in HBM.  Action sampling uses the Philox rule of oracle/philox.py
(stream ACT, counter = number of `step` calls) instead of torch's global generator.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops

_F32 = torch.float32


def _mlp(sizes):
    layers = []
    for j in range(len(sizes) - 1):
        layers += [nn.Linear(sizes[j], sizes[j + 1]), nn.Tanh() if j < len(sizes) - 2 else nn.Identity()]
    return nn.Sequential(*layers)


class MLPLambdaNet(nn.Module):
    """rmabppo_core.py:108-115."""

    def __init__(self, obs_dim, hidden_sizes=(8, 8)):
        super().__init__()
        self.lambda_net = _mlp([obs_dim] + list(hidden_sizes) + [1])

    def forward(self, obs):
        return torch.squeeze(self.lambda_net(obs), -1)

    def packed(self):
        """W1 b1 W2 b2 W3 b3 flattened (the layout of rmab_lambda_forward)."""
        lin = [m for m in self.lambda_net if isinstance(m, nn.Linear)]
        return torch.cat([t.detach().reshape(-1) for m in lin for t in (m.weight, m.bias)]).to(_F32)


class _ArmNetView(nn.Module):
    """One arm's MLP as a torch module whose Linear parameters are views into the packed `[N,P]` row."""

    def __init__(self, row, in_dim, h, out):
        super().__init__()
        self.net = _mlp([in_dim, h, h, out])
        o = 0
        for m in self.net:
            if isinstance(m, nn.Linear):
                fo, fi = m.weight.shape
                m.weight = nn.Parameter(row[o:o + fo * fi].view(fo, fi)); o += fo * fi
                m.bias = nn.Parameter(row[o:o + fo]); o += fo
        assert o == row.numel()


class _ArmActor(_ArmNetView):
    """MLPCategoricalActor (rmabppo_core.py:73-86)."""

    def _distribution(self, obs):
        return torch.distributions.Categorical(logits=self.net(obs.to(self.net[0].weight.device)))

    def _log_prob_from_distribution(self, pi, act):
        return pi.log_prob(act.to(pi.logits.device))

    def forward(self, obs, act=None):
        pi = self._distribution(obs)
        return pi, (self._log_prob_from_distribution(pi, act) if act is not None else None)


class _ArmCritic(_ArmNetView):
    """MLPCritic (rmabppo_core.py:89-95)."""

    def forward(self, obs):
        return torch.squeeze(self.net(obs.to(self.net[0].weight.device)), -1)


class _ArmList:
    def __init__(self, W, cls, in_dim, h, out):
        self.W, self.cls, self.shape = W, cls, (in_dim, h, out)
        self._cache = {}

    def __len__(self):
        return self.W.shape[0]

    def __getitem__(self, i):
        i = int(i)
        if i not in self._cache:
            self._cache[i] = self.cls(self.W[i], *self.shape)
        return self._cache[i]


class MLPActorCriticRMAB(nn.Module):

    def __init__(self, observation_space, action_space, hidden_sizes=(64, 64), C=None, N=None, B=None, strat_ind=0,
                 one_hot_encode=True, non_ohe_obs_dim=None, state_norm=None, activation=nn.Tanh, n_envs=1,
                 device="cuda", arm0=0, n_total=None):
        super().__init__()
        if activation is not nn.Tanh:
            raise ValueError("the packed kernels implement tanh hidden layers (every reference config uses nn.Tanh)")
        hs = list(hidden_sizes)
        if len(hs) != 2 or hs[0] != hs[1]:
            raise ValueError("packed per-arm nets need two hidden layers of equal width (8, 16, 32 or 64)")
        self.n_states = int(np.asarray(observation_space).shape[0])
        self.obs_dim = self.n_states if one_hot_encode else non_ohe_obs_dim
        self.act_type = 'd'
        self.non_ohe_obs_dim = non_ohe_obs_dim
        self.one_hot_encode = bool(one_hot_encode)
        self.act_dim = int(np.asarray(action_space).shape[0])
        self.N, self.C, self.B = int(N), np.asarray(C), B
        self.hidden_sizes, self.activation, self.state_norm = hs, activation, state_norm
        self.n_envs, self.arm0 = int(n_envs), int(arm0)
        self.n_total = int(n_total if n_total is not None else N)
        self.device_ = torch.device(device)
        if self.device_.type != "cuda":
            raise RuntimeError("MLPActorCriticRMAB runs on CUDA only (no CPU fallback)")
        h = hs[0]
        self.in_dim = self.obs_dim + 1
        self.Ppi, self.Pv = ops.n_params(self.in_dim, h, self.act_dim), ops.n_params(self.in_dim, h, 1)
        self.Wpi = self._default_init(self.N, self.in_dim, h, self.act_dim).to(self.device_)
        self.Wv = self._default_init(self.N, self.in_dim, h, 1).to(self.device_)
        self.lambda_net = MLPLambdaNet(self.n_total, [8, 8])
        self.name = "RMABPPO"
        self.ind = strat_ind
        self._steps = 0
        self._seed = 0
        self._views()

    # -- plumbing ------------------------------------------------------------------------------
    @staticmethod
    def _default_init(N, in_dim, h, out):
        """torch.nn.Linear default init per arm (uniform(+-1/sqrt(fan_in)), rmabppo_core.py:24-29), torch global RNG
        like the reference's module constructors."""
        parts = []
        for fo, fi in ((h, in_dim), (h, h), (out, h)):
            b = 1.0 / math.sqrt(fi)
            parts += [torch.empty(N, fo * fi).uniform_(-b, b), torch.empty(N, fo).uniform_(-b, b)]
        return torch.cat(parts, dim=1).contiguous()

    def _views(self):
        h = self.hidden_sizes[0]
        self.pi_list = _ArmList(self.Wpi, _ArmActor, self.in_dim, h, self.act_dim)
        self.v_list = _ArmList(self.Wv, _ArmCritic, self.in_dim, h, 1)
        self.q_list = None  # the reference allocates N unused Q nets (:153); nothing reads them

    def _net(self, E):
        return ops.net_desc(self.N, E, self.n_states, self.act_dim, self.hidden_sizes[0], self.one_hot_encode,
                            float(self.state_norm) if self.state_norm else 1.0)

    def seed(self, seed):
        self._seed, self._steps = int(seed), 0

    def __repr__(self):
        return "%s_%i" % (self.name, self.ind)

    def __getstate__(self):
        d = dict(self.__dict__)
        d["Wpi"], d["Wv"] = self.Wpi.detach().cpu(), self.Wv.detach().cpu()
        d.pop("pi_list", None); d.pop("v_list", None)
        d["device_"] = str(self.device_)
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.device_ = torch.device(d["device_"] if torch.cuda.is_available() else "cpu")
        self.Wpi, self.Wv = self.Wpi.to(self.device_), self.Wv.to(self.device_)
        self._views()

    def _obs_dev(self, obs, E):
        if isinstance(obs, torch.Tensor):
            t = obs.detach().to(self.device_)
        else:
            t = torch.as_tensor(np.ascontiguousarray(np.asarray(obs)), device=self.device_)
        return t.reshape(self.N, E).to(torch.uint8).contiguous()

    def _lamb_dev(self, lamb, E):
        if isinstance(lamb, torch.Tensor):
            t = lamb.detach().to(self.device_, _F32)
        else:
            t = torch.as_tensor(np.asarray(lamb, dtype=np.float32), device=self.device_)
        return t.reshape(-1).expand(E).contiguous() if t.numel() == 1 else t.reshape(E).contiguous()

    # -- reference API -------------------------------------------------------------------------
    def step_device(self, s, lamb, want_probs=False):
        """s `[N,E]` u8, lamb `[E]` fp32 on device -> device (a, v, logp, p1, probs)."""
        E = s.shape[1]
        r = ops.rng(self._seed, _lib.STREAM_ACT, self._steps, arm0=self.arm0)
        self._steps += 1
        return ops.ac_step(self._net(E), self.Wpi, self.Wv, s, lamb, r=r, want_probs=want_probs)

    def step(self, obs, lamb):
        E = self.n_envs
        a, v, logp, p1, _ = self.step_device(self._obs_dev(obs, E), self._lamb_dev(lamb, E))
        sq = (lambda t: t[:, 0]) if E == 1 else (lambda t: t)
        a_np = sq(a).cpu().numpy().astype(int)
        return (a_np, sq(v).cpu().numpy().astype(np.float64), sq(logp).cpu().numpy().astype(np.float64),
                np.zeros(a_np.shape), sq(p1).cpu().numpy().astype(np.float64))

    def act(self, obs, lamb):
        return self.step(obs, lamb)[0]

    def get_probs_for_all(self, obs, lamb):
        E = self.n_envs
        _, _, _, p1, _ = self.step_device(self._obs_dev(obs, E), self._lamb_dev(lamb, E))
        self._steps -= 1  # no draw is consumed by a pure probability query
        return (p1[:, 0] if E == 1 else p1).cpu().numpy().astype(np.float64)

    def get_lambda_device(self, s):
        scale = 1.0 if self.one_hot_encode else 1.0 / float(self.state_norm)
        params = self.lambda_net.packed().to(self.device_)
        lamb, _ = ops.lambda_forward(params, s, self.n_total, self.arm0, scale)
        return lamb

    def get_lambda(self, obs):
        E = self.n_envs
        lamb = self.get_lambda_device(self._obs_dev(np.asarray(obs).reshape(-1) if E == 1 else obs, E))
        out = lamb.cpu().numpy()
        return out[0] if E == 1 else out

    def act_test_device(self, s):
        """Budget-constrained greedy selection (act_test_deterministic_multiaction, :338-445) for every env."""
        E = s.shape[1]
        lamb = self.get_lambda_device(s)
        _, _, _, p1, probs = self.step_device(s, lamb, want_probs=True)
        self._steps -= 1
        Cc = torch.as_tensor(self.C.astype(np.float32), device=self.device_)
        if self.act_dim == 2 and float(self.C[0]) == 0.0:
            return ops.budget_select_binary(p1, float(self.C[1]), float(self.B), positive_only=True)
        return ops.budget_select_multi(probs, Cc, float(self.B))

    def shard(self, arm0, arm1):
        """The same strategy restricted to arms [arm0, arm1): a shallow copy that holds only those arms' nets (the
        lambda-net stays whole) and selects its share of the GLOBAL budget with `act_test_device_sharded`."""
        import copy
        sh = copy.copy(self)
        sh.N, sh.arm0, sh.n_total = int(arm1) - int(arm0), self.arm0 + int(arm0), self.n_total
        sh.Wpi, sh.Wv = self.Wpi[arm0:arm1].contiguous(), self.Wv[arm0:arm1].contiguous()
        sh._views()
        return sh

    def act_test_device_sharded(self, s_local, rank, world, group=None):
        """`act_test_device` when the arms are sharded over ranks: s_local `[N_local, E]` holds this rank's arms.
        lambda = lambda-net of ALL states (first-layer partial sums, one all-reduce of [E,8]); the budgeted selection is the
        global top-k by radix select with all-reduced digit histograms (ops.budget_select_binary_sharded).  Binary actions."""
        import torch.distributed as dist
        if not (self.act_dim == 2 and float(self.C[0]) == 0.0):
            raise NotImplementedError("sharded act_test covers binary actions; gather the states and use act_test_device")
        scale = 1.0 if self.one_hot_encode else 1.0 / float(self.state_norm)
        params = self.lambda_net.packed().to(self.device_)
        _, h1 = ops.lambda_forward(params, s_local, self.n_total, self.arm0, scale, want_h1=True, want_lamb=False)
        if world > 1:
            dist.all_reduce(h1, group=group)
        lamb, _ = ops.lambda_forward(params, s_local, self.n_total, self.arm0, scale, h1_in=h1)
        _, _, _, p1, _ = self.step_device(s_local, lamb)
        self._steps -= 1
        return ops.budget_select_binary_sharded(p1, float(self.C[1]), float(self.B), self.n_total, rank, world, group,
                                                positive_only=True)

    def act_test(self, obs):
        E = self.n_envs
        a = self.act_test_device(self._obs_dev(np.asarray(obs).reshape(-1) if E == 1 else obs, E))
        return (a[:, 0] if E == 1 else a).cpu().numpy().astype(int)

    def act_test_deterministic(self, obs):
        """rmabppo_core.py:289-334 (binary actions only): action 1 for the top floor(B / C[1]) arms by p(a = 1), with NO
        positivity filter -- unlike the multi-action sweep behind `act_test`, an arm with p1 == 0 is still selected when
        fewer than k arms have p1 > 0."""
        if self.act_dim != 2:
            raise NotImplementedError("act_test_deterministic is only implemented for binary actions (rmabppo_core.py:288)")
        E = self.n_envs
        s = self._obs_dev(np.asarray(obs).reshape(-1) if E == 1 else obs, E)
        lamb = self.get_lambda_device(s)
        _, _, _, p1, _ = self.step_device(s, lamb)
        self._steps -= 1
        a = ops.budget_select_binary(p1, float(self.C[1]), float(self.B), positive_only=False)
        return (a[:, 0] if E == 1 else a).cpu().numpy().astype(int)

    def act_test_deterministic_multiaction(self, obs):
        return self.act_test(obs)

    def return_large_lambda_loss(self, obs, gamma):
        disc_cost = 2 * self.B / (1 - gamma)
        lamb = self.lambda_net(torch.as_tensor(obs, dtype=torch.float32))
        return lamb * (self.B / (1 - gamma) - disc_cost)

    def reset_actor_and_critic_networks(self):
        h = self.hidden_sizes[0]
        self.Wpi.copy_(self._default_init(self.N, self.in_dim, h, self.act_dim))
        self.Wv.copy_(self._default_init(self.N, self.in_dim, h, 1))
