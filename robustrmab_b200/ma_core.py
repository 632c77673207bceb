"""Drop-in MA-RMABPPO actor-critic (rollout side): `RMABLambdaNatureOracle` of
robust_rmab/algos/ma_rmabppo/ma_rmabppo_core.py:148-331.

Kept: constructor signature, `step(obs, lamb)` -> the reference's 9-tuple (a_agent, v_agent, logp_agent, q_agent=0,
a_nature, v_nature, logp_nature, q_nature=0, p1_agent), `get_nature_action(obs)` (deterministic mu, :266-271),
`bound_nature_actions`, `pi_nature` / `v_nature` / `lambda_net` as torch modules, `__repr__ == "MA-RMABPPO_%i"`.
Device mapping: the per-arm agent actors / critics are packed `[N,P]` tensors evaluated by rmab_ac_step; the agent
critics' first-layer block over the nature inputs is a `[N,h,D]` tensor applied to all arms and envs as ONE dense
product (rmab_gemm3h: the package's TMA-fed tcgen05 kernel) whose result the kernel adds to the first-layer
pre-activation; the Gaussian nature actor's sample / log-prob is rmab_gaussian_sample; the dense nature nets
(N -> h -> h -> D, 2N -> h -> h -> 1) are evaluated and trained by the rmab_nat_* kernels on a packed copy of the
torch modules the drop-in exposes (nature_nets.py).  The MA update lives in nature_oracle.MARolloutTrainer.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from .core import MLPLambdaNet, MLPActorCriticRMAB, _mlp

_F32 = torch.float32


class MLPGaussianActor(nn.Module):
    """ma_rmabppo_core.py:84-107."""

    def __init__(self, obs_dim, act_dim, hidden_sizes):
        super().__init__()
        self.log_std = nn.Parameter(torch.as_tensor(-0.5 * np.ones(act_dim, dtype=np.float32)))
        self.mu_net = _mlp([obs_dim] + list(hidden_sizes) + [act_dim])


class MLPCritic(nn.Module):
    def __init__(self, obs_dim, hidden_sizes):
        super().__init__()
        self.v_net = _mlp([obs_dim] + list(hidden_sizes) + [1])

    def forward(self, obs):
        return torch.squeeze(self.v_net(obs), -1)


class RMABLambdaNatureOracle(nn.Module):

    def __init__(self, observation_space, action_space_agent, nature_parameter_ranges, action_dim_nature, env,
                 hidden_sizes=(64, 64), C=None, N=None, B=None, one_hot_encode=True, non_ohe_obs_dim=None, state_norm=1,
                 nature_state_norm=1, strat_ind=0, activation=nn.Tanh, n_envs=1, device="cuda", arm_shard=None):
        super().__init__()
        self.observation_space, self.action_space_agent = observation_space, action_space_agent
        self.nature_parameter_ranges = nature_parameter_ranges
        self.n_states = int(np.asarray(observation_space).shape[0])
        self.obs_dim = self.n_states if one_hot_encode else non_ohe_obs_dim
        self.act_type = 'd'
        self.non_ohe_obs_dim, self.one_hot_encode = non_ohe_obs_dim, bool(one_hot_encode)
        self.act_dim_agent = int(np.asarray(action_space_agent).shape[0])
        self.act_dim_nature = int(action_dim_nature)
        self.N, self.C, self.B = int(N), np.asarray(C), B
        self.hidden_sizes, self.activation = list(hidden_sizes), activation
        self.state_norm, self.nature_state_norm = state_norm, nature_state_norm
        self.env = env
        self.n_envs = int(n_envs)
        self.device_ = torch.device(device)
        h = self.hidden_sizes[0]
        self.in_dim = self.obs_dim + 1
        D = self.act_dim_nature
        self.pi_nature = MLPGaussianActor(self.N, D, self.hidden_sizes).to(self.device_)
        self.v_nature = MLPCritic(self.N + self.N, self.hidden_sizes).to(self.device_)
        self.Wpi = MLPActorCriticRMAB._default_init(self.N, self.in_dim, h, self.act_dim_agent).to(self.device_)
        # agent critic i = MLP(in + D -> h -> h -> 1) (:203): nn.Linear init bound uses fan_in = in + D for layer 1
        b = 1.0 / math.sqrt(self.in_dim + D)
        Wv = MLPActorCriticRMAB._default_init(self.N, self.in_dim, h, 1)
        Wv[:, :h * self.in_dim + h].uniform_(-b, b)
        self.Wv = Wv.to(self.device_)
        self.W1n = torch.empty(self.N, h, D).uniform_(-b, b).to(self.device_)
        # arm shard: the per-arm nets of arms [arm0, arm0 + n_local) only; the dense nature nets and the
        # lambda net stay whole on every rank.  All ranks draw the same full tensors (same torch seed) and keep a slice.
        self.arm0, self.n_local = 0, self.N
        if arm_shard is not None:
            a0, a1 = int(arm_shard[0]), int(arm_shard[1])
            self.arm0, self.n_local = a0, a1 - a0
            self.Wpi, self.Wv, self.W1n = (self.Wpi[a0:a1].clone(), self.Wv[a0:a1].clone(), self.W1n[a0:a1].clone())
        self.lambda_net = MLPLambdaNet(self.N, [8, 8])
        self.name = "MA-RMABPPO"
        self.ind = strat_ind
        self._steps, self._seed = 0, 0

    def __repr__(self):
        return "%s_%i" % (self.name, self.ind)

    def seed(self, seed):
        self._seed, self._steps = int(seed), 0

    def counter_base(self, dev_word=None):
        """As environments.counter_base: route the step counter of the NATURE / ACT streams through a device word."""
        self._t_word, self._t_origin = dev_word, (self._steps if dev_word is not None else 0)

    def _step_rng(self, stream, arm0=0):
        word = getattr(self, "_t_word", None)
        if word is None:
            return ops.rng(self._seed, stream, self._steps, arm0=arm0)
        return ops.rng(self._seed, stream, self._steps - self._t_origin, arm0=arm0, t_base=word)

    def _net(self, E, extra):
        return ops.net_desc(self.n_local, E, self.n_states, self.act_dim_agent, self.hidden_sizes[0], self.one_hot_encode,
                            float(self.state_norm) if self.state_norm else 1.0, n_extra=self.act_dim_nature if extra else 0)

    # -- nature ------------------------------------------------------------------------------------
    def _nature_obs(self, s):
        """[E,N] float input of the nature nets: obs (/state_norm if not one-hot) / nature_state_norm (:283-289)."""
        x = s.t().to(_F32)
        if not self.one_hot_encode:
            x = x / float(self.state_norm)
        return x / float(self.nature_state_norm)

    def bound_nature_actions(self, a_nature, state=None, reshape=True):
        return self.env.bound_nature_actions(a_nature, state=state, reshape=reshape)

    def bound_nature_device(self, a_nature, s):
        """tanh-bound raw nature actions `[E,D]` into the env's sampled ranges for states s `[N,E]`; returns `[E,D]`."""
        rg = self.env._ranges()
        E = a_nature.shape[0]
        if rg.dim() == 4:                                   # ARMMAN: [N,S,A,2] looked up at the arm's state
            sel = rg[torch.arange(self.N, device=rg.device)[:, None], s.long()]      # [N,E,A,2]
            lo = sel[..., 0].permute(1, 0, 2).reshape(E, -1).contiguous()
            hi = sel[..., 1].permute(1, 0, 2).reshape(E, -1).contiguous()
        else:                                               # counterexample [N,2], SIS [N,4,2]: state-independent bounds
            key = (E, rg.data_ptr(), rg._version)
            if getattr(self, "_bounds_key", None) != key:
                self._bounds = (rg[..., 0].reshape(1, -1).expand(E, -1).contiguous(),
                                rg[..., 1].reshape(1, -1).expand(E, -1).contiguous())
                self._bounds_key = key
            lo, hi = self._bounds
        return ops.bound_nature(a_nature.contiguous(), lo, hi)

    def get_nature_action_device(self, s):
        return self.nature_nets().actor_mu(s.contiguous(), self.nature_scales()[0])

    def get_nature_action(self, obs):
        s = torch.as_tensor(np.asarray(obs).reshape(self.N, -1).astype(np.uint8), device=self.device_)
        mu = self.get_nature_action_device(s).cpu().numpy()
        return mu[0] if mu.shape[0] == 1 else mu

    def nature_critic_input(self, s, a):
        """[obs, a_agent] rows of the nature critic (ma_rmabppo_core.py:318-325): s, a `[N,E]` u8 -> `[E,2N]`."""
        return torch.cat([s.t().to(_F32) / (1.0 if self.one_hot_encode else float(self.state_norm)), a.t().to(_F32)], dim=1)

    # -- step --------------------------------------------------------------------------------------
    def nature_nets(self):
        """The dense nature nets in packed device form (nature_nets.NatureNets), freshly packed from the torch modules."""
        from .nature_nets import NatureNets
        return NatureNets(self, self.arm0, self.n_local)

    def nature_scales(self):
        """(actor input scale in step(), critic obs scale in step()): obs (/state_norm if not one-hot) / nature_state_norm
        (:283-289) and obs (/state_norm if not one-hot) (:318-325)."""
        c = 1.0 if self.one_hot_encode else 1.0 / float(self.state_norm)
        return c / float(self.nature_state_norm), c

    def step_device(self, s, lamb, eps=None, u=None, w1n_split=None, s_full=None, nat_nets=None, want_v_nature=True):
        """s `[N,E]` u8, lamb `[E]`.  Returns dict of device tensors: a_agent, v_agent, logp_agent, p1 `[N,E]`;
        a_nature, a_nature_bounded `[E,D]`; logp_nature, v_nature `[E]`.  `w1n_split` = ops.split16_rows(W1n.reshape(-1, D))
        when the caller keeps it across steps (W1n is constant during a rollout); `nat_nets` likewise the packed nature
        nets.  Arm-sharded: s holds the local arms, `s_full` `[N_total,E]` all of them (input of the nature actor);
        v_nature needs every arm's action and is left to the caller (None), as it is with want_v_nature = False."""
        E = s.shape[1]
        sharded = s_full is not None
        s_all = s_full if sharded else s
        nets = nat_nets if nat_nets is not None else self.nature_nets()
        sc_a, sc_c = self.nature_scales()
        with torch.no_grad():
            mu = nets.actor_mu(s_all.contiguous(), sc_a)
            r_n = self._step_rng(_lib.STREAM_NATURE)
            log_std = nets.actor.ls(nets.actor.p)
            a_nat, logp_nat = ops.gaussian_sample(mu, log_std, r=None if eps is not None else r_n, eps=eps)
            nat_b = self.bound_nature_device(a_nat, s_all)
            # first-layer nature term of every agent critic as one dense product: [N*h, D] x [D, E] -> [N, h, E]
            D = self.act_dim_nature
            if w1n_split is not None and len(w1n_split) == 2:        # tf32 (hi, lo) pair: the library-GEMM A/B arm
                extra = ops.matmul_3xtf32(w1n_split, nat_b.t()).reshape(self.n_local, -1, E).contiguous()
            else:
                w1 = w1n_split if w1n_split is not None else ops.split16_rows(self.W1n.reshape(-1, D))
                extra = ops.gemm3h(w1, ops.split16_rows(nat_b), D, n_fast=True).reshape(self.n_local, -1, E)
            r_a = self._step_rng(_lib.STREAM_ACT, arm0=self.arm0)
            a, v, logp, p1, _ = ops.ac_step(self._net(E, True), self.Wpi, self.Wv, s, lamb,
                                            r=None if u is not None else r_a, u=u, extra=extra)
            v_nat = None
            if want_v_nature and not sharded:
                v_nat = nets.critic_values(s.view(-1, 1, E), a.view(-1, 1, E), sc_c, 1, E)
        self._steps += 1
        return dict(a_agent=a, v_agent=v, logp_agent=logp, p1=p1, a_nature=a_nat, a_nature_bounded=nat_b,
                    logp_nature=logp_nat, v_nature=v_nat)

    def step(self, obs, lamb):
        E = self.n_envs
        o = obs.detach().cpu().numpy() if isinstance(obs, torch.Tensor) else np.asarray(obs)
        s = torch.as_tensor(np.ascontiguousarray(o.reshape(self.N, E).astype(np.uint8)), device=self.device_)
        lam = torch.as_tensor(np.asarray(lamb.detach().cpu() if isinstance(lamb, torch.Tensor) else lamb,
                                         dtype=np.float32).reshape(-1), device=self.device_)
        lam = lam.expand(E).contiguous() if lam.numel() == 1 else lam
        d = self.step_device(s, lam)
        n = lambda t: t.cpu().numpy().astype(np.float64)
        sq = (lambda t: t[:, 0]) if E == 1 else (lambda t: t)
        a = sq(d["a_agent"]).cpu().numpy().astype(int)
        a_nat, v_nat, lp_nat = n(d["a_nature"]), n(d["v_nature"]), n(d["logp_nature"])
        if E == 1:
            a_nat, v_nat, lp_nat = a_nat[0], v_nat[0], lp_nat[0]
        return (a, n(sq(d["v_agent"])), n(sq(d["logp_agent"])), np.zeros(a.shape), a_nat, v_nat, lp_nat, 0,
                n(sq(d["p1"])))
