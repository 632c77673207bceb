"""MA-RMABPPO nature-oracle epoch loop restatement on the CPU (oracle, test-only).

Follows robust_rmab/algos/ma_rmabppo/nature_oracle.py best_response_per_cpu:
  :664       lambda = lambda_net(o) once per epoch
  :673-735   per step: ac.step, bound, env.step, 50 counterfactual env steps with agent_pol.act_test, r_nature
  :754-781   bootstrap ac.step, buf.finish_path(v_agent, last_costs, v_nature), env.reset
  :486-576   update: agent pi / v loops every epoch; lambda SGD + nature pi / v loops every lamb_update_freq epochs
  MA_RMABPPO_Buffer.finish_path / get  :107-193
Randomness: the Philox convention of the CUDA path (ACT / ENV draws of step t share a block, NATURE normals with
t = step, CF draws with t = step * 50 + trial, RESET with t = epoch).
"""
import copy

import numpy as np
import torch

from . import buffer as obuf
from . import envs as oenv
from . import ma as oma
from . import nets, philox
from . import ppo as oppo

F32 = np.float32
NUM_TEST_POLICY_RUNS = 50


def _dense(sizes):
    layers = []
    for j in range(len(sizes) - 1):
        layers += [torch.nn.Linear(sizes[j], sizes[j + 1]), torch.nn.Tanh() if j < len(sizes) - 2 else torch.nn.Identity()]
    return torch.nn.Sequential(*layers)


class CpuMARMABPPO:
    def __init__(self, domain, Wpi, Wv, W1n, lam_params, mu_net, log_std, v_nature, ranges, R, C, n_envs, T, hidden, B,
                 gamma, lam_other, clip_ratio, pi_lr_agent, pi_lr_nature, vf_lr_agent, vf_lr_nature, lm_lr, pi_iters,
                 v_iters, epochs, lamb_update_freq, final_train_lambdas, ent0, ent1, seed, pop=None, one_hot=True,
                 state_norm=1.0, nature_state_norm=1.0):
        self.domain, self.pop = domain, pop
        self.S = {oenv.DOMAIN_CE: 2, oenv.DOMAIN_ARMMAN: 3}.get(domain, (pop or 0) + 1)
        self.A = len(C)
        self.one_hot, self.state_norm, self.nsn = bool(one_hot), float(state_norm if state_norm else 1.0), float(nature_state_norm)
        self.in_dim = (self.S if self.one_hot else 1) + 1
        self.N, self.E, self.T, self.h = Wpi.shape[0], n_envs, T, hidden
        self.D = W1n.shape[2]
        self.opt = oppo.ArmOptim(Wpi, Wv, pi_lr_agent, vf_lr_agent)
        self.W1n = torch.nn.Parameter(torch.as_tensor(np.array(W1n, F32)))
        self.opt_v = torch.optim.Adam([self.opt.Wv, self.W1n], lr=vf_lr_agent)
        self.lam_params = [torch.nn.Parameter(torch.as_tensor(np.array(p, F32))) for p in lam_params]
        self.mu_net, self.v_nature = copy.deepcopy(mu_net), copy.deepcopy(v_nature)
        self.log_std = torch.nn.Parameter(torch.as_tensor(np.array(log_std, F32)))
        self.opt_pi_nat = torch.optim.Adam(list(self.mu_net.parameters()) + [self.log_std], lr=pi_lr_nature)
        self.opt_v_nat = torch.optim.Adam(self.v_nature.parameters(), lr=vf_lr_nature)
        self.ranges, self.R, self.C = np.asarray(ranges, F32), np.asarray(R, F32), np.asarray(C, F32)
        self.B, self.gamma, self.lam_other, self.clip, self.lm_lr = B, gamma, lam_other, clip_ratio, lm_lr
        self.pi_iters, self.v_iters, self.epochs = pi_iters, v_iters, epochs
        self.freq, self.final, self.ent0, self.ent1 = lamb_update_freq, final_train_lambdas, ent0, ent1
        self.seed, self.epoch_idx, self.t_global = seed, 0, 0
        self.s = oenv.reset_states(seed, 0, self.S, self.E, self.N)

    def _lambda(self, s, scale=1.0):
        """Raw observation for the rollout's prediction (nature_oracle.py:664); obs / state_norm inside
        compute_loss_lambda for non-one-hot nets (:447-459)."""
        x = torch.as_tensor(s.T.astype(F32)) * F32(scale)
        p = self.lam_params
        return (torch.tanh(torch.tanh(x @ p[0].T + p[1]) @ p[2].T + p[3]) @ p[4].T + p[5])[:, 0]

    def _bound(self, a_nat, s):
        """[E,D] raw -> bounded, ranges looked up like the env (CE [N,2]; ARMMAN [N,S,A,2] at the arm's state)."""
        N, E = s.shape
        if self.ranges.ndim == 2:
            lo, hi = np.broadcast_to(self.ranges[:, 0], (E, N)), np.broadcast_to(self.ranges[:, 1], (E, N))
        elif self.ranges.ndim == 4:
            sel = self.ranges[np.arange(N)[:, None], s.astype(np.int64)]            # [N,E,A,2]
            lo = sel[..., 0].transpose(1, 0, 2).reshape(E, -1)
            hi = sel[..., 1].transpose(1, 0, 2).reshape(E, -1)
        else:
            lo = np.broadcast_to(self.ranges[..., 0].reshape(1, -1), (E, self.D))
            hi = np.broadcast_to(self.ranges[..., 1].reshape(1, -1), (E, self.D))
        return oenv.bound_nature(a_nat, lo, hi).astype(F32)

    def _env_nature(self, nat_b):
        E, D = nat_b.shape
        return nat_b.T.copy() if D == self.N else nat_b.T.reshape(self.N, D // self.N, E).copy()

    def _env_step(self, s, a, nat_env, u):
        if self.domain == oenv.DOMAIN_CE:
            return oenv.ce_step(s, a, nat_env, self.ranges, self.R, u)
        if self.domain == oenv.DOMAIN_ARMMAN:
            return oenv.armman_step(s, a, nat_env, self.ranges, self.R, u, nature_is_table=False)
        return oenv.sis_step(self.pop, s, a, nat_env, self.R, u)

    def epoch(self, act_test):
        """act_test(s [N,E]) -> actions [N,E] of the sampled agent strategy."""
        N, E, T, S, D, h = self.N, self.E, self.T, self.S, self.D, self.h
        A, oh, sn = self.A, self.one_hot, self.state_norm
        ep = self.epoch_idx
        s0 = self.s.copy()
        with torch.no_grad():
            lamb = self._lambda(s0).numpy().astype(F32)
        Wpi, Wv, W1n = self.opt.Wpi.detach().numpy(), self.opt.Wv.detach().numpy(), self.W1n.detach().numpy()
        s_traj = np.zeros((N, T + 1, E), np.uint8); s_traj[:, 0] = s0
        a = np.zeros((N, T, E), np.uint8)
        logp = np.zeros((N, T, E), F32); val = np.zeros((N, T, E), F32)
        a_nat = np.zeros((T, E, D), F32); nat_b = np.zeros((T, E, D), F32)
        logp_nat = np.zeros((T, E), F32); val_nat = np.zeros((T, E), F32); rew_nat = np.zeros((T, E), F32)

        def step(s, t, sample=True):
            with torch.no_grad():
                # ma_rmabppo_core.py:276-289: obs / state_norm (non-one-hot only), then / nature_state_norm for the actor
                obs_ac = torch.as_tensor(s.T.astype(F32)) / F32(1.0 if oh else sn)
                mu = self.mu_net(obs_ac / F32(self.nsn)).numpy()
                eps = oma.normal_eps(self.seed, t, E, D)
                an, lpn = oma.gaussian_sample(mu, self.log_std.detach().numpy(), eps)
                nb = self._bound(an, s)
                ua = philox.uniform_grid(self.seed, philox.STREAM_ACT, t, E, N)
                aa, _, lp, p1, _ = nets.ac_step(Wpi, Wv, s, lamb, ua, S, A, h, oh, sn)
                v = oma.agent_critic_extra(Wv, W1n, s, lamb, nb, S, h, oh, sn)
                x = torch.cat([obs_ac, torch.as_tensor(aa.T.astype(F32))], dim=1)              # :322-323
                vn = self.v_nature(x)[:, 0].numpy()
            return aa, v, lp, an, nb, lpn, vn

        for t in range(T):
            tg = self.t_global + t
            s = s_traj[:, t]
            aa, v, lp, an, nb, lpn, vn = step(s, tg)
            nat_env = self._env_nature(nb)
            ue = philox.uniform_grid(self.seed, philox.STREAM_ENV, tg, E, N)
            s_next, r = self._env_step(s, aa, nat_env, ue)
            a_t = act_test(s)
            r_test = oma.counterfactual_mean(self.domain, s, a_t, nat_env, self.ranges, self.R, self.seed, tg,
                                             NUM_TEST_POLICY_RUNS, pop=self.pop) if self.domain != oenv.DOMAIN_ARMMAN else \
                _cf_armman(s, a_t, nat_env, self.ranges, self.R, self.seed, tg)
            rew_nat[t] = r.sum(0) - r_test.sum(0) - lamb * self.C[aa].sum(0)
            a[:, t], logp[:, t], val[:, t], s_traj[:, t + 1] = aa, lp, v, s_next
            a_nat[t], nat_b[t], logp_nat[t], val_nat[t] = an, nb, lpn, vn
        _, lv, _, _, _, _, lvn = step(s_traj[:, T], self.t_global + T)
        # ---- finish_path + get ----
        rew = np.stack([np.take_along_axis(self.R, s_traj[:, t + 1].astype(np.int64), axis=1) for t in range(T)], 1)
        adv, ret, _, cdcost = obuf.finish_path(rew, self.C[a], val, lv, lamb, self.gamma, self.lam_other)
        adv = obuf.normalize_adv(adv)
        adv_n, ret_n, _ = obuf.finish_path_stream(rew_nat, val_nat, lvn, self.gamma, self.lam_other)
        m, sd = obuf.statistics_scalar(adv_n)
        adv_n = ((adv_n - m) / sd).astype(F32)
        # ---- agent pi ----
        ent = oppo.entropy_coeff_for(ep, self.epochs, self.ent0, self.ent1, self.freq, self.final)
        oppo.update(self.opt, s_traj[:, :T], a, adv, logp, ret, lamb, S, A, h, self.clip, ent, self.pi_iters, 0,
                    one_hot=oh, state_norm=sn)
        # ---- agent critics over [x_i, bounded nature vector] ----
        x = np.stack([nets.features(s_traj[:, t], lamb, S, oh, sn) for t in range(T)], axis=1).reshape(N, T * E, self.in_dim)
        xt, rt = torch.as_tensor(x), torch.as_tensor(ret.reshape(N, T * E))
        natM = torch.as_tensor(nat_b.reshape(T * E, D))
        for _ in range(self.v_iters):
            self.opt_v.zero_grad()
            W1, b1, W2, b2, W3, b3 = oppo._unpack_t(self.opt.Wv, self.in_dim, h, 1)
            z1 = torch.bmm(xt, W1.transpose(1, 2)) + b1[:, None, :] + torch.matmul(natM, self.W1n.transpose(1, 2))
            a2 = torch.tanh(torch.bmm(torch.tanh(z1), W2.transpose(1, 2)) + b2[:, None, :])
            v = (torch.bmm(a2, W3.transpose(1, 2)) + b3[:, None, :])[..., 0]
            ((v - rt) ** 2).mean(dim=1).sum().backward()
            self.opt_v.step()
        # ---- lambda + nature nets ----
        if ep % self.freq == 0 and ep > 0:
            if (self.epochs - ep) > self.final:
                for p in self.lam_params:
                    p.grad = None
                coef = torch.as_tensor((self.B / (1.0 - self.gamma) - cdcost[0]).astype(F32))
                (self._lambda(s0, 1.0 if oh else 1.0 / sn) * coef).mean().backward()
                with torch.no_grad():
                    for p in self.lam_params:
                        p -= self.lm_lr * p.grad
            # compute_loss_pi_nature / compute_loss_v_nature divide the RAW buffer observations by nature_state_norm when
            # the nets are not one-hot, and by nothing otherwise (:359-364, :421-428) -- unlike the rollout's step()
            obs = torch.as_tensor(np.ascontiguousarray(s_traj[:, :T].transpose(1, 2, 0)).reshape(T * E, N).astype(F32))
            if not oh:
                obs = obs / F32(self.nsn)
            act = torch.as_tensor(a_nat.reshape(T * E, D))
            advn, lpo, retn = torch.as_tensor(adv_n.reshape(-1)), torch.as_tensor(logp_nat.reshape(-1)), torch.as_tensor(ret_n.reshape(-1))
            for _ in range(self.pi_iters):
                self.opt_pi_nat.zero_grad()
                pi = torch.distributions.Normal(self.mu_net(obs), torch.exp(self.log_std))
                ratio = torch.exp(pi.log_prob(act).sum(axis=-1) - lpo)
                (-(torch.min(ratio * advn, torch.clamp(ratio, 1 - self.clip, 1 + self.clip) * advn)).mean()).backward()
                self.opt_pi_nat.step()
            xa = torch.cat([obs, torch.as_tensor(np.ascontiguousarray(a.transpose(1, 2, 0)).reshape(T * E, N).astype(F32))], dim=1)
            for _ in range(self.v_iters):
                self.opt_v_nat.zero_grad()
                ((self.v_nature(xa)[:, 0] - retn) ** 2).mean().backward()
                self.opt_v_nat.step()
        self.epoch_idx += 1
        self.t_global += T
        self.s = oenv.reset_states(self.seed, self.epoch_idx, S, E, N)
        return dict(lamb=lamb, s_traj=s_traj, a=a, rew_nat=rew_nat, val=val, val_nat=val_nat, a_nat=a_nat, adv_n=adv_n,
                    ret_agent=rew.sum(axis=(0, 1)))


def _cf_armman(s, a, nat_env, ranges, R, seed, t):
    N, E = s.shape
    acc = np.zeros((N, E), F32)
    for tr in range(NUM_TEST_POLICY_RUNS):
        u = philox.uniform_grid(seed, philox.STREAM_CF, t * NUM_TEST_POLICY_RUNS + tr, E, N)
        _, r = oenv.armman_step(s, a, nat_env, ranges, R, u, nature_is_table=False)
        acc = (acc + r).astype(F32)
    return (acc / F32(NUM_TEST_POLICY_RUNS)).astype(F32)
