"""RMABPPO epoch loop restatement on the CPU (oracle, test-only): E env replicas x N arms.

Follows robust_rmab/algos/rmabppo/agent_oracle.py best_response_per_cpu:
  :497      lambda = lambda_net(o) once per epoch (no grad)
  :506-537  for t: ac.step -> nature action -> env.step -> buf.store
  :553-563  bootstrap ac.step on the last observation, buf.finish_path(v)
  :568      env.reset()
  :399-454  update(): buf.get() (per-arm adv normalisation), pi_iters + v_iters Adam steps per arm,
            lambda-net SGD step when `lambda_update_due`
Randomness follows the Philox convention of oracle/philox.py (stream ACT / ENV with the global step
counter as t, stream RESET with the epoch index), the same one the CUDA path uses, so the two can be
compared state for state.  The E replicas are what the reference runs as MPI ranks (one env each); the
lambda loss is averaged over replicas as mpi_avg_grads averages gradients over ranks.
"""
import numpy as np
import torch

from . import buffer as obuf
from . import envs as oenv
from . import nets, philox
from . import ppo as oppo

F32 = np.float32


class CpuRMABPPO:
    def __init__(self, domain, Wpi, Wv, lam_params, nature, ranges, R, C, n_envs, steps_per_epoch, hidden, B,
                 gamma=0.9, lam_other=0.97, clip_ratio=2.0, pi_lr=2e-3, vf_lr=2e-3, lm_lr=2e-3, train_pi_iters=20,
                 train_v_iters=20, epochs=20, lamb_update_freq=4, final_train_lambdas=2, start_entropy_coeff=0.5,
                 end_entropy_coeff=0.0, seed=0, arm0=0, pop=None, one_hot=True, state_norm=1.0):
        self.domain, self.pop = domain, pop
        self.S = {oenv.DOMAIN_CE: 2, oenv.DOMAIN_ARMMAN: 3}.get(domain, (pop or 0) + 1)
        self.A = len(C)
        self.one_hot, self.state_norm = bool(one_hot), float(state_norm if state_norm else 1.0)
        self.in_dim = (self.S if self.one_hot else 1) + 1
        self.N, self.E, self.T, self.h = Wpi.shape[0], n_envs, steps_per_epoch, hidden
        self.opt = oppo.ArmOptim(Wpi, Wv, pi_lr, vf_lr)
        self.lam_params = [torch.nn.Parameter(torch.as_tensor(np.array(p, F32))) for p in lam_params]
        self.nature, self.ranges, self.R, self.C = nature, ranges, np.asarray(R, F32), np.asarray(C, F32)
        self.B, self.gamma, self.lam_other, self.clip = B, gamma, lam_other, clip_ratio
        self.lm_lr, self.pi_iters, self.v_iters, self.epochs = lm_lr, train_pi_iters, train_v_iters, epochs
        self.freq, self.final = lamb_update_freq, final_train_lambdas
        self.ent0, self.ent1 = start_entropy_coeff, end_entropy_coeff
        self.seed, self.arm0 = seed, arm0
        self.epoch_idx, self.t_global = 0, 0
        self.s = oenv.reset_states(seed, 0, self.S, self.E, self.N, arm0)

    def _lambda(self, s, scale=1.0):
        """lambda_net on the observation times `scale`: 1 for the rollout's prediction (raw obs, agent_oracle.py:497),
        1/state_norm inside compute_loss_lambda when the nets are not one-hot (:364-367)."""
        x = torch.as_tensor(s.T.astype(F32)) * F32(scale)
        p = self.lam_params
        return (torch.tanh(torch.tanh(x @ p[0].T + p[1]) @ p[2].T + p[3]) @ p[4].T + p[5])[:, 0]

    def epoch(self):
        N, E, T, S = self.N, self.E, self.T, self.S
        ep = self.epoch_idx
        s0 = self.s.copy()
        with torch.no_grad():
            lamb = self._lambda(s0).numpy().astype(F32)
        Wpi, Wv = self.opt.Wpi.detach().numpy(), self.opt.Wv.detach().numpy()
        s_traj = np.zeros((N, T + 1, E), np.uint8)
        a = np.zeros((N, T, E), np.uint8)
        logp = np.zeros((N, T, E), F32)
        val = np.zeros((N, T, E), F32)
        s_traj[:, 0] = s0
        if self.domain == oenv.DOMAIN_SIS:
            step = lambda s_, a_, nat, rg, R_, u_: oenv.sis_step(self.pop, s_, a_, nat, R_, u_)
        else:
            step = oenv.ce_step if self.domain == oenv.DOMAIN_CE else oenv.armman_step
        for t in range(T):
            ua = philox.uniform_grid(self.seed, philox.STREAM_ACT, self.t_global + t, E, N, self.arm0)
            ue = philox.uniform_grid(self.seed, philox.STREAM_ENV, self.t_global + t, E, N, self.arm0)
            a[:, t], val[:, t], logp[:, t], _, _ = nets.ac_step(Wpi, Wv, s_traj[:, t], lamb, ua, S, self.A, self.h,
                                                                self.one_hot, self.state_norm)
            s_traj[:, t + 1], _ = step(s_traj[:, t], a[:, t], self.nature, self.ranges, self.R, ue)
        x = nets.features(s_traj[:, T], lamb, S, self.one_hot, self.state_norm)
        last_val = nets.mlp_forward(Wv, x, self.in_dim, self.h, 1)[..., 0].astype(F32)
        rew = np.stack([np.take_along_axis(self.R, s_traj[:, t + 1].astype(np.int64), axis=1) for t in range(T)], 1)
        cost = self.C[a]
        adv, ret, _, cdcost = obuf.finish_path(rew, cost, val, last_val, lamb, self.gamma, self.lam_other)
        adv_n = obuf.normalize_adv(adv)
        ent = oppo.entropy_coeff_for(ep, self.epochs, self.ent0, self.ent1, self.freq, self.final)
        lp, lv = oppo.update(self.opt, s_traj[:, :T], a, adv_n, logp, ret, lamb, S, self.A, self.h, self.clip, ent,
                             self.pi_iters, self.v_iters, one_hot=self.one_hot, state_norm=self.state_norm)
        if oppo.lambda_update_due(ep, self.epochs, self.freq, self.final):
            for p in self.lam_params:
                p.grad = None
            coef = torch.as_tensor((self.B / (1.0 - self.gamma) - cdcost[0]).astype(F32))
            scale = 1.0 if self.one_hot else 1.0 / self.state_norm
            (self._lambda(s0, scale) * coef).mean().backward()    # compute_loss_lambda (:360-376), mean over replicas
            with torch.no_grad():
                for p in self.lam_params:
                    p -= self.lm_lr * p.grad
        self.epoch_idx += 1
        self.t_global += T
        self.s = oenv.reset_states(self.seed, self.epoch_idx, S, E, N, self.arm0)
        return dict(lamb=lamb, s_traj=s_traj, a=a, logp=logp, val=val, last_val=last_val, adv=adv_n, ret=ret,
                    cdcost=cdcost, loss_pi=lp, loss_v=lv, rew_sum=rew.sum(axis=(0, 1)))


def simulate_reward_cpu(domain, Wpi, Wv, lam_params, nature, ranges, R, C, B, hidden, n_rollouts, steps, gamma, seed):
    """AgentOracle.simulate_reward (agent_oracle.py:599-652) with the `epochs` rollouts as env replicas and the
    Philox draws of the device path: reset stream t = 0, transition stream t = step.  Binary-action table domains.
    Returns the per-rollout discounted returns `[n_rollouts]` (float64)."""
    from . import select as osel
    S = 2 if domain == oenv.DOMAIN_CE else 3
    N, E = Wpi.shape[0], n_rollouts
    s = oenv.reset_states(seed, 0, S, E, N)
    step = oenv.ce_step if domain == oenv.DOMAIN_CE else oenv.armman_step
    ret = np.zeros(E)
    for t in range(steps):
        lamb = nets.lambda_net_forward(lam_params, s.T.astype(F32))
        x = nets.features(s, lamb, S, True)
        probs, _ = nets.softmax_logits(nets.mlp_forward(Wpi, x, S + 1, hidden, 2))
        a = np.stack([osel.topk_binary(probs[:, e, 1], C, B) for e in range(E)], axis=1).astype(np.uint8)
        u = philox.uniform_grid(seed, philox.STREAM_ENV, t, E, N)
        s, r = step(s, a, nature, ranges, R, u)
        ret += (gamma ** t) * r.astype(np.float64).sum(axis=0)
    return ret
