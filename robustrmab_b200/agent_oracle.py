"""Drop-in `AgentOracle` (robust_rmab/algos/rmabppo/agent_oracle.py:164-652): same constructor, `best_response`,
`best_response_per_cpu` keyword arguments and `simulate_reward`, with the training loop on the device.

Two device loops implement `best_response_per_cpu`'s `for epoch` body (:492-593):
  * fused    -- table domains (counterexample / ARMMAN) against baseline nature policies (objects with a
                `param_setting` table): `trainer.RMABPPOTrainer`, one launch for rollout+GAE+statistics, one per net
                for all Adam iterations.
  * stepwise -- every other case (SIS, learned nature policies with `get_nature_action_device`): one `ac.step`,
                nature action, `env.step` launch per time step into `[N,T,E]` buffers, then `rmab_gae`,
                `rmab_ppo_update` (per-sample rows), `rmab_cost_togo`, `rmab_lambda_sgd`.
Both follow the reference's schedule: lambda once per epoch (:497), nature policy resampled from the mixed strategy
every `lamb_update_freq` epochs (:501-503), bootstrap value + finish_path at the epoch end (:553-563), env.reset
(:568), update with the entropy schedule (:402-409, :578) and the lambda step when due (:440-454).
The returned object is a `core.MLPActorCriticRMAB` (packed weights) usable as an agent strategy (`act_test`).
"""
import numpy as np
import torch

from . import _lib, ops
from .baselines import nature_param_table
from .core import MLPActorCriticRMAB
from .environments import ARMMANRobustEnv, CounterExampleRobustEnv, SISRobustEnv
from .evaluate import simulate_reward as _simulate_reward
from .trainer import RMABPPOTrainer

_F32 = torch.float32


def lambda_pretrain(lam_params, s0, n_total, arm0, B, gamma, lr, n, allreduce=None):
    """`init_lambda_trains` SGD steps (agent_oracle.py:465-477, nature_oracle.py:605-617) on
    loss = lambda_net(o) * (B/(1-gamma) - 2B/(1-gamma)), o = the first observation, unnormalised."""
    E = s0.shape[1]
    coef = torch.full((E,), -float(B) / (1.0 - float(gamma)), dtype=_F32, device=s0.device)
    for _ in range(int(n)):
        _, h1 = ops.lambda_forward(lam_params, s0, n_total, arm0, 1.0, want_h1=True, want_lamb=False)
        if allreduce is not None:
            allreduce(h1)
        ops.lambda_sgd(lam_params, s0, n_total, arm0, 1.0, h1, coef, lr)


def check_episode_length(max_ep_len, steps_per_epoch):
    """The reference cuts a path (finish_path + env.reset) whenever ep_len == max_ep_len (:541-568).  The device loops run
    one path per epoch, which is what every shipped config does (max_ep_len defaults to 1000 > steps_per_epoch)."""
    if int(max_ep_len) < int(steps_per_epoch):
        raise NotImplementedError(f"max_ep_len={max_ep_len} < steps_per_epoch={steps_per_epoch}: mid-epoch path cuts are "
                                  "not implemented on the device loops")


class StepwiseRMABPPO:
    """The general device loop (any domain / nature policy): per-step launches, per-sample update rows."""

    def __init__(self, env, ac, T, gamma, lam_other, clip_ratio, pi_lr, vf_lr, lm_lr, pi_iters, v_iters, epochs,
                 lamb_update_freq, final_train_lambdas, ent0, ent1, seed):
        self.env, self.ac, self.T = env, ac, int(T)
        self.gamma, self.lam_other, self.clip = gamma, lam_other, clip_ratio
        self.pi_lr, self.vf_lr, self.lm_lr = pi_lr, vf_lr, lm_lr
        self.pi_iters, self.v_iters, self.epochs = int(pi_iters), int(v_iters), int(epochs)
        self.freq, self.final, self.ent0, self.ent1 = int(lamb_update_freq), int(final_train_lambdas), ent0, ent1
        N, E, dev = env.N, env.n_envs, env.device
        self.N, self.E, self.dev = N, E, dev
        self.s_traj = torch.empty((N, self.T + 1, E), dtype=torch.uint8, device=dev)
        self.a = torch.empty((N, self.T, E), dtype=torch.uint8, device=dev)
        self.logp = torch.empty((N, self.T, E), dtype=_F32, device=dev)
        self.val = torch.empty((N, self.T, E), dtype=_F32, device=dev)
        self.adam_pi = torch.zeros((N, 2, ac.Ppi), dtype=_F32, device=dev)
        self.adam_v = torch.zeros((N, 2, ac.Pv), dtype=_F32, device=dev)
        self.steps_pi = self.steps_v = 0
        self.epoch_idx = 0
        self.R = torch.as_tensor(np.asarray(env.R), dtype=_F32, device=dev).contiguous()
        self.C = torch.as_tensor(np.asarray(env.C), dtype=_F32, device=dev).contiguous()
        self.lam_params = ac.lambda_net.packed().to(dev)
        self.obs_scale = 1.0 if ac.one_hot_encode else 1.0 / float(ac.state_norm)
        ac.seed(seed)
        env.seed(seed)
        self.s = env.reset_device()

    def init_lambda_trains(self, n):
        """agent_oracle.py:465-477: n SGD steps on return_large_lambda_loss = lambda(o) (B/(1-gamma) - 2B/(1-gamma)) at the
        first observation (raw, rmabppo_core.py:188-195)."""
        lambda_pretrain(self.lam_params, self.s.contiguous(), self.N, 0, float(self.ac.B), self.gamma, self.lm_lr, n)

    def _entropy(self, ep):
        head = float(np.linspace(self.ent0, self.ent1, self.epochs)[min(ep, self.epochs - 1)])
        if (self.epochs - ep) > self.final:
            return float(np.linspace(head, 0, self.freq)[ep % self.freq])
        return 0.0

    def _nature(self, nature_pol, s):
        if hasattr(nature_pol, "get_nature_action_device"):
            raw = nature_pol.get_nature_action_device(s)                       # [E,D] raw, bounded by the policy's env
            return nature_pol.bound_nature_device(raw, s), True
        tab = nature_param_table(nature_pol, self.env)
        if tab is None:
            raise NotImplementedError(f"{nature_pol!r}: neither a device policy (get_nature_action_device) nor a "
                                      "deterministic per-arm parameter table")
        return torch.as_tensor(tab, device=self.dev), False

    def epoch(self, nature_pol):
        env, ac, T, N, E = self.env, self.ac, self.T, self.N, self.E
        ep = self.epoch_idx
        s0 = self.s.contiguous()
        # the rollout's lambda is predicted from the RAW observation (agent_oracle.py:497), while compute_loss_lambda
        # (:360-376) and get_lambda (rmabppo_core.py:281-287) divide by state_norm first -- kept as the reference has it
        lamb, h1 = ops.lambda_forward(self.lam_params, s0, N, 0, 1.0, want_h1=True)
        self.s_traj[:, 0] = s0
        ret = torch.zeros((E,), dtype=_F32, device=self.dev)
        for t in range(T):
            s = self.s_traj[:, t].contiguous()
            a, v, logp, _, _ = ac.step_device(s, lamb)
            nat, per_env = self._nature(nature_pol, s)
            if per_env:                                                     # [E,D] -> the env's per-env layout [N,...,E]
                nat = nat.t().reshape((N, -1, E)).contiguous() if nat.shape[1] != N else nat.t().contiguous()
            env._state = s
            s_next, r = env.step_device(a, nat)
            env._t += 1
            self.a[:, t], self.logp[:, t], self.val[:, t], self.s_traj[:, t + 1] = a, logp, v, s_next
            ret += r.sum(dim=0)
        s_last = self.s_traj[:, T].contiguous()
        _, last_val, _, _, _ = ac.step_device(s_last, lamb)                 # bootstrap value (:553)
        ac._steps -= 1                                                      # its action draw is discarded
        adv, rtg, _, _ = ops.gae(self.s_traj, self.a, self.val, last_val, lamb, self.C, self.R, self.gamma, self.lam_other,
                                 normalize=True)
        hp = ops.hyper(self.clip, self._entropy(ep), self.pi_lr, self.vf_lr, self.pi_iters, self.v_iters, self.steps_pi,
                       self.steps_v)
        ops.ppo_update(ac._net(E), hp, ac.Wpi, ac.Wv, self.adam_pi, self.adam_v, self.s_traj, self.a, adv, self.logp, rtg,
                       lamb)
        self.steps_pi += self.pi_iters
        self.steps_v += self.v_iters
        if ep % self.freq == 0 and ep > 0 and (self.epochs - ep) > self.final:
            cd = ops.cost_togo(self.a, self.C, self.gamma)
            coef = (float(ac.B) / (1.0 - self.gamma) - cd[0]).contiguous()
            if self.obs_scale != 1.0:
                _, h1 = ops.lambda_forward(self.lam_params, s0, N, 0, self.obs_scale, want_h1=True, want_lamb=False)
            ops.lambda_sgd(self.lam_params, s0, N, 0, self.obs_scale, h1, coef, self.lm_lr)
        self.epoch_idx += 1
        self.s = env.reset_device()
        return {"EpActualRet": ret, "Lamb": lamb, "VVals": self.val.mean().reshape(1)}

    def export_lambda(self):
        """Write the trained lambda-net parameters back into ac.lambda_net (a torch module on the host)."""
        flat = self.lam_params.detach().cpu()
        o = 0
        with torch.no_grad():
            for p in self.ac.lambda_net.parameters():
                p.copy_(flat[o:o + p.numel()].reshape(p.shape))
                o += p.numel()


class AgentOracle:

    def __init__(self, data, N, S, A, B, seed, REWARD_BOUND, agent_kwargs=dict(), home_dir="", exp_name="",
                 sampled_nature_parameter_ranges=None, robust_keyword="", pop_size=0, one_hot_encode=True,
                 non_ohe_obs_dim=None, state_norm=None, n_envs=1, device="cuda"):
        self.data, self.home_dir, self.exp_name, self.REWARD_BOUND = data, home_dir, exp_name, REWARD_BOUND
        self.N, self.S, self.A, self.B, self.seed = N, S, A, B, seed
        self.sampled_nature_parameter_ranges = sampled_nature_parameter_ranges
        self.robust_keyword, self.pop_size = robust_keyword, pop_size
        self.one_hot_encode, self.non_ohe_obs_dim, self.state_norm = one_hot_encode, non_ohe_obs_dim, state_norm
        self.n_envs, self.device = int(n_envs), device
        if data == 'armman':
            self.env_fn = lambda n_envs=self.n_envs: ARMMANRobustEnv(N, B, seed, n_envs=n_envs, device=device)
        elif data == 'counterexample':
            self.env_fn = lambda n_envs=self.n_envs: CounterExampleRobustEnv(N, B, seed, n_envs=n_envs, device=device)
        elif data == 'sis':
            self.env_fn = lambda n_envs=self.n_envs: SISRobustEnv(N, B, pop_size, seed, n_envs=n_envs, device=device)
        else:
            raise ValueError("data must be 'counterexample', 'armman' or 'sis' (the legacy non-robust envs of "
                             "bandit_env.py are outside the accelerated path)")
        self.actor_critic = MLPActorCriticRMAB
        self.agent_kwargs = agent_kwargs
        self.strat_ind = 0
        self.env = self.env_fn()
        self.env.seed(seed)
        self.env.sampled_parameter_ranges = self.sampled_nature_parameter_ranges
        self.history = []

    def best_response(self, nature_strats, nature_eq, add_to_seed):
        self.strat_ind += 1
        return self.best_response_per_cpu(nature_strats, nature_eq, add_to_seed, seed=self.seed, **self.agent_kwargs)

    def best_response_per_cpu(self, nature_strats, nature_eq, add_to_seed, actor_critic=None, ac_kwargs=dict(), seed=0,
                              steps_per_epoch=4000, epochs=50, gamma=0.99, clip_ratio=0.2, pi_lr=3e-4, vf_lr=1e-3,
                              qf_lr=1e-3, lm_lr=5e-2, train_pi_iters=80, train_v_iters=80, train_q_iters=80,
                              lam_OTHER=0.97, start_entropy_coeff=0.0, end_entropy_coeff=0.0, max_ep_len=1000,
                              target_kl=0.01, logger_kwargs=dict(), save_freq=10, lamb_update_freq=10,
                              init_lambda_trains=0, final_train_lambdas=0):
        env = self.env
        check_episode_length(max_ep_len, steps_per_epoch)
        hidden = list(ac_kwargs.get("hidden_sizes", (64, 64)))
        ac = MLPActorCriticRMAB(env.observation_space, env.action_space, hidden_sizes=hidden, N=env.N, C=env.C, B=env.B,
                                strat_ind=self.strat_ind, one_hot_encode=self.one_hot_encode,
                                non_ohe_obs_dim=self.non_ohe_obs_dim, state_norm=self.state_norm, n_envs=self.n_envs,
                                device=self.device)
        nature_eq = np.array(nature_eq, dtype=np.float64)
        nature_eq[nature_eq < 0] = 0
        nature_eq = nature_eq / nature_eq.sum()
        nature_pol = np.random.choice(nature_strats, p=nature_eq)                      # :489
        tables = {id(p): nature_param_table(p, env) for p in nature_strats}
        fused = self.data in ("counterexample", "armman") and all(t is not None for t in tables.values()) \
            and self.one_hot_encode
        hist = []
        if fused:
            tr = RMABPPOTrainer(self.data, env.N, self.n_envs, steps_per_epoch, hidden=hidden[0], B=env.B, gamma=gamma,
                                lam_other=lam_OTHER, clip_ratio=clip_ratio, pi_lr=pi_lr, vf_lr=vf_lr, lm_lr=lm_lr,
                                train_pi_iters=train_pi_iters, train_v_iters=train_v_iters, epochs=epochs,
                                lamb_update_freq=lamb_update_freq, final_train_lambdas=final_train_lambdas,
                                start_entropy_coeff=start_entropy_coeff, end_entropy_coeff=end_entropy_coeff, seed=seed,
                                ranges=np.asarray(env.sampled_parameter_ranges), nature=tables[id(nature_pol)],
                                device=self.device)
            tr.Wpi.copy_(ac.Wpi)                                # the nets the reference would have built (torch RNG)
            tr.Wv.copy_(ac.Wv)
            tr.lam_params.copy_(ac.lambda_net.packed().to(tr.dev))
            lambda_pretrain(tr.lam_params, tr._s0(), tr.N_total, tr.arm0, tr.B, gamma, lm_lr, init_lambda_trains)
            for epoch in range(epochs):
                if epoch % lamb_update_freq == 0 and epoch > 0:                         # :501-503
                    nature_pol = np.random.choice(nature_strats, p=nature_eq)
                    tr.nature.copy_(torch.as_tensor(tables[id(nature_pol)]))
                hist.append({k: v.detach().cpu().numpy() for k, v in tr.epoch().items()})
            ac.Wpi.copy_(tr.Wpi)
            ac.Wv.copy_(tr.Wv)
            flat, o = tr.lam_params.detach().cpu(), 0
            with torch.no_grad():
                for p in ac.lambda_net.parameters():
                    p.copy_(flat[o:o + p.numel()].reshape(p.shape))
                    o += p.numel()
        else:
            sw = StepwiseRMABPPO(env, ac, steps_per_epoch, gamma, lam_OTHER, clip_ratio, pi_lr, vf_lr, lm_lr, train_pi_iters,
                                 train_v_iters, epochs, lamb_update_freq, final_train_lambdas, start_entropy_coeff,
                                 end_entropy_coeff, seed)
            sw.init_lambda_trains(init_lambda_trains)
            for epoch in range(epochs):
                if epoch % lamb_update_freq == 0 and epoch > 0:
                    nature_pol = np.random.choice(nature_strats, p=nature_eq)
                hist.append({k: v.detach().cpu().numpy() for k, v in sw.epoch(nature_pol).items()})
            sw.export_lambda()
        self.history = hist
        return ac

    def simulate_reward(self, agent_pol, nature_pol, seed=0, steps_per_epoch=100, epochs=100, gamma=0.99):
        """agent_oracle.py:599-652 with the `epochs` rollouts as env replicas of one batched rollout."""
        env = self.env_fn(n_envs=epochs)
        env.sampled_parameter_ranges = self.sampled_nature_parameter_ranges
        tab = nature_param_table(nature_pol, env)
        if tab is not None:
            nature = torch.as_tensor(tab, device=env.device)
        else:
            def nature(s, pol=nature_pol):
                nb = pol.bound_nature_device(pol.get_nature_action_device(s), s)          # [E,D]
                N, E = s.shape
                return nb.t().reshape((N, -1, E)).contiguous() if nb.shape[1] != N else nb.t().contiguous()
        # the caller's strategy object is left untouched: act_test_device takes E from the state tensor
        return _simulate_reward(agent_pol, env, nature, steps_per_epoch=steps_per_epoch, epochs=epochs, gamma=gamma,
                                seed=seed)
