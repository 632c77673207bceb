"""Drop-in `NatureOracle` (robust_rmab/algos/ma_rmabppo/nature_oracle.py:197-806): MA-RMABPPO best response of the
nature player against a mixed agent strategy, with the rollout and the agent-side update on the device.

Per time step (nature_oracle.py:673-735): `ac.step` (Gaussian nature action, per-arm agent actions, agent critics over
the bounded nature vector, nature critic), `env.step`, then the counterfactual reward estimate -- 50 env steps from the
same state with the sampled agent strategy's `act_test` action -- and
`r_nature = sum r_agent - mean_test - lambda * sum cost` (:707-719).
Epoch end: bootstrap `ac.step`, `MA_RMABPPO_Buffer.finish_path` (:107-167: the agent GAE per arm, the nature GAE on
the scalar stream), `get` (:168-193: per-arm and nature advantage normalisation), `update` (:486-576):
  agent policies   train_pi_iters Adam steps (rmab_ppo_update, policy loss only)
  agent critics    train_v_iters Adam steps; every iteration = dense GEMM (nature block of the first layer), one
                   rmab_critic_extra_iter launch (everything else + Adam of the packed block), dense GEMM for the nature
                   block's gradient, rmab_adam_step; the dense products are the package's TMA-fed tcgen05 kernel
                   (rmab_gemm3h, csrc/gemm3h.cu)
  every lamb_update_freq epochs: lambda-net SGD step (rmab_lambda_sgd), nature policy and nature critic
                   train_pi_iters / train_v_iters Adam steps (dense N -> h -> h -> D nets: rmab_nat_* kernels on packed
                   parameters, nature_nets.py; layer 1 and the actor's output rows are sharded by arm).
"""
import os

import numpy as np
import torch

from . import _lib, ops
from .environments import ARMMANRobustEnv, CounterExampleRobustEnv, SISRobustEnv
from .ma_core import RMABLambdaNatureOracle

_F32 = torch.float32
NUM_TEST_POLICY_RUNS = 50      # nature_oracle.py:647


def nature_env_layout(nat_b, N):
    """Bounded nature actions `[E,D]` -> the env kernels' per-env layout `[N,E]` (D = N) or `[N,D/N,E]`."""
    E, D = nat_b.shape
    return nat_b.t().contiguous() if D == N else nat_b.t().reshape(N, D // N, E).contiguous()


class _Stages:
    """CUDA-event stage timers (off unless `MARolloutTrainer.profile` is set): name -> accumulated ms."""

    def __init__(self, on):
        self.on, self.ms, self._open = on, {}, None

    def mark(self, name):
        if not self.on:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if self._open is not None:
            prev_name, prev_ev = self._open
            self.ms.setdefault(prev_name, []).append((prev_ev, ev))
        self._open = (name, ev) if name else None

    def report(self):
        torch.cuda.synchronize()
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in self.ms.items()}


class MARolloutTrainer:
    """Device state and loops of one `best_response_per_cpu` call.

    Arm-sharded over `world_size` ranks when `ac` was built with `arm_shard`: every rank keeps all E envs
    of its arm range -- env transitions, per-arm agent actors / critics (and the nature block W1n of its critics), the
    50 counterfactual steps, GAE and the agent updates are local.  Exchanged over NCCL: the state vector (all-gather of
    `[N,E]` u8, once per step: input of the nature actor's forward, which every rank evaluates identically, and of the
    agent strategy), per-step reward / cost partial sums (`[T,3,E]`, one all-reduce per epoch), the nature critic's
    layer-1 partial sums (`[(T+1)E, h]`, once per epoch), the lambda-net's first-layer partial sums and the discounted
    cost at t = 0 as in the RMABPPO trainer.  The nature nets are TRAINED sharded: per Adam step three small
    all-reduces (`[TE,h]`, `[TE]`, `[TE,h]`; critic: one), and one exchange of the updated slices per update."""
    profile = False
    # A/B switch: RMAB_MA_LIBRARY_GEMM=1 evaluates the nature-block products as three library TF32 GEMMs (round 1)
    library_gemm = os.environ.get("RMAB_MA_LIBRARY_GEMM", "0") == "1"
    # RMAB_MA_GRAPHS=0: launch every rollout step eagerly (default: the per-step sequence is captured as CUDA graphs in the
    # second rollout and replayed afterwards, one graph per time step)
    use_graphs = os.environ.get("RMAB_MA_GRAPHS", "1") == "1"

    def __init__(self, env, ac, T, gamma, lam_other, clip_ratio, pi_lr_agent, pi_lr_nature, vf_lr_agent, vf_lr_nature,
                 lm_lr, pi_iters, v_iters, epochs, lamb_update_freq, final_train_lambdas, ent0, ent1, seed,
                 rank=0, world_size=1, process_group=None):
        self.ac, self.T = ac, int(T)
        self.gamma, self.lam_other, self.clip = float(gamma), float(lam_other), float(clip_ratio)
        self.pi_lr_agent, self.vf_lr_agent, self.lm_lr = pi_lr_agent, vf_lr_agent, lm_lr
        self.pi_iters, self.v_iters, self.epochs = int(pi_iters), int(v_iters), int(epochs)
        self.freq, self.final, self.ent0, self.ent1 = int(lamb_update_freq), int(final_train_lambdas), ent0, ent1
        self.rank, self.world, self.pg = int(rank), int(world_size), process_group
        self.N_total, self.arm0 = ac.N, ac.arm0
        N, E, dev, D = ac.n_local, env.n_envs, env.device, ac.act_dim_nature
        if self.world > 1:
            assert N < ac.N or self.world == 1, "build the actor-critic with arm_shard=mpi_tools.arm_shard(N, rank, world)"
            self.env = env.shard(ac.arm0, ac.arm0 + N)   # the rank's arms; `env` itself keeps the global ranges for ac
            self.per = (self.N_total + self.world - 1) // self.world
        else:
            self.env = env
        env = self.env
        self.N, self.E, self.D, self.dev = N, E, D, dev
        T = self.T
        u8 = torch.uint8
        self.s_traj = torch.empty((N, T + 1, E), dtype=u8, device=dev)
        self.a = torch.empty((N, T, E), dtype=u8, device=dev)
        self.logp = torch.empty((N, T, E), dtype=_F32, device=dev)
        self.val = torch.empty((N, T, E), dtype=_F32, device=dev)
        self.a_nat = torch.empty((T, E, D), dtype=_F32, device=dev)       # raw nature actions (buffer act_nature)
        self.nat_b = torch.empty((T, E, D), dtype=_F32, device=dev)       # bounded (critic inputs, :398-399)
        self.logp_nat = torch.empty((T, E), dtype=_F32, device=dev)
        self.val_nat = torch.empty((T, E), dtype=_F32, device=dev)
        self.rew_nat = torch.empty((T, E), dtype=_F32, device=dev)
        self.part = torch.zeros((T, 3, E), dtype=_F32, device=dev)        # per step: sum r, sum counterfactual r, sum cost
        self.cf = torch.empty((T, N, E), dtype=_F32, device=dev)          # counterfactual reward estimates per step
        self.adam_pi = torch.zeros((N, 2, ac.Wpi.shape[1]), dtype=_F32, device=dev)
        self.adam_v = torch.zeros((N, 2, ac.Wv.shape[1]), dtype=_F32, device=dev)
        self.m_w1n, self.v_w1n = torch.zeros_like(ac.W1n), torch.zeros_like(ac.W1n)
        self.steps_pi = self.steps_v = 0
        self.pi_lr_nature, self.vf_lr_nature = float(pi_lr_nature), float(vf_lr_nature)
        self._gpad = self._gout = None
        self._pol_shard = (None, None)
        # static inputs of the (graph-captured) rollout step
        self.lamb_buf = torch.empty((E,), dtype=_F32, device=dev)
        self.a_boot = torch.empty((N, E), dtype=torch.uint8, device=dev)
        self.v_boot = torch.empty((N, E), dtype=_F32, device=dev)
        self.t_words = torch.zeros((2, 1), dtype=torch.int32, device=dev)
        self._w1n_split = None
        self.graphs, self.graph_pol, self.graph_pool, self.graph_ok, self.rollouts_done = None, None, None, None, 0
        self.nets = ac.nature_nets()       # packed nature actor / critic + their Adam state (nature_nets.py)
        self.R = torch.as_tensor(np.asarray(env.R), dtype=_F32, device=dev).contiguous()
        self.C = torch.as_tensor(np.asarray(env.C), dtype=_F32, device=dev).contiguous()
        self.lam_params = ac.lambda_net.packed().to(dev)
        self.obs_scale = 1.0 if ac.one_hot_encode else 1.0 / float(ac.state_norm)
        self.epoch_idx = 0
        ac.seed(seed)
        env.seed(seed)
        self.seed = int(seed)
        self.s = env.reset_device()

    def _entropy(self, ep):
        head = float(np.linspace(self.ent0, self.ent1, self.epochs)[min(ep, self.epochs - 1)])
        if (self.epochs - ep) > self.final:
            return float(np.linspace(head, 0, self.freq)[ep % self.freq])
        return 0.0

    # -- cross-rank plumbing (no-ops on one rank) --------------------------------------------------
    def _allreduce(self, t):
        if self.world > 1:
            torch.distributed.all_reduce(t, group=self.pg)
        return t

    def _gather_arms(self, x):
        """Local `[N_local, E]` u8 rows of every rank -> `[N_total, E]` in global arm order (ragged last shard padded).
        The buffers are allocated once: a fresh tensor per step would keep the caching allocator busy behind NCCL."""
        if self.world == 1:
            return x
        if self._gpad is None or self._gpad.shape[1:] != x.shape[1:]:
            tail = tuple(x.shape[1:])
            self._gpad = torch.zeros((self.per,) + tail, dtype=x.dtype, device=x.device)
            self._gout = torch.empty((self.world * self.per,) + tail, dtype=x.dtype, device=x.device)
        self._gpad[:x.shape[0]] = x
        torch.distributed.all_gather_into_tensor(self._gout, self._gpad, group=self.pg)
        return self._gout[:self.N_total]

    def _lambda(self, s0, scale=1.0):
        """lambda [E] and the (all-reduced) first-layer sums.  The rollout predicts lambda from the RAW observation
        (nature_oracle.py:664); compute_loss_lambda (:447-459) divides by state_norm first -- `scale` says which."""
        if self.world == 1:
            return ops.lambda_forward(self.lam_params, s0, self.N_total, 0, scale, want_h1=True)
        _, h1 = ops.lambda_forward(self.lam_params, s0, self.N_total, self.arm0, scale, want_h1=True, want_lamb=False)
        self._allreduce(h1)
        lamb, _ = ops.lambda_forward(self.lam_params, s0, self.N_total, self.arm0, scale, h1_in=h1)
        return lamb, h1

    def init_lambda_trains(self, n):
        """nature_oracle.py:605-617."""
        from .agent_oracle import lambda_pretrain
        lambda_pretrain(self.lam_params, self.s.contiguous(), self.N_total, self.arm0, float(self.ac.B), self.gamma,
                        self.lm_lr, n, allreduce=self._allreduce if self.world > 1 else None)

    # ---------------------------------------------------------------------------------------------
    # One time step of the rollout (nature_oracle.py:673-735) on static buffers: everything it reads that changes between
    # epochs (lambda, the W1n split, the Philox counters) lives in preallocated device memory, so that the step can be
    # captured once as a CUDA graph and replayed (the host side of ~35 small launches per step is what bounds the sharded
    # rollout).  Sharded: `s_full` is the static all-gather buffer, filled eagerly before the step.
    def _step(self, t, agent_pol):
        env, ac, T, N = self.env, self.ac, self.T, self.N
        sharded = self.world > 1
        sl = slice(self.arm0, self.arm0 + N)
        s = self.s_traj[:, t].contiguous()
        s_full = self._gout[:self.N_total] if sharded else None
        d = ac.step_device(s, self.lamb_buf, w1n_split=self._w1n_split, s_full=s_full, nat_nets=self.nets, want_v_nature=False)
        ac._steps -= 1                                                          # _advance() owns the host counters
        if t == T:                                                              # the bootstrap ac.step (:754)
            self.a_boot.copy_(d["a_agent"])
            self.v_boot.copy_(d["v_agent"])
            return
        self.a[:, t] = d["a_agent"]
        nat_env = nature_env_layout(d["a_nature_bounded"], self.N_total)
        if sharded:
            nat_env = nat_env[sl].contiguous()
        env._state = s
        s_next, _ = env.step_device(d["a_agent"], nat_env, want_reward=False)   # rewards are R[arm, s'] (summed after the loop)
        if self._sharded_policy(agent_pol):
            # a learned strategy evaluates only this rank's arms and takes its share of the global top-k
            a_test = self._agent_shard(agent_pol).act_test_device_sharded(s, self.rank, self.world, self.pg)
        else:
            a_test = agent_pol.act_test_device(s_full if sharded else s)
            if sharded:
                a_test = a_test[sl].contiguous()
        word = getattr(env, "_t_word", None)
        rng = ops.rng(self.seed, _lib.STREAM_CF, env._t - (env._t_origin if word is not None else 0), arm0=self.arm0, t_base=word)
        ops.env_counterfactual(env.domain, s, a_test, nat_env, env._ranges() if env.domain != _lib.DOMAIN_SIS else self.R,
                               self.R, rng, NUM_TEST_POLICY_RUNS, nature_per_env=True, out=self.cf[t])
        self.logp[:, t], self.val[:, t], self.s_traj[:, t + 1] = d["logp_agent"], d["v_agent"], s_next
        self.a_nat[t], self.nat_b[t] = d["a_nature"], d["a_nature_bounded"]
        self.logp_nat[t] = d["logp_nature"]

    def _sharded_policy(self, agent_pol):
        """A learned binary-action strategy under arm sharding selects through the sharded radix select (collectives inside
        the step, so such rollouts are launched eagerly)."""
        return (self.world > 1 and hasattr(agent_pol, "act_test_device_sharded") and getattr(agent_pol, "act_dim", 0) == 2
                and float(np.asarray(agent_pol.C)[0]) == 0.0)

    def _agent_shard(self, agent_pol):
        key = id(agent_pol)
        if self._pol_shard[0] != key:
            self._pol_shard = (key, agent_pol.shard(self.arm0, self.arm0 + self.N))
        return self._pol_shard[1]

    def _advance(self, t):
        """Host-side counters of one step (what the eager step would have advanced)."""
        if t < self.T:
            self.env._t += 1
            self.ac._steps += 1

    def rollout(self, agent_pol):
        env, ac, T, N, E = self.env, self.ac, self.T, self.N, self.E
        sharded = self.world > 1
        s0 = self.s.contiguous()
        lamb, h1 = self._lambda(s0)
        self.lamb_buf.copy_(lamb)
        self.s_traj[:, 0] = s0
        # W1n is constant over the rollout: split it once (fp16 hi/lo rows for gemm3h, or tf32 hi/lo for the library arm)
        if self.library_gemm:
            self._w1n_split = ops.split_tf32(ac.W1n.reshape(-1, self.D))
        else:
            self._w1n_split = ops.split16_rows(ac.W1n.reshape(-1, self.D), out=self._w1n_split)
        # Philox counters of this rollout through device words (see environments.counter_base)
        self.t_words[0].fill_(env._t)
        self.t_words[1].fill_(ac._steps)
        env.counter_base(self.t_words[0])
        ac.counter_base(self.t_words[1])
        use_graphs = (self.use_graphs and not self.library_gemm and self.graph_ok is not False
                      and not self._sharded_policy(agent_pol))
        replay = use_graphs and self.graphs is not None and self.graph_pol is agent_pol
        capture = use_graphs and not replay and self.rollouts_done >= 1
        if capture:
            self.graphs, self.graph_pol = [], agent_pol
        t_env0, t_ac0 = env._t, ac._steps
        try:
            for t in range(T + 1):
                if sharded:
                    self._gather_arms(self.s_traj[:, t].contiguous())
                if replay:
                    self.graphs[t].replay()
                elif capture:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=self.graph_pool):
                        self._step(t, agent_pol)
                    if self.graph_pool is None:
                        self.graph_pool = g.pool()
                    self.graphs.append(g)
                    g.replay()                           # capture does not execute
                else:
                    self._step(t, agent_pol)
                self._advance(t)
        except Exception:
            if not capture:
                raise
            # the agent strategy (or something else in the step) cannot be captured, e.g. it synchronises: run this and
            # every later rollout eagerly
            self.graph_ok, self.graphs = False, None
            torch.cuda.synchronize()
            env._t, ac._steps = t_env0, t_ac0
            for t in range(T + 1):
                if sharded:
                    self._gather_arms(self.s_traj[:, t].contiguous())
                self._step(t, agent_pol)
                self._advance(t)
        env.counter_base(None)
        ac.counter_base(None)
        self.rollouts_done += 1
        # per-step sums over this rank's arms, all T steps at once: rewards R[arm, s_{t+1}] (:521-527), the counterfactual
        # estimates, the action costs
        rew = torch.gather(self.R, 1, self.s_traj[:, 1:].reshape(N, -1).long()).view(N, T, E)
        self.part[:, 0].copy_(rew.sum(dim=0))
        self.part[:, 1].copy_(self.cf.sum(dim=1))
        self.part[:, 2].copy_(self.C[self.a.long()].sum(dim=0))
        ret_local = self.part[:, 0].sum(dim=0)
        if sharded:
            self._allreduce(self.part)
            self._allreduce(ret_local)
        # nature critic over [obs, a_agent] of ALL arms and all T + 1 steps (:318-325) in one batch: every rank's layer-1
        # partial over its own arms, one all-reduce of [(T+1) E, h]
        a_ext = torch.cat([self.a, self.a_boot[:, None]], dim=1).contiguous()
        vn = self.nets.critic_values(self.s_traj, a_ext, ac.nature_scales()[1], T + 1, E,
                                     allreduce=self._allreduce if sharded else None).view(T + 1, E)
        self.val_nat.copy_(vn[:T])
        last_v_nature = vn[T].contiguous()
        torch.sub(self.part[:, 0] - self.part[:, 1], lamb * self.part[:, 2], out=self.rew_nat)   # lamb_adjusted_r_nature (:707-719)
        return lamb, h1, s0, self.v_boot.clone(), last_v_nature, ret_local

    def update(self, lamb, h1, s0, last_v_agent, last_v_nature):
        ac, T, N, E, D = self.ac, self.T, self.N, self.E, self.D
        ep = self.epoch_idx
        st = self.stages = _Stages(self.profile)
        st.mark("gae")
        a = self.a
        # ---- MA_RMABPPO_Buffer.finish_path + get (:107-193) ----
        adv, ret, _, _ = ops.gae(self.s_traj, a, self.val, last_v_agent.contiguous(), lamb, self.C, self.R, self.gamma,
                                 self.lam_other, normalize=True)
        zero = torch.zeros((1, T, E), dtype=_F32, device=self.dev)
        adv_n, ret_n, _, _ = ops.gae_explicit(self.rew_nat[None].contiguous(), zero, self.val_nat[None].contiguous(),
                                              last_v_nature.reshape(1, E).contiguous(), lamb, self.gamma, self.lam_other,
                                              normalize=True)
        # ---- agent policies (:491-502) ----
        st.mark("agent_policies")
        hp = ops.hyper(self.clip, self._entropy(ep), self.pi_lr_agent, self.vf_lr_agent, self.pi_iters, 0, self.steps_pi,
                       self.steps_v)
        ops.ppo_update(ac._net(E, False), hp, ac.Wpi, ac.Wv, self.adam_pi, self.adam_v, self.s_traj, a, adv, self.logp,
                       ret, lamb)
        self.steps_pi += self.pi_iters
        # ---- agent critics over [x_i, bounded nature vector] (:505-513) ----
        st.mark("agent_critics")
        net = ac._net(E, False)
        h = ac.hidden_sizes[0]
        M = T * E
        nat2d = self.nat_b.reshape(M, D)
        if self.library_gemm:
            natM = ops.split_tf32(nat2d)                                         # sample m = t*E + e; (hi, lo) for 3xTF32
        else:
            # operands of the package's own product (csrc/gemm3h.cu), constant over the iterations: nature rows [M, D]
            # for z1 = nat . W1n^T and the transposed split [D, M] for dW1n = dz1^T . nat
            nat_rows, nat_cols = ops.split16_rows(nat2d), ops.split16_cols(nat2d)
        for _ in range(self.v_iters):
            # nature block of all N first layers as ONE dense product: [T*E, D] x [D, N*h] -> sample-major [T*E, N, h]
            st.mark("critics.z1_gemm")
            if self.library_gemm:
                w1h, w1l = ops.split_tf32(ac.W1n.reshape(N * h, D))
                z1 = ops.matmul_3xtf32(natM, (w1h.t(), w1l.t())).reshape(M, N, h)
                del w1h, w1l
            else:
                z1 = ops.gemm3h(nat_rows, ops.split16_rows(ac.W1n.reshape(N * h, D)), D, n_fast=False).reshape(M, N, h)
            st.mark("critics.kernel")
            hpv = ops.hyper(self.clip, 0.0, self.pi_lr_agent, self.vf_lr_agent, 0, 1, self.steps_pi, self.steps_v)
            dz1, _ = ops.critic_extra_iter(net, hpv, ac.Wv, self.adam_v, self.s_traj, ret, lamb, z1, sample_major=True)
            del z1
            st.mark("critics.grad_gemm")
            if self.library_gemm:
                dh, dl = ops.split_tf32(dz1.reshape(M, N * h))
                gW1n = ops.matmul_3xtf32((dh.t(), dl.t()), natM).reshape(N, h, D)    # dW1_nature = dz1^T . nature, one GEMM
                del dh, dl, dz1
                st.mark("critics.adam")
                self.steps_v += 1
                ops.adam_step(ac.W1n, gW1n, self.m_w1n, self.v_w1n, self.vf_lr_agent, self.steps_v)
                del gW1n
            else:
                # dW1n = dz1^T . nature and its Adam step in one kernel: the gradient is consumed in the product's epilogue
                self.steps_v += 1
                ops.gemm3h_adam(ops.split16_cols(dz1.reshape(M, N * h)), nat_cols, M, ac.W1n.view(N * h, D),
                                self.m_w1n.view(N * h, D), self.v_w1n.view(N * h, D), self.vf_lr_agent, self.steps_v, n_fast=True)
                del dz1
        # ---- lambda net and the nature nets, every lamb_update_freq epochs (:516-576) ----
        st.mark("lambda_and_nature_nets")
        if ep % self.freq == 0 and ep > 0:
            sharded = self.world > 1
            if (self.epochs - ep) > self.final:
                cd0 = ops.cost_togo(a, self.C, self.gamma)[0].contiguous()
                self._allreduce(cd0)
                coef = (float(ac.B) / (1.0 - self.gamma) - cd0).contiguous()
                if self.obs_scale != 1.0:
                    _, h1 = self._lambda(s0, self.obs_scale)
                ops.lambda_sgd(self.lam_params, s0, self.N_total, self.arm0, self.obs_scale, h1, coef, self.lm_lr)
            # compute_loss_pi_nature / compute_loss_v_nature (:359-364, :421-428) divide the RAW buffer observations by
            # nature_state_norm when the nets are not one-hot and by nothing otherwise -- NOT by state_norm as step() does
            # (ma_rmabppo_core.py:276-289); kept as the reference has it
            sc = 1.0 if ac.one_hot_encode else 1.0 / float(ac.nature_state_norm)
            ar = self._allreduce if sharded else None
            act = self.a_nat.reshape(T * E, D)
            advn, lpo, retn = adv_n.reshape(-1).contiguous(), self.logp_nat.reshape(-1), ret_n.reshape(-1).contiguous()
            for _ in range(self.pi_iters):      # entropy_coeff = 0 for nature (:544)
                self.nets.actor_step(self.s_traj, T, E, sc, act, lpo, advn, self.clip, self.pi_lr_nature, allreduce=ar)
            for _ in range(self.v_iters):
                self.nets.critic_step(self.s_traj, a, T, E, sc, retn, self.vf_lr_nature, allreduce=ar)
            if sharded:
                self.nets.exchange_shards(self._allreduce)
            self.nets.unpack()                  # the drop-in's torch modules follow the packed parameters
        st.mark(None)
        self.epoch_idx += 1

    def epoch(self, agent_pol):
        lamb, h1, s0, lv, lvn, ret_agent = self.rollout(agent_pol)
        vv = self._allreduce(self.val.sum().reshape(1)) / float(self.N_total * self.T * self.E)
        stats = {"EpActualRetAgent": ret_agent, "EpLambAdjRetNature": self.rew_nat.sum(dim=0), "Lamb": lamb,
                 "VVals_agent": vv, "VVals_nature": self.val_nat.mean().reshape(1)}
        self.update(lamb, h1, s0, lv, lvn)
        self.s = self.env.reset_device()                                      # :779
        return stats

    def export_lambda(self):
        flat = self.lam_params.detach().clone()
        if self.world > 1:   # each rank trained the first-layer columns of its own arms
            w1 = flat[:8 * self.N_total].view(8, self.N_total)
            own = torch.zeros_like(w1)
            own[:, self.arm0:self.arm0 + self.N] = w1[:, self.arm0:self.arm0 + self.N]
            self._allreduce(own)
            w1.copy_(own)
        flat, o = flat.cpu(), 0
        with torch.no_grad():
            for p in self.ac.lambda_net.parameters():
                p.copy_(flat[o:o + p.numel()].reshape(p.shape))
                o += p.numel()


class NatureOracle:

    def __init__(self, data, N, S, A, B, seed, REWARD_BOUND, nature_kwargs=dict(), home_dir="", exp_name="",
                 sampled_nature_parameter_ranges=None, pop_size=0, one_hot_encode=True, non_ohe_obs_dim=None, state_norm=1,
                 nature_state_norm=1, n_envs=1, device="cuda", rank=0, world_size=1, process_group=None):
        self.rank, self.world, self.pg = int(rank), int(world_size), process_group
        self.data, self.home_dir, self.exp_name, self.REWARD_BOUND = data, home_dir, exp_name, REWARD_BOUND
        self.N, self.S, self.A, self.B, self.seed = N, S, A, B, seed
        self.sampled_nature_parameter_ranges = sampled_nature_parameter_ranges
        self.nature_state_norm, self.pop_size = nature_state_norm, pop_size
        self.one_hot_encode, self.non_ohe_obs_dim, self.state_norm = one_hot_encode, non_ohe_obs_dim, state_norm
        self.n_envs, self.device = int(n_envs), device
        if data == 'armman':
            self.env_fn = lambda: ARMMANRobustEnv(N, B, seed, n_envs=self.n_envs, device=device)
        elif data == 'counterexample':
            self.env_fn = lambda: CounterExampleRobustEnv(N, B, seed, n_envs=self.n_envs, device=device)
        elif data == 'sis':
            self.env_fn = lambda: SISRobustEnv(N, B, pop_size, seed, n_envs=self.n_envs, device=device)
        else:
            raise ValueError("data must be 'counterexample', 'armman' or 'sis'")
        self.ma_actor_critic = RMABLambdaNatureOracle
        self.nature_kwargs = nature_kwargs
        self.strat_ind = -1
        self.env = self.env_fn()
        self.env.seed(seed)
        self.env.sampled_parameter_ranges = self.sampled_nature_parameter_ranges
        self.history = []

    def best_response(self, nature_strats, nature_eq, add_to_seed):
        self.strat_ind += 1
        return self.best_response_per_cpu(nature_strats, nature_eq, add_to_seed, seed=self.seed, **self.nature_kwargs)

    def best_response_per_cpu(self, agent_strats, agent_eq, add_to_seed, ma_actor_critic=None, ac_kwargs=dict(), seed=0,
                              steps_per_epoch=4000, epochs=50, gamma=0.99, clip_ratio=0.2, pi_lr_agent=3e-4,
                              pi_lr_nature=3e-4, vf_lr_agent=1e-3, vf_lr_nature=1e-3, qf_lr=1e-3, lm_lr=5e-2,
                              train_pi_iters=80, train_v_iters=80, train_q_iters=80, lam_OTHER=0.97, max_ep_len=1000,
                              start_entropy_coeff=0.0, end_entropy_coeff=0.0, target_kl=0.01, logger_kwargs=dict(),
                              save_freq=10, lamb_update_freq=10, init_lambda_trains=0, final_train_lambdas=0):
        env = self.env
        from .agent_oracle import check_episode_length
        check_episode_length(max_ep_len, steps_per_epoch)
        shard = None
        if self.world > 1:   # arm-sharded ranks must draw the same nets and pick the same agent strategy
            from .mpi_tools import arm_shard
            a0, n_loc = arm_shard(env.N, self.rank, self.world)
            shard = (a0, a0 + n_loc)
            torch.manual_seed(int(seed) + 7919 * (self.strat_ind + 1))
            np.random.seed((int(seed) + 7919 * (self.strat_ind + 1)) % (2 ** 31))
        ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges,
                                    env.action_dim_nature, env=env, N=env.N, C=env.C, B=env.B, strat_ind=self.strat_ind,
                                    one_hot_encode=self.one_hot_encode, non_ohe_obs_dim=self.non_ohe_obs_dim,
                                    state_norm=self.state_norm, nature_state_norm=self.nature_state_norm,
                                    n_envs=self.n_envs, device=self.device, arm_shard=shard, **ac_kwargs)
        agent_eq = np.array(agent_eq, dtype=np.float64)
        agent_eq[agent_eq < 0] = 0
        agent_eq = agent_eq / agent_eq.sum()
        agent_pol = np.random.choice(agent_strats, p=agent_eq)                   # :651
        tr = MARolloutTrainer(env, ac, steps_per_epoch, gamma, lam_OTHER, clip_ratio, pi_lr_agent, pi_lr_nature,
                              vf_lr_agent, vf_lr_nature, lm_lr, train_pi_iters, train_v_iters, epochs, lamb_update_freq,
                              final_train_lambdas, start_entropy_coeff, end_entropy_coeff, seed, rank=self.rank,
                              world_size=self.world, process_group=self.pg)
        tr.init_lambda_trains(init_lambda_trains)
        self.history, self.agent_pol_history = [], []
        for epoch in range(epochs):
            # the agent strategy is re-drawn from the mixed equilibrium every lamb_update_freq epochs (:669-670); under
            # world > 1 np.random was seeded identically on every rank above, so all ranks draw the same strategy
            if epoch % lamb_update_freq == 0 and epoch > 0:
                agent_pol = np.random.choice(agent_strats, p=agent_eq)
            self.agent_pol_history.append(agent_pol)
            self.history.append({k: v.detach().cpu().numpy() for k, v in tr.epoch(agent_pol).items()})
        tr.export_lambda()
        return ac
