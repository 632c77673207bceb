"""Batched RMABPPO epoch on device: the rollout-and-update loop of AgentOracle.best_response_per_cpu
(robust_rmab/algos/rmabppo/agent_oracle.py:456-593) for N arms x E env replicas, all state resident in HBM.

One `epoch()` =  lambda-net forward (:497)  ->  T steps of ac.step / nature / env.step / buf.store (:506-537)
              ->  bootstrap value + finish_path (:553-563)  ->  env.reset (:568)  ->  update() (:399-454).
Arms shard across ranks: every rank holds all E envs of its contiguous arm range;
the only exchanges are the lambda-net first-layer partial sums [E,8], the arm-summed discounted cost
[T,E] and the logging scalars, each one NCCL all-reduce on the compute stream.

Layout: states [N,E] u8, trajectories [N,T,E], packed per-arm weights [N,P] (see include/rmab.h).
"""
import math

import numpy as np
import torch

from . import _lib, ops

DOMAINS = {"counterexample": _lib.DOMAIN_CE, "armman": _lib.DOMAIN_ARMMAN, "sis": _lib.DOMAIN_SIS}
_DOMAIN_S = {_lib.DOMAIN_CE: 2, _lib.DOMAIN_ARMMAN: 3}


def linear_default_init_(W, in_dim, h, out, generator):
    """torch.nn.Linear default init (uniform(+-1/sqrt(fan_in)) for weight and bias, rmabppo_core.py:24-29)
    written straight into a packed [N,P] tensor."""
    o = 0
    for fan_out, fan_in in ((h, in_dim), (h, h), (out, h)):
        b = 1.0 / math.sqrt(fan_in)
        n = fan_out * fan_in + fan_out
        W[:, o:o + n].uniform_(-b, b, generator=generator)
        o += n
    assert o == W.shape[1]
    return W


class _Stage:
    def __init__(self, timers, name):
        self.t, self.name = timers, name

    def __enter__(self):
        self.a = torch.cuda.Event(enable_timing=True)
        self.b = torch.cuda.Event(enable_timing=True)
        self.a.record()

    def __exit__(self, *exc):
        self.b.record()
        self.t.setdefault(self.name, []).append((self.a, self.b))
        return False


class _Null:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _Null()


class RMABPPOTrainer:
    """Device-resident RMABPPO for the table domains (counterexample, ARMMAN) under a per-arm nature table
    (the baseline nature policies: nature_baselines_counterexample.py:255-256, nature_baselines_armman.py:224-232).
    """

    def __init__(self, domain, n_arms, n_envs, steps_per_epoch, hidden=16, B=None, gamma=0.9, lam_other=0.97,
                 clip_ratio=2.0, pi_lr=2e-3, vf_lr=2e-3, lm_lr=2e-3, train_pi_iters=20, train_v_iters=20,
                 epochs=20, lamb_update_freq=4, final_train_lambdas=2, start_entropy_coeff=0.5,
                 end_entropy_coeff=0.0, seed=0, ranges=None, nature=None, device="cuda", rank=0, world_size=1,
                 process_group=None, keep_trajectory=False):
        """keep_trajectory: also materialise the [N,T+1,E] / [N,T,E] state / action byte arrays in HBM (`s_traj`, `a`);
        by default the fused epoch keeps the trajectory on chip and only the loss statistics leave the SM."""
        self.domain = DOMAINS[domain] if isinstance(domain, str) else int(domain)
        if self.domain not in _DOMAIN_S:
            raise ValueError("RMABPPOTrainer covers the table domains; SIS goes through ops.env_step_sis")
        self.S, self.A = _DOMAIN_S[self.domain], 2
        self.N_total, self.E, self.T, self.h = int(n_arms), int(n_envs), int(steps_per_epoch), int(hidden)
        self.rank, self.world, self.pg = int(rank), int(world_size), process_group
        per = (self.N_total + self.world - 1) // self.world
        self.arm0 = min(self.rank * per, self.N_total)
        self.N = min(per, self.N_total - self.arm0)
        self.B = float(B if B is not None else self.N_total // 3)
        self.gamma, self.lam_other, self.clip_ratio = float(gamma), float(lam_other), float(clip_ratio)
        self.pi_lr, self.vf_lr, self.lm_lr = float(pi_lr), float(vf_lr), float(lm_lr)
        self.pi_iters, self.v_iters, self.epochs = int(train_pi_iters), int(train_v_iters), int(epochs)
        self.lamb_update_freq, self.final_train_lambdas = int(lamb_update_freq), int(final_train_lambdas)
        self.ent0, self.ent1 = float(start_entropy_coeff), float(end_entropy_coeff)
        self.seed = int(seed)
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("RMABPPOTrainer runs on CUDA only (no CPU fallback)")
        N, E, T, S, A, h = self.N, self.E, self.T, self.S, self.A, self.h
        dev = self.dev
        f32, u8 = torch.float32, torch.uint8
        self.net = ops.net_desc(N, E, S, A, h, True, 1.0)
        self.Ppi, self.Pv = ops.n_params(S + 1, h, A), ops.n_params(S + 1, h, 1)
        # nets are drawn in blocks of 1024 arms from generators keyed by (seed, block), so a rank only draws the
        # blocks it owns and any sharding gives every arm the same initial weights
        self.Wpi = self._init_arm_nets(S + 1, h, A, 1).to(dev)
        self.Wv = self._init_arm_nets(S + 1, h, 1, 2).to(dev)
        g = torch.Generator(device="cpu").manual_seed(self.seed)
        self.lam_params = self._init_lambda_net(g).to(dev)
        self.adam_pi = torch.zeros((N, 2, self.Ppi), dtype=f32, device=dev)
        self.adam_v = torch.zeros((N, 2, self.Pv), dtype=f32, device=dev)
        self.steps_pi = self.steps_v = 0
        rs = np.random.RandomState(self.seed)
        if ranges is None:
            ranges = self.default_ranges(self.domain, self.N_total, rs)
        if nature is None:
            lo, hi = ranges[..., 0], ranges[..., 1]
            nature = lo + rs.rand(*lo.shape) * (hi - lo)
        sl = slice(self.arm0, self.arm0 + N)
        self.ranges = torch.as_tensor(np.ascontiguousarray(ranges[sl]), dtype=f32).to(dev)
        self.nature = torch.as_tensor(np.ascontiguousarray(nature[sl]), dtype=f32).to(dev)
        Rrow = [0.0, 1.0] if self.domain == _lib.DOMAIN_CE else [1.0, 0.5, 0.0]
        self.R = torch.tensor(Rrow, dtype=f32).repeat(N, 1).contiguous().to(dev)
        self.C = torch.tensor([0.0, 1.0], dtype=f32, device=dev)
        self.keep_trajectory = bool(keep_trajectory)
        # row 0 = the epoch's start states; rows 1..T (and `a`) exist only when the trajectory is kept in HBM
        self.s_traj = torch.empty((N, T + 1 if self.keep_trajectory else 1, E), dtype=u8, device=dev)
        self.a = torch.empty((N, T, E), dtype=u8, device=dev) if self.keep_trajectory else None
        self.last_val = torch.empty((N, E), dtype=f32, device=dev)
        self.stats = torch.empty((N, ops.stat_rows(self.domain), E), dtype=f32, device=dev)
        self.arm_stats = torch.empty((N, 4), dtype=f32, device=dev)
        # one reduction buffer per epoch: per-env reward sum [E], per-env discounted cost-to-go [E], VVals sum [1]
        self.red_buf = torch.zeros((2 * E + 1,), dtype=f32, device=dev)
        self.env_sums = self.red_buf[:2 * E].view(2, E)
        self.ws = torch.empty((max(int(_lib.lib().rmab_epoch_ws_bytes(N, E)), 8) + 7) // 8, dtype=torch.float64,
                              device=dev)
        self.adv = self.ret = None
        self.epoch_idx = 0
        self.t_global = 0
        self.s_traj[:, 0] = ops.reset_states(N, E, S, ops.rng(self.seed, _lib.STREAM_RESET, 0, arm0=self.arm0))
        self.last_stats = None
        self._timers = None
        self._timer_epoch0 = 0

    @property
    def timers(self):
        return self._timers

    @timers.setter
    def timers(self, v):
        self._timers = v
        self._timer_epoch0 = self.epoch_idx

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def default_ranges(domain, N, rs):
        """sampled_parameter_ranges.  Counterexample: PARAMETER_RANGES itself, no sub-sampling
        (bandit_env_robust.py:790-824); ARMMAN: a sorted U(0,1) pair scaled into PARAMETER_RANGES (:468-481)."""
        if domain == _lib.DOMAIN_CE:
            return np.array([[0.0, 1.0], [0.05, 0.9], [0.1, 0.95]])[np.arange(N) % 3]
        else:
            rA = [[[0.0, 1.0], [0.0, 1.0]], [[0.5, 1.0], [0.5, 1.0]], [[0.35, 0.85], [0.35, 0.85]]]
            rB = [[[0.0, 1.0], [0.0, 1.0]], [[0.35, 0.85], [0.15, 0.65]], [[0.35, 0.85], [0.35, 0.85]]]
            rC = [[[0.0, 1.0], [0.0, 1.0]], [[0.35, 0.85], [0.0, 0.5]], [[0.35, 0.85], [0.35, 0.85]]]
            nA, nB = int(N * 0.2), int(N * 0.2)
            base = np.empty((N, 3, 2, 2))
            base[:nA], base[nA:nA + nB], base[nA + nB:] = rA, rB, rC
        draw = np.sort(rs.rand(*base.shape), axis=-1)
        lo = base.min(-1)[..., None]
        return draw * (base.max(-1)[..., None] - lo) + lo

    def _init_arm_nets(self, in_dim, h, out, salt, block=1024):
        P = ops.n_params(in_dim, h, out)
        W = torch.empty((self.N, P))
        b0, b1 = self.arm0 // block, (self.arm0 + self.N + block - 1) // block
        for b in range(b0, b1):
            g = torch.Generator(device="cpu").manual_seed((self.seed * 1000003 + salt) * 1000003 + b)
            blk = linear_default_init_(torch.empty((block, P)), in_dim, h, out, g)
            lo, hi = max(b * block, self.arm0), min((b + 1) * block, self.arm0 + self.N)
            W[lo - self.arm0:hi - self.arm0] = blk[lo - b * block:hi - b * block]
        return W.contiguous()

    def _init_lambda_net(self, g):
        """MLPLambdaNet N -> 8 -> 8 -> 1 (rmabppo_core.py:108-115, :170-171), packed W1 b1 W2 b2 W3 b3."""
        Nt = self.N_total
        P = 8 * Nt + 8 + 64 + 8 + 8 + 1
        p = torch.empty((1, P))
        o = 0
        for fan_out, fan_in in ((8, Nt), (8, 8), (1, 8)):
            b = 1.0 / math.sqrt(fan_in)
            n = fan_out * fan_in + fan_out
            p[:, o:o + n].uniform_(-b, b, generator=g)
            o += n
        return p[0].contiguous()

    def _allreduce(self, t):
        if self.world > 1:
            torch.distributed.all_reduce(t, group=self.pg)
        return t

    def entropy_coeff(self, epoch):
        """agent_oracle.py:402-409 with head = linspace(start, end, epochs)[epoch] (:491, :578)."""
        head = float(np.linspace(self.ent0, self.ent1, self.epochs)[min(epoch, self.epochs - 1)])
        if (self.epochs - epoch) > self.final_train_lambdas:
            return float(np.linspace(head, 0, self.lamb_update_freq)[epoch % self.lamb_update_freq])
        return 0.0

    def lambda_update_due(self, epoch):
        return epoch % self.lamb_update_freq == 0 and epoch > 0 and (self.epochs - epoch) > self.final_train_lambdas

    # ------------------------------------------------------------------------------------------
    def compute_lambda(self):
        s0 = self._s0()
        if self.world == 1:
            lamb, h1 = ops.lambda_forward(self.lam_params, s0, self.N_total, self.arm0, 1.0, want_h1=True)
        else:
            _, h1 = ops.lambda_forward(self.lam_params, s0, self.N_total, self.arm0, 1.0, want_h1=True,
                                       want_lamb=False)
            self._allreduce(h1)
            lamb, _ = ops.lambda_forward(self.lam_params, s0, self.N_total, self.arm0, 1.0, h1_in=h1)
        return lamb, h1, s0

    def _lambda_step(self, s0, h1, lr):
        coef = self.B / (1.0 - self.gamma) - self.env_sums[1]           # compute_loss_lambda (:360-376)
        ops.lambda_sgd(self.lam_params, s0, self.N_total, self.arm0, 1.0, h1, coef, lr)

    def prewarm(self):
        """Run the once-every-lamb_update_freq-epochs branch with a zero learning rate (no state change), so that a timed
        region that contains its first occurrence does not also contain its one-off kernel loads."""
        lamb, h1, s0 = self.compute_lambda()
        self._allreduce(self.red_buf)
        self._lambda_step(s0, h1, 0.0)
        torch.cuda.synchronize()

    def _s0(self):
        # s_traj[:, 0] is a strided view; the lambda kernels want [N,E] contiguous
        return self.s_traj[:, 0].contiguous()

    # ------------------------------------------------------------------------------------------
    def _stage(self, name):
        """Context manager recording a CUDA-event pair around one stage on the launching stream
        (only when `self.timers` is a dict, i.e. inside bench.py's timed region)."""
        return _Stage(self.timers, name) if self.timers is not None else _NULL

    def stage_times(self):
        """Average ms per epoch of each recorded stage (synchronises)."""
        torch.cuda.synchronize()
        out = {}
        for name, pairs in (self.timers or {}).items():
            out[name] = sum(a.elapsed_time(b) for a, b in pairs) / max(1, self.epoch_idx - self._timer_epoch0)
        return out

    def buffer_bytes(self):
        N, T, E = self.N, self.T, self.E
        traj = N * (T + 1) * E + N * T * E if self.keep_trajectory else N * E
        return traj + self.stats.numel() * 4 + (self.Wpi.numel() + self.Wv.numel()) * 4 * 3

    def algorithmic_bytes(self):
        """Algorithmic HBM bytes per launch of each stage: SURVEY.md section 8d per-unit figures x N*T*E
        (rollout 10 B + GAE 14 B per arm-step; update 14 B per arm-step + 24 B per parameter)."""
        n = self.N * self.T * self.E
        return {"rollout_gae_stats": 24 * n, "ppo_update": 14 * n + 24 * self.N * (self.Ppi + self.Pv)}

    def dominant_kernel(self, stage_ms):
        if not stage_ms:
            return "none", 0.0, 0
        name = max(stage_ms, key=stage_ms.get)
        return name, stage_ms[name], self.algorithmic_bytes().get(name, 0)

    def epoch_host(self, s0_host, nature_host):
        """End-to-end epoch through host buffers: H2D of the start states and the nature policy's parameter
        table (pinned), the epoch, D2H of the per-env results."""
        if self.s_traj.shape[1] == 1:                       # [N,1,E]: row 0 is a contiguous [N,E] view, one H2D copy
            self.s_traj[:, 0].copy_(s0_host, non_blocking=True)
        else:
            self.s_traj[:, 0].copy_(s0_host.to(self.dev, non_blocking=True))
        self.nature.copy_(nature_host, non_blocking=True)
        st = self.epoch(want_stats=True)
        # the per-env results packed into ONE device buffer and read back by ONE copy into pinned memory (three copies into
        # pageable memory before: each of those is synchronous); the returned arrays are views of that pinned buffer, valid
        # until the next call
        keys = list(st.keys())
        sizes = [int(st[k].numel()) for k in keys]
        total = sum(sizes)
        if getattr(self, "_host_out", None) is None or self._host_out.numel() != total:
            self._dev_out = torch.empty(total, dtype=torch.float32, device=self.dev)
            self._host_out = torch.empty(total, dtype=torch.float32).pin_memory()
        torch.cat([st[k].reshape(-1).to(torch.float32) for k in keys], out=self._dev_out)
        self._host_out.copy_(self._dev_out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out, o = {}, 0
        for k, n_ in zip(keys, sizes):
            out[k] = self._host_out[o:o + n_].view(st[k].shape)
            o += n_
        return out

    def epoch(self, want_stats=True, keep_buffers=False):
        """One rollout + update epoch.  Returns a dict of device tensors (no host sync):
        EpActualRet [E] (sum over arms and steps of the reward, agent_oracle.py:521-527), Lamb [E], VVals [1].
        keep_buffers: also materialise the normalised advantages / returns / last values of buf.get()."""
        ep = self.epoch_idx
        N, E, T, S, A = self.N, self.E, self.T, self.S, self.A
        with self._stage("lambda"):
            lamb, h1, s0 = self.compute_lambda()
        if keep_buffers and not self.keep_trajectory:
            raise ValueError("keep_buffers needs a trainer built with keep_trajectory=True")
        if keep_buffers and self.adv is None:
            self.adv = torch.empty((N, T, E), dtype=torch.float32, device=self.dev)
            self.ret = torch.empty((N, T, E), dtype=torch.float32, device=self.dev)
        with self._stage("rollout_gae_stats"):
            ops.epoch_table(self.domain, self.net, T, self.Wpi, self.Wv, lamb, self.nature, self.ranges, self.R, self.C,
                            self.gamma, self.lam_other, self.seed, self.t_global, self.t_global,
                            self.s_traj if self.keep_trajectory else self.s_traj[:, 0], self.a,
                            stats=self.stats, arm_stats=self.arm_stats, env_sums=self.env_sums, ws=self.ws,
                            adv=self.adv if keep_buffers else None, ret=self.ret if keep_buffers else None,
                            last_val=self.last_val, env0=0, arm0=self.arm0)
        hp = ops.hyper(self.clip_ratio, self.entropy_coeff(ep), self.pi_lr, self.vf_lr, self.pi_iters, self.v_iters,
                       self.steps_pi, self.steps_v)
        with self._stage("ppo_update"):
            ops.ppo_update_table(self.net, hp, T, self.Wpi, self.Wv, self.adam_pi, self.adam_v, self.stats,
                                 self.arm_stats, lamb)
        self.steps_pi += self.pi_iters
        self.steps_v += self.v_iters
        with self._stage("lambda_sgd_stats"):
            due = self.lambda_update_due(ep)
            if want_stats:
                SA = S * A
                self.red_buf[2 * E] = (self.stats[:, 3 * SA:3 * SA + S] * self.stats[:, 3 * SA + 2 * S:3 * SA + 3 * S]).sum()
            if want_stats or due:
                self._allreduce(self.red_buf)                           # the epoch's only cross-rank exchange besides h1
            if due:
                self._lambda_step(s0, h1, self.lm_lr)
            stats = None
            if want_stats:
                stats = {"EpActualRet": self.env_sums[0].clone(), "Lamb": lamb,
                         "VVals": self.red_buf[2 * E:] / float(self.N_total * T * E)}
        # env.reset() at the end of the epoch (:568)
        self.epoch_idx += 1
        self.t_global += T
        self.s_traj[:, 0] = ops.reset_states(N, E, self.S, ops.rng(self.seed, _lib.STREAM_RESET, self.epoch_idx,
                                                                   arm0=self.arm0))
        self.last_stats = stats
        return stats


class EnvShardedRMABPPOTrainer(RMABPPOTrainer):
    """The reference's own MPI layout (SpinningUp `mpi_fork`: every process all N arms, its own environments; gradients
    averaged with `mpi_avg_grads`, advantage statistics with `mpi_statistics_scalar` -- utils/mpi_pytorch.py:19-30,
    mpi_tools.py:76-97) on k GPUs: rank r simulates envs [r E/k, (r+1) E/k) of ALL arms.  What crosses NVLink per epoch:

      * the raw advantage moments `[N,2]` fp64 (all-reduce) -- get()'s per-arm normalisation is over the envs of all ranks;
        the fused rollout kernel runs twice (export moments, then normalise with the global ones: same Philox counters, same
        trajectory) instead of writing the trajectory to HBM;
      * the loss statistics, re-sharded from envs to ARMS (all-to-all of `[N/k, K, E/k]` blocks): averaging per-sample
        gradients over all ranks' envs IS the full-batch gradient, and the statistics are the batch -- so a
        rank then runs the fused update kernels for its N/k arms over all E envs; no `[N,P]` gradient all-reduce per Adam
        iteration (3.7 GB each at BASELINE configs[2]);
      * the updated nets of every rank's arms (all-gather `[N,P]`), lambda `[E]` (all-gather), the lambda-net gradient
        (all-reduce) and the logging sums.

    The small-N / large-E fallback of SURVEY.md 8(e); the arm-sharded RMABPPOTrainer is the fast path."""

    def __init__(self, domain, n_arms, n_envs, steps_per_epoch, rank=0, world_size=1, process_group=None, **kw):
        kw.pop("keep_trajectory", None)
        E_total = int(n_envs)
        if E_total % int(world_size):
            raise ValueError("env sharding needs n_envs divisible by the number of ranks")
        self.E_total, self.erank, self.eworld, self.epg = E_total, int(rank), int(world_size), process_group
        self.env0 = self.erank * (E_total // self.eworld)
        super().__init__(domain, n_arms, E_total // self.eworld, steps_per_epoch, rank=0, world_size=1, **kw)
        N, dev = self.N, self.dev
        per = (N + self.eworld - 1) // self.eworld
        self.per = per
        self.a0 = min(self.erank * per, N)
        self.a1 = min(self.a0 + per, N)
        self.mom = torch.zeros((N, 2), dtype=torch.float64, device=dev)
        self.mom_in = torch.zeros((N, 2), dtype=torch.float32, device=dev)
        K = self.stats.shape[1]
        self.stats_send = torch.zeros((self.eworld * per, K, self.E), dtype=torch.float32, device=dev)
        self.stats_recv = torch.empty_like(self.stats_send)
        self.net_upd = ops.net_desc(self.a1 - self.a0, E_total, self.S, self.A, self.h, True, 1.0)
        self.s_traj[:, 0] = ops.reset_states(N, self.E, self.S, ops.rng(self.seed, _lib.STREAM_RESET, 0, env0=self.env0))

    def _ar(self, t):
        if self.eworld > 1:
            torch.distributed.all_reduce(t, group=self.epg)
        return t

    def _gather_rows(self, x_local):
        """Rows [a0, a1) of every rank -> the full [N, ...] tensor (ragged last shard padded)."""
        if self.eworld == 1:
            return x_local
        pad = torch.zeros((self.per,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=self.dev)
        pad[:x_local.shape[0]] = x_local
        out = torch.empty((self.eworld * self.per,) + tuple(x_local.shape[1:]), dtype=x_local.dtype, device=self.dev)
        torch.distributed.all_gather_into_tensor(out, pad, group=self.epg)
        return out[:self.N]

    @staticmethod
    def reshard_env_to_arm(send, recv, k, per, group=None):
        """send `[k * per, K, E_loc]` = this rank's env slice of ALL arms (rows padded to k * per) -> `[per, K, k * E_loc]` = ALL
        envs (rank order = env order) of this rank's `per` arms: one all-to-all of `[per, K, E_loc]` blocks."""
        torch.distributed.all_to_all_single(recv, send, group=group)
        E_loc = send.shape[2]
        # recv block j = this rank's arm rows as simulated by rank j: [k, per, K, E_loc] -> [per, K, k, E_loc]
        return recv.view(k, per, -1, E_loc).permute(1, 2, 0, 3).reshape(per, -1, k * E_loc)

    def epoch(self, want_stats=True, keep_buffers=False):
        if keep_buffers:
            raise ValueError("the env-sharded trainer keeps the trajectory on chip")
        ep, N, E, T, k = self.epoch_idx, self.N, self.E, self.T, self.eworld
        s0 = self._s0()
        lamb, h1 = ops.lambda_forward(self.lam_params, s0, N, 0, 1.0, want_h1=True)      # all arms are local
        args = (self.domain, self.net, T, self.Wpi, self.Wv, lamb, self.nature, self.ranges, self.R, self.C, self.gamma,
                self.lam_other, self.seed, self.t_global, self.s_traj[:, 0], self.stats, self.arm_stats, self.env_sums, self.ws,
                self.last_val, self.env0)
        ops.epoch_table_envshard(*args, mom_out=self.mom)
        self._ar(self.mom)
        cnt = float(T) * float(self.E_total)
        mean = self.mom[:, 0] / cnt
        var = (self.mom[:, 1] / cnt - mean * mean).clamp_min(0.0)
        self.mom_in[:, 0], self.mom_in[:, 1] = mean.float(), var.sqrt().float()
        ops.epoch_table_envshard(*args, mom_in=self.mom_in)
        as2 = self.arm_stats[:, 2].contiguous()
        self._ar(as2)
        self.arm_stats[:, 2] = as2
        # ---- statistics: env-sharded -> arm-sharded ----
        if k > 1:
            self.stats_send[:N] = self.stats
            st_arm = self.reshard_env_to_arm(self.stats_send, self.stats_recv, k, self.per, self.epg)[:self.a1 - self.a0].contiguous()
            lam_all = torch.empty((self.E_total,), dtype=torch.float32, device=self.dev)
            torch.distributed.all_gather_into_tensor(lam_all, lamb.contiguous(), group=self.epg)
        else:
            st_arm, lam_all = self.stats, lamb
        hp = ops.hyper(self.clip_ratio, self.entropy_coeff(ep), self.pi_lr, self.vf_lr, self.pi_iters, self.v_iters,
                       self.steps_pi, self.steps_v)
        sl = slice(self.a0, self.a1)
        if self.a1 > self.a0:
            wpi, wv = self.Wpi[sl].contiguous(), self.Wv[sl].contiguous()
            ops.ppo_update_table(self.net_upd, hp, T, wpi, wv, self.adam_pi[sl], self.adam_v[sl], st_arm,
                                 self.arm_stats[sl].contiguous(), lam_all)
        else:
            wpi, wv = self.Wpi[sl], self.Wv[sl]
        self.steps_pi += self.pi_iters
        self.steps_v += self.v_iters
        self.Wpi.copy_(self._gather_rows(wpi))
        self.Wv.copy_(self._gather_rows(wv))
        # ---- lambda-net SGD step (compute_loss_lambda, agent_oracle.py:360-376): mean over ALL envs = average of the ranks' ----
        if self.lambda_update_due(ep):
            coef = self.B / (1.0 - self.gamma) - self.env_sums[1]
            grad = ops.lambda_sgd(self.lam_params, s0, N, 0, 1.0, h1, coef.contiguous(), 0.0, want_grad=True)
            self._ar(grad)
            self.lam_params.add_(grad, alpha=-self.lm_lr / k)
        stats = None
        if want_stats:
            ret_all = torch.empty((self.E_total,), dtype=torch.float32, device=self.dev) if k > 1 else None
            if k > 1:
                torch.distributed.all_gather_into_tensor(ret_all, self.env_sums[0].contiguous(), group=self.epg)
            stats = {"EpActualRet": ret_all if k > 1 else self.env_sums[0].clone(), "Lamb": lam_all}
        self.epoch_idx += 1
        self.t_global += T
        self.s_traj[:, 0] = ops.reset_states(N, E, self.S, ops.rng(self.seed, _lib.STREAM_RESET, self.epoch_idx, env0=self.env0))
        return stats
