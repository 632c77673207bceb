"""Golden fixture for the whole RMABPPO epoch loop: runs the UNMODIFIED reference
`AgentOracle.best_response` (robust_rmab/algos/rmabppo/agent_oracle.py:215-593) on a tiny counterexample /
ARMMAN problem with injected randomness and records initial weights, final weights and the logged per-epoch
values, so that oracle/train.py (and through it the CUDA trainer) is pinned to the reference's own training loop.

    python -m oracle.gen_golden train
"""
import os
import shutil
import tempfile

import numpy as np

from . import nets, philox, ref_shims

OUT = os.path.join(os.path.dirname(ref_shims.HERE), "tests", "golden")
F32 = np.float32


class _ActSource(ref_shims.UniformSource):
    """Uniforms for Categorical.sample keyed like the oracle: the T rollout steps of epoch e use stream ACT with
    t = e*T + step; the extra bootstrap ac.step at the end of every epoch (agent_oracle.py:553) uses t = e*T + T."""

    def __init__(self, seed, n_arms, T):
        super().__init__(seed, philox.STREAM_ACT, n_arms)
        self.T, self.sweep = T, 0

    def next(self):
        k = self.sweep % (self.T + 1)
        ep = self.sweep // (self.T + 1)
        # k == T is the bootstrap step: it draws with the time counter of the next epoch's first step (the device
        # path and oracle/*train.py do the same; the draw only matters for MA-RMABPPO's nature critic input)
        u = philox.uniform(self.seed, self.stream, ep * self.T + k, 0, np.uint32(self.i))
        self.i += 1
        if self.i == self.n:
            self.i, self.sweep = 0, self.sweep + 1
        return np.float32(u)


_REC = {}


def _recorder_class(core):
    """Subclass of the reference actor-critic that records its initial weights (module-level so that the
    reference's logger can pickle the instance, utils/logx.py:253-275)."""
    if "Recorder" not in globals():
        class Recorder(core.MLPActorCriticRMAB):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                _REC["Wpi0"] = np.stack([nets.pack_torch_mlp(self.pi_list[i].logits_net) for i in range(self.N)])
                _REC["Wv0"] = np.stack([nets.pack_torch_mlp(self.v_list[i].v_net) for i in range(self.N)])
                _REC["lam0"] = [p.detach().numpy().copy() for p in self.lambda_net.parameters()]
        Recorder.__qualname__ = "Recorder"
        globals()["Recorder"] = Recorder
    return globals()["Recorder"]


def _run(data, N, B, T, epochs, hidden, seed, kw, pop=0):
    import torch
    from robust_rmab.algos.rmabppo import agent_oracle as ao
    from robust_rmab.algos.rmabppo import rmabppo_core as core
    A, extra = 2, {}
    if data == "counterexample":
        from robust_rmab.environments.bandit_env_robust import CounterExampleRobustEnv as Env
        from robust_rmab.baselines.nature_baselines_counterexample import SampledRandomNaturePolicy
        S = 2
    elif data == "armman":
        from robust_rmab.environments.bandit_env_robust import ARMMANRobustEnv as Env
        from robust_rmab.baselines.nature_baselines_armman import SampledRandomNaturePolicy
        S = 3
    else:   # SIS: S = pop + 1 states, 3 actions, scalar state input s / state_norm (agent_oracle.py:775-785)
        from robust_rmab.environments.bandit_env_robust import SISRobustEnv
        from robust_rmab.baselines.nature_baselines_sis import SampledRandomNaturePolicy
        Env = lambda N_, B_, seed_: SISRobustEnv(N_, B_, pop, seed_)
        S, A = pop + 1, 3
        extra = dict(pop_size=pop, one_hot_encode=False, non_ohe_obs_dim=1, state_norm=pop)
    rec = _REC
    rec.clear()
    Recorder = _recorder_class(core)

    home = tempfile.mkdtemp(prefix="rmab_golden_")
    try:
        env0 = Env(N, B, seed)
        ranges = env0.sample_parameter_ranges().astype(F32).astype(np.float64)
        agent_kwargs = dict(steps_per_epoch=T, epochs=epochs, ac_kwargs=dict(hidden_sizes=[hidden, hidden]),
                            actor_critic=Recorder, **kw)
        orc = ao.AgentOracle(data, N, S, A, B, seed, 1, agent_kwargs=agent_kwargs, home_dir=home, exp_name="golden",
                             sampled_nature_parameter_ranges=ranges, robust_keyword="sample_random", **extra)
        step_src = ref_shims.UniformSource(seed, philox.STREAM_ENV, N)
        reset_src = ref_shims.UniformSource(seed, philox.STREAM_RESET, N)
        orc.env.random_stream = ref_shims.InjectedStream(step_src, reset_src, S)
        nature = SampledRandomNaturePolicy(ranges, 0)
        nature.sample_param_setting(seed)
        nature.param_setting = nature.param_setting.astype(F32).astype(np.float64)
        torch.manual_seed(seed)
        np.random.seed(seed)
        with ref_shims.patch_categorical_sample(_ActSource(seed, N, T)), ref_shims.quiet():
            ac = orc.best_response([nature], [1.0], 0)
        prog = None
        for dp, _, fs in os.walk(home):
            if "progress.txt" in fs:
                prog = os.path.join(dp, "progress.txt")
        rows = [l.rstrip("\n").split("\t") for l in open(prog)]
        cols = {c: np.array([float(r[i]) for r in rows[1:]]) for i, c in enumerate(rows[0])}
        out = dict(meta=np.array([N, S, T, epochs, hidden, seed]), B=np.array(B), ranges=ranges, pop=np.array(pop),
                   nature=nature.param_setting, Wpi0=rec["Wpi0"], Wv0=rec["Wv0"],
                   Wpi1=np.stack([nets.pack_torch_mlp(ac.pi_list[i].logits_net) for i in range(N)]),
                   Wv1=np.stack([nets.pack_torch_mlp(ac.v_list[i].v_net) for i in range(N)]),
                   AverageLamb=cols["Lamb"], AverageEpActualRet=cols["AverageEpActualRet"],
                   AverageVVals=cols["AverageVVals"])
        for i, p in enumerate(rec["lam0"]):
            out[f"lam0_{i}"] = p
        for i, p in enumerate(ac.lambda_net.parameters()):
            out[f"lam1_{i}"] = p.detach().numpy().copy()
        out["hyper_names"] = np.array(sorted(kw))
        out["hyper_values"] = np.array([float(kw[k]) for k in sorted(kw)])
        return out
    finally:
        shutil.rmtree(home, ignore_errors=True)


KW = dict(gamma=0.9, clip_ratio=2.0, pi_lr=2e-3, vf_lr=2e-3, lm_lr=2e-3, train_pi_iters=5, train_v_iters=5,
          lam_OTHER=0.97, start_entropy_coeff=0.5, end_entropy_coeff=0.0, lamb_update_freq=2, final_train_lambdas=1)


def gen_train():
    out = {}
    for name, data, N, B, T, epochs, hidden in [("ce", "counterexample", 6, 2.0, 12, 6, 16), ("armman", "armman", 5, 2.0, 9, 5, 8)]:
        r = _run(data, N, B, T, epochs, hidden, 3, KW)
        out.update({f"{name}_{k}": v for k, v in r.items()})
    np.savez_compressed(os.path.join(OUT, "train.npz"), **out)


def gen_train_sis():
    """AgentOracle.best_response(data='sis') -- S = pop + 1, A = 3, C = [0,1,2], non-one-hot scalar inputs -- at hidden 16
    and 64 (the widths of the per-sample update kernels): tests/golden/train_sis.npz."""
    out = {}
    for name, N, B, T, epochs, hidden, pop in [("h16", 5, 3.0, 8, 5, 16, 10), ("h64", 4, 2.0, 7, 4, 64, 20)]:
        r = _run("sis", N, B, T, epochs, hidden, 5, KW, pop=pop)
        out.update({f"{name}_{k}": v for k, v in r.items()})
    np.savez_compressed(os.path.join(OUT, "train_sis.npz"), **out)


def gen_ma():
    """RMABLambdaNatureOracle.step (ma_rmabppo_core.py:274-331) on the unmodified reference with injected
    Categorical / Normal draws, for the counterexample (D = N) and ARMMAN (D = 2N, state-dependent bounds) domains."""
    import contextlib
    import torch
    from torch.distributions.normal import Normal
    from robust_rmab.algos.ma_rmabppo.ma_rmabppo_core import RMABLambdaNatureOracle
    from robust_rmab.environments.bandit_env_robust import ARMMANRobustEnv, CounterExampleRobustEnv

    @contextlib.contextmanager
    def patch_normal(eps_rows):
        orig = Normal.sample
        it = iter(eps_rows)

        def sample(self, sample_shape=torch.Size()):
            return self.loc + self.scale * torch.as_tensor(next(it))
        Normal.sample = sample
        try:
            yield
        finally:
            Normal.sample = orig

    def lin(seq):
        return [m for m in seq if isinstance(m, torch.nn.Linear)]

    out = {}
    rs = np.random.RandomState(31)
    for name, Env, N, S, h in [("ce", CounterExampleRobustEnv, 6, 2, 16), ("armman", ARMMANRobustEnv, 5, 3, 8)]:
        env = Env(N, 2.0, 0)
        env.sampled_parameter_ranges = env.sample_parameter_ranges().astype(F32).astype(np.float64)
        D = env.action_dim_nature
        torch.manual_seed(1)
        ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, D, env,
                                    hidden_sizes=[h, h], C=env.C, N=N, B=env.B)
        in_dim = S + 1
        K = 6
        obs = rs.randint(0, S, size=(K, N))
        lamb = rs.uniform(0, 2, size=K).astype(F32)
        eps = rs.randn(K, D).astype(F32)
        src = ref_shims.UniformSource(41, philox.STREAM_ACT, N)
        res = {k: [] for k in "a v logp an vn lpn p1 anb".split()}
        with ref_shims.patch_categorical_sample(src), patch_normal(list(eps)):
            for k in range(K):
                o = torch.as_tensor(obs[k], dtype=torch.float32)
                a, v, logp, q, an, vn, lpn, qn, p1 = ac.step(o, torch.as_tensor(lamb[k]))
                res["a"].append(a); res["v"].append(v); res["logp"].append(logp); res["an"].append(an)
                res["vn"].append(vn); res["lpn"].append(lpn); res["p1"].append(p1)
                res["anb"].append(np.asarray(env.bound_nature_actions(an, state=obs[k], reshape=False)).reshape(-1))
        Wv_full = [lin(ac.v_list_agent[i].v_net) for i in range(N)]
        first = np.stack([l[0].weight.detach().numpy() for l in Wv_full])          # [N,h,in+D]
        packed_v = []
        for i, l in enumerate(Wv_full):
            parts = [first[i][:, :in_dim].reshape(-1), l[0].bias.detach().numpy()]
            for m in l[1:]:
                parts += [m.weight.detach().numpy().reshape(-1), m.bias.detach().numpy()]
            packed_v.append(np.concatenate(parts))
        out.update({f"{name}_meta": np.array([N, S, h, D]), f"{name}_obs": obs, f"{name}_lamb": lamb, f"{name}_eps": eps,
                    f"{name}_u": np.array(src.log, F32).reshape(K, N), f"{name}_ranges": env.sampled_parameter_ranges,
                    f"{name}_Wpi": np.stack([nets.pack_torch_mlp(ac.pi_list_agent[i].logits_net) for i in range(N)]),
                    f"{name}_Wv": np.stack(packed_v).astype(F32), f"{name}_W1n": first[:, :, in_dim:].astype(F32),
                    f"{name}_log_std": ac.pi_nature.log_std.detach().numpy()})
        for tag, seq in (("mu", ac.pi_nature.mu_net), ("vn", ac.v_nature.v_net)):
            for j, m in enumerate(lin(seq)):
                out[f"{name}_{tag}_W{j}"] = m.weight.detach().numpy()
                out[f"{name}_{tag}_b{j}"] = m.bias.detach().numpy()
        out.update({f"{name}_out_{k}": np.array(v) for k, v in res.items()})
    np.savez_compressed(os.path.join(OUT, "ma.npz"), **out)


class _MAEnvStream:
    """`env.random_stream` shim for the nature oracle's rollout: every time step makes 51 env.step calls -- the real
    one (stream ENV, t = step) and NUM_TEST_POLICY_RUNS = 50 counterfactual ones (stream CF, t = step*50 + trial) --
    of N multinomial draws each; resets draw from stream RESET (t = reset index)."""

    def __init__(self, seed, n_arms, n_states, trials=50):
        self.seed, self.N, self.S, self.trials = seed, n_arms, n_states, trials
        self.calls, self.resets = 0, 0

    def seed(self, seed=None):
        return None

    def multinomial(self, n, p):
        step, within = divmod(self.calls, self.N * (1 + self.trials))
        block, arm = divmod(within, self.N)
        self.calls += 1
        if block == 0:
            u = philox.uniform(self.seed, philox.STREAM_ENV, step, 0, np.uint32(arm))
        else:
            u = philox.uniform(self.seed, philox.STREAM_CF, step * self.trials + block - 1, 0, np.uint32(arm))
        k = int(philox.categorical(np.asarray(p, dtype=np.float32), np.float32(u)))
        out = np.zeros(len(p), dtype=np.int64)
        out[k] = 1
        return out

    def _reset_states(self, size):
        u = philox.uniform(self.seed, philox.STREAM_RESET, self.resets, 0, np.arange(size, dtype=np.uint32))
        self.resets += 1
        return np.minimum((u * np.float32(self.S)).astype(np.int64), self.S - 1)

    def choice(self, space, size):
        return np.asarray(space)[self._reset_states(size)]

    def randint(self, low=0, high=None, size=None):
        return low + self._reset_states(size)


_MA_REC = {}


def _ma_recorder_class(core):
    if "MARecorder" not in globals():
        import torch

        class MARecorder(core.RMABLambdaNatureOracle):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                lin = lambda seq: [m for m in seq if isinstance(m, torch.nn.Linear)]
                N, in_dim = self.N, self.obs_dim + 1
                _MA_REC["Wpi0"] = np.stack([nets.pack_torch_mlp(self.pi_list_agent[i].logits_net) for i in range(N)])
                _MA_REC["Wv0"], _MA_REC["W1n0"] = _pack_ma_critics(self, in_dim)
                _MA_REC["lam0"] = [p.detach().numpy().copy() for p in self.lambda_net.parameters()]
                _MA_REC["mu0"] = [t.detach().numpy().copy() for m in lin(self.pi_nature.mu_net) for t in (m.weight, m.bias)]
                _MA_REC["vn0"] = [t.detach().numpy().copy() for m in lin(self.v_nature.v_net) for t in (m.weight, m.bias)]
                _MA_REC["log_std0"] = self.pi_nature.log_std.detach().numpy().copy()
        MARecorder.__qualname__ = "MARecorder"
        globals()["MARecorder"] = MARecorder
    return globals()["MARecorder"]


def _pack_ma_critics(ac, in_dim):
    import torch
    packed, w1n = [], []
    for i in range(ac.N):
        l = [m for m in ac.v_list_agent[i].v_net if isinstance(m, torch.nn.Linear)]
        first = l[0].weight.detach().numpy()
        parts = [first[:, :in_dim].reshape(-1), l[0].bias.detach().numpy()]
        for m in l[1:]:
            parts += [m.weight.detach().numpy().reshape(-1), m.bias.detach().numpy()]
        packed.append(np.concatenate(parts))
        w1n.append(first[:, in_dim:].copy())
    return np.stack(packed).astype(F32), np.stack(w1n).astype(F32)


def gen_ma_train_sis():
    """NatureOracle.best_response on SISRobustEnv (BASELINE configs[3]'s domain: S = pop + 1, A = 3, D = 4N nature
    parameters, scalar state inputs) at hidden 64: tests/golden/ma_train_sis.npz."""
    gen_ma_train(epochs=5, out_name="ma_train_sis.npz", data="sis", N=6, pop=10, h=64, T=6, B=3.0, seed=6)


def gen_ma_train(epochs=5, out_name="ma_train.npz", final_train_lambdas=1, data="counterexample", N=6, pop=0, h=16, T=7,
                 B=2.0, seed=4):
    """The UNMODIFIED reference NatureOracle.best_response (nature_oracle.py:252-806) against a Pessimist agent on a
    tiny counterexample problem with every random draw injected from the Philox convention of the CUDA path."""
    import contextlib
    import torch
    from torch.distributions.normal import Normal
    from robust_rmab.algos.ma_rmabppo import ma_rmabppo_core as core
    from robust_rmab.algos.ma_rmabppo import nature_oracle as no
    from robust_rmab.baselines.agent_baselines import PessimisticAgentPolicy
    from robust_rmab.environments.bandit_env_robust import CounterExampleRobustEnv, SISRobustEnv
    from . import ma as oma
    sis = data == "sis"
    S, A = (pop + 1, 3) if sis else (2, 2)
    extra = dict(pop_size=pop, one_hot_encode=False, non_ohe_obs_dim=1, state_norm=pop, nature_state_norm=1) if sis else {}
    kw = dict(gamma=0.9, clip_ratio=2.0, pi_lr_agent=2e-3, pi_lr_nature=2e-3, vf_lr_agent=2e-3, vf_lr_nature=2e-3, lm_lr=2e-3,
              train_pi_iters=3, train_v_iters=3, lam_OTHER=0.97, start_entropy_coeff=0.5, end_entropy_coeff=0.0,
              lamb_update_freq=2, final_train_lambdas=final_train_lambdas)
    env0 = SISRobustEnv(N, B, pop, seed) if sis else CounterExampleRobustEnv(N, B, seed)
    ranges = env0.sample_parameter_ranges().astype(F32).astype(np.float64)
    D = env0.action_dim_nature
    sweep = {"k": 0}

    @contextlib.contextmanager
    def patch_normal():
        orig = Normal.sample

        def sample(self, sample_shape=torch.Size()):
            k, ep = sweep["k"] % (T + 1), sweep["k"] // (T + 1)
            sweep["k"] += 1
            eps = oma.normal_eps(seed, ep * T + k, 1, D)      # k == T: the bootstrap step (t of the next epoch's first)
            return self.loc + self.scale * torch.as_tensor(eps[0])
        Normal.sample = sample
        try:
            yield
        finally:
            Normal.sample = orig

    home = tempfile.mkdtemp(prefix="rmab_golden_ma_")
    try:
        _MA_REC.clear()
        Rec = _ma_recorder_class(core)
        nature_kwargs = dict(steps_per_epoch=T, epochs=epochs, ac_kwargs=dict(hidden_sizes=[h, h]), ma_actor_critic=Rec, **kw)
        orc = no.NatureOracle(data, N, S, A, B, seed, 1, nature_kwargs=nature_kwargs, home_dir=home,
                              exp_name="golden_ma", sampled_nature_parameter_ranges=ranges, **extra)
        orc.env.random_stream = _MAEnvStream(seed, N, S)
        torch.manual_seed(seed)
        np.random.seed(seed)
        with ref_shims.patch_categorical_sample(_ActSource(seed, N, T)), patch_normal(), ref_shims.quiet():
            ac = orc.best_response([PessimisticAgentPolicy(N, 0)], [1.0], 0)
        prog = None
        for dp, _, fs in os.walk(home):
            if "progress.txt" in fs:
                prog = os.path.join(dp, "progress.txt")
        rows = [l.rstrip("\n").split("\t") for l in open(prog)]
        cols = {c: np.array([float(r[i]) for r in rows[1:]]) for i, c in enumerate(rows[0])}
        lin = lambda seq: [m for m in seq if isinstance(m, torch.nn.Linear)]
        out = dict(meta=np.array([N, S, T, epochs, h, seed, D]), B=np.array(B), ranges=ranges, pop=np.array(pop),
                   Wpi0=_MA_REC["Wpi0"], Wv0=_MA_REC["Wv0"], W1n0=_MA_REC["W1n0"], log_std0=_MA_REC["log_std0"],
                   Wpi1=np.stack([nets.pack_torch_mlp(ac.pi_list_agent[i].logits_net) for i in range(N)]),
                   log_std1=ac.pi_nature.log_std.detach().numpy().copy(),
                   Lamb=cols["Lamb"], EpActualRetAgent=cols["AverageEpActualRetAgent"],
                   EpLambAdjRetNature=cols["AverageEpLambAdjRetNature"], VVals_agent=cols["AverageVVals_agent"],
                   VVals_nature=cols["AverageVVals_nature"])
        out["Wv1"], out["W1n1"] = _pack_ma_critics(ac, ac.obs_dim + 1)
        for tag, lst in (("lam0", _MA_REC["lam0"]), ("mu0", _MA_REC["mu0"]), ("vn0", _MA_REC["vn0"])):
            for i, p in enumerate(lst):
                out[f"{tag}_{i}"] = p
        for i, p in enumerate(ac.lambda_net.parameters()):
            out[f"lam1_{i}"] = p.detach().numpy().copy()
        for i, t in enumerate([t for m in lin(ac.pi_nature.mu_net) for t in (m.weight, m.bias)]):
            out[f"mu1_{i}"] = t.detach().numpy().copy()
        for i, t in enumerate([t for m in lin(ac.v_nature.v_net) for t in (m.weight, m.bias)]):
            out[f"vn1_{i}"] = t.detach().numpy().copy()
        out["hyper_names"] = np.array(sorted(kw))
        out["hyper_values"] = np.array([float(kw[k]) for k in sorted(kw)])
        np.savez_compressed(os.path.join(OUT, out_name) if os.sep not in out_name else out_name, **out)
    finally:
        shutil.rmtree(home, ignore_errors=True)
