"""Generate tests/golden/*.npz by running the UNMODIFIED reference classes (test infrastructure).

Run once in the build container (where /root/reference exists):
    python -m oracle.gen_golden
The fixtures are committed; the GPU box only reads the .npz files.  torch-version dependent parts
(default nn.Linear init under torch.manual_seed) are captured as data (packed weights), so the
fixtures stay valid across torch versions.
"""
import os
import sys

import numpy as np

from . import nets, philox, ref_shims

OUT = os.path.join(os.path.dirname(ref_shims.HERE), "tests", "golden")
F32 = np.float32


def _f32_in(rs, lo, hi, shape):
    """fp32-representable float64 samples (so fp64 reference arithmetic on them rounds like fp32)."""
    return rs.uniform(lo, hi, size=shape).astype(F32).astype(np.float64)


def gen_env():
    from robust_rmab.environments.bandit_env_robust import (
        ARMMANRobustEnv, CounterExampleRobustEnv, SISRobustEnv)
    rs = np.random.RandomState(123)
    out = {}
    K = 60

    # ---- counterexample: nature params drawn wider than the range to exercise the clamp ----
    N = 6
    env = CounterExampleRobustEnv(N, 2.0, 0)
    env.sampled_parameter_ranges = env.sample_parameter_ranges()
    step_src = ref_shims.UniformSource(11, philox.STREAM_ENV, N)
    reset_src = ref_shims.UniformSource(11, philox.STREAM_RESET, N)
    env.random_stream = ref_shims.InjectedStream(step_src, reset_src, env.S)
    s0 = env.reset().reshape(-1).copy()
    S_, A_, P_, SN_, R_ = [], [], [], [], []
    with ref_shims.quiet():
        for t in range(K):
            s = env.current_full_state.copy()
            a = rs.randint(0, 2, size=N)
            p = _f32_in(rs, -0.1, 1.1, N)
            sn, r, _, _ = env.step(a, p)
            S_.append(s); A_.append(a); P_.append(p); SN_.append(sn.reshape(-1)); R_.append(r)
    out.update(ce_s=np.array(S_), ce_a=np.array(A_), ce_nature=np.array(P_), ce_s_next=np.array(SN_),
               ce_r=np.array(R_), ce_u=np.array(step_src.log, F32).reshape(K, N), ce_s0=s0,
               ce_ranges=env.sampled_parameter_ranges, ce_R=env.R.astype(np.float64),
               ce_PARAMETER_RANGES=env.PARAMETER_RANGES)
    raw = rs.randn(N).astype(F32)
    out.update(ce_bound_raw=raw, ce_bound_out=env.bound_nature_actions(raw, state=s0, reshape=True))

    # ---- ARMMAN ----
    N = 10
    env = ARMMANRobustEnv(N, 2.0, 0)
    out["armman_PARAMETER_RANGES"] = env.PARAMETER_RANGES
    env.sampled_parameter_ranges = env.sample_parameter_ranges().astype(F32).astype(np.float64)
    step_src = ref_shims.UniformSource(12, philox.STREAM_ENV, N)
    reset_src = ref_shims.UniformSource(12, philox.STREAM_RESET, N)
    env.random_stream = ref_shims.InjectedStream(step_src, reset_src, env.S)
    s0 = env.reset().reshape(-1).copy()
    S_, A_, P_, SN_, R_ = [], [], [], [], []
    with ref_shims.quiet():
        for t in range(K):
            s = env.current_full_state.copy()
            a = rs.randint(0, 2, size=N)
            p = _f32_in(rs, -0.1, 1.1, (N, 2))
            sn, r, _, _ = env.step(a, p)
            S_.append(s); A_.append(a); P_.append(p); SN_.append(sn.reshape(-1)); R_.append(r)
    out.update(armman_s=np.array(S_), armman_a=np.array(A_), armman_nature=np.array(P_),
               armman_s_next=np.array(SN_), armman_r=np.array(R_),
               armman_u=np.array(step_src.log, F32).reshape(K, N), armman_s0=s0,
               armman_ranges=env.sampled_parameter_ranges, armman_R=env.R.astype(np.float64))
    raw = rs.randn(N * 2).astype(F32)
    out.update(armman_bound_raw=raw, armman_bound_state=s0,
               armman_bound_out=env.bound_nature_actions(raw, state=s0, reshape=True))

    # ---- SIS ----
    N, pop = 5, 10
    env = SISRobustEnv(N, 2.0, pop, 0)
    env.sampled_parameter_ranges = env.sample_parameter_ranges().astype(F32).astype(np.float64)
    step_src = ref_shims.UniformSource(13, philox.STREAM_ENV, N)
    reset_src = ref_shims.UniformSource(13, philox.STREAM_RESET, N)
    env.random_stream = ref_shims.InjectedStream(step_src, reset_src, env.S)
    s0 = env.reset().reshape(-1).copy()
    lo, hi = env.sampled_parameter_ranges[..., 0], env.sampled_parameter_ranges[..., 1]
    S_, A_, P_, SN_, R_ = [], [], [], [], []
    for t in range(K):
        s = env.current_full_state.copy()
        a = rs.randint(0, 3, size=N)
        w = rs.uniform(0, 1, size=(N, 4))
        p = np.clip((lo + w * (hi - lo)).astype(F32).astype(np.float64), lo, hi)
        sn, r, _, _ = env.step(a, p)
        S_.append(s); A_.append(a); P_.append(p); SN_.append(sn.reshape(-1)); R_.append(r)
    out.update(sis_s=np.array(S_), sis_a=np.array(A_), sis_nature=np.array(P_), sis_s_next=np.array(SN_),
               sis_r=np.array(R_), sis_u=np.array(step_src.log, F32).reshape(K, N), sis_s0=s0,
               sis_ranges=env.sampled_parameter_ranges, sis_R=env.R.astype(np.float64), sis_pop=pop)
    # full pmf table for the last parameter setting (compute_distro for every (arm, s, a))
    out["sis_T"] = env.get_T_for_a_nature(P_[-1])
    raw = rs.randn(N * 4).astype(F32)
    out.update(sis_bound_raw=raw, sis_bound_out=env.bound_nature_actions(raw, state=s0, reshape=True))
    # a bigger pmf table (pop 50, the config-4 size) for two arms
    env50 = SISRobustEnv(2, 1.0, 50, 0)
    env50.sampled_parameter_ranges = env50.sample_parameter_ranges()
    p50 = env50.sampled_parameter_ranges.mean(axis=-1).astype(F32).astype(np.float64)
    out.update(sis50_params=p50, sis50_T=env50.get_T_for_a_nature(p50))
    np.savez_compressed(os.path.join(OUT, "env.npz"), **out)


def _pack_ac(ac, N):
    Wpi = np.stack([nets.pack_torch_mlp(ac.pi_list[i].logits_net) for i in range(N)])
    Wv = np.stack([nets.pack_torch_mlp(ac.v_list[i].v_net) for i in range(N)])
    return Wpi, Wv


def _pack_lambda(ac):
    import torch
    return [m_.detach().numpy().copy() for m in ac.lambda_net.lambda_net if isinstance(m, torch.nn.Linear)
            for m_ in (m.weight, m.bias)]


def gen_ac():
    import torch
    from robust_rmab.algos.rmabppo.rmabppo_core import MLPActorCriticRMAB
    out = {}
    rs = np.random.RandomState(5)
    cases = [("oh3", 5, 3, 2, 16, True, None, [0, 1]), ("oh2", 6, 2, 2, 8, True, None, [0, 1]),
             ("sis", 4, 11, 3, 16, False, 10, [0, 1, 2]), ("oh3h64", 3, 3, 2, 64, True, None, [0, 1])]
    for name, N, S, A, h, one_hot, norm, C in cases:
        torch.manual_seed(0)
        ac = MLPActorCriticRMAB(np.arange(S), np.arange(A), hidden_sizes=[h, h], C=np.array(C), N=N, B=2.0,
                                one_hot_encode=one_hot, non_ohe_obs_dim=1, state_norm=norm)
        Wpi, Wv = _pack_ac(ac, N)
        K = 8
        src = ref_shims.UniformSource(21, philox.STREAM_ACT, N)
        obs = rs.randint(0, S, size=(K, N))
        lamb = rs.uniform(-0.5, 2.0, size=K).astype(F32)
        res = {k: [] for k in "a v logp p1".split()}
        with ref_shims.patch_categorical_sample(src):
            for k in range(K):
                a, v, logp, q, p1 = ac.step(torch.as_tensor(obs[k], dtype=torch.float32), torch.as_tensor(lamb[k]))
                res["a"].append(a); res["v"].append(v); res["logp"].append(logp); res["p1"].append(p1)
        out.update({f"{name}_Wpi": Wpi, f"{name}_Wv": Wv, f"{name}_obs": obs, f"{name}_lamb": lamb,
                    f"{name}_u": np.array(src.log, F32).reshape(K, N),
                    f"{name}_meta": np.array([N, S, A, h, int(one_hot), norm or 1])})
        out.update({f"{name}_{k}": np.array(v) for k, v in res.items()})
        lam_params = _pack_lambda(ac)
        for i, p in enumerate(lam_params):
            out[f"{name}_lam{i}"] = p
        out[f"{name}_lam_out"] = np.array([ac.get_lambda(obs[k].astype(np.float64)) for k in range(K)])
    np.savez_compressed(os.path.join(OUT, "ac.npz"), **out)


def gen_buffer():
    from robust_rmab.algos.rmabppo.agent_oracle import RMABPPO_Buffer
    rs = np.random.RandomState(9)
    out = {}
    for name, N, T, S, A, gamma, lam in [("a", 4, 12, 3, 2, 0.9, 0.97), ("b", 3, 30, 2, 2, 0.99, 0.95)]:
        buf = RMABPPO_Buffer(S, A, N, 'd', T, one_hot_encode=True, gamma=gamma, lam_OTHER=lam)
        obs = rs.randint(0, S, size=(T, N)); act = rs.randint(0, A, size=(T, N))
        rew = rs.uniform(0, 1, size=(T, N)); cost = act.astype(np.float64)
        val = rs.randn(T, N); logp = -rs.uniform(0.01, 2, size=(T, N))
        lamb = np.full(T, rs.uniform(0, 2)); last = rs.randn(N)
        for t in range(T):
            buf.store(obs[t], act[t], rew[t], cost[t], val[t], np.zeros(N), lamb[t], logp[t])
        buf.finish_path(last, np.zeros((50, N)))
        adv_raw = buf.adv_buf.copy()
        data = {k: v.numpy() for k, v in buf.get().items()}
        out.update({f"{name}_obs": obs, f"{name}_act": act, f"{name}_rew": rew, f"{name}_cost": cost,
                    f"{name}_val": val, f"{name}_logp": logp, f"{name}_lamb": lamb, f"{name}_last": last,
                    f"{name}_adv_raw": adv_raw, f"{name}_hyper": np.array([gamma, lam])})
        out.update({f"{name}_get_{k}": v for k, v in data.items()})
    np.savez_compressed(os.path.join(OUT, "buffer.npz"), **out)


class _FixedPi:
    """Stands in for pi_list[i]: `_distribution(x).probs` returns a fixed row."""

    def __init__(self, row):
        import torch
        self.row = torch.as_tensor(row, dtype=torch.float32)

    def _distribution(self, x):
        class D:
            probs = self.row
        return D


def gen_select():
    import torch
    from robust_rmab.algos.rmabppo.rmabppo_core import MLPActorCriticRMAB
    rs = np.random.RandomState(17)
    out = {}
    cases = []
    # the survey's known answers
    known = np.array([[.05, .05, .9], [.2, .7, .1], [.6, .3, .1], [.3, .4, .3]])
    for B in (1, 3, 4):
        cases.append((known, [0, 1, 2], B))
    for N, A, C in [(12, 3, [0, 1, 2]), (16, 3, [0, 1, 2]), (9, 2, [0, 1]), (40, 2, [0, 1]), (30, 3, [0, 1, 2])]:
        for B in (0, 1, 2, 3, 5, 8, 13, 100):
            p = rs.dirichlet(np.ones(A) * rs.choice([0.3, 1.0, 3.0]), size=N)
            cases.append((p, C, B))
    for ci, (p, C, B) in enumerate(cases):
        N, A = p.shape
        torch.manual_seed(0)
        ac = MLPActorCriticRMAB(np.arange(A), np.arange(A), hidden_sizes=[4, 4], C=np.array(C), N=N, B=B)
        p32 = p.astype(F32)
        for i in range(N):
            ac.pi_list[i] = _FixedPi(p32[i])
        obs = np.zeros(N)
        multi = ac.act_test_deterministic_multiaction(obs)
        out[f"c{ci}_probs"] = p32
        out[f"c{ci}_C"] = np.array(C)
        out[f"c{ci}_B"] = np.array(B)
        out[f"c{ci}_multi"] = multi
        if A == 2:
            out[f"c{ci}_binary"] = ac.act_test_deterministic(obs)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "select.npz"), **out)


def gen_vi():
    import mdptoolbox.mdp
    import mdptoolbox.example
    from robust_rmab.environments.bandit_env_robust import (
        ARMMANRobustEnv, CounterExampleRobustEnv, SISRobustEnv)
    out = {}
    # docstring known answers (mdp.py:1248-1290) -- valid for the span ('fast') rule
    P, R = mdptoolbox.example.forest()
    vi = mdptoolbox.mdp.ValueIteration(P, R, 0.96, stop_criterion='fast'); vi.run()
    out.update(forest_P=P, forest_R=R, forest_V_fast=np.array(vi.V), forest_pol_fast=np.array(vi.policy),
               forest_iter_fast=np.array(vi.iter))
    vi = mdptoolbox.mdp.ValueIteration(P, R, 0.96); vi.run()
    out.update(forest_V_full=np.array(vi.V), forest_pol_full=np.array(vi.policy), forest_iter_full=np.array(vi.iter))

    def solve_all(T, R, C, lambdas, gamma, eps):
        N, S, A, _ = T.shape
        L = len(lambdas)
        V = np.full((N, L, S), np.nan); pol = np.zeros((N, L, S), np.int64)
        it = np.full((N, L), -1); mi = np.full((N, L), -1)
        for i in range(N):
            for l, lam in enumerate(lambdas):
                T_i = np.swapaxes(T[i], 0, 1)                      # simulator.py:506
                R_i = np.zeros(T_i.shape)
                for x in range(R_i.shape[0]):
                    for y in range(R_i.shape[1]):
                        R_i[x, :, y] = R[i] - lam * C[x]            # :508-510 with the lambda charge
                try:
                    m = mdptoolbox.mdp.ValueIteration(T_i, R_i, discount=gamma, stop_criterion='full', epsilon=eps)
                    m.run()
                except (ZeroDivisionError, ValueError, OverflowError):
                    continue
                V[i, l] = m.V; pol[i, l] = m.policy; it[i, l] = m.iter; mi[i, l] = m.max_iter
        return V, pol, it, mi

    lambdas = np.linspace(0, 10, 9)
    env = CounterExampleRobustEnv(6, 2.0, 0)
    env.sampled_parameter_ranges = env.sample_parameter_ranges()
    T = env.get_T_for_a_nature(env.sampled_parameter_ranges.mean(axis=-1))
    for eps in (1e-1, 1e-8):
        V, pol, it, mi = solve_all(T, env.R.astype(float), env.C, lambdas, 0.9, eps)
        out.update({f"ce_V_{eps}": V, f"ce_pol_{eps}": pol, f"ce_it_{eps}": it, f"ce_mi_{eps}": mi})
    out.update(ce_T=T, ce_R=env.R.astype(float), ce_C=env.C, lambdas=lambdas)

    env = ARMMANRobustEnv(10, 2.0, 0)
    env.sampled_parameter_ranges = env.sample_parameter_ranges()
    T = env.get_T_for_a_nature(env.sampled_parameter_ranges.mean(axis=-1))
    for eps in (1e-4, 1e-8):
        V, pol, it, mi = solve_all(T, env.R.astype(float), env.C, lambdas, 0.9, eps)
        out.update({f"armman_V_{eps}": V, f"armman_pol_{eps}": pol, f"armman_it_{eps}": it, f"armman_mi_{eps}": mi})
    out.update(armman_T=T, armman_R=env.R.astype(float), armman_C=env.C)

    env = SISRobustEnv(3, 2.0, 10, 0)
    env.sampled_parameter_ranges = env.sample_parameter_ranges()
    T = env.get_T_for_a_nature(env.sampled_parameter_ranges.mean(axis=-1))
    V, pol, it, mi = solve_all(T, env.R.astype(float), env.C, lambdas[:5], 0.9, 1e-4)
    out.update(sis_T=T, sis_R=env.R.astype(float), sis_C=env.C, sis_V=V, sis_pol=pol, sis_it=it, sis_mi=mi)
    np.savez_compressed(os.path.join(OUT, "vi.npz"), **out)


def gen_baselines():
    """Nature baseline policies (robust_rmab/baselines/nature_baselines_*.py): their answers at every constant state
    vector and at a mixed one, for ranges of each domain's shape -- what baselines.nature_param_table must reproduce."""
    from robust_rmab.baselines import nature_baselines_armman as nb_a
    from robust_rmab.baselines import nature_baselines_counterexample as nb_c
    from robust_rmab.baselines import nature_baselines_sis as nb_s
    rs = np.random.RandomState(31)
    out = {}
    N, S, A = 7, 3, 2
    shapes = dict(ce=(N, 2), armman=(N, S, A, 2), sis=(N, 4, 2))
    mods = dict(ce=nb_c, armman=nb_a, sis=nb_s)
    for dom, shp in shapes.items():
        ranges = np.sort(rs.rand(*shp), axis=-1)
        out[f"{dom}_ranges"] = ranges
        pert = rs.rand(*shp[:-1])
        out[f"{dom}_pert"] = pert
        m = mods[dom]
        pols = dict(pess=m.PessimisticNaturePolicy(ranges, 0), opt=m.OptimisticNaturePolicy(ranges, 1),
                    mid=m.MiddleNaturePolicy(ranges, 2), midp=m.MiddleNaturePolicy(ranges, 3, perturbations=pert,
                                                                                   perturbation_size=0.2),
                    samp=m.SampledRandomNaturePolicy(ranges, 4))
        pols["samp"].sample_param_setting(5)
        n_s = S if dom == "armman" else 2
        mixed = np.arange(N) % n_s
        for k, pol in pols.items():
            with ref_shims.quiet():
                rows = np.stack([np.asarray(pol.get_nature_action(np.full(N, s_, dtype=int)), dtype=np.float64)
                                 for s_ in range(n_s)])
                out[f"{dom}_{k}_const"] = rows
                out[f"{dom}_{k}_mixed"] = np.asarray(pol.get_nature_action(mixed), dtype=np.float64)
            out[f"{dom}_{k}_repr"] = np.array(repr(pol))
    np.savez_compressed(os.path.join(OUT, "baselines.npz"), **out)


def main(which=None):
    ref_shims.import_reference()
    os.makedirs(OUT, exist_ok=True)
    gens = dict(env=gen_env, ac=gen_ac, buffer=gen_buffer, select=gen_select, vi=gen_vi, baselines=gen_baselines)
    try:
        from .gen_golden_train import gen_train, gen_ma, gen_ma_train, gen_train_sis, gen_ma_train_sis
        gens.update(train=gen_train, ma=gen_ma, ma_train=gen_ma_train, train_sis=gen_train_sis, ma_train_sis=gen_ma_train_sis)
    except ImportError:
        pass
    for k, f in gens.items():
        if which and k not in which:
            continue
        f()
        print("wrote", k)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
