"""MA-RMABPPO rollout pieces on the device: Gaussian nature sample, agent critics with nature inputs, the drop-in
RMABLambdaNatureOracle.step against the reference golden, and the counterfactual reward estimate.  GPU only."""
import numpy as np
import pytest

from oracle import envs as oenv
from oracle import ma as oma
from oracle import nets as onets
from oracle import philox

pytestmark = pytest.mark.gpu
F32 = np.float32


@pytest.fixture(scope="module")
def K():
    import torch
    assert torch.cuda.is_available()
    from robustrmab_b200 import _lib, ops

    class NS:
        pass
    ns = NS()
    ns.torch, ns.ops, ns.lib = torch, ops, _lib
    ns.dev = torch.device("cuda:0")
    ns.t = lambda x, dt=None: torch.as_tensor(np.ascontiguousarray(x if dt is None else np.asarray(x).astype(dt)),
                                              device=ns.dev)
    ns.n = lambda x: x.detach().cpu().numpy()
    return ns


@pytest.mark.parametrize("E,D", [(1, 6), (7, 33), (64, 8192)])
def test_gaussian_sample(K, E, D):
    rs = np.random.RandomState(D)
    mu = rs.randn(E, D).astype(F32)
    log_std = rs.uniform(-1.5, 0.2, D).astype(F32)
    eps = rs.randn(E, D).astype(F32)
    a, lp = K.ops.gaussian_sample(K.t(mu), K.t(log_std), eps=K.t(eps))
    wa, wlp = oma.gaussian_sample(mu, log_std, eps)
    np.testing.assert_allclose(K.n(a), wa, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(K.n(lp), wlp, rtol=2e-05, atol=1e-05)
    # Philox path: same normals as the oracle's Box-Muller on the NATURE stream
    a2, lp2 = K.ops.gaussian_sample(K.t(mu), K.t(log_std), r=K.ops.rng(99, K.lib.STREAM_NATURE, 5, env0=2))
    weps = oma.normal_eps(99, 5, E, D, env0=2)
    wa2, wlp2 = oma.gaussian_sample(mu, log_std, weps)
    np.testing.assert_allclose(K.n(a2), wa2, rtol=1e-05, atol=2e-06)      # sincos / log ulps between libm and CUDA
    np.testing.assert_allclose(K.n(lp2), wlp2, rtol=1e-05, atol=5e-4 * max(1, D / 1000))
    if E * D > 10000:
        z = (K.n(a2) - mu) / np.exp(log_std)
        assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01


@pytest.mark.parametrize("name,domain", [("ce", 0), ("armman", 1)])
def test_ma_step_dropin_matches_reference(K, golden, name, domain):
    torch = K.torch
    from robustrmab_b200.environments import ARMMANRobustEnv, CounterExampleRobustEnv
    from robustrmab_b200.ma_core import RMABLambdaNatureOracle
    g = golden("ma")
    N, S, h, D = [int(x) for x in g[name + "_meta"]]
    env = (CounterExampleRobustEnv if domain == 0 else ARMMANRobustEnv)(N, 2.0, 0)
    env.sampled_parameter_ranges = g[name + "_ranges"]
    ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, D, env,
                                hidden_sizes=[h, h], C=env.C, N=N, B=env.B)
    assert repr(ac) == "MA-RMABPPO_0" and ac.act_dim_nature == D
    ac.Wpi.copy_(K.t(g[name + "_Wpi"])); ac.Wv.copy_(K.t(g[name + "_Wv"])); ac.W1n.copy_(K.t(g[name + "_W1n"]))
    with torch.no_grad():
        for tag, seq in (("mu", ac.pi_nature.mu_net), ("vn", ac.v_nature.v_net)):
            for j, m in enumerate([m for m in seq if isinstance(m, torch.nn.Linear)]):
                m.weight.copy_(K.t(g[f"{name}_{tag}_W{j}"])); m.bias.copy_(K.t(g[f"{name}_{tag}_b{j}"]))
        ac.pi_nature.log_std.copy_(K.t(g[name + "_log_std"]))
    for k in range(g[name + "_obs"].shape[0]):
        s = K.t(g[name + "_obs"][k][:, None], np.uint8)
        d = ac.step_device(s, K.t(g[name + "_lamb"][k:k + 1]), eps=K.t(g[name + "_eps"][k][None]),
                           u=K.t(g[name + "_u"][k][:, None]))
        assert np.array_equal(K.n(d["a_agent"])[:, 0], g[name + "_out_a"][k])
        np.testing.assert_allclose(K.n(d["a_nature"])[0], g[name + "_out_an"][k], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(K.n(d["a_nature_bounded"])[0], g[name + "_out_anb"][k], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(K.n(d["logp_nature"])[0], g[name + "_out_lpn"][k], rtol=2e-5)
        np.testing.assert_allclose(K.n(d["v_nature"])[0], g[name + "_out_vn"][k], rtol=1e-05, atol=2e-07)
        np.testing.assert_allclose(K.n(d["v_agent"])[:, 0], g[name + "_out_v"][k], rtol=1e-05, atol=5e-07)
        np.testing.assert_allclose(K.n(d["logp_agent"])[:, 0], g[name + "_out_logp"][k], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(K.n(d["p1"])[:, 0], g[name + "_out_p1"][k], rtol=1e-5, atol=1e-7)
    out = ac.step(g[name + "_obs"][0].astype(np.float32), 0.5)      # the reference's 9-tuple and shapes
    assert len(out) == 9 and out[0].shape == (N,) and out[4].shape == (D,) and np.ndim(out[5]) == 0
    mu = ac.get_nature_action(g[name + "_obs"][0])
    assert mu.shape == (D,)


def test_agent_critic_extra_batched(K):
    rs = np.random.RandomState(3)
    N, E, S, h, D = 9, 40, 3, 16, 18
    Wpi = onets.init_linear_default(rs, N, S + 1, h, 2)
    Wv = onets.init_linear_default(rs, N, S + 1, h, 1)
    W1n = rs.uniform(-0.3, 0.3, (N, h, D)).astype(F32)
    s = rs.randint(0, S, (N, E)).astype(np.uint8)
    lamb = rs.uniform(0, 2, E).astype(F32)
    nat = rs.uniform(0, 1, (E, D)).astype(F32)
    u = rs.rand(N, E).astype(F32)
    extra = (K.t(W1n).reshape(-1, D) @ K.t(nat).t()).reshape(N, h, E).contiguous()
    net = K.ops.net_desc(N, E, S, 2, h, n_extra=D)
    a, v, logp, p1, _ = K.ops.ac_step(net, K.t(Wpi), K.t(Wv), K.t(s), K.t(lamb), u=K.t(u), extra=extra)
    want = oma.agent_critic_extra(Wv, W1n, s, lamb, nat, S, h)
    np.testing.assert_allclose(K.n(v), want, rtol=1e-05, atol=1e-06)
    with pytest.raises(K.lib.RmabError):
        K.ops.ac_step(net, K.t(Wpi), K.t(Wv), K.t(s), K.t(lamb), u=K.t(u))      # n_extra > 0 without the addend


@pytest.mark.parametrize("N,T,E,S", [(3, 6, 8, 11), (2, 20, 40, 51), (4, 3, 50, 7), (2, 8, 32, 11)])   # last: T * E = 256, aligned tiles
def test_critic_extra_iter_tcgen05_vs_per_sample_kernel(K, N, T, E, S):
    """rmab_critic_extra_iter at hidden 64 with the scalar state input: the sample-major call runs the tcgen05 kernel,
    the arm-major call the per-sample FFMA kernel (itself pinned to the oracle by the MA loop tests); same loss, dLoss/dz1
    and critic after three Adam steps."""
    torch = K.torch
    rs = np.random.RandomState(N * 100 + E)
    h = 64
    Wv0 = onets.init_linear_default(rs, N, 2, h, 1)
    s = rs.randint(0, S, size=(N, T + 1, E)).astype(np.uint8)
    ret = (rs.randn(N, T, E) * 2).astype(F32)
    lamb = rs.uniform(0, 2, size=E).astype(F32)
    net = K.ops.net_desc(N, E, S, 3, h, False, float(S - 1))
    z1 = (rs.randn(N, T * E, h) * 0.3).astype(F32)
    outs = []
    for sample_major in (False, True):
        Wv = K.t(Wv0.copy())
        adam = torch.zeros((N, 2, Wv0.shape[1]), device=K.dev)
        z = K.t(z1.transpose(1, 0, 2).copy()) if sample_major else K.t(z1)
        losses, dz = [], None
        for it in range(3):
            hp = K.ops.hyper(0.2, 0.0, 2e-3, 2e-3, 0, 1, 0, it)
            dz, loss = K.ops.critic_extra_iter(net, hp, Wv, adam, K.t(s), K.t(ret), K.t(lamb), z, want_loss=True,
                                               sample_major=sample_major)
            losses.append(K.n(loss)[:, 0])
        dz = K.n(dz)
        outs.append((np.array(losses), dz.transpose(1, 0, 2) if sample_major else dz, K.n(Wv)))
    (l0, d0, w0), (l1, d1, w1) = outs
    np.testing.assert_allclose(l1, l0, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(d1, d0, rtol=1e-3, atol=2e-8)
    np.testing.assert_allclose(w1, w0, rtol=0, atol=1e-05)


@pytest.mark.parametrize("domain,pop", [(0, None), (1, None), (2, 10), (2, 50)])
def test_counterfactual(K, domain, pop):
    rs = np.random.RandomState(7 + domain)
    N, E, trials = 8, 12, 50
    if domain == 2:
        S = pop + 1
        rg = oenv.sis_param_ranges(N)
        nature = (rg[..., 0] + rs.rand(N, 4) * (rg[..., 1] - rg[..., 0])).astype(F32)
        ranges = np.zeros((1,), F32)
        a = rs.randint(0, 3, (N, E)).astype(np.uint8)
    else:
        S = 2 if domain == 0 else 3
        base = oenv.ce_param_ranges(N) if domain == 0 else oenv.armman_param_ranges(N)
        ranges = oenv.sample_sub_ranges(base, rs).astype(F32)
        nature = (ranges[..., 0] + rs.rand(*ranges.shape[:-1]) * (ranges[..., 1] - ranges[..., 0])).astype(F32)
        a = rs.randint(0, 2, (N, E)).astype(np.uint8)
    s = rs.randint(0, S, (N, E)).astype(np.uint8)
    R = oenv.rewards_table(domain, N, S)
    got = K.n(K.ops.env_counterfactual(domain, K.t(s), K.t(a), K.t(nature), K.t(ranges), K.t(R),
                                       K.ops.rng(5, K.lib.STREAM_CF, 3), trials))
    want = oma.counterfactual_mean(domain, s, a, nature, ranges, R, 5, 3, trials if domain != 2 or pop <= 10 else trials,
                                   pop=pop)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("data,domain", [("counterexample", 0), ("armman", 1)])
def test_ma_epoch_loop_matches_oracle(K, data, domain):
    """Device MA-RMABPPO nature-oracle loop (rollout with counterfactual estimate, agent pi / critic-with-nature-inputs
    updates, lambda step, nature PPO) against the CPU restatement of nature_oracle.py's loop (oracle/ma_train.py)."""
    import copy
    torch = K.torch
    from oracle.ma_train import CpuMARMABPPO
    from robustrmab_b200.environments import ARMMANRobustEnv, CounterExampleRobustEnv
    from robustrmab_b200.evaluate import PessimisticAgentPolicy
    from robustrmab_b200.ma_core import RMABLambdaNatureOracle
    from robustrmab_b200.nature_oracle import MARolloutTrainer, NatureOracle
    N, E, T, h, B, epochs = 6, 4, 8, 16, 2.0, 4
    S = 2 if domain == 0 else 3
    env = (CounterExampleRobustEnv if domain == 0 else ARMMANRobustEnv)(N, B, 9, n_envs=E)
    env.sampled_parameter_ranges = env.sample_parameter_ranges()
    torch.manual_seed(2)
    ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges,
                                env.action_dim_nature, env, hidden_sizes=[h, h], C=env.C, N=N, B=B, n_envs=E)
    hyp = dict(gamma=0.9, lam_other=0.97, clip_ratio=2.0, pi_lr_agent=2e-3, pi_lr_nature=2e-3, vf_lr_agent=2e-3,
               vf_lr_nature=2e-3, lm_lr=2e-3, pi_iters=3, v_iters=3, epochs=epochs, lamb_update_freq=2,
               final_train_lambdas=1, ent0=0.5, ent1=0.0, seed=9)
    lam = [p.detach().numpy().copy() for p in ac.lambda_net.parameters()]
    cpu = CpuMARMABPPO(domain, K.n(ac.Wpi), K.n(ac.Wv), K.n(ac.W1n), lam, copy.deepcopy(ac.pi_nature.mu_net).cpu(),
                       K.n(ac.pi_nature.log_std), copy.deepcopy(ac.v_nature.v_net).cpu(), env.sampled_parameter_ranges,
                       oenv.rewards_table(domain, N, S), np.array([0, 1], F32), E, T, h, B, hyp["gamma"], hyp["lam_other"],
                       hyp["clip_ratio"], hyp["pi_lr_agent"], hyp["pi_lr_nature"], hyp["vf_lr_agent"], hyp["vf_lr_nature"],
                       hyp["lm_lr"], hyp["pi_iters"], hyp["v_iters"], epochs, hyp["lamb_update_freq"],
                       hyp["final_train_lambdas"], hyp["ent0"], hyp["ent1"], hyp["seed"])
    tr = MARolloutTrainer(env, ac, T, hyp["gamma"], hyp["lam_other"], hyp["clip_ratio"], hyp["pi_lr_agent"],
                          hyp["pi_lr_nature"], hyp["vf_lr_agent"], hyp["vf_lr_nature"], hyp["lm_lr"], hyp["pi_iters"],
                          hyp["v_iters"], epochs, hyp["lamb_update_freq"], hyp["final_train_lambdas"], hyp["ent0"],
                          hyp["ent1"], hyp["seed"])
    pess = PessimisticAgentPolicy(N)
    for ep in range(3):                                   # epoch 2 runs the lambda step and the nature PPO updates
        st = tr.epoch(pess)
        w = cpu.epoch(lambda s: np.zeros_like(s))
        np.testing.assert_allclose(K.n(st["Lamb"]), w["lamb"], rtol=1e-05, atol=1e-06)
        agree = (K.n(tr.a) == w["a"]).mean()
        if ep == 0:
            assert agree == 1.0 and np.array_equal(K.n(tr.s_traj), w["s_traj"])
            np.testing.assert_allclose(K.n(tr.a_nat), w["a_nat"], rtol=1e-05, atol=2e-06)
            np.testing.assert_allclose(K.n(tr.rew_nat), w["rew_nat"], rtol=1e-05, atol=1e-05)
            np.testing.assert_allclose(K.n(tr.val), w["val"], rtol=1e-05, atol=1e-06)
            np.testing.assert_allclose(K.n(tr.val_nat), w["val_nat"], rtol=1e-05, atol=1e-06)
            assert np.array_equal(K.n(st["EpActualRetAgent"]), w["ret_agent"])
        else:
            assert agree > 0.95, (ep, agree)
    np.testing.assert_allclose(K.n(ac.Wpi), cpu.opt.Wpi.detach().numpy(), atol=0.0001)
    np.testing.assert_allclose(K.n(ac.Wv), cpu.opt.Wv.detach().numpy(), atol=0.0001)
    np.testing.assert_allclose(K.n(ac.W1n), cpu.W1n.detach().numpy(), atol=0.0001)
    np.testing.assert_allclose(K.n(ac.pi_nature.log_std), cpu.log_std.detach().numpy(), atol=0.0001)
    for p, q in zip(ac.pi_nature.mu_net.parameters(), cpu.mu_net.parameters()):
        np.testing.assert_allclose(K.n(p), q.detach().numpy(), atol=2e-4)
    # the drop-in class end to end
    orc = NatureOracle(data, N, S, 2, B, 9, 1, nature_kwargs=dict(steps_per_epoch=T, epochs=2, gamma=0.9, clip_ratio=2.0,
                       train_pi_iters=2, train_v_iters=2, lamb_update_freq=1, ac_kwargs=dict(hidden_sizes=[h, h])),
                       sampled_nature_parameter_ranges=env.sampled_parameter_ranges, n_envs=E)
    nat_ac = orc.best_response([pess], [1.0], 0)
    assert repr(nat_ac) == "MA-RMABPPO_0" and len(orc.history) == 2
    assert nat_ac.get_nature_action(np.zeros((N, E))).shape == (E, nat_ac.act_dim_nature)


def test_ma_device_loop_matches_reference_nature_oracle(K, golden):
    """Device MA-RMABPPO loop (E = 1) against the values logged by the UNMODIFIED reference NatureOracle.best_response
    (vs a Pessimist agent) run with every draw injected (tests/golden/ma_train.npz)."""
    torch = K.torch
    from robustrmab_b200.environments import CounterExampleRobustEnv
    from robustrmab_b200.evaluate import PessimisticAgentPolicy
    from robustrmab_b200.ma_core import RMABLambdaNatureOracle
    from robustrmab_b200.nature_oracle import MARolloutTrainer
    g = golden("ma_train")
    N, S, T, epochs, h, seed, D = [int(x) for x in g["meta"]]
    kw = dict(zip([str(k) for k in g["hyper_names"]], [float(v) for v in g["hyper_values"]]))
    env = CounterExampleRobustEnv(N, float(g["B"]), seed)
    env.sampled_parameter_ranges = g["ranges"]
    ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, D, env,
                                hidden_sizes=[h, h], C=env.C, N=N, B=env.B)
    ac.Wpi.copy_(K.t(g["Wpi0"])); ac.Wv.copy_(K.t(g["Wv0"])); ac.W1n.copy_(K.t(g["W1n0"]))
    with torch.no_grad():
        for tag, seq in (("mu0", ac.pi_nature.mu_net), ("vn0", ac.v_nature.v_net)):
            for j, m in enumerate([m for m in seq if isinstance(m, torch.nn.Linear)]):
                m.weight.copy_(K.t(g[f"{tag}_{2 * j}"])); m.bias.copy_(K.t(g[f"{tag}_{2 * j + 1}"]))
        ac.pi_nature.log_std.copy_(K.t(g["log_std0"]))
        for i, p in enumerate(ac.lambda_net.parameters()):
            p.copy_(torch.as_tensor(g[f"lam0_{i}"]))
    tr = MARolloutTrainer(env, ac, T, kw["gamma"], kw["lam_OTHER"], kw["clip_ratio"], kw["pi_lr_agent"], kw["pi_lr_nature"],
                          kw["vf_lr_agent"], kw["vf_lr_nature"], kw["lm_lr"], int(kw["train_pi_iters"]),
                          int(kw["train_v_iters"]), epochs, int(kw["lamb_update_freq"]), int(kw["final_train_lambdas"]),
                          kw["start_entropy_coeff"], kw["end_entropy_coeff"], seed)
    pess = PessimisticAgentPolicy(N)
    lamb, ret, rn, vv, vn = [], [], [], [], []
    for _ in range(epochs):
        st = tr.epoch(pess)
        lamb.append(float(st["Lamb"][0])); ret.append(float(st["EpActualRetAgent"][0]))
        rn.append(float(st["EpLambAdjRetNature"][0])); vv.append(float(st["VVals_agent"][0])); vn.append(float(st["VVals_nature"][0]))
    assert np.array_equal(np.array(ret), g["EpActualRetAgent"])                   # same trajectories as the reference
    np.testing.assert_allclose(lamb, g["Lamb"], rtol=1e-05, atol=2e-07)
    np.testing.assert_allclose(rn, g["EpLambAdjRetNature"], rtol=1e-05, atol=1e-05)
    np.testing.assert_allclose(vv, g["VVals_agent"], rtol=0.0001, atol=1e-05)
    np.testing.assert_allclose(vn, g["VVals_nature"], rtol=0.0001, atol=1e-05)
    np.testing.assert_allclose(K.n(ac.Wpi), g["Wpi1"], atol=5e-5)
    np.testing.assert_allclose(K.n(ac.Wv), g["Wv1"], atol=5e-5)
    np.testing.assert_allclose(K.n(ac.W1n), g["W1n1"], atol=5e-5)
    tr.export_lambda()
    for i, p in enumerate(ac.lambda_net.parameters()):
        np.testing.assert_allclose(p.detach().numpy(), g[f"lam1_{i}"], rtol=1e-05, atol=1e-07)


def test_nature_oracle_resamples_agent_strategy(K):
    """nature_oracle.py:651, :669-670: the agent strategy nature trains against is drawn from the mixed equilibrium at the
    start AND re-drawn every lamb_update_freq epochs -- with a mixed agent_eq nature must meet both strategies."""
    torch = K.torch
    from robustrmab_b200.baselines import PessimisticAgentPolicy
    from robustrmab_b200.core import MLPActorCriticRMAB
    from robustrmab_b200.nature_oracle import NatureOracle
    N, E, T, epochs = 6, 4, 5, 8
    kw = dict(steps_per_epoch=T, epochs=epochs, gamma=0.9, clip_ratio=2.0, pi_lr_agent=2e-3, pi_lr_nature=2e-3,
              vf_lr_agent=2e-3, vf_lr_nature=2e-3, lm_lr=2e-3, train_pi_iters=1, train_v_iters=1, lamb_update_freq=1,
              final_train_lambdas=0, ac_kwargs=dict(hidden_sizes=[16, 16]))
    tmp = NatureOracle("counterexample", N, 2, 2, 2.0, 3, 1, nature_kwargs=kw, n_envs=E)
    orc = NatureOracle("counterexample", N, 2, 2, 2.0, 3, 1, nature_kwargs=kw, n_envs=E,
                       sampled_nature_parameter_ranges=tmp.env.sample_parameter_ranges())
    pess = PessimisticAgentPolicy(N)
    ac = MLPActorCriticRMAB(np.arange(2), np.arange(2), hidden_sizes=[16, 16], C=np.array([0, 1]), N=N, B=2.0, n_envs=1)
    strats, eq = [pess, ac], [0.5, 0.5]
    np.random.seed(123)
    orc.best_response(strats, eq, 0)
    np.random.seed(123)
    want = [np.random.choice(strats, p=np.array(eq))]
    for epoch in range(1, epochs):
        want.append(np.random.choice(strats, p=np.array(eq)))
    assert [id(p) for p in orc.agent_pol_history] == [id(p) for p in want]
    assert {id(p) for p in orc.agent_pol_history} == {id(pess), id(ac)}
    assert ac.n_envs == 1                                   # the caller's strategy object is not mutated
    assert all(np.isfinite(h["EpLambAdjRetNature"]).all() for h in orc.history)


# ---------------------------------------------------------------------------------------------------------------------
# SIS (BASELINE configs[3]'s domain: S = pop + 1, A = 3, C = [0,1,2], D = 4N nature parameters, scalar state inputs)
# Tolerances (set from measured headroom, profiles/r2_tolerance_headroom.json): the integer path (states, actions) is
# compared bit for bit on the first epoch, i.e. while both sides hold identical nets; per-step float outputs of that epoch
# are held to north_star's 1e-5 relative.  Nets AFTER several Adam steps: Adam divides by sqrt(v) + 1e-8, so a gradient
# entry near zero turns an fp32 rounding difference into a full +-lr step (lr = 2e-3 here; at hidden 64 the products are
# 3xTF32, relative error 2.4e-7 against 6e-8 for FFMA) -- the critics' bound of 3e-3 is one such step plus drift
# (measured worst: 2.3e-3 on W1n), everything else is held to 1e-4 / 2e-4.
def _sis_problem(K, N, E, pop, h, seed, B):
    from robustrmab_b200.environments import SISRobustEnv
    from robustrmab_b200.ma_core import RMABLambdaNatureOracle
    env = SISRobustEnv(N, B, pop, seed, n_envs=E)
    env.sampled_parameter_ranges = env.sample_parameter_ranges()
    K.torch.manual_seed(seed)
    ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges,
                                env.action_dim_nature, env, hidden_sizes=[h, h], C=env.C, N=N, B=B, n_envs=E,
                                one_hot_encode=False, non_ohe_obs_dim=1, state_norm=pop, nature_state_norm=1)
    return env, ac


@pytest.mark.parametrize("pop,h", [(10, 16), (20, 64), (10, 64)])
def test_ma_epoch_loop_sis_matches_oracle(K, pop, h):
    """Device MA-RMABPPO loop on SIS against oracle/ma_train.py (itself pinned to the unmodified reference on SIS by
    tests/golden/ma_train_sis.npz), E = 4 envs, incl. the lambda step and the nature PPO updates of epoch 2."""
    import copy
    from oracle.ma_train import CpuMARMABPPO
    from robustrmab_b200.baselines import PessimisticAgentPolicy
    from robustrmab_b200.nature_oracle import MARolloutTrainer
    N, E, T, B, epochs = 6, 4, 6, 3.0, 4
    S = pop + 1
    env, ac = _sis_problem(K, N, E, pop, h, 13, B)
    hyp = dict(gamma=0.9, lam_other=0.97, clip_ratio=2.0, pi_lr_agent=2e-3, pi_lr_nature=2e-3, vf_lr_agent=2e-3,
               vf_lr_nature=2e-3, lm_lr=2e-3, pi_iters=3, v_iters=3, epochs=epochs, lamb_update_freq=2,
               final_train_lambdas=1, ent0=0.5, ent1=0.0, seed=13)
    lam = [p.detach().numpy().copy() for p in ac.lambda_net.parameters()]
    cpu = CpuMARMABPPO(oenv.DOMAIN_SIS, K.n(ac.Wpi), K.n(ac.Wv), K.n(ac.W1n), lam, copy.deepcopy(ac.pi_nature.mu_net).cpu(),
                       K.n(ac.pi_nature.log_std), copy.deepcopy(ac.v_nature.v_net).cpu(), env.sampled_parameter_ranges,
                       oenv.rewards_table(2, N, S), np.array([0, 1, 2], F32), E, T, h, B, hyp["gamma"], hyp["lam_other"],
                       hyp["clip_ratio"], hyp["pi_lr_agent"], hyp["pi_lr_nature"], hyp["vf_lr_agent"], hyp["vf_lr_nature"],
                       hyp["lm_lr"], hyp["pi_iters"], hyp["v_iters"], epochs, hyp["lamb_update_freq"],
                       hyp["final_train_lambdas"], hyp["ent0"], hyp["ent1"], hyp["seed"], pop=pop, one_hot=False,
                       state_norm=pop, nature_state_norm=1.0)
    tr = MARolloutTrainer(env, ac, T, hyp["gamma"], hyp["lam_other"], hyp["clip_ratio"], hyp["pi_lr_agent"],
                          hyp["pi_lr_nature"], hyp["vf_lr_agent"], hyp["vf_lr_nature"], hyp["lm_lr"], hyp["pi_iters"],
                          hyp["v_iters"], epochs, hyp["lamb_update_freq"], hyp["final_train_lambdas"], hyp["ent0"],
                          hyp["ent1"], hyp["seed"])
    pess = PessimisticAgentPolicy(N)
    for ep in range(3):
        st = tr.epoch(pess)
        w = cpu.epoch(lambda s: np.zeros_like(s))
        np.testing.assert_allclose(K.n(st["Lamb"]), w["lamb"], rtol=1e-05, atol=1e-06)
        if ep == 0:
            assert np.array_equal(K.n(tr.a), w["a"]) and np.array_equal(K.n(tr.s_traj), w["s_traj"])
            np.testing.assert_allclose(K.n(tr.a_nat), w["a_nat"], rtol=1e-05, atol=2e-06)
            np.testing.assert_allclose(K.n(tr.rew_nat), w["rew_nat"], rtol=1e-05, atol=1e-05)
            np.testing.assert_allclose(K.n(tr.val), w["val"], rtol=1e-05, atol=1e-06)
            np.testing.assert_allclose(K.n(tr.val_nat), w["val_nat"], rtol=1e-05, atol=1e-06)
            np.testing.assert_allclose(K.n(st["EpActualRetAgent"]), w["ret_agent"], rtol=1e-6)
        else:
            assert (K.n(tr.a) == w["a"]).mean() > 0.9, ep
    np.testing.assert_allclose(K.n(ac.Wpi), cpu.opt.Wpi.detach().numpy(), atol=0.0001)
    np.testing.assert_allclose(K.n(ac.Wv), cpu.opt.Wv.detach().numpy(), atol=3e-3)
    np.testing.assert_allclose(K.n(ac.W1n), cpu.W1n.detach().numpy(), atol=3e-3)
    np.testing.assert_allclose(K.n(ac.pi_nature.log_std), cpu.log_std.detach().numpy(), atol=0.0001)
    for p, q in zip(ac.pi_nature.mu_net.parameters(), cpu.mu_net.parameters()):
        np.testing.assert_allclose(K.n(p), q.detach().numpy(), atol=2e-4)
    for p, q in zip(ac.v_nature.v_net.parameters(), cpu.v_nature.parameters()):
        np.testing.assert_allclose(K.n(p), q.detach().numpy(), atol=2e-4)


def test_ma_device_loop_sis_matches_reference_nature_oracle(K, golden):
    """Device MA-RMABPPO loop on SIS (hidden 64, E = 1) against the values logged by the UNMODIFIED reference
    NatureOracle.best_response on SISRobustEnv with every draw injected (tests/golden/ma_train_sis.npz)."""
    torch = K.torch
    from robustrmab_b200.baselines import PessimisticAgentPolicy
    from robustrmab_b200.environments import SISRobustEnv
    from robustrmab_b200.ma_core import RMABLambdaNatureOracle
    from robustrmab_b200.nature_oracle import MARolloutTrainer
    g = golden("ma_train_sis")
    N, S, T, epochs, h, seed, D = [int(x) for x in g["meta"]]
    pop = int(g["pop"])
    kw = dict(zip([str(k) for k in g["hyper_names"]], [float(v) for v in g["hyper_values"]]))
    env = SISRobustEnv(N, float(g["B"]), pop, seed)
    env.sampled_parameter_ranges = g["ranges"]
    ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, D, env,
                                hidden_sizes=[h, h], C=env.C, N=N, B=env.B, one_hot_encode=False, non_ohe_obs_dim=1,
                                state_norm=pop, nature_state_norm=1)
    ac.Wpi.copy_(K.t(g["Wpi0"])); ac.Wv.copy_(K.t(g["Wv0"])); ac.W1n.copy_(K.t(g["W1n0"]))
    with torch.no_grad():
        for tag, seq in (("mu0", ac.pi_nature.mu_net), ("vn0", ac.v_nature.v_net)):
            for j, m in enumerate([m for m in seq if isinstance(m, torch.nn.Linear)]):
                m.weight.copy_(K.t(g[f"{tag}_{2 * j}"])); m.bias.copy_(K.t(g[f"{tag}_{2 * j + 1}"]))
        ac.pi_nature.log_std.copy_(K.t(g["log_std0"]))
        for i, p in enumerate(ac.lambda_net.parameters()):
            p.copy_(torch.as_tensor(g[f"lam0_{i}"]))
    tr = MARolloutTrainer(env, ac, T, kw["gamma"], kw["lam_OTHER"], kw["clip_ratio"], kw["pi_lr_agent"], kw["pi_lr_nature"],
                          kw["vf_lr_agent"], kw["vf_lr_nature"], kw["lm_lr"], int(kw["train_pi_iters"]),
                          int(kw["train_v_iters"]), epochs, int(kw["lamb_update_freq"]), int(kw["final_train_lambdas"]),
                          kw["start_entropy_coeff"], kw["end_entropy_coeff"], seed)
    pess = PessimisticAgentPolicy(N)
    lamb, ret, rn, vv, vn = [], [], [], [], []
    for _ in range(epochs):
        st = tr.epoch(pess)
        lamb.append(float(st["Lamb"][0])); ret.append(float(st["EpActualRetAgent"][0]))
        rn.append(float(st["EpLambAdjRetNature"][0])); vv.append(float(st["VVals_agent"][0])); vn.append(float(st["VVals_nature"][0]))
    # With ONE env and six arms a single draw that lands on the other side of a cdf edge (nets equal to ~1e-6 after an
    # Adam step; hidden 64 runs 3xTF32 products) changes the whole remaining trajectory, so values are compared over the
    # leading epochs whose trajectories are identical to the reference's (at least the first); the chain to the
    # reference for the later epochs is golden <-> oracle (tests/test_oracle_golden.py, 2e-5) <-> device
    # (test_ma_epoch_loop_sis_matches_oracle, E = 4).
    want = g["EpActualRetAgent"]
    same = 0
    while same < epochs and abs(ret[same] - want[same]) <= 1e-6 * max(1.0, abs(want[same])):
        same += 1
    assert same >= 1, (ret, want)
    k = min(same + 1, epochs)
    np.testing.assert_allclose(lamb[:k], g["Lamb"][:k], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rn[:same], g["EpLambAdjRetNature"][:same], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(vv[:same], g["VVals_agent"][:same], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(vn[:same], g["VVals_nature"][:same], rtol=1e-4, atol=1e-5)
    if same == epochs:
        np.testing.assert_allclose(K.n(ac.Wpi), g["Wpi1"], atol=3e-3)
        np.testing.assert_allclose(K.n(ac.Wv), g["Wv1"], atol=3e-3)
        np.testing.assert_allclose(K.n(ac.W1n), g["W1n1"], atol=3e-3)
        tr.export_lambda()
        for i, p in enumerate(ac.lambda_net.parameters()):
            np.testing.assert_allclose(p.detach().numpy(), g[f"lam1_{i}"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("N,T,E,S", [(3, 6, 8, 11), (2, 20, 40, 51)])
def test_critic_extra_iter_h64_vs_autograd_oracle(K, N, T, E, S):
    """rmab_critic_extra_iter at hidden 64 in its tcgen05 (sample-major) mode DIRECTLY against torch autograd on the
    reference's critic (compute_loss_v_agent, nature_oracle.py:388-416, as restated in oracle/ma_train.py): loss before
    the step, dLoss/dz1, and the packed critic after three Adam steps.  Contract: loss 1e-5 relative (north_star);
    dz1 (of the THIRD step, i.e. after two Adam steps on either side) 1e-3 relative + 1e-8 -- entries are O(1e-6), the
    mean over T*E samples; weights 2e-4 absolute after 3 Adam steps of lr 2e-3."""
    import torch
    from oracle import ppo as oppo
    rs = np.random.RandomState(N * 10 + E)
    h, in_dim = 64, 2
    Wv0 = onets.init_linear_default(rs, N, in_dim, h, 1)
    s = rs.randint(0, S, size=(N, T + 1, E)).astype(np.uint8)
    ret = (rs.randn(N, T, E) * 2).astype(F32)
    lamb = rs.uniform(0, 2, size=E).astype(F32)
    z1 = (rs.randn(T * E, N, h) * 0.3).astype(F32)                               # sample-major extra pre-activation
    net = K.ops.net_desc(N, E, S, 3, h, False, float(S - 1))
    # ---- oracle: autograd + torch Adam on the same loss ----
    W = torch.nn.Parameter(torch.as_tensor(Wv0.copy()))
    opt = torch.optim.Adam([W], lr=2e-3)
    x = np.stack([onets.features(s[:, t], lamb, S, False, float(S - 1)) for t in range(T)], axis=1).reshape(N, T * E, in_dim)
    xt, rt = torch.as_tensor(x), torch.as_tensor(ret.reshape(N, T * E))
    zt = torch.as_tensor(z1.transpose(1, 0, 2).copy()).requires_grad_(True)       # [N, T*E, h]
    want_loss, want_dz = [], None
    for it in range(3):
        opt.zero_grad()
        zt.grad = None
        W1, b1, W2, b2, W3, b3 = oppo._unpack_t(W, in_dim, h, 1)
        pre = torch.bmm(xt, W1.transpose(1, 2)) + b1[:, None, :] + zt
        a2 = torch.tanh(torch.bmm(torch.tanh(pre), W2.transpose(1, 2)) + b2[:, None, :])
        v = (torch.bmm(a2, W3.transpose(1, 2)) + b3[:, None, :])[..., 0]
        loss = ((v - rt) ** 2).mean(dim=1)
        loss.sum().backward()
        want_loss.append(loss.detach().numpy().copy())
        want_dz = zt.grad.detach().numpy().copy()
        opt.step()
    # ---- device ----
    Wv = K.t(Wv0.copy())
    adam = K.torch.zeros((N, 2, Wv0.shape[1]), device=K.dev)
    got_loss = []
    for it in range(3):
        hp = K.ops.hyper(0.2, 0.0, 2e-3, 2e-3, 0, 1, 0, it)
        dz, loss = K.ops.critic_extra_iter(net, hp, Wv, adam, K.t(s), K.t(ret), K.t(lamb), K.t(z1), want_loss=True,
                                           sample_major=True)
        got_loss.append(K.n(loss)[:, 0])
    np.testing.assert_allclose(np.array(got_loss)[0], want_loss[0], rtol=1e-5)
    np.testing.assert_allclose(np.array(got_loss), np.array(want_loss), rtol=1e-05)     # after 1-2 Adam steps
    np.testing.assert_allclose(K.n(dz).transpose(1, 0, 2), want_dz, rtol=1e-3, atol=1e-8)
    np.testing.assert_allclose(K.n(Wv), W.detach().numpy(), rtol=0, atol=2e-4)


@pytest.mark.parametrize("M,N,Kd", [(256, 256, 64), (300, 520, 200), (128, 1000, 72), (1000, 130, 1032), (5120, 1024, 2048)])
def test_gemm3h_matches_fp64_product(K, M, N, Kd):
    """The package's dense product for the agent critics' nature block (csrc/gemm3h.cu: fp16 hi/lo split with per-row
    power-of-two scales, tcgen05 kind::f16 x 3, TMA-fed CTA pairs) against the fp64 product of the same fp32 inputs.
    Tolerance: the scheme keeps 22 significant bits per operand and accumulates in fp32, so the error is bounded by a few
    2^-22 of sum_k |a||b|; rows with very different magnitudes and an all-zero row exercise the per-row scales."""
    import torch
    ops = K.ops
    g = torch.Generator(device="cuda").manual_seed(M + N + Kd)
    a = torch.randn((M, Kd), device="cuda", generator=g) * torch.logspace(-6, 3, M, device="cuda")[:, None]
    b = torch.randn((N, Kd), device="cuda", generator=g) * torch.logspace(2, -8, N, device="cuda")[:, None]
    a[M // 2] = 0.0
    want = a.double() @ b.double().t()
    bound = a.double().abs() @ b.double().abs().t()
    for n_fast in (True, False):
        got = ops.gemm3h(ops.split16_rows(a), ops.split16_rows(b), Kd, n_fast=n_fast)
        err = ((got.double() - want).abs() / (bound + 1e-300)).max().item()
        assert err < 2e-6, (n_fast, err)
    # the transposed split: (a^T)^T . b^T from a^T stored [K, M]
    at = a.t().contiguous()
    got = ops.gemm3h(ops.split16_cols(at), ops.split16_rows(b), Kd)
    err = ((got.double() - want).abs() / (bound + 1e-300)).max().item()
    assert err < 2e-6, err
    # an output view with a row stride (ldc > N)
    big = torch.zeros((M, N + 8), device="cuda")
    ops.gemm3h(ops.split16_rows(a), ops.split16_rows(b), Kd, out=big[:, :N])
    assert torch.equal(big[:, :N], ops.gemm3h(ops.split16_rows(a), ops.split16_rows(b), Kd)) and float(big[:, N:].abs().max()) == 0.0


def _nature_problem(K, N, k, h, T, E, seed):
    """A tiny actor-critic shell with torch nature modules + random trajectories for the nature-net kernel tests."""
    import torch
    from types import SimpleNamespace
    torch.manual_seed(seed)
    dev = K.dev
    mk = lambda i, o: torch.nn.Sequential(torch.nn.Linear(i, h), torch.nn.Tanh(), torch.nn.Linear(h, h), torch.nn.Tanh(),
                                          torch.nn.Linear(h, o)).to(dev)
    D = k * N
    ac = SimpleNamespace(N=N, act_dim_nature=D, hidden_sizes=[h, h], device_=dev,
                         pi_nature=SimpleNamespace(mu_net=mk(N, D), log_std=torch.nn.Parameter(torch.full((D,), -0.5, device=dev))),
                         v_nature=SimpleNamespace(v_net=mk(2 * N, 1)))
    g = torch.Generator(device=dev).manual_seed(seed)
    s_traj = torch.randint(0, 5, (N, T + 1, E), dtype=torch.uint8, device=dev, generator=g)
    a = torch.randint(0, 3, (N, T, E), dtype=torch.uint8, device=dev, generator=g)
    act = torch.randn((T * E, D), device=dev, generator=g)
    logp_old = torch.randn((T * E,), device=dev, generator=g) * 0.3 - 0.9 * D
    adv = torch.randn((T * E,), device=dev, generator=g)
    ret = torch.randn((T * E,), device=dev, generator=g)
    return ac, s_traj, a, act, logp_old, adv, ret


@pytest.mark.parametrize("N,k,h,T,E", [(7, 1, 8, 3, 5), (13, 4, 16, 5, 9), (70, 2, 64, 4, 33)])
def test_nature_net_kernels_vs_torch_autograd(K, N, k, h, T, E):
    """rmab_nat_* (csrc/nature.cu) against the torch modules they replace: forward (mu, nature value), then three Adam steps
    of compute_loss_pi_nature (nature_oracle.py:359-386, PPO clip on the summed Normal log-prob) and of compute_loss_v_nature
    (:421-428) -- unsharded, and as two arm shards whose all-reduces are emulated by sums.  fp32 against fp32 with a different
    summation order; Adam divides by sqrt(v), so an element whose gradient is near zero turns a 1e-7 gradient difference
    into a visible fraction of a step: bound = 1e-4 absolute = 1 % of one step of lr 1e-2 (measured: 4e-5 on 7 of 17 944
    elements at the largest shape, < 2e-5 elsewhere), on weights of magnitude 0.1-0.5."""
    import copy
    import torch
    from robustrmab_b200.nature_nets import NatureNets
    ac, s_traj, a, act, logp_old, adv, ret = _nature_problem(K, N, k, h, T, E, N + h)
    M, D, sc, clip, lr = T * E, k * N, 0.25, 0.3, 1e-2
    logp_old = logp_old * 0 + torch.distributions.Normal(ac.pi_nature.mu_net((s_traj[:, :T].permute(1, 2, 0).reshape(M, N).float() * sc)),
                                                         torch.exp(ac.pi_nature.log_std)).log_prob(act).sum(-1).detach() + 0.2 * adv
    # ---- torch reference: the modules, autograd, torch.optim.Adam ----
    mu_ref, v_ref = copy.deepcopy(ac.pi_nature.mu_net), copy.deepcopy(ac.v_nature.v_net)
    ls_ref = torch.nn.Parameter(ac.pi_nature.log_std.detach().clone())
    opt_pi = torch.optim.Adam(list(mu_ref.parameters()) + [ls_ref], lr=lr)
    opt_v = torch.optim.Adam(v_ref.parameters(), lr=lr)
    obs = s_traj[:, :T].permute(1, 2, 0).reshape(M, N).float() * sc
    xa = torch.cat([obs, a.permute(1, 2, 0).reshape(M, N).float()], dim=1)
    nets = NatureNets(ac)
    # forward parity first
    s0 = s_traj[:, 0].contiguous()
    np.testing.assert_allclose(K.n(nets.actor_mu(s0, sc)), K.n(mu_ref(s0.t().float() * sc)), rtol=1e-5, atol=2e-6)
    vn = nets.critic_values(s_traj, torch.cat([a, a[:, :1]], dim=1).contiguous(), sc, T, E)
    np.testing.assert_allclose(K.n(vn), K.n(v_ref(xa)[:, 0]), rtol=1e-5, atol=2e-6)
    for _ in range(3):
        opt_pi.zero_grad()
        pi = torch.distributions.Normal(mu_ref(obs), torch.exp(ls_ref))
        ratio = torch.exp(pi.log_prob(act).sum(-1) - logp_old)
        (-(torch.min(ratio * adv, torch.clamp(ratio, 1 - clip, 1 + clip) * adv)).mean()).backward()
        opt_pi.step()
        opt_v.zero_grad()
        ((v_ref(xa)[:, 0] - ret) ** 2).mean().backward()
        opt_v.step()
    # ---- kernels, one rank ----
    for _ in range(3):
        nets.actor_step(s_traj, T, E, sc, act, logp_old, adv, clip, lr)
        nets.critic_step(s_traj, a, T, E, sc, ret, lr)
    # ---- kernels, two arm shards on one GPU (all-reduce = sum of the two shards' buffers) ----
    cut = N // 2 + 1
    ac2 = _nature_problem(K, N, k, h, T, E, N + h)[0]
    sh = [NatureNets(ac2, 0, cut), NatureNets(ac2, cut, N - cut)]
    st = [s_traj[:cut].contiguous(), s_traj[cut:].contiguous()]
    aa = [a[:cut].contiguous(), a[cut:].contiguous()]

    # the two shards run in lockstep on two threads: an "all-reduce" parks its tensor, and when both have arrived the
    # tensors are summed in place
    import threading
    bar = threading.Barrier(2)
    slots = [None, None]

    def make_ar(r):
        def ar(t):
            torch.cuda.synchronize()
            slots[r] = t
            bar.wait()
            if r == 0:
                tot = slots[0] + slots[1]
                slots[0].copy_(tot)
                slots[1].copy_(tot)
                torch.cuda.synchronize()
            bar.wait()
            return t
        return ar

    def worker(r):
        torch.cuda.set_device(K.dev)
        for _ in range(3):
            sh[r].actor_step(st[r], T, E, sc, act, logp_old, adv, clip, lr, allreduce=make_ar(r))
            sh[r].critic_step(st[r], aa[r], T, E, sc, ret, lr, allreduce=make_ar(r))
        sh[r].exchange_shards(make_ar(r))
    th = [threading.Thread(target=worker, args=(r,)) for r in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    for name, packed_list in (("one rank", [nets]), ("two shards", sh)):
        for pk in packed_list:
            ref_actor = torch.cat([mu_ref[0].weight.detach().t().reshape(-1), mu_ref[0].bias.detach(), mu_ref[2].weight.detach().reshape(-1),
                                   mu_ref[2].bias.detach(), mu_ref[4].weight.detach().reshape(-1), mu_ref[4].bias.detach(), ls_ref.detach()])
            ref_critic = torch.cat([v_ref[0].weight.detach().t().reshape(-1), v_ref[0].bias.detach(), v_ref[2].weight.detach().reshape(-1),
                                    v_ref[2].bias.detach(), v_ref[4].weight.detach().reshape(-1), v_ref[4].bias.detach()])
            np.testing.assert_allclose(K.n(pk.actor.p), K.n(ref_actor), atol=1e-4, rtol=0, err_msg=name + " actor")
            np.testing.assert_allclose(K.n(pk.critic.p), K.n(ref_critic), atol=1e-4, rtol=0, err_msg=name + " critic")
    assert torch.equal(sh[0].actor.p, sh[1].actor.p) and torch.equal(sh[0].critic.p, sh[1].critic.p)


def test_ma_rollout_graphs_equal_eager(K):
    """The CUDA-graph replay of the rollout step (Philox counters through device words) gives exactly the eager rollout:
    two trainers on the same problem, one with graphs, five epochs with updates in between -- every trajectory byte and
    every net bit-equal."""
    import torch
    from robustrmab_b200.environments import SISRobustEnv
    from robustrmab_b200.evaluate import PessimisticAgentPolicy
    from robustrmab_b200.ma_core import RMABLambdaNatureOracle
    from robustrmab_b200.nature_oracle import MARolloutTrainer
    N, E, T, h, pop = 9, 6, 5, 64, 10
    out = []
    for graphs in (False, True):
        env = SISRobustEnv(N, 2.0, pop, 4, n_envs=E)
        env.random_stream.seed(4)
        env.sampled_parameter_ranges = env.sample_parameter_ranges()
        torch.manual_seed(7)
        ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, env.action_dim_nature,
                                    env, hidden_sizes=[h, h], C=env.C, N=N, B=env.B, one_hot_encode=False, non_ohe_obs_dim=1,
                                    state_norm=pop, nature_state_norm=1, n_envs=E)
        tr = MARolloutTrainer(env, ac, T, 0.9, 0.97, 2.0, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, 2, 2, 8, 2, 1, 0.5, 0.0, 4)
        tr.use_graphs = graphs
        pess = PessimisticAgentPolicy(N)
        trace = []
        for _ in range(5):
            tr.epoch(pess)
            trace.append((tr.s_traj.clone(), tr.a.clone(), tr.a_nat.clone(), tr.rew_nat.clone(), tr.val.clone()))
        assert (tr.graphs is not None) == graphs
        out.append((trace, ac.Wpi.clone(), ac.Wv.clone(), ac.W1n.clone(), tr.nets.actor.p.clone(), tr.nets.critic.p.clone()))
    for (ta, tb) in zip(out[0][0], out[1][0]):
        for x, y in zip(ta, tb):
            assert torch.equal(x, y)
    for x, y in zip(out[0][1:], out[1][1:]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("M,N,Kd", [(300, 520, 200), (1024, 768, 512)])
def test_gemm3h_adam_epilogue_equals_separate_adam(K, M, N, Kd):
    """rmab_gemm3h_adam (the gradient consumed in the product's epilogue) == rmab_gemm3h followed by rmab_adam_step, bit
    for bit, over three steps (parameters and both moments)."""
    import torch
    ops = K.ops
    g = torch.Generator(device="cuda").manual_seed(M + Kd)
    p0 = torch.randn((M, N), device="cuda", generator=g) * 0.05
    pa, ma, va = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    pb, mb, vb = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in (1, 2, 3):
        a = torch.randn((M, Kd), device="cuda", generator=g) * 1e-3
        b = torch.randn((N, Kd), device="cuda", generator=g)
        sa, sb = ops.split16_rows(a), ops.split16_rows(b)
        grad = ops.gemm3h(sa, sb, Kd)
        ops.adam_step(pa, grad, ma, va, 2e-3, step)
        ops.gemm3h_adam(sa, sb, Kd, pb, mb, vb, 2e-3, step)
        assert torch.equal(pa, pb) and torch.equal(ma, mb) and torch.equal(va, vb), step
