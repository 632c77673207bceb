"""Pin the CPU oracle against golden vectors generated from the reference's own classes
(oracle/gen_golden.py) and against the mdptoolbox docstring known answers.  CPU only."""
import numpy as np
import pytest

from oracle.golden_setup import ma_train_setup, sis_ma_oracle, sis_rmabppo_oracle
from oracle import buffer as obuf
from oracle import envs as oenv
from oracle import nets as onets
from oracle import philox
from oracle import select as osel
from oracle import vi as ovi

F32 = np.float32


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kat:
        got = philox.philox4x32_10(*c, *k)
        assert tuple(int(x) for x in got) == want


def test_categorical_rule():
    p = np.array([[0.2, 0.3, 0.5], [0.0, 1.0, 0.0], [0.5, 0.5, 0.0]], F32)
    assert philox.categorical(p, np.array([0.1, 0.999, 0.75], F32)).tolist() == [0, 1, 1]
    assert philox.categorical(p[0], F32(0.2)) == 1          # cdf <= u moves on
    assert philox.categorical(p[0], F32(0.99999994)) == 2


def test_env_ce(golden):
    g = golden("env")
    K, N = g["ce_s"].shape
    for t in range(K):
        sn, r = oenv.ce_step(g["ce_s"][t][:, None], g["ce_a"][t][:, None], g["ce_nature"][t],
                             g["ce_ranges"], g["ce_R"], g["ce_u"][t][:, None])
        assert np.array_equal(sn[:, 0], g["ce_s_next"][t])
        assert np.array_equal(r[:, 0], g["ce_r"][t].astype(F32))
    assert np.allclose(oenv.ce_param_ranges(N), g["ce_PARAMETER_RANGES"])
    s0 = oenv.reset_states(11, 0, 2, 1, N)[:, 0]
    assert np.array_equal(s0, g["ce_s0"])


def test_env_armman(golden):
    g = golden("env")
    K, N = g["armman_s"].shape
    for t in range(K):
        nat = np.moveaxis(g["armman_nature"][t][None], 0, 2)        # [N,A,E=1]
        sn, r = oenv.armman_step(g["armman_s"][t][:, None], g["armman_a"][t][:, None], nat,
                                 g["armman_ranges"], g["armman_R"], g["armman_u"][t][:, None],
                                 nature_is_table=False)
        assert np.array_equal(sn[:, 0], g["armman_s_next"][t])
        assert np.array_equal(r[:, 0], g["armman_r"][t].astype(F32))
    assert np.allclose(oenv.armman_param_ranges(N), g["armman_PARAMETER_RANGES"])
    assert np.array_equal(oenv.reset_states(12, 0, 3, 1, N)[:, 0], g["armman_s0"])


def test_env_sis(golden):
    g = golden("env")
    K, N = g["sis_s"].shape
    pop = int(g["sis_pop"])
    for t in range(K):
        sn, r = oenv.sis_step(pop, g["sis_s"][t][:, None], g["sis_a"][t][:, None], g["sis_nature"][t],
                              g["sis_R"], g["sis_u"][t][:, None])
        assert np.array_equal(sn[:, 0], g["sis_s_next"][t])
        assert np.allclose(r[:, 0], g["sis_r"][t], rtol=1e-7)
    T = g["sis_T"]
    for n in range(N):
        for s in range(pop + 1):
            for a in range(3):
                assert np.array_equal(oenv.sis_distro(pop, s, a, g["sis_nature"][-1][n]), T[n, s, a])
    T50 = g["sis50_T"]
    for s in (0, 1, 17, 49, 50):
        for a in range(3):
            assert np.array_equal(oenv.sis_distro(50, s, a, g["sis50_params"][0]), T50[0, s, a])
    assert np.array_equal(oenv.reset_states(13, 0, pop + 1, 1, N)[:, 0], g["sis_s0"])


def test_bound_nature(golden):
    g = golden("env")
    r = g["ce_ranges"]
    assert np.array_equal(oenv.bound_nature(g["ce_bound_raw"], r[:, 0], r[:, 1]), g["ce_bound_out"])
    r = g["armman_ranges"]
    N = r.shape[0]
    s = g["armman_bound_state"].astype(int)
    lo, hi = r[np.arange(N), s, :, 0], r[np.arange(N), s, :, 1]
    got = oenv.bound_nature(g["armman_bound_raw"].reshape(N, 2), lo, hi)
    assert np.array_equal(got, g["armman_bound_out"])
    r = g["sis_ranges"]
    got = oenv.bound_nature(g["sis_bound_raw"].reshape(-1, 4), r[..., 0], r[..., 1])
    assert np.array_equal(got, g["sis_bound_out"])


@pytest.mark.parametrize("name", ["oh3", "oh2", "sis", "oh3h64"])
def test_ac_step(golden, name):
    g = golden("ac")
    N, S, A, h, one_hot, norm = [int(x) for x in g[name + "_meta"]]
    K = g[name + "_obs"].shape[0]
    for k in range(K):
        a, v, logp, p1, probs = onets.ac_step(g[name + "_Wpi"], g[name + "_Wv"], g[name + "_obs"][k][:, None],
                                             g[name + "_lamb"][k:k + 1], g[name + "_u"][k][:, None],
                                             S, A, h, bool(one_hot), norm)
        np.testing.assert_allclose(p1[:, 0], g[name + "_p1"][k], rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(v[:, 0], g[name + "_v"][k], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(logp[:, 0], g[name + "_logp"][k], rtol=2e-6, atol=1e-6)
        assert np.array_equal(a[:, 0], g[name + "_a"][k])
    lam = [g[f"{name}_lam{i}"] for i in range(6)]
    obs = g[name + "_obs"].astype(F32) / (1.0 if one_hot else norm)
    np.testing.assert_allclose(onets.lambda_net_forward(lam, obs), g[name + "_lam_out"], rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["a", "b"])
def test_buffer(golden, name):
    g = golden("buffer")
    gamma, lam = g[name + "_hyper"]
    tr = lambda z: np.ascontiguousarray(np.asarray(z, F32).T[:, :, None])     # [T,N] -> [N,T,1]
    adv, ret, q, cdcost = obuf.finish_path(tr(g[name + "_rew"]), tr(g[name + "_cost"]), tr(g[name + "_val"]),
                                           np.asarray(g[name + "_last"])[:, None],
                                           np.asarray(g[name + "_lamb"], F32)[:, None], gamma, lam)
    assert np.array_equal(adv[:, :, 0].T, g[name + "_adv_raw"])
    assert np.array_equal(ret[:, :, 0].T, g[name + "_get_ret"])
    assert np.array_equal(q[:, :, 0].T, g[name + "_get_qs"])
    assert np.array_equal(cdcost[:, 0], g[name + "_get_costs"])
    np.testing.assert_allclose(obuf.normalize_adv(adv)[:, :, 0].T, g[name + "_get_adv"], rtol=1e-6, atol=1e-7)


def test_select(golden):
    g = golden("select")
    for ci in range(int(g["n_cases"])):
        p, C, B = g[f"c{ci}_probs"], g[f"c{ci}_C"], float(g[f"c{ci}_B"])
        want = g[f"c{ci}_multi"]
        assert np.array_equal(osel.greedy_multiaction(p, C, B), want), ci
        assert np.array_equal(osel.greedy_multiaction_stepwise(p, C, B), want), ci
        assert (C[want].sum() <= B) or not want.any()
        if p.shape[1] == 2:
            assert np.array_equal(osel.topk_binary(p[:, 1], C, B), g[f"c{ci}_binary"]), ci
            assert np.array_equal(want, g[f"c{ci}_binary"]), ci     # binary multi == top-B


def test_select_known_answers():
    p = np.array([[.05, .05, .9], [.2, .7, .1], [.6, .3, .1], [.3, .4, .3]])
    C = np.array([0, 1, 2])
    assert osel.greedy_multiaction(p, C, 1).tolist() == [0, 0, 0, 0]
    assert osel.greedy_multiaction(p, C, 3).tolist() == [2, 1, 0, 0]
    assert osel.greedy_multiaction(p, C, 4).tolist() == [2, 1, 0, 1]


def test_select_stepwise_property():
    rs = np.random.RandomState(3)
    for trial in range(150):
        N = rs.randint(1, 25)
        A = rs.randint(2, 5)
        C = np.arange(A)
        p = rs.dirichlet(np.ones(A) * rs.choice([0.2, 1, 5]), size=N).astype(F32)
        if trial % 5 == 0:                       # exact ties
            p = np.round(p * 4) / 4
        B = rs.randint(0, 2 * N + 2)
        a1 = osel.greedy_multiaction(p, C, B)
        a2 = osel.greedy_multiaction_stepwise(p, C, B)
        assert np.array_equal(a1, a2), (trial, p, B)
        assert C[a1].sum() <= B or not a1.any()


def test_vi_docstring_known_answers(golden):
    g = golden("vi")
    P, R = g["forest_P"], g["forest_R"]
    Rm = np.stack([R[:, a] for a in range(2)])
    V, pol, it, _ = ovi.value_iteration(P, Rm, 0.96, 0.01, stop="fast")
    assert np.allclose(V, (5.93215488, 9.38815488, 13.38815488), atol=1e-8) and it == 4 and pol.tolist() == [0, 0, 0]
    P2 = np.array([[[0.5, 0.5], [0.8, 0.2]], [[0, 1], [0.1, 0.9]]])
    R2 = np.array([[5, 10], [-1, 2]], float)
    V, pol, it, _ = ovi.value_iteration(P2, R2.T.copy(), 0.9, 0.01, stop="fast")
    assert np.allclose(V, (40.048625392716815, 33.65371175967546), atol=1e-12) and it == 26 and pol.tolist() == [1, 0]
    V, pol, it, _ = ovi.value_iteration(P, Rm, 0.96, 0.01, stop="full")
    assert np.allclose(V, g["forest_V_full"], atol=1e-10) and it == int(g["forest_iter_full"])


@pytest.mark.parametrize("dom,eps", [("ce", 1e-1), ("ce", 1e-8), ("armman", 1e-4), ("armman", 1e-8), ("sis", None)])
def test_vi_probes(golden, dom, eps):
    g = golden("vi")
    lambdas = g["lambdas"] if dom != "sis" else g["lambdas"][:5]
    sfx = f"_{eps}" if eps is not None else ""
    V, pol, it, mi = ovi.batched_vi(g[dom + "_T"], g[dom + "_R"], g[dom + "_C"], lambdas, 0.9, eps or 1e-4)
    assert np.array_equal(it, g[f"{dom}_it{sfx}"])
    assert np.array_equal(mi, g[f"{dom}_mi{sfx}"])
    ok = it >= 0
    np.testing.assert_allclose(V[ok], g[f"{dom}_V{sfx}"][ok], rtol=1e-12, atol=1e-12)
    assert np.array_equal(pol[ok], g[f"{dom}_pol{sfx}"][ok])


@pytest.mark.parametrize("name,domain", [("ce", 0), ("armman", 1)])
def test_epoch_loop_matches_reference_best_response(golden, name, domain):
    """oracle/train.py against the unmodified reference AgentOracle.best_response run with the same injected
    uniforms (oracle/gen_golden_train.py): per-epoch logged values and the weights after all epochs."""
    from oracle.train import CpuRMABPPO
    g = golden("train")
    N, S, T, epochs, hidden, seed = [int(x) for x in g[name + "_meta"]]
    kw = dict(zip([str(k) for k in g[name + "_hyper_names"]], [float(v) for v in g[name + "_hyper_values"]]))
    lam0 = [g[f"{name}_lam0_{i}"] for i in range(6)]
    R = oenv.rewards_table(domain, N, S)
    cpu = CpuRMABPPO(domain, g[name + "_Wpi0"], g[name + "_Wv0"], lam0, g[name + "_nature"].astype(F32),
                     g[name + "_ranges"].astype(F32), R, np.array([0, 1], F32), 1, T, hidden, float(g[name + "_B"]),
                     gamma=kw["gamma"], lam_other=kw["lam_OTHER"], clip_ratio=kw["clip_ratio"], pi_lr=kw["pi_lr"],
                     vf_lr=kw["vf_lr"], lm_lr=kw["lm_lr"], train_pi_iters=int(kw["train_pi_iters"]),
                     train_v_iters=int(kw["train_v_iters"]), epochs=epochs, lamb_update_freq=int(kw["lamb_update_freq"]),
                     final_train_lambdas=int(kw["final_train_lambdas"]), start_entropy_coeff=kw["start_entropy_coeff"],
                     end_entropy_coeff=kw["end_entropy_coeff"], seed=seed)
    lamb, ret, vv = [], [], []
    for _ in range(epochs):
        w = cpu.epoch()
        lamb.append(w["lamb"][0]); ret.append(w["rew_sum"][0]); vv.append(w["val"].mean())
    assert np.array_equal(np.array(ret), g[name + "_AverageEpActualRet"])      # integer path: identical trajectories
    np.testing.assert_allclose(lamb, g[name + "_AverageLamb"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(vv, g[name + "_AverageVVals"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(cpu.opt.Wpi.detach().numpy(), g[name + "_Wpi1"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(cpu.opt.Wv.detach().numpy(), g[name + "_Wv1"], rtol=0, atol=2e-5)
    for i, p in enumerate(cpu.lam_params):
        np.testing.assert_allclose(p.detach().numpy(), g[f"{name}_lam1_{i}"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", ["h16", "h64"])
def test_epoch_loop_sis_matches_reference_best_response(golden, name):
    """oracle/train.py on the SIS domain (S = pop + 1, A = 3, C = [0,1,2], scalar state input s / state_norm) against the
    unmodified reference AgentOracle.best_response(data='sis') with every draw injected (gen_train_sis)."""
    g = golden("train_sis")
    cpu, epochs = sis_rmabppo_oracle(g, name)
    lamb, ret, vv = [], [], []
    for _ in range(epochs):
        w = cpu.epoch()
        lamb.append(w["lamb"][0]); ret.append(w["rew_sum"][0]); vv.append(w["val"].mean())
    np.testing.assert_allclose(ret, g[name + "_AverageEpActualRet"], rtol=1e-6)   # same trajectories (R = linspace: fp32 sums)
    np.testing.assert_allclose(lamb, g[name + "_AverageLamb"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(vv, g[name + "_AverageVVals"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(cpu.opt.Wpi.detach().numpy(), g[name + "_Wpi1"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(cpu.opt.Wv.detach().numpy(), g[name + "_Wv1"], rtol=0, atol=2e-5)
    for i, p in enumerate(cpu.lam_params):
        np.testing.assert_allclose(p.detach().numpy(), g[f"{name}_lam1_{i}"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name,domain", [("ce", 0), ("armman", 1)])
def test_ma_step_matches_reference(golden, name, domain):
    """oracle/ma.py against RMABLambdaNatureOracle.step of the unmodified reference (injected Normal / Categorical draws)."""
    from oracle import ma as oma
    g = golden("ma")
    N, S, h, D = [int(x) for x in g[name + "_meta"]]
    mu_p = [g[f"{name}_mu_{k}{j}"] for j in range(3) for k in ("W", "b")]
    vn_p = [g[f"{name}_vn_{k}{j}"] for j in range(3) for k in ("W", "b")]
    rg = g[name + "_ranges"]
    for k in range(g[name + "_obs"].shape[0]):
        obs = g[name + "_obs"][k]
        s = obs[:, None].astype(np.uint8)
        lamb = g[name + "_lamb"][k:k + 1]
        mu = oma.dense_mlp(mu_p, obs[None].astype(F32))
        a_nat, lp_nat = oma.gaussian_sample(mu, g[name + "_log_std"], g[name + "_eps"][k][None])
        np.testing.assert_allclose(a_nat[0], g[name + "_out_an"][k], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(lp_nat[0], g[name + "_out_lpn"][k], rtol=1e-5)
        if domain == 0:
            lo, hi = rg[:, 0], rg[:, 1]
        else:
            lo, hi = rg[np.arange(N), obs, :, 0].reshape(-1), rg[np.arange(N), obs, :, 1].reshape(-1)
        nat_b = oenv.bound_nature(a_nat[0], lo, hi)
        np.testing.assert_allclose(nat_b, g[name + "_out_anb"][k], rtol=1e-6, atol=1e-7)
        a, _, logp, p1, _ = onets.ac_step(g[name + "_Wpi"], g[name + "_Wv"], s, lamb, g[name + "_u"][k][:, None], S, 2, h)
        assert np.array_equal(a[:, 0], g[name + "_out_a"][k])
        np.testing.assert_allclose(logp[:, 0], g[name + "_out_logp"][k], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(p1[:, 0], g[name + "_out_p1"][k], rtol=2e-6, atol=1e-7)
        v = oma.agent_critic_extra(g[name + "_Wv"], g[name + "_W1n"], s, lamb, nat_b[None].astype(F32), S, h)
        np.testing.assert_allclose(v[:, 0], g[name + "_out_v"][k], rtol=1e-5, atol=2e-6)
        vn = oma.dense_mlp(vn_p, np.concatenate([obs, a[:, 0]])[None].astype(F32))[0, 0]
        np.testing.assert_allclose(vn, g[name + "_out_vn"][k], rtol=1e-5, atol=1e-6)


def test_ma_epoch_loop_matches_reference_nature_oracle(golden):
    """oracle/ma_train.py against the UNMODIFIED reference NatureOracle.best_response (vs a Pessimist agent) run with
    every draw injected (oracle/gen_golden_train.py::gen_ma_train): logged per-epoch values and the final nets."""
    from oracle.ma_train import CpuMARMABPPO
    g = golden("ma_train")
    N, S, T, epochs, h, seed, D, kw, dense = ma_train_setup(g)
    cpu = CpuMARMABPPO(0, g["Wpi0"], g["Wv0"], g["W1n0"], [g[f"lam0_{i}"] for i in range(6)], dense("mu0", [N, h, h, D]),
                       g["log_std0"], dense("vn0", [2 * N, h, h, 1]), g["ranges"], oenv.rewards_table(0, N, S),
                       np.array([0, 1], F32), 1, T, h, float(g["B"]), kw["gamma"], kw["lam_OTHER"], kw["clip_ratio"],
                       kw["pi_lr_agent"], kw["pi_lr_nature"], kw["vf_lr_agent"], kw["vf_lr_nature"], kw["lm_lr"],
                       int(kw["train_pi_iters"]), int(kw["train_v_iters"]), epochs, int(kw["lamb_update_freq"]),
                       int(kw["final_train_lambdas"]), kw["start_entropy_coeff"], kw["end_entropy_coeff"], seed)
    lamb, ret, rn, vv, vn = [], [], [], [], []
    for _ in range(epochs):
        w = cpu.epoch(lambda s: np.zeros_like(s))
        lamb.append(w["lamb"][0]); ret.append(w["ret_agent"][0]); rn.append(w["rew_nat"].sum())
        vv.append(w["val"].mean()); vn.append(w["val_nat"].mean())
    assert np.array_equal(np.array(ret), g["EpActualRetAgent"])                   # same trajectories
    np.testing.assert_allclose(lamb, g["Lamb"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(rn, g["EpLambAdjRetNature"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(vv, g["VVals_agent"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(vn, g["VVals_nature"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(cpu.opt.Wpi.detach().numpy(), g["Wpi1"], atol=2e-5)
    np.testing.assert_allclose(cpu.opt.Wv.detach().numpy(), g["Wv1"], atol=2e-5)
    np.testing.assert_allclose(cpu.W1n.detach().numpy(), g["W1n1"], atol=2e-5)
    np.testing.assert_allclose(cpu.log_std.detach().numpy(), g["log_std1"], atol=2e-5)
    for i, p in enumerate(cpu.mu_net.parameters()):
        np.testing.assert_allclose(p.detach().numpy(), g[f"mu1_{i}"], atol=2e-5)
    for i, p in enumerate(cpu.v_nature.parameters()):
        np.testing.assert_allclose(p.detach().numpy(), g[f"vn1_{i}"], atol=2e-5)
    for i, p in enumerate(cpu.lam_params):
        np.testing.assert_allclose(p.detach().numpy(), g[f"lam1_{i}"], rtol=1e-5, atol=1e-7)


def test_ma_epoch_loop_sis_matches_reference_nature_oracle(golden):
    """BASELINE configs[3]'s domain: oracle/ma_train.py on SIS (S = pop + 1, A = 3, D = 4N, hidden 64, scalar inputs)
    against the UNMODIFIED reference NatureOracle.best_response with every draw injected (gen_ma_train_sis)."""
    g = golden("ma_train_sis")
    cpu, epochs = sis_ma_oracle(g)
    lamb, ret, rn, vv, vn = [], [], [], [], []
    for _ in range(epochs):
        w = cpu.epoch(lambda s: np.zeros_like(s))
        lamb.append(w["lamb"][0]); ret.append(w["ret_agent"][0]); rn.append(w["rew_nat"].sum())
        vv.append(w["val"].mean()); vn.append(w["val_nat"].mean())
    np.testing.assert_allclose(ret, g["EpActualRetAgent"], rtol=1e-6)             # same trajectories
    np.testing.assert_allclose(lamb, g["Lamb"], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(rn, g["EpLambAdjRetNature"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(vv, g["VVals_agent"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(vn, g["VVals_nature"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(cpu.opt.Wpi.detach().numpy(), g["Wpi1"], atol=2e-5)
    np.testing.assert_allclose(cpu.opt.Wv.detach().numpy(), g["Wv1"], atol=2e-5)
    np.testing.assert_allclose(cpu.W1n.detach().numpy(), g["W1n1"], atol=2e-5)
    np.testing.assert_allclose(cpu.log_std.detach().numpy(), g["log_std1"], atol=2e-5)
    for i, p in enumerate(cpu.mu_net.parameters()):
        np.testing.assert_allclose(p.detach().numpy(), g[f"mu1_{i}"], atol=2e-5)
    for i, p in enumerate(cpu.v_nature.parameters()):
        np.testing.assert_allclose(p.detach().numpy(), g[f"vn1_{i}"], atol=2e-5)
    for i, p in enumerate(cpu.lam_params):
        np.testing.assert_allclose(p.detach().numpy(), g[f"lam1_{i}"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("seed", [0])
def test_whittle_bisect_pinned_to_index_lp(seed):
    """The indifference-point index (oracle bisection) against a HiGHS restatement of the reference's index LP
    (lp_methods.py:111-212, oracle/vi.py::whittle_lp_highs; Gurobi itself is absent).  Wherever the LP's optimum is the
    Bellman solution the two agree to 1e-6 (all of the counterexample domain); every other entry is one of the
    documented ways the LP leaves the Bellman solution -- never an unexplained difference."""
    from oracle import vi as ovi
    tot = {}
    for name, T, R, C, lim in ovi.whittle_pin_instances(seed):
        idx = ovi.whittle_bisect(T, R, C, 0.9, 1e-10, 0.0, lim, 36, stop="abs")
        cnt, bad = ovi.whittle_lp_pin(idx, T, R, C, 0.9, lim)
        assert cnt["mismatch"] == 0, (name, bad)
        if name == "ce":
            assert cnt["agree"] == idx.size - cnt["quirk"]
        for k, v in cnt.items():
            tot[k] = tot.get(k, 0) + v
    assert tot["agree"] >= 0.75 * (sum(tot.values()) - tot["quirk"])
