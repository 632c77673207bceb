"""Fused table-domain epoch (rmab_epoch_table) and the statistics-based PPO update (rmab_ppo_update_table)
against the per-sample CPU oracle, and the device trainer against the CPU epoch loop.  GPU only."""
import numpy as np
import pytest

from oracle import buffer as obuf
from oracle import envs as oenv
from oracle import nets as onets
from oracle import ppo as oppo
from oracle.train import CpuRMABPPO

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


def _problem(domain, N, E, h, seed):
    rs = np.random.RandomState(seed)
    S = 2 if domain == 0 else 3
    base = oenv.ce_param_ranges(N) if domain == 0 else oenv.armman_param_ranges(N)
    ranges = oenv.sample_sub_ranges(base, rs).astype(F32)
    nature = (ranges[..., 0] + rs.rand(*ranges.shape[:-1]) * (ranges[..., 1] - ranges[..., 0])).astype(F32)
    Wpi = onets.init_linear_default(rs, N, S + 1, h, 2)
    Wv = onets.init_linear_default(rs, N, S + 1, h, 1)
    b = 1 / np.sqrt(N)
    lam = [rs.uniform(-b, b, (8, N)).astype(F32), rs.uniform(-b, b, 8).astype(F32),
           rs.uniform(-.35, .35, (8, 8)).astype(F32), rs.uniform(-.35, .35, 8).astype(F32),
           rs.uniform(-.35, .35, (1, 8)).astype(F32), rs.uniform(0.2, .5, 1).astype(F32)]
    R = oenv.rewards_table(domain, N, S)
    C = np.array([0, 1], F32)
    return S, ranges, nature, Wpi, Wv, lam, R, C


@pytest.mark.parametrize("domain,h,N,E,T", [(0, 16, 9, 40, 25), (1, 64, 5, 128, 10), (1, 8, 700, 3, 6),
                                            (0, 32, 4, 300, 12),
                                            # hidden 64: the table pre-pass runs on tcgen05 (table_fwd_umma64.cu); ragged tiles
                                            (0, 64, 4, 96, 12), (1, 64, 3, 171, 7), (0, 64, 2, 1, 9),
                                            # on-chip trajectory words: T = 32 / 33 / 64 / 100 cross 32-step word boundaries
                                            (0, 16, 3, 33, 32), (1, 16, 3, 33, 33), (1, 8, 2, 70, 64), (0, 16, 2, 256, 100),
                                            (1, 16, 2, 64, 131),
                                            # too long for shared memory: the trajectory is re-read from HBM
                                            (1, 8, 2, 300, 3000)])
def test_epoch_table_vs_oracle(K, domain, h, N, E, T):
    S, ranges, nature, Wpi, Wv, lam, R, C = _problem(domain, N, E, h, 11 + h)
    cpu = CpuRMABPPO(domain, Wpi, Wv, lam, nature, ranges, R, C, E, T, h, N // 3, train_pi_iters=0, train_v_iters=0,
                     seed=77, arm0=3)
    w = cpu.epoch()
    torch = K.torch
    net = K.ops.net_desc(N, E, S, 2, h)
    s_traj = torch.zeros((N, T + 1, E), dtype=torch.uint8, device=K.dev)
    s_traj[:, 0] = K.t(w["s_traj"][:, 0])
    a = torch.empty((N, T, E), dtype=torch.uint8, device=K.dev)
    adv = torch.empty((N, T, E), dtype=torch.float32, device=K.dev)
    ret = torch.empty((N, T, E), dtype=torch.float32, device=K.dev)
    last = torch.empty((N, E), dtype=torch.float32, device=K.dev)
    stats, arm_stats, env_sums = K.ops.epoch_table(domain, net, T, K.t(Wpi), K.t(Wv), K.t(w["lamb"]), K.t(nature),
                                                   K.t(ranges), K.t(R), K.t(C), 0.9, 0.97, 77, 0, 0, s_traj, a,
                                                   adv=adv, ret=ret, last_val=last, arm0=3)
    assert np.array_equal(K.n(s_traj), w["s_traj"])                # integer outputs bit-exact
    assert np.array_equal(K.n(a), w["a"])
    np.testing.assert_allclose(K.n(last), w["last_val"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(K.n(ret), w["ret"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(K.n(adv), w["adv"], rtol=1e-05, atol=2e-06)
    # statistics rows against the oracle's regrouping of its own per-sample arrays
    st = oppo.loss_statistics(w["s_traj"][:, :T], w["a"], w["adv"], w["ret"], S, 2)
    g = K.n(stats)
    SA = S * 2
    np.testing.assert_allclose(g[:, 0:SA], st["Ap"], rtol=1e-05, atol=2e-05)
    np.testing.assert_allclose(g[:, SA:2 * SA], st["Am"], rtol=1e-05, atol=2e-05)
    assert np.array_equal(g[:, 3 * SA + 3 * S:], st["cnt"])
    assert np.array_equal(g[:, 3 * SA:3 * SA + S], st["cntS"])
    np.testing.assert_allclose(g[:, 3 * SA + S:3 * SA + 2 * S], st["mean_ret"], rtol=1e-5, atol=1e-5)
    # logp_old / v_old rows reproduce the per-sample logp / val of the rollout
    lp_rows = g[:, 2 * SA:3 * SA]
    cell = (w["s_traj"][:, :T].astype(np.int64) * 2 + w["a"]).astype(np.int64)
    np.testing.assert_allclose(np.take_along_axis(lp_rows, cell, axis=1), w["logp"], rtol=1e-5, atol=2e-6)
    v_rows = g[:, 3 * SA + 2 * S:3 * SA + 3 * S]
    np.testing.assert_allclose(np.take_along_axis(v_rows, w["s_traj"][:, :T].astype(np.int64), axis=1), w["val"],
                               rtol=1e-5, atol=2e-6)
    raw_adv, _, _, cd = obuf.finish_path(
        np.stack([np.take_along_axis(R, w["s_traj"][:, t + 1].astype(np.int64), axis=1) for t in range(T)], 1),
        C[w["a"]], w["val"], w["last_val"], w["lamb"], 0.9, 0.97)
    am = K.n(arm_stats)
    np.testing.assert_allclose(am[:, 0], raw_adv.reshape(N, -1).astype(np.float64).mean(1), rtol=1e-05, atol=1e-06)
    np.testing.assert_allclose(am[:, 1], raw_adv.reshape(N, -1).astype(np.float64).std(1), rtol=1e-5)
    np.testing.assert_allclose(am[:, 2], st["ret_var_sum"], rtol=1e-05, atol=1e-05)
    es = K.n(env_sums)
    assert np.array_equal(es[0], w["rew_sum"])                     # rewards are multiples of 0.5: exact
    np.testing.assert_allclose(es[1], w["cdcost"][0], rtol=1e-6)


@pytest.mark.parametrize("domain,h,N,E,T", [(0, 16, 7, 256, 100), (1, 64, 3, 1024, 10), (1, 16, 5, 77, 95), (0, 8, 4, 33, 1)])
def test_epoch_table_onchip_trajectory_equals_materialised(K, domain, h, N, E, T):
    """a == NULL (trajectory only in shared-memory bit planes, s_traj = the [N,E] start states) gives bit-identical
    statistics, per-arm moments and per-env sums to the launch that also writes the byte trajectory to HBM."""
    torch = K.torch
    S, ranges, nature, Wpi, Wv, lam, R, C = _problem(domain, N, E, h, 3 + h)
    rs = np.random.RandomState(T)
    s0 = rs.randint(0, S, (N, E)).astype(np.uint8)
    lamb = rs.uniform(0, 2, E).astype(F32)
    net = K.ops.net_desc(N, E, S, 2, h)
    s_traj = torch.zeros((N, T + 1, E), dtype=torch.uint8, device=K.dev)
    s_traj[:, 0] = K.t(s0)
    a = torch.empty((N, T, E), dtype=torch.uint8, device=K.dev)
    args = (domain, net, T, K.t(Wpi), K.t(Wv), K.t(lamb), K.t(nature), K.t(ranges), K.t(R), K.t(C), 0.9, 0.97, 5, 8, 8)
    last1 = torch.empty((N, E), dtype=torch.float32, device=K.dev)
    last2 = torch.empty((N, E), dtype=torch.float32, device=K.dev)
    st1, as1, es1 = K.ops.epoch_table(*args, s_traj, a, last_val=last1, arm0=2)
    st2, as2, es2 = K.ops.epoch_table(*args, K.t(s0), None, last_val=last2, arm0=2)
    assert torch.equal(st1, st2) and torch.equal(as1, as2) and torch.equal(es1, es2) and torch.equal(last1, last2)
    assert int(s_traj.max()) < S and int(a.max()) <= 1


@pytest.mark.parametrize("S,h,N,T,E", [(2, 16, 4, 30, 8), (3, 64, 3, 10, 32), (3, 8, 3, 7, 5), (3, 32, 2, 12, 70),
                                       (3, 64, 2, 6, 100), (2, 64, 2, 5, 200), (3, 64, 2, 4, 128), (3, 64, 2, 9, 1), (2, 64, 3, 7, 3),
                                       (2, 64, 1, 2, 6500)])   # lambda [E] no longer fits beside the tiles: read from global
def test_ppo_update_table_equals_per_sample(K, S, h, N, T, E):
    """The statistics-based update against the oracle's per-sample torch-autograd update on the same batch."""
    rs = np.random.RandomState(S * 10 + h)
    A = 2
    Wpi = onets.init_linear_default(rs, N, S + 1, h, A)
    Wv = onets.init_linear_default(rs, N, S + 1, h, 1)
    s = rs.randint(0, S, size=(N, T, E)).astype(np.uint8)
    s[0, :, 0] = 0                                                # a state that is never visited in one env
    a = rs.randint(0, A, size=(N, T, E)).astype(np.uint8)
    lamb = rs.uniform(0, 2, size=E).astype(F32)
    adv = obuf.normalize_adv(rs.randn(N, T, E).astype(F32) * 3 + 1)
    ret = (rs.randn(N, T, E) * 2 + 3).astype(F32)
    # logp_old must be a function of (arm, env, s, a), as it is after a rollout: evaluate the old policy
    x = onets.features(s.reshape(N, T * E), np.tile(lamb, T), S, True)
    _, lp_all = onets.softmax_logits(onets.mlp_forward(Wpi, x, S + 1, h, A))
    logp_old = np.take_along_axis(lp_all, a.reshape(N, T * E, 1).astype(np.int64), axis=2)[..., 0].reshape(N, T, E)
    logp_old = (logp_old + rs.uniform(-0.3, 0.3, size=(N, 1, E)) * (s == 1)).astype(F32)   # off-policy ratio != 1
    clip, ent, lr, iters = 0.2, 0.3, 2e-3, 6
    opt = oppo.ArmOptim(Wpi, Wv, lr, lr)
    wlp, wlv = oppo.update(opt, s, a, adv, logp_old, ret, lamb, S, A, h, clip, ent, iters, iters)
    st = oppo.loss_statistics(s, a, adv, ret, S, A)
    SA = S * A
    rows = np.zeros((N, 4 * SA + 3 * S, E), F32)
    rows[:, 0:SA], rows[:, SA:2 * SA] = st["Ap"], st["Am"]
    for si in range(S):
        for ai in range(A):
            m = (s == si) & (a == ai)
            with np.errstate(invalid="ignore"):
                v = np.where(m.any(1), (logp_old * m).sum(1) / np.maximum(m.sum(1), 1), 0.0)
            rows[:, 2 * SA + si * A + ai] = v
    rows[:, 3 * SA:3 * SA + S] = st["cntS"]
    rows[:, 3 * SA + S:3 * SA + 2 * S] = st["mean_ret"]
    rows[:, 3 * SA + 3 * S:] = st["cnt"]
    arm_stats = np.zeros((N, 4), F32)
    arm_stats[:, 2] = st["ret_var_sum"]
    torch = K.torch
    net = K.ops.net_desc(N, E, S, A, h)
    hp = K.ops.hyper(clip, ent, lr, lr, iters, iters)
    dWpi, dWv = K.t(Wpi), K.t(Wv)
    adam_pi = torch.zeros((N, 2, Wpi.shape[1]), device=K.dev)
    adam_v = torch.zeros((N, 2, Wv.shape[1]), device=K.dev)
    lp, lv = K.ops.ppo_update_table(net, hp, T, dWpi, dWv, adam_pi, adam_v, K.t(rows), K.t(arm_stats), K.t(lamb),
                                    want_loss=True)
    np.testing.assert_allclose(K.n(lp)[:, 0], wlp[0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(K.n(lv)[:, 0], wlv[0], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(K.n(lp).T, wlp, rtol=2e-05, atol=2e-06)
    np.testing.assert_allclose(K.n(lv).T, wlv, rtol=2e-05, atol=2e-06)
    np.testing.assert_allclose(K.n(dWpi), opt.Wpi.detach().numpy(), rtol=0, atol=2e-4)
    np.testing.assert_allclose(K.n(dWv), opt.Wv.detach().numpy(), rtol=0, atol=2e-4)


@pytest.mark.parametrize("domain,h", [(0, 16), (1, 64)])
def test_trainer_matches_cpu_epoch_loop(K, domain, h):
    """Device trainer vs the CPU restatement of best_response_per_cpu's loop over several epochs, including a
    lambda-net SGD step (epoch 2 with lamb_update_freq = 2)."""
    from robustrmab_b200.trainer import RMABPPOTrainer
    N, E, T = 12, 16, 20
    kw = dict(gamma=0.9, lam_other=0.97, clip_ratio=2.0, pi_lr=2e-3, vf_lr=2e-3, lm_lr=2e-3, train_pi_iters=5,
              train_v_iters=5, epochs=6, lamb_update_freq=2, final_train_lambdas=1, start_entropy_coeff=0.5,
              end_entropy_coeff=0.0, seed=5)
    tr = RMABPPOTrainer(domain, N, E, T, hidden=h, B=N // 3, device=K.dev, keep_trajectory=True, **kw)
    S = tr.S
    lam_flat = K.n(tr.lam_params)
    lam = [lam_flat[:8 * N].reshape(8, N), lam_flat[8 * N:8 * N + 8], lam_flat[8 * N + 8:8 * N + 72].reshape(8, 8),
           lam_flat[8 * N + 72:8 * N + 80], lam_flat[8 * N + 80:8 * N + 88].reshape(1, 8), lam_flat[8 * N + 88:]]
    cpu = CpuRMABPPO(domain, K.n(tr.Wpi), K.n(tr.Wv), lam, K.n(tr.nature), K.n(tr.ranges), K.n(tr.R), K.n(tr.C), E, T, h,
                     N // 3, **kw)
    agree = []
    for ep in range(4):
        st = tr.epoch()
        w = cpu.epoch()
        np.testing.assert_allclose(K.n(st["Lamb"]), w["lamb"], rtol=1e-05, atol=1e-06)
        assert np.array_equal(K.n(tr.s_traj[:, 0]), oenv.reset_states(5, ep + 1, S, E, N))
        agree.append((K.n(tr.a) == w["a"]).mean())
        if ep == 0:
            assert np.array_equal(K.n(tr.a), w["a"])
            assert np.array_equal(K.n(st["EpActualRet"]), w["rew_sum"])
    # after an update the two fp32 paths differ in the last bits of the weights, so a vanishing fraction of the
    # draws may land on the other side of a cdf edge; anything systematic would show up as a large mismatch
    assert min(agree) > 0.995, agree
    np.testing.assert_allclose(K.n(tr.Wpi), cpu.opt.Wpi.detach().numpy(), atol=0.0005)
    np.testing.assert_allclose(K.n(tr.Wv), cpu.opt.Wv.detach().numpy(), atol=2e-4)
    lam_cpu = np.concatenate([p.detach().numpy().reshape(-1) for p in cpu.lam_params])
    np.testing.assert_allclose(K.n(tr.lam_params), lam_cpu, rtol=0.0001, atol=1e-06)


def test_learned_policy_returns_statistically_indistinguishable(K):
    """north_star's third parity criterion: after training, the device trainer's returns are statistically
    indistinguishable from the CPU restatement's.  Three seeds x 24 epochs each; mean per-env return over the last six
    epochs compared with the spread over envs and epochs (the two runs share the Philox convention but drift apart
    after the first updates, so this is a comparison of learned policies, not of identical trajectories)."""
    from robustrmab_b200.trainer import RMABPPOTrainer
    N, E, T, h, epochs = 12, 32, 20, 16, 24
    diffs, ses, gains = [], [], []
    for seed in (1, 2, 3):
        kw = dict(gamma=0.9, lam_other=0.97, clip_ratio=2.0, pi_lr=2e-3, vf_lr=2e-3, lm_lr=2e-3, train_pi_iters=10,
                  train_v_iters=10, epochs=epochs, lamb_update_freq=4, final_train_lambdas=2, start_entropy_coeff=0.5,
                  end_entropy_coeff=0.0, seed=seed)
        tr = RMABPPOTrainer(0, N, E, T, hidden=h, B=N // 3, device=K.dev, **kw)
        lam_flat = K.n(tr.lam_params)
        lam = [lam_flat[:8 * N].reshape(8, N), lam_flat[8 * N:8 * N + 8], lam_flat[8 * N + 8:8 * N + 72].reshape(8, 8),
               lam_flat[8 * N + 72:8 * N + 80], lam_flat[8 * N + 80:8 * N + 88].reshape(1, 8), lam_flat[8 * N + 88:]]
        cpu = CpuRMABPPO(0, K.n(tr.Wpi), K.n(tr.Wv), lam, K.n(tr.nature), K.n(tr.ranges), K.n(tr.R), K.n(tr.C), E, T, h,
                         N // 3, **kw)
        rg, rc = [], []
        for ep in range(epochs):
            rg.append(K.n(tr.epoch()["EpActualRet"]))
            rc.append(cpu.epoch()["rew_sum"])
        rg, rc = np.array(rg), np.array(rc)                      # [epochs, E]
        tail_g, tail_c = rg[-6:], rc[-6:]
        diffs.append(tail_g.mean() - tail_c.mean())
        ses.append(np.sqrt(tail_g.var() / tail_g.size + tail_c.var() / tail_c.size))
        gains.append((tail_g.mean() - rg[:3].mean(), tail_c.mean() - rc[:3].mean()))
    diffs, ses = np.array(diffs), np.array(ses)
    assert (np.abs(diffs) < 3.0 * ses).all(), (diffs, ses)      # every seed within three standard errors
    assert abs(diffs.mean()) < 3.0 * np.sqrt((ses ** 2).sum()) / len(ses), (diffs, ses)
    # and both actually learned the same amount (return change from the first to the last epochs)
    for gg, gc in gains:
        assert abs(gg - gc) < 3.0 * ses.max() * np.sqrt(2), gains


@pytest.mark.parametrize("name,domain", [("ce", 0), ("armman", 1)])
def test_trainer_matches_reference_best_response(K, golden, name, domain):
    """Device trainer (E = 1) against the values logged by the UNMODIFIED reference AgentOracle.best_response run
    with the same injected uniforms (tests/golden/train.npz, oracle/gen_golden_train.py)."""
    from robustrmab_b200.trainer import RMABPPOTrainer
    g = golden("train")
    N, S, T, epochs, hidden, seed = [int(x) for x in g[name + "_meta"]]
    kw = dict(zip([str(k) for k in g[name + "_hyper_names"]], [float(v) for v in g[name + "_hyper_values"]]))
    tr = RMABPPOTrainer(domain, N, 1, T, hidden=hidden, B=float(g[name + "_B"]), gamma=kw["gamma"],
                        lam_other=kw["lam_OTHER"], clip_ratio=kw["clip_ratio"], pi_lr=kw["pi_lr"], vf_lr=kw["vf_lr"],
                        lm_lr=kw["lm_lr"], train_pi_iters=int(kw["train_pi_iters"]), train_v_iters=int(kw["train_v_iters"]),
                        epochs=epochs, lamb_update_freq=int(kw["lamb_update_freq"]),
                        final_train_lambdas=int(kw["final_train_lambdas"]), start_entropy_coeff=kw["start_entropy_coeff"],
                        end_entropy_coeff=kw["end_entropy_coeff"], seed=seed, ranges=g[name + "_ranges"],
                        nature=g[name + "_nature"], device=K.dev)
    tr.Wpi.copy_(K.t(g[name + "_Wpi0"]))
    tr.Wv.copy_(K.t(g[name + "_Wv0"]))
    tr.lam_params.copy_(K.t(np.concatenate([g[f"{name}_lam0_{i}"].reshape(-1) for i in range(6)])))
    lamb, ret, vv = [], [], []
    for _ in range(epochs):
        st = tr.epoch()
        lamb.append(float(st["Lamb"][0])); ret.append(float(st["EpActualRet"][0])); vv.append(float(st["VVals"][0]))
    assert np.array_equal(np.array(ret), g[name + "_AverageEpActualRet"])       # same trajectories as the reference
    np.testing.assert_allclose(lamb, g[name + "_AverageLamb"], rtol=5e-5, atol=2e-6)
    np.testing.assert_allclose(vv, g[name + "_AverageVVals"], rtol=5e-05, atol=5e-06)
    np.testing.assert_allclose(K.n(tr.Wpi), g[name + "_Wpi1"], rtol=0, atol=5e-4)
    np.testing.assert_allclose(K.n(tr.Wv), g[name + "_Wv1"], rtol=0, atol=5e-05)
    lam1 = np.concatenate([g[f"{name}_lam1_{i}"].reshape(-1) for i in range(6)])
    np.testing.assert_allclose(K.n(tr.lam_params), lam1, rtol=2e-5, atol=1e-6)


def test_full_size_config3_armman_properties(K):
    """BASELINE configs[2] shape on one GPU -- ARMMAN, N = 100 000 arms x 1024 envs, T = 10, hidden [64,64] -- checked
    through size-independent properties (the oracle cannot run this size): state / action ranges, visit counts that
    add up to T, normalised advantages with zero mean and unit variance per arm, budget feasibility of act_test."""
    torch = K.torch
    from robustrmab_b200.trainer import RMABPPOTrainer
    N, E, T = 100000, 1024, 10
    tr = RMABPPOTrainer("armman", N, E, T, hidden=64, B=0.2 * N, train_pi_iters=1, train_v_iters=1, epochs=4, seed=1,
                        device=K.dev, keep_trajectory=True)
    w0 = tr.Wpi[:64].clone()
    st = tr.epoch(keep_buffers=False)
    torch.cuda.synchronize()
    assert int(tr.s_traj.max()) <= 2 and int(tr.a.max()) <= 1
    SA, S = 6, 3
    cnt = tr.stats[:, 3 * SA + 3 * S:]                               # n[s,a] rows
    assert torch.equal(cnt.sum(dim=1), torch.full((N, E), float(T), device=K.dev))
    cntS = tr.stats[:, 3 * SA:3 * SA + S]
    assert torch.equal(cntS.sum(dim=1), torch.full((N, E), float(T), device=K.dev))
    # sum of A+ and A- over all cells and envs = sum of the normalised advantages of the arm = 0; their spread = T*E
    tot = (tr.stats[:, :SA] + tr.stats[:, SA:2 * SA]).sum(dim=(1, 2))
    assert float(tot.abs().max()) < 0.5                               # of T*E = 10240 unit-variance samples
    assert torch.isfinite(tr.arm_stats[:, :3]).all() and float(tr.arm_stats[:, 1].min()) > 0
    assert torch.isfinite(tr.Wpi).all() and torch.isfinite(tr.Wv).all() and not torch.equal(tr.Wpi[:64], w0)
    assert torch.isfinite(st["Lamb"]).all() and st["EpActualRet"].shape == (E,)
    # rewards are in [0, 1]: per-env return of N arms over T steps
    assert float(st["EpActualRet"].min()) >= 0 and float(st["EpActualRet"].max()) <= N * T
    # act_test at full size: exactly floor(B) arms per env get action 1 (binary top-k over 100k arms)
    p1 = torch.rand((N, 64), device=K.dev)
    a = K.ops.budget_select_binary(p1, 1.0, 0.2 * N)
    assert torch.equal(a.sum(dim=0, dtype=torch.int64), torch.full((64,), int(0.2 * N), device=K.dev))
    thr = torch.topk(p1[:, 0], int(0.2 * N)).values.min()
    assert torch.equal(a[:, 0].bool(), p1[:, 0] >= thr)
    del tr
    torch.cuda.empty_cache()


def test_full_size_config5_vi_properties(K):
    """BASELINE configs[4] shape: counterexample, N = 99 999 arms x 64 lambda-probes."""
    torch = K.torch
    N, L = 99999, 64
    rg = oenv.ce_param_ranges(N)
    p = rg.mean(-1)
    T = np.zeros((N, 2, 2, 2))
    T[:, 0] = 0.5
    T[:, 1, 0, 0] = 1.0
    T[:, 1, 1, 1], T[:, 1, 1, 0] = p, 1 - p
    R = np.tile(np.array([0.0, 1.0]), (N, 1))
    C = np.array([0.0, 1.0])
    lambdas = np.linspace(0, 10, L)
    V, pol, it, mi, Q = K.ops.vi_batched(K.t(T), K.t(R), K.t(C), K.t(lambdas), 0.9, 1e-8, want_q=True)
    ok = it >= 0
    assert bool(ok.all()) and torch.equal(it, mi)                     # the reference always stops at _boundIter's bound
    assert bool((V[:, 1:] <= V[:, :-1] + 1e-12).all())                # value is non-increasing in the charge lambda
    assert bool((pol[:, -1] == 0).all())                              # at lambda = lambda_lim nobody acts
    # arms repeat with period 3 (ranges cycle by arm % 3): identical MDPs give identical results
    assert torch.equal(V[0], V[3]) and torch.equal(V[1], V[99997])
    from oracle import vi as ovi
    wV, wpol, wit, _ = ovi.batched_vi(T[:3], R[:3], C, lambdas, 0.9, 1e-8)
    np.testing.assert_allclose(K.n(V[:3]), wV, rtol=1e-12, atol=1e-12)
    assert np.array_equal(K.n(it[:3]), wit)
    idx = K.ops.whittle_bisect(K.t(T), K.t(R), K.t(C), 0.9, 1e-8, 0.0, 10.0, 30)
    assert bool(((idx >= 0) & (idx <= 10)).all())
    assert float(idx[:, 0].max()) < 1e-6                              # state 0 ignores the action: zero index
    widx = ovi.whittle_bisect(T[:3], R[:3], C, 0.9, 1e-8, 0.0, 10.0, 30)
    np.testing.assert_allclose(K.n(idx[:3]), widx, atol=1e-9)
