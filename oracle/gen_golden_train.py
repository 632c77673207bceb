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
    t = e*T + step; the extra bootstrap ac.step at the end of every epoch (agent_oracle.py:553) consumes draws that
    do not influence anything and gets a throw-away stream."""

    def __init__(self, seed, n_arms, T):
        super().__init__(seed, philox.STREAM_ACT, n_arms)
        self.T, self.sweep = T, 0

    def next(self):
        k = self.sweep % (self.T + 1)
        ep = self.sweep // (self.T + 1)
        if k == self.T:
            u = np.float32(0.5)
        else:
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


def _run(data, N, B, T, epochs, hidden, seed, kw):
    import torch
    from robust_rmab.algos.rmabppo import agent_oracle as ao
    from robust_rmab.algos.rmabppo import rmabppo_core as core
    if data == "counterexample":
        from robust_rmab.environments.bandit_env_robust import CounterExampleRobustEnv as Env
        from robust_rmab.baselines.nature_baselines_counterexample import SampledRandomNaturePolicy
        S = 2
    else:
        from robust_rmab.environments.bandit_env_robust import ARMMANRobustEnv as Env
        from robust_rmab.baselines.nature_baselines_armman import SampledRandomNaturePolicy
        S = 3
    rec = _REC
    rec.clear()
    Recorder = _recorder_class(core)

    home = tempfile.mkdtemp(prefix="rmab_golden_")
    try:
        env0 = Env(N, B, seed)
        ranges = env0.sample_parameter_ranges().astype(F32).astype(np.float64)
        agent_kwargs = dict(steps_per_epoch=T, epochs=epochs, ac_kwargs=dict(hidden_sizes=[hidden, hidden]),
                            actor_critic=Recorder, **kw)
        orc = ao.AgentOracle(data, N, S, 2, B, seed, 1, agent_kwargs=agent_kwargs, home_dir=home, exp_name="golden",
                             sampled_nature_parameter_ranges=ranges, robust_keyword="sample_random")
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
        out = dict(meta=np.array([N, S, T, epochs, hidden, seed]), B=np.array(B), ranges=ranges,
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


def gen_ma():
    raise NotImplementedError("MA-RMABPPO golden vectors: next round")
