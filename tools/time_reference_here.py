"""Times the UNMODIFIED reference classes (AgentOracle.best_response, robust_rmab/algos/rmabppo/agent_oracle.py) and the
oracle port on the same small problem, on THIS container's CPU (the reference is pure Python under /root/reference and
cannot travel to the GPU box: `pip install --target` of it yields an empty wheel, its setup.py packages nothing).
Writes profiles/r2_reference_classes_cpu.json, which bench.py quotes next to the port's rate.

    python tools/time_reference_here.py
"""
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "stubs"))
sys.path.insert(0, REF)

import torch  # noqa: E402

from oracle import ref_shims  # noqa: E402

torch.set_num_threads(1)
N, T, epochs, hidden, seed = 30, 100, 3, 16, 0
KW = dict(gamma=0.9, clip_ratio=2.0, pi_lr=2e-3, vf_lr=2e-3, lm_lr=2e-3, train_pi_iters=20, train_v_iters=20, lam_OTHER=0.97,
          start_entropy_coeff=0.5, end_entropy_coeff=0.0, lamb_update_freq=4, final_train_lambdas=2)


def run_reference():
    from robust_rmab.algos.rmabppo import agent_oracle as ao
    from robust_rmab.baselines.nature_baselines_counterexample import SampledRandomNaturePolicy
    from robust_rmab.environments.bandit_env_robust import CounterExampleRobustEnv
    home = tempfile.mkdtemp(prefix="rmab_time_")
    try:
        env0 = CounterExampleRobustEnv(N, N // 3, seed)
        ranges = env0.sample_parameter_ranges()
        agent_kwargs = dict(steps_per_epoch=T, epochs=epochs, ac_kwargs=dict(hidden_sizes=[hidden, hidden]), **KW)
        orc = ao.AgentOracle("counterexample", N, 2, 2, N // 3, seed, 1, agent_kwargs=agent_kwargs, home_dir=home,
                             exp_name="timing", sampled_nature_parameter_ranges=ranges, robust_keyword="sample_random")
        nature = SampledRandomNaturePolicy(ranges, 0)
        nature.sample_param_setting(seed)
        torch.manual_seed(seed)
        np.random.seed(seed)
        t0 = time.perf_counter()
        with ref_shims.quiet():
            orc.best_response([nature], [1.0], 0)
        return time.perf_counter() - t0
    finally:
        shutil.rmtree(home, ignore_errors=True)


def run_port():
    import argparse
    import bench
    a = argparse.Namespace(domain="counterexample", hidden=hidden, horizon=T, pi_iters=20, v_iters=20)
    from oracle.train import CpuRMABPPO
    dom, S, Wpi, Wv, lam, nature, ranges, R, C = bench._cpu_problem(a, N, 1, seed)
    tr = CpuRMABPPO(dom, Wpi, Wv, lam, nature, ranges, R, C, 1, T, hidden, N // 3, train_pi_iters=20, train_v_iters=20, seed=seed)
    tr.epoch()
    t0 = time.perf_counter()
    for _ in range(epochs):
        tr.epoch()
    return time.perf_counter() - t0


dt_ref = run_reference()
dt_port = run_port()
steps = N * T * epochs
out = {"problem": f"counterexample, N={N} arms, 1 env, T={T}, hidden [{hidden},{hidden}], 20+20 iters, {epochs} epochs, 1 thread",
       "reference_classes_arm_steps_per_s": steps / dt_ref, "reference_s": dt_ref,
       "oracle_port_arm_steps_per_s": steps / dt_port, "port_s": dt_port, "port_over_reference": dt_ref / dt_port,
       "host": "build container CPU (not the GPU box)", "how": "tools/time_reference_here.py"}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_reference_classes_cpu.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
