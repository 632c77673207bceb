"""BASELINE configs[3] shape: SIS epidemic domain with the MA-RMABPPO nature oracle, N = 2048 arms, pop = 50
(S = 51, A = 3, D = 4N = 8192 nature parameters), E = 256 envs, T = 20, hidden [64,64], Pessimist agent strategy.
One-off scale check with per-stage CUDA-event timings (python tools/ma_scale_run.py [N] [E])."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from robustrmab_b200.environments import SISRobustEnv          # noqa: E402
from robustrmab_b200.evaluate import PessimisticAgentPolicy    # noqa: E402
from robustrmab_b200.ma_core import RMABLambdaNatureOracle     # noqa: E402
from robustrmab_b200.nature_oracle import MARolloutTrainer     # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
pop, T, h, v_iters, pi_iters = 50, 20, 64, 4, 4
env = SISRobustEnv(N, 0.2 * N, pop, 0, n_envs=E)
env.sampled_parameter_ranges = env.sample_parameter_ranges()
torch.manual_seed(0)
ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, env.action_dim_nature,
                            env, hidden_sizes=[h, h], C=env.C, N=N, B=env.B, one_hot_encode=False, non_ohe_obs_dim=1,
                            state_norm=pop, nature_state_norm=1, n_envs=E)
tr = MARolloutTrainer(env, ac, T, 0.9, 0.97, 2.0, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, pi_iters, v_iters, 4, 1, 0, 0.5, 0.0, 0)
pess = PessimisticAgentPolicy(N)
tr.profile = True


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


for ep in range(3):
    (lamb, h1, s0, lv, lvn, ret), ms_roll = timed(lambda: tr.rollout(pess))
    _, ms_upd = timed(lambda: tr.update(lamb, h1, s0, lv, lvn))
    tr.s = env.reset_device()
    ok = bool(torch.isfinite(ac.Wpi).all() and torch.isfinite(ac.Wv).all() and torch.isfinite(ac.W1n).all())
    print(f"epoch {ep}: rollout {ms_roll:.1f} ms ({N * E * T / ms_roll / 1e3:.1f} M arm-steps/s incl. 50 counterfactual steps each), "
          f"update ({pi_iters}+{v_iters} iters, nature nets every epoch) {ms_upd:.1f} ms, finite={ok}, "
          f"mean return/env {float(ret.mean()):.1f}, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    print("   update stages (ms):", {k: round(v, 1) for k, v in tr.stages.report().items()})
    assert ok and int(tr.s_traj.max()) <= pop
