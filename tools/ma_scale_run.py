"""BASELINE configs[3] shape: SIS epidemic domain with the MA-RMABPPO nature oracle, N = 2048 arms, pop = 50
(S = 51, A = 3, D = 4N = 8192 nature parameters), E = 256 envs, T = 20, hidden [64,64], Pessimist agent strategy.
Scale check with per-stage CUDA-event timings:   python tools/ma_scale_run.py [N] [E]
Arm-sharded over k GPUs (strong scaling, N fixed): python -m torch.distributed.run --nnodes=1 --nproc-per-node k
    --master-addr 127.0.0.1 --master-port 29551 tools/ma_scale_run.py [N] [E]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustrmab_b200.environments import SISRobustEnv          # noqa: E402
from robustrmab_b200.evaluate import PessimisticAgentPolicy    # noqa: E402
from robustrmab_b200.ma_core import RMABLambdaNatureOracle     # noqa: E402
from robustrmab_b200.mpi_tools import arm_shard                # noqa: E402
from robustrmab_b200.nature_oracle import MARolloutTrainer     # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = 0
dev = torch.device("cuda", 0)
if world > 1:
    dist.init_process_group("nccl")
    rank = dist.get_rank()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
torch.cuda.set_device(dev)
pop, T, h, v_iters, pi_iters = 50, 20, 64, 4, 4
env = SISRobustEnv(N, 0.2 * N, pop, 0, n_envs=E, device=dev)
env.random_stream.seed(0)
env.sampled_parameter_ranges = env.sample_parameter_ranges()
torch.manual_seed(0)
a0, n_loc = arm_shard(N, rank, world)
ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, env.action_dim_nature,
                            env, hidden_sizes=[h, h], C=env.C, N=N, B=env.B, one_hot_encode=False, non_ohe_obs_dim=1,
                            state_norm=pop, nature_state_norm=1, n_envs=E, device=dev,
                            arm_shard=(a0, a0 + n_loc) if world > 1 else None)
tr = MARolloutTrainer(env, ac, T, 0.9, 0.97, 2.0, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, pi_iters, v_iters, 4, 1, 0, 0.5, 0.0, 0,
                      rank=rank, world_size=world)
pess = PessimisticAgentPolicy(N)
tr.profile = True


def timed(fn):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return out, float(ms.item())


for ep in range(int(os.environ.get("MA_EPOCHS", "3"))):
    (lamb, h1, s0, lv, lvn, ret), ms_roll = timed(lambda: tr.rollout(pess))
    _, ms_upd = timed(lambda: tr.update(lamb, h1, s0, lv, lvn))
    tr.s = tr.env.reset_device()
    ok = bool(torch.isfinite(ac.Wpi).all() and torch.isfinite(ac.Wv).all() and torch.isfinite(ac.W1n).all())
    if rank == 0:
        print(f"[{world} GPU] epoch {ep}: rollout {ms_roll:.1f} ms ({N * E * T / ms_roll / 1e3:.1f} M arm-steps/s incl. 50 counterfactual "
              f"steps each), update ({pi_iters}+{v_iters} iters, nature nets every epoch) {ms_upd:.1f} ms, "
              f"epoch {N * E * T / (ms_roll + ms_upd) / 1e3:.2f} M arm-steps/s, finite={ok}, mean return/env {float(ret.mean()):.1f}, "
              f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        print("   update stages (ms, rank 0):", {k: round(v, 1) for k, v in tr.stages.report().items()}, flush=True)
    assert ok and int(tr.s_traj.max()) <= pop
if world > 1:
    dist.destroy_process_group()
