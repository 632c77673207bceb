"""Arm-sharded MA-RMABPPO == unsharded, on real GPUs.  Under torchrun (one rank per GPU) every rank trains a sharded
MARolloutTrainer and, beside it, an unsharded one from the same seeds, and compares trajectories, nature rewards and nets.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ma_shard_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustrmab_b200.environments import ARMMANRobustEnv, CounterExampleRobustEnv, SISRobustEnv   # noqa: E402
from robustrmab_b200.evaluate import PessimisticAgentPolicy    # noqa: E402
from robustrmab_b200.ma_core import RMABLambdaNatureOracle     # noqa: E402
from robustrmab_b200.mpi_tools import arm_shard                # noqa: E402
from robustrmab_b200.nature_oracle import MARolloutTrainer     # noqa: E402


def make(domain, N, E, dev, shard, rank, world):
    if domain == "sis":
        env = SISRobustEnv(N, 0.2 * N, 10, 0, n_envs=E, device=dev)
        kw = dict(one_hot_encode=False, non_ohe_obs_dim=1, state_norm=10)
    elif domain == "armman":
        env = ARMMANRobustEnv(N, 0.2 * N, 0, n_envs=E, device=dev)
        kw = dict(one_hot_encode=True)
    else:
        env = CounterExampleRobustEnv(N, N // 3, 0, n_envs=E, device=dev)
        kw = dict(one_hot_encode=True)
    env.random_stream.seed(3)
    env.sampled_parameter_ranges = env.sample_parameter_ranges()
    torch.manual_seed(11)
    ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges,
                                env.action_dim_nature, env, hidden_sizes=[16, 16], C=env.C, N=N, B=env.B, n_envs=E,
                                device=dev, arm_shard=shard, **kw)
    tr = MARolloutTrainer(env, ac, 6, 0.9, 0.97, 2.0, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, 3, 3, 4, 1, 0, 0.5, 0.0, 0,
                          rank=rank if shard else 0, world_size=world if shard else 1)
    return env, ac, tr


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    ok = True
    for domain, N, E in (("counterexample", 15, 8), ("armman", 21, 16), ("sis", 19, 12)):
        a0, n = arm_shard(N, rank, world)
        sl = slice(a0, a0 + n)
        _, ac_s, sh = make(domain, N, E, dev, (a0, a0 + n), rank, world)
        _, ac_f, full = make(domain, N, E, dev, None, rank, world)
        pess = PessimisticAgentPolicy(N)
        if domain == "counterexample":
            # a learned RMABPPO strategy as the agent: under sharding it evaluates its local arms only and takes its share of
            # the GLOBAL budget through the sharded radix select (act_test_device_sharded); unsharded it sees all arms
            from robustrmab_b200.core import MLPActorCriticRMAB
            torch.manual_seed(5)
            env_ = full.env
            pess = MLPActorCriticRMAB(env_.observation_space, env_.action_space, hidden_sizes=[16, 16], C=env_.C, N=N, B=env_.B,
                                      n_envs=E, device=dev)
        for ep in range(3):
            st_s, st_f = sh.epoch(pess), full.epoch(pess)
            same_s = torch.equal(sh.s_traj, full.s_traj[sl])
            same_a = torch.equal(sh.a, full.a[sl])
            d_rn = float((sh.rew_nat - full.rew_nat).abs().max())
            d_lam = float((st_s["Lamb"] - st_f["Lamb"]).abs().max())
            d_ret = float((st_s["EpActualRetAgent"] - st_f["EpActualRetAgent"]).abs().max())
            d_pi = float((ac_s.Wpi - ac_f.Wpi[sl]).abs().max())
            d_w1n = float((ac_s.W1n - ac_f.W1n[sl]).abs().max())
            d_nat = max(float((p - q).abs().max()) for p, q in zip(ac_s.pi_nature.parameters(), ac_f.pi_nature.parameters()))
            d_vn = max(float((p - q).abs().max()) for p, q in zip(ac_s.v_nature.parameters(), ac_f.v_nature.parameters()))
            print(f"[rank {rank}] {domain} N={N} E={E} epoch {ep}: states equal {same_s}, actions equal {same_a}, |d r_nature| {d_rn:.3g}, "
                  f"|d lambda| {d_lam:.3g}, |d return| {d_ret:.3g}, |d Wpi| {d_pi:.3g}, |d W1n| {d_w1n:.3g}, "
                  f"|d pi_nature| {d_nat:.3g}, |d v_nature| {d_vn:.3g}", flush=True)
            ok &= same_s and same_a and d_rn < 1e-3 and d_lam < 1e-5 and d_ret < 1e-3 and d_pi < 2e-4 and d_w1n < 2e-4 \
                and d_nat < 2e-4 and d_vn < 2e-4
        sh.export_lambda(); full.export_lambda()
        d_l = max(float((p - q).abs().max()) for p, q in zip(ac_s.lambda_net.parameters(), ac_f.lambda_net.parameters()))
        print(f"[rank {rank}] {domain}: |d exported lambda-net| {d_l:.3g}", flush=True)
        ok &= d_l < 1e-4
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MA SHARD CHECK", "PASSED" if t.item() == 1.0 else "FAILED", f"(world {world})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
