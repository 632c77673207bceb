"""Arm-sharding check: under torchrun with k ranks, every rank runs (a) its shard of a k-way
arm-sharded RMABPPOTrainer (NCCL all-reduces for the lambda-net partial sums and the per-env sums) and (b) the
unsharded trainer of the same problem on its own GPU, and compares its arms' trajectories and nets.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustrmab_b200.trainer import RMABPPOTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for domain, hidden, N, E, T in (("counterexample", 16, 3000, 64, 25), ("armman", 64, 1001, 96, 10)):
    kw = dict(hidden=hidden, B=N // 3, train_pi_iters=4, train_v_iters=4, epochs=8, lamb_update_freq=2, final_train_lambdas=1,
              seed=11, device=dev)
    sh = RMABPPOTrainer(domain, N, E, T, rank=rank, world_size=world, keep_trajectory=True, **kw)
    full = RMABPPOTrainer(domain, N, E, T, keep_trajectory=True, **kw)
    sl = slice(sh.arm0, sh.arm0 + sh.N)
    assert torch.equal(sh.Wpi, full.Wpi[sl]) and torch.equal(sh.nature, full.nature[sl]), "initial shard differs"
    for ep in range(4):
        a, b = sh.epoch(), full.epoch()
        same_a = bool(torch.equal(sh.a, full.a[sl]))
        same_s = bool(torch.equal(sh.s_traj, full.s_traj[sl]))
        ret = float((a["EpActualRet"] - b["EpActualRet"]).abs().max())
        lam = float((a["Lamb"] - b["Lamb"]).abs().max())
        dW = float((sh.Wpi - full.Wpi[sl]).abs().max())
        print(f"[rank {rank}] {domain} h{hidden} epoch {ep}: actions equal {same_a}, states equal {same_s}, "
              f"|dEpActualRet| {ret:.3g}, |dLamb| {lam:.3g}, |dWpi| {dW:.3g}", flush=True)
        ok &= same_a and same_s and ret == 0.0 and lam < 1e-5 and dW < 1e-4
    # lambda-net after its SGD steps: rank r updates only the W1 columns of its own arms (the rest of the net on all)
    W1 = slice(0, 8 * N)
    cols = torch.zeros(8 * N, dtype=torch.bool, device=dev)
    cols.view(8, N)[:, sl] = True
    lam_diff = float(torch.cat([(sh.lam_params[W1] - full.lam_params[W1])[cols], sh.lam_params[8 * N:] - full.lam_params[8 * N:]]).abs().max())
    print(f"[rank {rank}] {domain}: |d lambda-net params| (own W1 columns + shared layers) {lam_diff:.3g}", flush=True)
    ok &= lam_diff < 1e-4
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARD CHECK", "PASSED" if int(flag.item()) else "FAILED", f"(world {world})")
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
