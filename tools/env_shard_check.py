"""Env-sharding check (the reference's MPI layout; SURVEY.md section 8e fallback): under torchrun with k ranks every rank
runs (a) its env slice of a k-way EnvShardedRMABPPOTrainer and (b) the unsharded trainer of the same problem on its own GPU,
and compares per-env returns, lambda and the nets of ALL arms after every epoch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/env_shard_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustrmab_b200.trainer import EnvShardedRMABPPOTrainer, RMABPPOTrainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for domain, hidden, N, E, T in (("counterexample", 16, 301, 64 * world, 25), ("armman", 64, 75, 96 * world, 10)):
    kw = dict(hidden=hidden, B=N // 3, train_pi_iters=4, train_v_iters=4, epochs=8, lamb_update_freq=2, final_train_lambdas=1,
              seed=11, device=dev)
    sh = EnvShardedRMABPPOTrainer(domain, N, E, T, rank=rank, world_size=world, **kw)
    full = RMABPPOTrainer(domain, N, E, T, **kw)
    assert torch.equal(sh.Wpi, full.Wpi), "initial nets differ"
    for ep in range(5):
        a, b = sh.epoch(), full.epoch()
        ret = float((a["EpActualRet"] - b["EpActualRet"]).abs().max())
        lam = float((a["Lamb"] - b["Lamb"]).abs().max())
        dpi = float((sh.Wpi - full.Wpi).abs().max())
        dv = float((sh.Wv - full.Wv).abs().max())
        dl = float((sh.lam_params - full.lam_params).abs().max())
        print(f"[rank {rank}] {domain} h{hidden} N={N} E={E} epoch {ep}: |dEpActualRet| {ret:.3g}, |dLamb| {lam:.3g}, |dWpi| {dpi:.3g}, "
              f"|dWv| {dv:.3g}, |d lambda-net| {dl:.3g}", flush=True)
        ok &= ret == 0.0 and lam < 1e-5 and dpi < 1e-4 and dv < 1e-4 and dl < 1e-4
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ENV SHARD CHECK", "PASSED" if int(flag.item()) else "FAILED", f"(world {world})")
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
