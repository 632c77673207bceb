"""world_size-2 gloo tests (CPU) of the collective drop-ins (robust_rmab/utils/mpi_tools.py, mpi_pytorch.py
semantics) and of the arm sharding host logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from robustrmab_b200 import mpi_tools as mt
    res = {}
    res["ids"] = (mt.proc_id(), mt.num_procs())
    res["avg"] = float(mt.mpi_avg(float(rank + 1)))
    x = np.arange(5, dtype=np.float64) + 10 * rank
    res["stats"] = [float(v) for v in mt.mpi_statistics_scalar(x, with_min_and_max=True)]
    torch.manual_seed(rank)                       # ranks start with different parameters and gradients
    m = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1))
    mt.sync_params(m)
    res["params"] = torch.cat([p.data.reshape(-1) for p in m.parameters()]).numpy().copy()
    (m(torch.full((2, 3), float(rank + 1))).sum()).backward()
    res["grad_local"] = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).numpy().copy()
    mt.mpi_avg_grads(m)
    res["grad_avg"] = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).numpy().copy()
    res["shard"] = mt.arm_shard(10)
    # arm-sharded lambda-net first layer: partial sums over each rank's arms add up to the full product
    N, E = 10, 3
    rs = np.random.RandomState(0)
    W1, s = rs.randn(8, N).astype(np.float32), rs.randint(0, 3, (N, E)).astype(np.float32)
    a0, n = res["shard"]
    part = torch.as_tensor(W1[:, a0:a0 + n] @ s[a0:a0 + n])
    mt.allreduce_(part)
    res["h1"] = part.numpy().copy()
    res["h1_full"] = W1 @ s
    out[rank] = res
    dist.destroy_process_group()


def test_collectives_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["ids"] == (0, 2) and r1["ids"] == (1, 2)
    assert r0["avg"] == r1["avg"] == 1.5
    allx = np.concatenate([np.arange(5) + 0.0, np.arange(5) + 10.0]).astype(np.float32)
    np.testing.assert_allclose(r0["stats"], [allx.mean(), allx.std(), 0.0, 14.0], rtol=1e-6)
    assert r0["stats"] == r1["stats"]
    assert np.array_equal(r0["params"], r1["params"])                       # sync_params: rank 0's values everywhere
    np.testing.assert_allclose(r0["grad_avg"], 0.5 * (r0["grad_local"] + r1["grad_local"]), rtol=1e-6)
    assert np.array_equal(r0["grad_avg"], r1["grad_avg"])
    assert r0["shard"] == (0, 5) and r1["shard"] == (5, 5)
    np.testing.assert_allclose(r0["h1"], r0["h1_full"], rtol=1e-5, atol=1e-5)


def test_single_process_defaults():
    from robustrmab_b200 import mpi_tools as mt
    assert mt.proc_id() == 0 and mt.num_procs() == 1
    assert mt.mpi_avg(3.0) == 3.0
    mean, std = mt.mpi_statistics_scalar(np.array([1.0, 2.0, 3.0]))
    assert abs(mean - 2.0) < 1e-6 and abs(std - np.std([1, 2, 3])) < 1e-6
    assert [mt.arm_shard(10, r, 3) for r in range(3)] == [(0, 4), (4, 4), (8, 2)]
    assert mt.arm_shard(2, 3, 4) == (2, 0)


def _ma_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from robustrmab_b200 import mpi_tools as mt
    from robustrmab_b200.nature_oracle import MARolloutTrainer
    N, E, T = 7, 3, 4                                   # ragged: 4 + 3 arms
    a0, n = mt.arm_shard(N, rank, world)
    me = SimpleNamespace(world=world, pg=None, per=(N + world - 1) // world, N_total=N, N=n, arm0=a0, _gpad=None, _gout=None)
    full = torch.arange(N * T * E, dtype=torch.int64).remainder(251).to(torch.uint8).reshape(N, T, E)
    got = MARolloutTrainer._gather_arms(me, full[a0:a0 + n].contiguous())
    res = {"gather_ok": bool(torch.equal(got, full))}
    part = torch.full((T, 3, E), float(rank + 1))
    MARolloutTrainer._allreduce(me, part)
    res["part"] = float(part[0, 0, 0])
    # export_lambda: every rank trained only its own first-layer columns; the merge takes each column from its owner
    lam = torch.zeros(8 * N + 89)
    lam[:8 * N].view(8, N)[:, a0:a0 + n] = float(rank + 1)
    lam[8 * N:] = 5.0
    net = torch.nn.Sequential(torch.nn.Linear(N, 8), torch.nn.Linear(8, 8), torch.nn.Linear(8, 1))
    me.lam_params, me.ac = lam, SimpleNamespace(lambda_net=net)
    me._allreduce = lambda t: MARolloutTrainer._allreduce(me, t)
    MARolloutTrainer.export_lambda(me)
    res["w1"] = net[0].weight.detach().numpy().copy()
    res["b1"] = net[0].bias.detach().numpy().copy()
    # sharded nature-net training: after the Adam steps every rank owns the rows of its arms (W1T, W3, b3, log_std of the
    # actor; both W1T blocks of the critic); exchange_shards must give every rank the same, complete parameters
    from robustrmab_b200.nature_nets import NatureNets
    h, k = 8, 2
    mk = lambda i, o: torch.nn.Sequential(torch.nn.Linear(i, h), torch.nn.Tanh(), torch.nn.Linear(h, h), torch.nn.Tanh(),
                                          torch.nn.Linear(h, o))
    torch.manual_seed(3)
    acn = SimpleNamespace(N=N, act_dim_nature=k * N, hidden_sizes=[h, h], device_=torch.device("cpu"),
                          pi_nature=SimpleNamespace(mu_net=mk(N, k * N), log_std=torch.nn.Parameter(torch.zeros(k * N))),
                          v_nature=SimpleNamespace(v_net=mk(2 * N, 1)))
    nets = NatureNets(acn, a0, n)
    before = nets.actor.p.clone()
    for net_ in (nets.actor, nets.critic):
        net_.p[:net_.ob1].view(-1, h)[a0:a0 + n] += 10.0 * (rank + 1)          # W1T rows of own arms
    nets.critic.p[:nets.critic.ob1].view(-1, h)[N + a0:N + a0 + n] += 10.0 * (rank + 1)
    nets.actor.W3(nets.actor.p, a0 * k, n * k).add_(10.0 * (rank + 1))
    nets.actor.b3(nets.actor.p, a0 * k, n * k).add_(10.0 * (rank + 1))
    nets.actor.ls(nets.actor.p, a0 * k, n * k).add_(10.0 * (rank + 1))
    nets.exchange_shards(lambda t: MARolloutTrainer._allreduce(me, t))
    own = torch.zeros(N)
    own[:4], own[4:] = 10.0, 20.0
    d = nets.actor.p - before
    res["nat_ok"] = bool(torch.equal(d[:nets.actor.ob1].view(N, h), own[:, None].expand(N, h)) and
                         float(d[nets.actor.ob1:nets.actor.oW3].abs().max()) == 0.0 and
                         torch.equal(d[nets.actor.oW3:nets.actor.ob3].view(N, k, h)[:, 0, 0], own) and
                         torch.equal(d[nets.actor.ols:].view(N, k)[:, 1], own))
    res["nat_p"] = nets.critic.p.numpy().copy()
    # env-sharded -> arm-sharded statistics (EnvShardedRMABPPOTrainer): rank r holds envs [r E_loc, (r+1) E_loc) of all arms
    from robustrmab_b200.trainer import EnvShardedRMABPPOTrainer
    Nn, Kk, El = 5, 3, 4                                 # ragged: per = 3 arms per rank
    per = (Nn + world - 1) // world
    full_st = torch.arange(Nn * Kk * El * world, dtype=torch.float32).reshape(Nn, Kk, El * world)
    send = torch.zeros((world * per, Kk, El))
    send[:Nn] = full_st[:, :, rank * El:(rank + 1) * El]
    got_st = EnvShardedRMABPPOTrainer.reshard_env_to_arm(send, torch.empty_like(send), world, per)
    lo_, hi_ = rank * per, min((rank + 1) * per, Nn)
    res["reshard_ok"] = bool(torch.equal(got_st[:hi_ - lo_], full_st[lo_:hi_]))
    out[rank] = res
    dist.destroy_process_group()


def test_ma_sharding_host_logic_world2():
    """MARolloutTrainer's cross-rank plumbing on CPU: ragged arm all-gather, partial-sum all-reduce, lambda-net merge."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ma_worker, args=(world, port, out), nprocs=world, join=True)
    for r in (0, 1):
        assert out[r]["gather_ok"]
        assert out[r]["part"] == 3.0
        w1 = out[r]["w1"]
        assert np.array_equal(w1[:, :4], np.full((8, 4), 1.0)) and np.array_equal(w1[:, 4:], np.full((8, 3), 2.0))
        assert np.array_equal(out[r]["b1"], np.full(8, 5.0))
        assert out[r]["nat_ok"]
        assert out[r]["reshard_ok"]
    assert np.array_equal(out[0]["nat_p"], out[1]["nat_p"])
