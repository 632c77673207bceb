"""Run-to-run determinism of the update kernels: the same launch on the same inputs must give bit-identical nets
(no atomics, fixed reduction orders).  A difference means a race.  GPU only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets as onets            # noqa: E402  (initialiser only)
from robustrmab_b200 import ops             # noqa: E402

dev = torch.device("cuda:0")
t = lambda x: torch.as_tensor(np.ascontiguousarray(x), device=dev)


def per_sample(N, T, E, S, A, h, one_hot, reps=6, iters=5):
    rs = np.random.RandomState(N + T + E + h)
    ind = (S if one_hot else 1) + 1
    Wpi0, Wv0 = onets.init_linear_default(rs, N, ind, h, A), onets.init_linear_default(rs, N, ind, h, 1)
    s = rs.randint(0, S, (N, T + 1, E)).astype(np.uint8)
    a = rs.randint(0, A, (N, T, E)).astype(np.uint8)
    adv = rs.randn(N, T, E).astype(np.float32)
    logp = (-rs.rand(N, T, E)).astype(np.float32)
    ret = rs.randn(N, T, E).astype(np.float32)
    lamb = rs.rand(E).astype(np.float32)
    net = ops.net_desc(N, E, S, A, h, one_hot, float(S - 1))
    outs = []
    for _ in range(reps):
        Wpi, Wv = t(Wpi0.copy()), t(Wv0.copy())
        ap = torch.zeros((N, 2, Wpi0.shape[1]), device=dev)
        av = torch.zeros((N, 2, Wv0.shape[1]), device=dev)
        hp = ops.hyper(2.0, 0.3, 2e-3, 2e-3, iters, iters)
        ops.ppo_update(net, hp, Wpi, Wv, ap, av, t(s), t(a), t(adv), t(logp), t(ret), t(lamb))
        torch.cuda.synchronize()
        outs.append((Wpi.cpu().numpy(), Wv.cpu().numpy()))
    same = all(np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) for o in outs)
    md = max(max(np.abs(o[0] - outs[0][0]).max(), np.abs(o[1] - outs[0][1]).max()) for o in outs)
    print(f"ppo_update N={N} T={T} E={E} S={S} A={A} h={h} one_hot={one_hot}: deterministic={same} max|diff|={md:.3g}", flush=True)
    return same


ok = True
for cfg in [(4, 7, 1, 21, 3, 64, False), (6, 6, 1, 11, 3, 64, False), (4, 8, 1, 11, 3, 16, False), (6, 6, 4, 11, 3, 64, False),
            (3, 20, 40, 51, 3, 64, False), (5, 12, 16, 3, 2, 64, True), (5, 12, 16, 2, 2, 16, True), (2, 30, 300, 51, 3, 64, False)]:
    ok &= per_sample(*cfg)
print("ALL DETERMINISTIC" if ok else "NONDETERMINISM FOUND")


def table(N, T, E, S, h, reps=5, iters=4):
    rs = np.random.RandomState(N + E + h)
    A = 2
    Wpi0, Wv0 = onets.init_linear_default(rs, N, S + 1, h, A), onets.init_linear_default(rs, N, S + 1, h, 1)
    K = 4 * S * A + 3 * S
    stats = rs.randn(N, K, E).astype(np.float32)
    stats[:, 3 * S * A:3 * S * A + S] = rs.randint(1, T, (N, S, E))
    stats[:, 2 * S * A:3 * S * A] = -np.abs(stats[:, 2 * S * A:3 * S * A])
    arm_stats = np.abs(rs.randn(N, 4)).astype(np.float32)
    lamb = rs.rand(E).astype(np.float32)
    net = ops.net_desc(N, E, S, A, h)
    outs = []
    for _ in range(reps):
        Wpi, Wv = t(Wpi0.copy()), t(Wv0.copy())
        ap = torch.zeros((N, 2, Wpi0.shape[1]), device=dev)
        av = torch.zeros((N, 2, Wv0.shape[1]), device=dev)
        hp = ops.hyper(2.0, 0.3, 2e-3, 2e-3, iters, iters)
        ops.ppo_update_table(net, hp, T, Wpi, Wv, ap, av, t(stats), t(arm_stats), t(lamb))
        torch.cuda.synchronize()
        outs.append((Wpi.cpu().numpy(), Wv.cpu().numpy()))
    same = all(np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) for o in outs)
    md = max(max(np.abs(o[0] - outs[0][0]).max(), np.abs(o[1] - outs[0][1]).max()) for o in outs)
    print(f"ppo_update_table N={N} T={T} E={E} S={S} h={h}: deterministic={same} max|diff|={md:.3g}", flush=True)
    return same


ok2 = True
for cfg in [(8, 10, 1024, 3, 64), (8, 10, 100, 3, 64), (8, 10, 300, 2, 64), (16, 100, 256, 2, 16), (16, 20, 200, 3, 16), (6, 10, 256, 3, 32)]:
    ok2 &= table(*cfg)
print("TABLE ALL DETERMINISTIC" if ok2 else "TABLE NONDETERMINISM FOUND")
