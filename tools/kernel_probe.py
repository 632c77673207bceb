"""One launch (after one warm-up) of each stand-alone kernel at the bench shapes, for `ncu --set full -k regex:...`:
gemm3h (configs[3] forward product at N = 512 arms: [5120, 2048] x [2048, 32768]), the split kernels, the multi-CTA radix
select, env step, GAE and batched value iteration at the bench_blocks shapes.   python tools/kernel_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from robustrmab_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
reps = 2
# ---- gemm3h + splits
Na, h, M = 512, 64, 5120
D, NH = 4 * Na, Na * h
nat = torch.rand((M, D), device=dev, generator=g) * 9 + 1
W = (torch.rand((NH, D), device=dev, generator=g) - 0.5) * 0.02
dz = torch.randn((M, NH), device=dev, generator=g) * 1e-4
out = torch.empty((M, NH), device=dev)
for _ in range(reps):
    a, b = ops.split16_rows(nat), ops.split16_rows(W)
    ops.gemm3h(a, b, D, out=out, n_fast=False)
    c = ops.split16_cols(dz)
del nat, W, dz, out, a, b, c
# ---- selection, env step, GAE at 100k x 1024
N, E, T = 100_000, 1024, 10
p1 = torch.rand((N, E), device=dev, generator=g)
for _ in range(reps):
    ops.budget_select_binary(p1, 1.0, float(N // 5))
del p1
from robustrmab_b200.trainer import RMABPPOTrainer  # noqa: E402
ranges = RMABPPOTrainer.default_ranges(_lib.DOMAIN_ARMMAN, N, np.random.RandomState(0))
nature = torch.as_tensor(ranges.mean(-1), dtype=torch.float32, device=dev).contiguous()
rg = torch.as_tensor(ranges, dtype=torch.float32, device=dev).contiguous()
R = torch.tensor([1.0, 0.5, 0.0], device=dev).repeat(N, 1).contiguous()
s = torch.randint(0, 3, (N, E), dtype=torch.uint8, device=dev, generator=g)
a = torch.randint(0, 2, (N, E), dtype=torch.uint8, device=dev, generator=g)
s2, rew = torch.empty_like(s), torch.empty((N, E), dtype=torch.float32, device=dev)
for k in range(reps):
    ops.env_step_table(_lib.DOMAIN_ARMMAN, s, a, nature, rg, R, r=ops.rng(0, _lib.STREAM_ENV, k), s_next=s2, rew=rew)
del rew, s2, s, a
Ng = N // 2
s_traj = torch.randint(0, 3, (Ng, T + 1, E), dtype=torch.uint8, device=dev, generator=g)
at = torch.randint(0, 2, (Ng, T, E), dtype=torch.uint8, device=dev, generator=g)
val = torch.randn((Ng, T, E), device=dev, generator=g)
last = torch.randn((Ng, E), device=dev, generator=g)
lamb = torch.rand((E,), device=dev, generator=g)
Cc = torch.tensor([0.0, 1.0], device=dev)
for _ in range(reps):
    ops.gae(s_traj, at, val, last, lamb, Cc, R[:Ng].contiguous(), 0.9, 0.97, normalize=True)
del s_traj, at, val
# ---- batched VI, counterexample tables, 99 999 arms x 64 probes, eps 1e-8
Nv, L = 99_999, 64
p = RMABPPOTrainer.default_ranges(_lib.DOMAIN_CE, Nv, None).mean(-1)
Tt = np.zeros((Nv, 2, 2, 2))
Tt[:, 0] = 0.5
Tt[:, 1, 0, 0] = 1.0
Tt[:, 1, 1, 1], Tt[:, 1, 1, 0] = p, 1 - p
Rv = np.tile(np.array([0.0, 1.0]), (Nv, 1))
dT, dR, dC, dL = (torch.as_tensor(x, device=dev) for x in (Tt, Rv, np.array([0.0, 1.0]), np.linspace(0, 10.0, L)))
for _ in range(reps):
    ops.vi_batched(dT, dR, dC, dL, 0.9, 1e-8)
    ops.whittle_bisect(dT, dR, dC, 0.9, 1e-8, 0.0, 10.0, 20)
torch.cuda.synchronize()
print("probe ok")
