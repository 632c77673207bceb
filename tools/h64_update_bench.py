"""Times rmab_ppo_update_table at the ARMMAN hidden-64 shape and checks the tcgen05 kernel against the mma.sync one.

    python tools/h64_update_bench.py [N] [E]          # run on a B200; re-executes itself with RMAB_H64_MMA_SYNC=1

Synthetic statistics rows (the kernels do not care where they come from); 20 + 20 Adam steps; CUDA-event timing.
"""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustrmab_b200 import ops  # noqa: E402


def run(N, E, out_path):
    dev = torch.device("cuda:0")
    S, A, h, T = 3, 2, 64, 10
    g = torch.Generator(device="cpu").manual_seed(5)
    SA = S * A
    P_pi, P_v = ops.n_params(S + 1, h, A), ops.n_params(S + 1, h, 1)
    Wpi = ((torch.rand((N, P_pi), generator=g) - 0.5) * 0.25).to(dev)
    Wv = ((torch.rand((N, P_v), generator=g) - 0.5) * 0.25).to(dev)
    rows = torch.zeros((N, 4 * SA + 3 * S, E))
    cnt = torch.randint(0, 4, (N, SA, E), generator=g).float()
    adv = torch.randn((N, SA, E), generator=g)
    rows[:, 0:SA] = cnt * adv.clamp(min=0)
    rows[:, SA:2 * SA] = cnt * adv.clamp(max=0)
    rows[:, 2 * SA:3 * SA] = -0.7 + 0.1 * torch.randn((N, SA, E), generator=g)
    rows[:, 3 * SA:3 * SA + S] = cnt.view(N, S, A, E).sum(2)
    rows[:, 3 * SA + S:3 * SA + 2 * S] = torch.randn((N, S, E), generator=g)
    rows[:, 3 * SA + 3 * S:] = cnt
    rows = rows.to(dev)
    arm_stats = torch.zeros((N, 4), device=dev)
    lamb = torch.rand(E, generator=g).to(dev)
    net = ops.net_desc(N, E, S, A, h)
    hp = ops.hyper(0.2, 0.1, 3e-4, 1e-3, 20, 20)
    adam_pi = torch.zeros((N, 2, P_pi), device=dev)
    adam_v = torch.zeros((N, 2, P_v), device=dev)
    W0pi, W0v = Wpi.clone(), Wv.clone()
    ms = []
    for rep in range(3):
        Wpi.copy_(W0pi); Wv.copy_(W0v); adam_pi.zero_(); adam_v.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        lp, lv = ops.ppo_update_table(net, hp, T, Wpi, Wv, adam_pi, adam_v, rows, arm_stats, lamb, want_loss=True)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    np.savez(out_path, Wpi=Wpi.cpu().numpy(), Wv=Wv.cpu().numpy(), lp=lp.cpu().numpy(), lv=lv.cpu().numpy(), ms=np.array(ms))
    return ms


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1480
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    if os.environ.get("H64_CHILD"):
        run(N, E, os.environ["H64_CHILD"])
        sys.exit(0)
    os.makedirs("gpurun_out", exist_ok=True)
    res = {}
    for name, env in (("tcgen05", {}), ("mma_sync", {"RMAB_H64_MMA_SYNC": "1"})):
        path = f"gpurun_out/h64_{name}.npz"
        subprocess.run([sys.executable, __file__, str(N), str(E)], check=True, env={**os.environ, **env, "H64_CHILD": path}, timeout=600)
        res[name] = np.load(path)
        print(f"{name:9s}: update of N={N} arms x E={E} envs, 20+20 Adam steps: {np.round(res[name]['ms'], 2)} ms", flush=True)
    a, b = res["tcgen05"], res["mma_sync"]
    for k in ("Wpi", "Wv", "lp", "lv"):
        d = np.abs(a[k] - b[k])
        print(f"  {k}: max |tcgen05 - mma.sync| = {d.max():.3g}   (max |value| {np.abs(b[k]).max():.3g})")
    flops = 3 * 2 * 64 * 64 * 3 * E * 40 * N
    print(f"  fp32-equivalent GEMM rate: tcgen05 {flops / a['ms'].min() / 1e9:.1f} TFLOP/s, mma.sync {flops / b['ms'].min() / 1e9:.1f} TFLOP/s")
