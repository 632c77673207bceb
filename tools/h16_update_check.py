"""Cross-check of the hidden-16 tcgen05 update kernel against the mma.sync one on synthetic statistics rows at shapes
the parity suite does not reach (large / ragged env counts: lambda read from global memory, partially filled last tile).

    python tools/h16_update_check.py          # on a B200; re-executes itself with RMAB_H16_MMA_SYNC=1
"""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustrmab_b200 import ops  # noqa: E402

SHAPES = [(64, 256, 2), (40, 2000, 2), (24, 4096, 3), (48, 1000, 3), (16, 16384, 2)]   # (arms, envs, states)


def run(out_path):
    dev = torch.device("cuda:0")
    res = {}
    for N, E, S in SHAPES:
        A, h, T = 2, 16, 10
        g = torch.Generator(device="cpu").manual_seed(7 + E)
        SA = S * A
        P_pi, P_v = ops.n_params(S + 1, h, A), ops.n_params(S + 1, h, 1)
        Wpi = ((torch.rand((N, P_pi), generator=g) - 0.5) * 0.5).to(dev)
        Wv = ((torch.rand((N, P_v), generator=g) - 0.5) * 0.5).to(dev)
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
        hp = ops.hyper(0.2, 0.1, 3e-4, 1e-3, 6, 6)
        adam_pi = torch.zeros((N, 2, P_pi), device=dev)
        adam_v = torch.zeros((N, 2, P_v), device=dev)
        lp, lv = ops.ppo_update_table(net, hp, T, Wpi, Wv, adam_pi, adam_v, rows, arm_stats, lamb, want_loss=True)
        torch.cuda.synchronize()
        for k, v in (("Wpi", Wpi), ("Wv", Wv), ("lp", lp), ("lv", lv)):
            res[f"{k}_{N}_{E}_{S}"] = v.cpu().numpy()
    np.savez(out_path, **res)


if __name__ == "__main__":
    if os.environ.get("H16_CHILD"):
        run(os.environ["H16_CHILD"])
        sys.exit(0)
    os.makedirs("gpurun_out", exist_ok=True)
    out = {}
    for name, env in (("tcgen05", {}), ("mma_sync", {"RMAB_H16_MMA_SYNC": "1"})):
        path = f"/tmp/h16_{name}.npz"
        subprocess.run([sys.executable, __file__], check=True, env={**os.environ, **env, "H16_CHILD": path}, timeout=600)
        out[name] = np.load(path)
    bad = False
    for N, E, S in SHAPES:
        line = f"N={N:3d} E={E:6d} S={S}:"
        for k in ("Wpi", "Wv", "lp", "lv"):
            a, b = out["tcgen05"][f"{k}_{N}_{E}_{S}"], out["mma_sync"][f"{k}_{N}_{E}_{S}"]
            d = float(np.abs(a - b).max())
            # weights after 6 Adam steps: a near-zero gradient component makes Adam's m / sqrt(v) amplify rounding noise to a
            # fraction of lr per step (4.4e-5 observed at E = 2000, identically for the round-2 kernel before the last pass)
            tol = 2e-4 if k.startswith("W") else 1e-5 * max(1.0, float(np.abs(b).max()))
            bad |= not np.isfinite(a).all() or d > tol
            line += f"  {k} max|d|={d:.2e}"
        print(line)
    print("H16 CROSS-CHECK", "FAILED" if bad else "OK")
    sys.exit(1 if bad else 0)
