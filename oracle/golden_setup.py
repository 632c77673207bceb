"""Builders of the CPU oracle loops on the problems stored in tests/golden/*.npz (test infrastructure; shared by the
CPU tests of the oracle and the GPU parity tests)."""
import numpy as np

from . import envs as oenv

F32 = np.float32


def sis_rmabppo_oracle(g, name, n_envs=1):
    """CpuRMABPPO on the problem of tests/golden/train_sis.npz[name] (shared with the GPU tests)."""
    from oracle.train import CpuRMABPPO
    N, S, T, epochs, hidden, seed = [int(x) for x in g[name + "_meta"]]
    pop = int(g[name + "_pop"])
    kw = dict(zip([str(k) for k in g[name + "_hyper_names"]], [float(v) for v in g[name + "_hyper_values"]]))
    lam0 = [g[f"{name}_lam0_{i}"] for i in range(6)]
    R = np.tile(np.linspace(0, 1, S).astype(F32), (N, 1))                    # bandit_env_robust.py:1181
    cpu = CpuRMABPPO(oenv.DOMAIN_SIS, g[name + "_Wpi0"], g[name + "_Wv0"], lam0, g[name + "_nature"].astype(F32),
                     g[name + "_ranges"].astype(F32), R, np.array([0, 1, 2], F32), n_envs, T, hidden, float(g[name + "_B"]),
                     gamma=kw["gamma"], lam_other=kw["lam_OTHER"], clip_ratio=kw["clip_ratio"], pi_lr=kw["pi_lr"],
                     vf_lr=kw["vf_lr"], lm_lr=kw["lm_lr"], train_pi_iters=int(kw["train_pi_iters"]),
                     train_v_iters=int(kw["train_v_iters"]), epochs=epochs, lamb_update_freq=int(kw["lamb_update_freq"]),
                     final_train_lambdas=int(kw["final_train_lambdas"]), start_entropy_coeff=kw["start_entropy_coeff"],
                     end_entropy_coeff=kw["end_entropy_coeff"], seed=seed, pop=pop, one_hot=False, state_norm=pop)
    return cpu, epochs


def ma_train_setup(g):
    import torch
    N, S, T, epochs, h, seed, D = [int(x) for x in g["meta"]]
    kw = dict(zip([str(k) for k in g["hyper_names"]], [float(v) for v in g["hyper_values"]]))

    def dense(prefix, sizes):
        layers = []
        for j in range(len(sizes) - 1):
            lin = torch.nn.Linear(sizes[j], sizes[j + 1])
            with torch.no_grad():
                lin.weight.copy_(torch.as_tensor(g[f"{prefix}_{2 * j}"])); lin.bias.copy_(torch.as_tensor(g[f"{prefix}_{2 * j + 1}"]))
            layers += [lin, torch.nn.Tanh() if j < len(sizes) - 2 else torch.nn.Identity()]
        return torch.nn.Sequential(*layers)
    return N, S, T, epochs, h, seed, D, kw, dense


def sis_ma_oracle(g, n_envs=1):
    """CpuMARMABPPO on the problem of tests/golden/ma_train_sis.npz (shared with the GPU tests)."""
    from oracle.ma_train import CpuMARMABPPO
    N, S, T, epochs, h, seed, D, kw, dense = ma_train_setup(g)
    pop = int(g["pop"])
    R = np.tile(np.linspace(0, 1, S).astype(F32), (N, 1))
    cpu = CpuMARMABPPO(oenv.DOMAIN_SIS, g["Wpi0"], g["Wv0"], g["W1n0"], [g[f"lam0_{i}"] for i in range(6)],
                       dense("mu0", [N, h, h, D]), g["log_std0"], dense("vn0", [2 * N, h, h, 1]), g["ranges"], R,
                       np.array([0, 1, 2], F32), n_envs, T, h, float(g["B"]), kw["gamma"], kw["lam_OTHER"], kw["clip_ratio"],
                       kw["pi_lr_agent"], kw["pi_lr_nature"], kw["vf_lr_agent"], kw["vf_lr_nature"], kw["lm_lr"],
                       int(kw["train_pi_iters"]), int(kw["train_v_iters"]), epochs, int(kw["lamb_update_freq"]),
                       int(kw["final_train_lambdas"]), kw["start_entropy_coeff"], kw["end_entropy_coeff"], seed, pop=pop,
                       one_hot=False, state_norm=pop, nature_state_norm=1.0)
    return cpu, epochs
