"""Model files for the packed actor-critic: a flat, pickle-free checkpoint.

The reference saves the whole `ac` object with `torch.save` (utils/logx.py:253-275) and loads it in
simulator.py:443-451, which breaks under torch >= 2.6 (`weights_only=True`) and keeps no optimiser state.
`save` writes one `.npz` (numpy archive, no pickled objects): packed nets `[N,P]`, the lambda-net parameters, the
constructor arguments, and optionally the Adam moments / step counters of a trainer so training can resume.
`load` rebuilds a `core.MLPActorCriticRMAB` (the drop-in that simulator.py / double_oracle.py call `act_test` on).
"""
import json

import numpy as np
import torch

from .core import MLPActorCriticRMAB

FORMAT = "robustrmab_b200.ac.v1"


def _lambda_arrays(ac, flat=None):
    """lambda-net parameters as named arrays; from the packed flat tensor (W1 b1 W2 b2 W3 b3) when given."""
    out, o = {}, 0
    for i, p in enumerate(ac.lambda_net.parameters()):
        if flat is None:
            out[f"lambda_{i}"] = p.detach().cpu().numpy()
        else:
            out[f"lambda_{i}"] = flat[o:o + p.numel()].reshape(tuple(p.shape)).numpy().copy()
            o += p.numel()
    return out


def save(path, ac, trainer=None):
    """With `trainer` (an `RMABPPOTrainer`), the nets, the lambda net, the Adam moments and the counters are ALL read
    from the trainer -- it owns its own tensors, so `ac` only supplies the constructor arguments and need not have been
    synchronised.  An arm-sharded trainer writes its shard (arm0, N recorded; only its lambda-net W1 columns are
    current) and `restore_trainer` refuses a trainer holding different arms."""
    meta = dict(format=FORMAT, n_states=ac.n_states, act_dim=ac.act_dim, hidden_sizes=list(ac.hidden_sizes),
                C=[float(c) for c in ac.C], N=ac.N, B=float(ac.B), strat_ind=int(ac.ind), one_hot_encode=ac.one_hot_encode,
                non_ohe_obs_dim=ac.non_ohe_obs_dim, state_norm=ac.state_norm, n_total=ac.n_total, arm0=ac.arm0)
    if trainer is None:
        Wpi, Wv, lam = ac.Wpi, ac.Wv, _lambda_arrays(ac)
    else:
        if trainer.N_total != ac.n_total:
            raise ValueError(f"trainer has {trainer.N_total} arms in total, ac {ac.n_total}")
        Wpi, Wv = trainer.Wpi, trainer.Wv
        lam = _lambda_arrays(ac, trainer.lam_params.detach().cpu())
        meta.update(N=int(trainer.N), arm0=int(trainer.arm0), n_total=int(trainer.N_total))
    arrays = {"Wpi": Wpi.detach().cpu().numpy(), "Wv": Wv.detach().cpu().numpy(),
              "meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    arrays.update(lam)
    if trainer is not None:
        arrays.update(adam_pi=trainer.adam_pi.detach().cpu().numpy(), adam_v=trainer.adam_v.detach().cpu().numpy(),
                      steps=np.array([trainer.steps_pi, trainer.steps_v, trainer.epoch_idx, trainer.t_global]),
                      shard=np.array([trainer.arm0, trainer.N, trainer.N_total]))
    np.savez(path, **arrays)


def load(path, device="cuda", n_envs=1):
    with np.load(path, allow_pickle=False) as z:
        meta = json.loads(bytes(z["meta"]).decode())
        if meta.get("format") != FORMAT:
            raise ValueError(f"{path}: not a {FORMAT} checkpoint")
        ac = MLPActorCriticRMAB(np.arange(meta["n_states"]), np.arange(meta["act_dim"]), hidden_sizes=meta["hidden_sizes"],
                                C=np.array(meta["C"]), N=meta["N"], B=meta["B"], strat_ind=meta["strat_ind"],
                                one_hot_encode=meta["one_hot_encode"], non_ohe_obs_dim=meta["non_ohe_obs_dim"],
                                state_norm=meta["state_norm"], n_envs=n_envs, device=device, arm0=meta["arm0"],
                                n_total=meta["n_total"])
        ac.Wpi.copy_(torch.as_tensor(z["Wpi"]))
        ac.Wv.copy_(torch.as_tensor(z["Wv"]))
        with torch.no_grad():
            for i, p in enumerate(ac.lambda_net.parameters()):
                p.copy_(torch.as_tensor(z[f"lambda_{i}"]))
        extra = {k: z[k] for k in ("adam_pi", "adam_v", "steps", "shard") if k in z.files}
    return ac, extra


def restore_trainer(trainer, ac, extra):
    """Put a loaded checkpoint back into an `RMABPPOTrainer` (same shape and arm shard) to resume training exactly:
    nets, lambda net, Adam state, counters, and the start states of the next epoch (the trainer draws them from
    (seed, RESET, epoch_idx), so they are regenerated rather than stored)."""
    from . import _lib, ops
    if "shard" in extra:
        a0, n, nt = (int(x) for x in extra["shard"])
        if (a0, n, nt) != (trainer.arm0, trainer.N, trainer.N_total):
            raise ValueError(f"checkpoint holds arms [{a0}, {a0 + n}) of {nt}; this trainer holds "
                             f"[{trainer.arm0}, {trainer.arm0 + trainer.N}) of {trainer.N_total}")
    trainer.Wpi.copy_(ac.Wpi)
    trainer.Wv.copy_(ac.Wv)
    trainer.lam_params.copy_(ac.lambda_net.packed().to(trainer.dev))
    if "adam_pi" in extra:
        trainer.adam_pi.copy_(torch.as_tensor(extra["adam_pi"]))
        trainer.adam_v.copy_(torch.as_tensor(extra["adam_v"]))
        trainer.steps_pi, trainer.steps_v, trainer.epoch_idx, trainer.t_global = [int(x) for x in extra["steps"]]
        trainer.s_traj[:, 0] = ops.reset_states(trainer.N, trainer.E, trainer.S,
                                                ops.rng(trainer.seed, _lib.STREAM_RESET, trainer.epoch_idx, arm0=trainer.arm0))
