"""Batched evaluation rollouts: `AgentOracle.simulate_reward` (robust_rmab/algos/rmabppo/agent_oracle.py:599-652).

The reference runs `epochs` independent rollouts of `steps_per_epoch` steps one after another (fresh random start
state each) and returns the mean discounted return; here the `epochs` rollouts are the env replicas of ONE batched
rollout: per step `act_test` (lambda-net -> per-arm probabilities -> budget-constrained selection), the nature
policy's parameters, `env.step`, and a discounted reward accumulation, all on device.  DoubleOracle.update_payoffs
(double_oracle.py:242-278) calls this O(|agent strats| x |nature strats|) times per iteration.
"""
import numpy as np
import torch

from . import ops


def simulate_reward(agent_pol, env, nature, steps_per_epoch=100, epochs=100, gamma=0.99, seed=0, return_all=False):
    """agent_pol: `core.MLPActorCriticRMAB` (or any object with `act_test_device(s[N,E]) -> a[N,E]` u8);
    env: a `robustrmab_b200.environments` env constructed with `n_envs=epochs` (its ranges already set);
    nature: what `env.step_device` accepts (per-arm parameter table of a baseline nature policy, or a callable
    `nature(s) -> parameters` for a learned nature policy).  Returns the mean over rollouts of sum_t gamma^t r_t."""
    if env.n_envs != int(epochs):
        raise ValueError("construct the evaluation env with n_envs = epochs (one replica per rollout)")
    env.seed(seed)
    s = env.reset_device()
    ret = torch.zeros((env.n_envs,), dtype=torch.float64, device=env.device)
    disc = 1.0
    for _ in range(int(steps_per_epoch)):
        a = agent_pol.act_test_device(s)
        nat = nature(s) if callable(nature) else nature
        s, r = env.step_device(a, nat)
        env._state = s
        env._t += 1
        ret += disc * r.sum(dim=0, dtype=torch.float64)
        disc *= gamma
    if return_all:
        return ret
    return float(ret.mean().item())


from .baselines import PessimisticAgentPolicy  # noqa: E402,F401  (kept importable from here)
