"""MA-RMABPPO rollout-step restatement (oracle, test-only).

Follows robust_rmab/algos/ma_rmabppo/ma_rmabppo_core.py:
  MLPGaussianActor              :84-107   mu_net(obs), std = exp(log_std), logp = Normal.log_prob(a).sum(-1)
  RMABLambdaNatureOracle.step   :274-331  nature sample (:289-291); per arm: agent actor as rmabppo_core.step,
                                          agent critic on [x_i, bounded nature vector] (:312-316);
                                          v_nature([obs, a_agent]) (:322-323)
and the counterfactual estimate of nature_oracle.py:685-701 (mean over 50 env steps from the same state).
"""
import numpy as np

from . import envs as oenv
from . import nets, philox

F32 = np.float32


def dense_mlp(params, x):
    """Linear-Tanh-Linear-Tanh-Linear on `[E, in]`; params = (W1, b1, W2, b2, W3, b3) in torch [out, in] layout."""
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, F32) for p in params]
    a1 = np.tanh(x.astype(F32) @ W1.T + b1)
    a2 = np.tanh(a1 @ W2.T + b2)
    return a2 @ W3.T + b3


def normal_eps(seed, t, n_envs, dim, env0=0):
    """Standard normals `[E, D]` of the CUDA path: Box-Muller on words 0/1 of block (NATURE, t, env, d // 2)."""
    env = (np.arange(n_envs, dtype=np.uint32) + np.uint32(env0))[:, None]
    idx = (np.arange(dim, dtype=np.uint32) // 2)[None, :]
    z0, z1 = philox.normal_pair(seed, philox.STREAM_NATURE, t, env, idx)
    return np.where((np.arange(dim) % 2)[None, :] == 1, z1, z0).astype(F32)


def gaussian_sample(mu, log_std, eps):
    """a = mu + exp(log_std) * eps ; logp = sum_d Normal(mu, std).log_prob(a) (torch formula, fp32)."""
    mu, log_std, eps = np.asarray(mu, F32), np.asarray(log_std, F32), np.asarray(eps, F32)
    std = np.exp(log_std).astype(F32)
    a = (mu + std * eps).astype(F32)
    var = (std * std).astype(F32)
    lp = -((a - mu) ** 2) / (F32(2) * var) - log_std - F32(np.log(np.sqrt(2 * np.pi)))
    return a, lp.astype(F32).sum(axis=-1).astype(F32)


def agent_critic_extra(Wv, W1n, s, lamb, nat_bounded, S, h, one_hot=True, state_norm=1.0):
    """Agent critics on [x_i, bounded nature vector] (:312-316).  Wv `[N,Pv]` packed over the (obs, lambda) inputs,
    W1n `[N,h,D]` the first-layer block over the nature inputs, nat_bounded `[E,D]`.  Returns v `[N,E]`."""
    in_dim = (S if one_hot else 1) + 1
    W1, b1, W2, b2, W3, b3 = [p.astype(F32) for p in nets.unpack(Wv, in_dim, h, 1)]
    x = nets.features(s, lamb, S, one_hot, state_norm)                        # [N,E,in]
    z1 = np.einsum('nei,nji->nej', x, W1) + b1[:, None, :] + np.einsum('njd,ed->nej', W1n.astype(F32), nat_bounded.astype(F32))
    a1 = np.tanh(z1)
    a2 = np.tanh(np.einsum('nek,njk->nej', a1, W2) + b2[:, None, :])
    return (np.einsum('nek,njk->nej', a2, W3) + b3[:, None, :])[..., 0].astype(F32)


def counterfactual_mean(domain, s, a, nature, ranges, R, seed, t, n_trials, pop=None):
    """Mean over n_trials env steps from the same state of R[arm, s'] per (arm, env); draws: stream CF,
    time counter t * n_trials + trial."""
    N, E = s.shape
    acc = np.zeros((N, E), F32)
    for tr in range(n_trials):
        u = philox.uniform_grid(seed, philox.STREAM_CF, t * n_trials + tr, E, N)
        if domain == oenv.DOMAIN_CE:
            _, r = oenv.ce_step(s, a, nature, ranges, R, u)
        elif domain == oenv.DOMAIN_ARMMAN:
            _, r = oenv.armman_step(s, a, nature, ranges, R, u)
        else:
            _, r = oenv.sis_step(pop, s, a, nature, R, u)
        acc = (acc + r).astype(F32)
    return (acc / F32(n_trials)).astype(F32)
