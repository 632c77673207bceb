"""Per-arm actor / critic MLP restatement on packed weights (oracle, test-only).

Follows robust_rmab/algos/rmabppo/rmabppo_core.py:
  mlp / MLPCategoricalActor / MLPCritic / MLPLambdaNet   :24-29, :73-95, :108-115
  MLPActorCriticRMAB.__init__ (N separate nets)           :131-174
  MLPActorCriticRMAB.step                                 :198-244
Packed layout per arm (row of `[N, P]`, torch.nn.Linear `[out, in]` row-major, layer by layer):
  W1[h,in] b1[h] W2[h,h] b2[h] W3[out,h] b3[out]      P = (in+1)h + (h+1)h + (h+1)out
Input features: one-hot(state) ++ [lambda]  (:213-216), or [state/state_norm, lambda] (:217-218);
the MA agent critic appends the bounded nature action vector (ma_rmabppo_core.py:312-316).
"""
import numpy as np

from . import philox

F32 = np.float32


def n_params(in_dim, h, out):
    return (in_dim + 1) * h + (h + 1) * h + (h + 1) * out


def unpack(W, in_dim, h, out):
    """Views `(W1[N,h,in], b1[N,h], W2[N,h,h], b2[N,h], W3[N,out,h], b3[N,out])` of packed `[N,P]`."""
    N = W.shape[0]
    o = 0
    W1 = W[:, o:o + h * in_dim].reshape(N, h, in_dim); o += h * in_dim
    b1 = W[:, o:o + h]; o += h
    W2 = W[:, o:o + h * h].reshape(N, h, h); o += h * h
    b2 = W[:, o:o + h]; o += h
    W3 = W[:, o:o + out * h].reshape(N, out, h); o += out * h
    b3 = W[:, o:o + out]; o += out
    assert o == W.shape[1]
    return W1, b1, W2, b2, W3, b3


def pack_torch_mlp(seq):
    """Flatten a reference `mlp(...)` nn.Sequential (Linear,Tanh,Linear,Tanh,Linear,Identity) to `[P]`."""
    import torch
    lin = [m for m in seq if isinstance(m, torch.nn.Linear)]
    assert len(lin) == 3, "packed layout supports exactly two hidden layers"
    parts = []
    for m in lin:
        parts += [m.weight.detach().numpy().reshape(-1), m.bias.detach().numpy().reshape(-1)]
    return np.concatenate(parts).astype(F32)


def init_linear_default(rs, N, in_dim, h, out):
    """nn.Linear default init (uniform(+-1/sqrt(fan_in)) for weight and bias) from a RandomState."""
    def lin(o, i):
        b = 1.0 / np.sqrt(i)
        return [rs.uniform(-b, b, size=(N, o * i)), rs.uniform(-b, b, size=(N, o))]
    return np.concatenate(lin(h, in_dim) + lin(h, h) + lin(out, h), axis=1).astype(F32)


def features(s, lamb, S, one_hot, state_norm=1.0, extra=None):
    """x `[N,E,in]` fp32 from states `[N,E]`, lambda `[E]` (and optional extra inputs `[E,X]`)."""
    s = np.asarray(s)
    N, E = s.shape
    if one_hot:
        oh = np.zeros((N, E, S), dtype=F32)
        np.put_along_axis(oh, s.astype(np.int64)[..., None], 1.0, axis=2)
        x = oh
    else:
        x = (s.astype(F32) / F32(state_norm))[..., None]
    lam = np.broadcast_to(np.asarray(lamb, F32)[None, :, None], (N, E, 1))
    parts = [x, lam]
    if extra is not None:
        parts.append(np.broadcast_to(np.asarray(extra, F32)[None], (N,) + np.asarray(extra).shape))
    return np.concatenate(parts, axis=2)


def mlp_forward(W, x, in_dim, h, out, dtype=F32):
    """Batched per-arm MLP: W `[N,P]`, x `[N,M,in]` -> `[N,M,out]` (tanh hidden, identity output)."""
    W1, b1, W2, b2, W3, b3 = [p.astype(dtype) for p in unpack(W, in_dim, h, out)]
    x = x.astype(dtype)
    a1 = np.tanh(np.einsum('nmi,nji->nmj', x, W1) + b1[:, None, :])
    a2 = np.tanh(np.einsum('nmk,njk->nmj', a1, W2) + b2[:, None, :])
    return np.einsum('nmk,njk->nmj', a2, W3) + b3[:, None, :]


def softmax_logits(logits):
    """torch Categorical(logits=...) normalisation (rmabppo_core.py:79-81): logits - logsumexp."""
    m = logits.max(axis=-1, keepdims=True)
    z = logits - m
    lse = np.log(np.exp(z).sum(axis=-1, keepdims=True))
    logp = z - lse
    return np.exp(logp), logp


def ac_step(Wpi, Wv, s, lamb, u, S, A, h, one_hot=True, state_norm=1.0, dtype=F32):
    """rmabppo_core.py:198-244 for all arms x envs.  Returns (a uint8, v, logp, p1, probs)."""
    in_dim = (S if one_hot else 1) + 1
    x = features(s, lamb, S, one_hot, state_norm)
    probs, logp_all = softmax_logits(mlp_forward(Wpi, x, in_dim, h, A, dtype))
    a = philox.categorical(probs.astype(F32), u)
    logp = np.take_along_axis(logp_all, a[..., None], axis=2)[..., 0]
    v = mlp_forward(Wv, x, in_dim, h, 1, dtype)[..., 0]
    return a.astype(np.uint8), v.astype(F32), logp.astype(F32), probs[..., 1].astype(F32), probs.astype(F32)


def lambda_net_forward(params, obs):
    """MLPLambdaNet (rmabppo_core.py:108-115,170-171): N -> 8 -> 8 -> 1, tanh hidden.

    params = (W1[8,N], b1[8], W2[8,8], b2[8], W3[1,8], b3[1]); obs `[E,N]` -> `[E]`.
    """
    W1, b1, W2, b2, W3, b3 = [np.asarray(p, F32) for p in params]
    a1 = np.tanh(obs.astype(F32) @ W1.T + b1)
    a2 = np.tanh(a1 @ W2.T + b2)
    return (a2 @ W3.T + b3)[:, 0]
