"""PPO-clip update restatement for N independent per-arm nets (oracle, test-only).

Follows robust_rmab/algos/rmabppo/agent_oracle.py:
  compute_loss_pi      :285-322  ratio = exp(logp - logp_old); -mean(min(ratio*adv, clamp(ratio,1-c,1+c)*adv))
                                 - entropy_coeff * mean(H(pi))
  compute_loss_v       :325-339  mean((v - ret)^2)
  compute_loss_lambda  :360-376  lambda_net(obs_0) * (B/(1-gamma) - cdcost_0)
  update               :399-454  train_pi_iters full-batch Adam steps, then train_v_iters; no KL early stop
  optimisers           :383-391  Adam(lr) with torch defaults (betas .9/.999, eps 1e-8), SGD(lm_lr) for lambda
  entropy schedule     :402-409, :491, :578
torch autograd on packed `[N,P]` weights: the summed loss over arms has the same per-arm gradients
as the reference's N separate backward() calls because arms share no parameters.  Adam is
elementwise, so one Adam over `[N,P]` equals N per-arm Adam objects.
"""
import numpy as np
import torch

from . import nets


def _unpack_t(W, in_dim, h, out):
    N = W.shape[0]
    o = 0
    W1 = W[:, o:o + h * in_dim].reshape(N, h, in_dim); o += h * in_dim
    b1 = W[:, o:o + h]; o += h
    W2 = W[:, o:o + h * h].reshape(N, h, h); o += h * h
    b2 = W[:, o:o + h]; o += h
    W3 = W[:, o:o + out * h].reshape(N, out, h); o += out * h
    b3 = W[:, o:o + out]
    return W1, b1, W2, b2, W3, b3


def mlp_t(W, x, in_dim, h, out):
    """W `[N,P]`, x `[N,M,in]` -> `[N,M,out]`."""
    W1, b1, W2, b2, W3, b3 = _unpack_t(W, in_dim, h, out)
    a1 = torch.tanh(torch.bmm(x, W1.transpose(1, 2)) + b1[:, None, :])
    a2 = torch.tanh(torch.bmm(a1, W2.transpose(1, 2)) + b2[:, None, :])
    return torch.bmm(a2, W3.transpose(1, 2)) + b3[:, None, :]


def loss_pi(Wpi, x, act, adv, logp_old, clip_ratio, entropy_coeff, in_dim, h, A):
    """Per-arm policy loss `[N]` (agent_oracle.py:301-309).  x `[N,M,in]`, act/adv/logp_old `[N,M]`."""
    logits = mlp_t(Wpi, x, in_dim, h, A)
    logp_all = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    logp = torch.gather(logp_all, 2, act.long()[..., None])[..., 0]
    p = torch.exp(logp_all)
    ent = -(p * logp_all).sum(-1).mean(dim=1)
    ratio = torch.exp(logp - logp_old)
    clip_adv = torch.clamp(ratio, 1 - clip_ratio, 1 + clip_ratio) * adv
    loss = -(torch.min(ratio * adv, clip_adv)).mean(dim=1)
    return loss - entropy_coeff * ent


def loss_v(Wv, x, ret, in_dim, h):
    """Per-arm value loss `[N]` (agent_oracle.py:337-338)."""
    v = mlp_t(Wv, x, in_dim, h, 1)[..., 0]
    return ((v - ret) ** 2).mean(dim=1)


class ArmOptim:
    """The 2N Adam optimisers of agent_oracle.py:383-388 as two Adams over packed weights."""

    def __init__(self, Wpi, Wv, pi_lr, vf_lr):
        self.Wpi = torch.nn.Parameter(torch.as_tensor(np.array(Wpi)).clone())
        self.Wv = torch.nn.Parameter(torch.as_tensor(np.array(Wv)).clone())
        self.opt_pi = torch.optim.Adam([self.Wpi], lr=pi_lr)
        self.opt_v = torch.optim.Adam([self.Wv], lr=vf_lr)


def update(opt, s, a, adv, logp_old, ret, lamb, S, A, h, clip_ratio, entropy_coeff,
           pi_iters, v_iters, one_hot=True, state_norm=1.0):
    """One `update()` call (agent_oracle.py:399-436) minus the lambda step.

    s, a `[N,T,E]` ints; adv (already normalised), logp_old, ret `[N,T,E]`; lamb `[E]`.
    Returns per-iteration losses (`[pi_iters,N]`, `[v_iters,N]`) evaluated BEFORE each step.
    """
    N, T, E = s.shape
    dt = opt.Wpi.dtype
    in_dim = (S if one_hot else 1) + 1
    x = np.stack([nets.features(s[:, t], lamb, S, one_hot, state_norm) for t in range(T)], axis=1)
    x = torch.as_tensor(x.reshape(N, T * E, in_dim)).to(dt)
    flat = lambda z: torch.as_tensor(np.ascontiguousarray(z).reshape(N, T * E))
    act, adv_t, lpo, ret_t = flat(a.astype(np.int64)), flat(adv).to(dt), flat(logp_old).to(dt), flat(ret).to(dt)
    lp_hist, lv_hist = [], []
    for _ in range(pi_iters):
        opt.opt_pi.zero_grad()
        l = loss_pi(opt.Wpi, x, act, adv_t, lpo, clip_ratio, entropy_coeff, in_dim, h, A)
        l.sum().backward()
        opt.opt_pi.step()
        lp_hist.append(l.detach().numpy().copy())
    for _ in range(v_iters):
        opt.opt_v.zero_grad()
        l = loss_v(opt.Wv, x, ret_t, in_dim, h)
        l.sum().backward()
        opt.opt_v.step()
        lv_hist.append(l.detach().numpy().copy())
    return np.array(lp_hist), np.array(lv_hist)


def entropy_coeff_for(epoch, epochs, start, end, lamb_update_freq, final_train_lambdas):
    """agent_oracle.py:402-409 with head = linspace(start,end,epochs)[epoch] (:491,:578)."""
    head = np.linspace(start, end, epochs)[epoch]
    if (epochs - epoch) > final_train_lambdas:
        return float(np.linspace(head, 0, lamb_update_freq)[epoch % lamb_update_freq])
    return 0.0


def lambda_update_due(epoch, epochs, lamb_update_freq, final_train_lambdas):
    """agent_oracle.py:440."""
    return epoch % lamb_update_freq == 0 and epoch > 0 and (epochs - epoch) > final_train_lambdas


def loss_statistics(s, a, adv_n, ret, S, A):
    """Per-(arm, env, state, action) sufficient statistics of compute_loss_pi / compute_loss_v for one-hot
    inputs (the regrouping the CUDA table update uses; csrc/epoch.cu).  s, a `[N,T,E]` ints; adv_n (normalised),
    ret `[N,T,E]`.  Returns dict of float64 arrays: Ap, Am, cnt `[N,S*A,E]`; cntS, mean_ret `[N,S,E]`;
    ret_var_sum `[N]` = sum_t (ret_t - mean_ret[s_t, e])^2.
    """
    s = np.asarray(s).astype(np.int64)
    a = np.asarray(a).astype(np.int64)
    N, T, E = s.shape
    x = np.asarray(adv_n, np.float64)
    r = np.asarray(ret, np.float64)
    Ap = np.zeros((N, S * A, E)); Am = np.zeros((N, S * A, E)); cnt = np.zeros((N, S * A, E))
    cntS = np.zeros((N, S, E)); Sr = np.zeros((N, S, E))
    for si in range(S):
        ms = s == si
        cntS[:, si] = ms.sum(1)
        Sr[:, si] = (r * ms).sum(1)
        for ai in range(A):
            m = ms & (a == ai)
            c = si * A + ai
            Ap[:, c] = (np.where(x > 0, x, 0.0) * m).sum(1)
            Am[:, c] = (np.where(x < 0, x, 0.0) * m).sum(1)
            cnt[:, c] = m.sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        mean_ret = np.where(cntS > 0, Sr / cntS, 0.0)
    mr = np.take_along_axis(mean_ret.astype(np.float32).astype(np.float64), s, axis=1)  # [N,T,E]: mean_ret[n, s_t, e]
    var_sum = ((r - mr) ** 2).sum(axis=(1, 2))
    return dict(Ap=Ap, Am=Am, cnt=cnt, cntS=cntS, mean_ret=mean_ret, ret_var_sum=var_sum)
