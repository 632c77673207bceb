"""Torch-tensor front end of the C ABI: validates dtype / device / contiguity, passes data_ptr() and the
current CUDA stream.  PyTorch is only the memory and stream plumbing here; all compute is in
librmab_b200.so.  Every function fails loudly on CPU tensors (there is no fallback path).

Layouts (arm-major, env fastest): states `[N,E]` uint8, trajectories `[N,T,E]`, packed weights `[N,P]`.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import RmabHyper, RmabNetDesc, RmabRng, check, lib

u8, f32, f64, i32 = torch.uint8, torch.float32, torch.float64, torch.int32


def _p(t, dtype=None, name="tensor"):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: robustrmab_b200 runs on CUDA tensors only (no CPU fallback)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rng(seed, stream, t, env0=0, arm0=0, t_base=None):
    """t_base: optional int32 CUDA tensor of one element that the kernel adds to t when it RUNS (a captured CUDA graph
    replays with the counter the tensor holds at replay time)."""
    r = RmabRng(int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream), int(t) & 0xFFFFFFFF, int(env0), int(arm0),
                None if t_base is None else t_base.data_ptr())
    r._keep = t_base   # the struct only holds the address
    return r


def net_desc(N, E, S, A, hidden, one_hot=True, state_norm=1.0, n_extra=0):
    return RmabNetDesc(int(N), int(E), int(S), int(A), int(hidden), int(bool(one_hot)),
                       float(state_norm if state_norm else 1.0), int(n_extra))


def n_params(in_dim, h, out):
    return (in_dim + 1) * h + (h + 1) * h + (h + 1) * out


def in_dim(net):
    return (net.n_states if net.one_hot else 1) + 1


def _rng_ptr(r):
    return C.byref(r) if r is not None else None


# ---------------------------------------------------------------------------------------------- env
def reset_states(N, E, S, r=None, u=None, out=None, device="cuda"):
    out = torch.empty((N, E), dtype=u8, device=device) if out is None else out
    check(lib().rmab_reset_states(E, N, S, _rng_ptr(r), _p(u, f32, "u"), _p(out, u8, "s"), _stream()))
    return out


def env_step_table(domain, s, a, nature, ranges, R, r=None, u=None, nature_per_env=False, s_next=None,
                   rew=None, want_reward=True):
    N, E = s.shape
    s_next = torch.empty_like(s) if s_next is None else s_next
    if rew is None and want_reward:
        rew = torch.empty((N, E), dtype=f32, device=s.device)
    check(lib().rmab_env_step_table(int(domain), E, N, _p(s, u8, "s"), _p(a, u8, "a"), _p(nature, f32, "nature"),
                                    int(bool(nature_per_env)), _p(ranges, f32, "ranges"), _p(R, f32, "R"),
                                    _rng_ptr(r), _p(u, f32, "u"), _p(s_next, u8, "s_next"), _p(rew, f32, "r"),
                                    _stream()))
    return s_next, rew


def env_step_sis(pop, s, a, params, R, r=None, u=None, nature_per_env=False, s_next=None, rew=None,
                 want_reward=True):
    N, E = s.shape
    s_next = torch.empty_like(s) if s_next is None else s_next
    if rew is None and want_reward:
        rew = torch.empty((N, E), dtype=f32, device=s.device)
    check(lib().rmab_env_step_sis(E, N, int(pop), _p(s, u8, "s"), _p(a, u8, "a"), _p(params, f32, "params"),
                                  int(bool(nature_per_env)), _p(R, f32, "R"), _rng_ptr(r), _p(u, f32, "u"),
                                  _p(s_next, u8, "s_next"), _p(rew, f32, "r"), _stream()))
    return s_next, rew


def sis_table(pop, params):
    N = params.shape[0]
    T = torch.empty((N, pop + 1, 3, pop + 1), dtype=f32, device=params.device)
    check(lib().rmab_sis_table(N, int(pop), _p(params, f32, "params"), _p(T, f32, "T"), _stream()))
    return T


def bound_nature(x, lo, hi):
    out = torch.empty_like(x)
    check(lib().rmab_bound_nature(x.numel(), _p(x, f32, "x"), _p(lo, f32, "lo"), _p(hi, f32, "hi"), _p(out, f32, "out"),
                                  _stream()))
    return out


# ------------------------------------------------------------------------------------- actor-critic
def ac_step(net, Wpi, Wv, s, lamb, r=None, u=None, want_probs=False, extra=None):
    N, E = s.shape
    dev = s.device
    a = torch.empty((N, E), dtype=u8, device=dev)
    v = torch.empty((N, E), dtype=f32, device=dev)
    logp = torch.empty((N, E), dtype=f32, device=dev)
    p1 = torch.empty((N, E), dtype=f32, device=dev)
    probs = torch.empty((N, net.n_actions, E), dtype=f32, device=dev) if want_probs else None
    check(lib().rmab_ac_step(C.byref(net), _p(Wpi, f32, "Wpi"), _p(Wv, f32, "Wv"), _p(s, u8, "s"),
                             _p(lamb, f32, "lamb"), _p(extra, f32, "extra"), _rng_ptr(r), _p(u, f32, "u"), _p(a, u8), _p(v, f32),
                             _p(logp, f32), _p(p1, f32), _p(probs, f32), _stream()))
    return a, v, logp, p1, probs


_LAMBDA_WS = {}


def lambda_forward(params, s, n_total, arm0=0, obs_scale=1.0, h1_in=None, want_h1=False, want_lamb=True):
    N, E = s.shape
    dev = s.device
    h1 = torch.empty((E, 8), dtype=f32, device=dev) if want_h1 else None
    lamb = torch.empty((E,), dtype=f32, device=dev) if want_lamb else None
    need = int(lib().rmab_lambda_ws_bytes(E, N))
    ws = _LAMBDA_WS.get((dev, need))
    if ws is None:
        ws = _LAMBDA_WS[(dev, need)] = torch.empty((need + 3) // 4, dtype=f32, device=dev)
    check(lib().rmab_lambda_forward(E, N, _p(params, f32, "params"), int(n_total), int(arm0), _p(s, u8, "s"),
                                    float(obs_scale), _p(h1_in, f32, "h1_in"), _p(h1, f32), _p(lamb, f32), _p(ws, f32),
                                    ws.numel() * 4, _stream()))
    return lamb, h1


def lambda_sgd(params, s, n_total, arm0, obs_scale, h1, coef, lr, want_grad=False):
    N, E = s.shape
    ws = torch.empty((E, 8), dtype=f32, device=s.device)
    grad = torch.zeros_like(params) if want_grad else None
    check(lib().rmab_lambda_sgd(E, N, _p(params, f32, "params"), int(n_total), int(arm0), _p(s, u8, "s"),
                                float(obs_scale), _p(h1, f32, "h1"), _p(coef, f32, "coef"), float(lr), _p(ws, f32),
                                _p(grad, f32), _stream()))
    return grad


def rollout_table(domain, net, T, Wpi, Wv, lamb, nature, ranges, seed, t_act0, t_env0, s_traj, a, logp, val,
                  last_val, env0=0, arm0=0):
    check(lib().rmab_rollout_table(int(domain), C.byref(net), int(T), _p(Wpi, f32, "Wpi"), _p(Wv, f32, "Wv"),
                                   _p(lamb, f32, "lamb"), _p(nature, f32, "nature"), _p(ranges, f32, "ranges"),
                                   int(seed) & 0xFFFFFFFFFFFFFFFF, int(t_act0), int(t_env0), int(env0), int(arm0),
                                   _p(s_traj, u8, "s_traj"), _p(a, u8, "a"), _p(logp, f32, "logp"), _p(val, f32, "val"),
                                   _p(last_val, f32, "last_val"), _stream()))


# ------------------------------------------------------------------------------ MA-RMABPPO rollout
def gaussian_sample(mu, log_std, r=None, eps=None):
    """a = mu + exp(log_std) * eps, logp = sum_d Normal.log_prob(a).  mu [E,D] -> (a [E,D], logp [E])."""
    E, D = mu.shape
    a = torch.empty_like(mu)
    logp = torch.empty((E,), dtype=f32, device=mu.device)
    check(lib().rmab_gaussian_sample(E, D, _p(mu, f32, "mu"), _p(log_std, f32, "log_std"), _rng_ptr(r),
                                     _p(eps, f32, "eps"), _p(a, f32), _p(logp, f32), _stream()))
    return a, logp


def env_counterfactual(domain, s, a, nature, ranges, R, r, n_trials, nature_per_env=False, out=None):
    """Mean over n_trials independent env steps from the same state of the reward, per (arm, env)."""
    N, E = s.shape
    out = torch.empty((N, E), dtype=f32, device=s.device) if out is None else out
    check(lib().rmab_env_counterfactual(int(domain), E, N, R.shape[1], _p(s, u8, "s"), _p(a, u8, "a"),
                                        _p(nature, f32, "nature"), int(bool(nature_per_env)), _p(ranges, f32, "ranges"),
                                        _p(R, f32, "R"), _rng_ptr(r), int(n_trials), _p(out, f32), _stream()))
    return out


def critic_extra_iter(net, hp, Wv, adam_v, s, ret, lamb, z1_extra, want_loss=False, sample_major=False):
    """One Adam iteration of the agent critics with nature inputs; returns (dz1, loss [N] or None).  z1_extra and dz1 are
    `[N,T*E,h]`, or `[T*E,N,h]` with sample_major (the row-major result of one dense GEMM over all arms)."""
    N, T, E = ret.shape
    h = net.hidden
    assert tuple(z1_extra.shape) == ((T * E, N, h) if sample_major else (N, T * E, h)), "z1_extra shape"
    dz1 = torch.zeros_like(z1_extra)
    loss = torch.zeros((N, 1), dtype=f32, device=ret.device) if want_loss else None
    check(lib().rmab_critic_extra_iter(C.byref(net), C.byref(hp), T, s.shape[1], _p(Wv, f32, "Wv"), _p(adam_v, f32, "adam_v"),
                                       _p(s, u8, "s"), _p(ret, f32, "ret"), _p(lamb, f32, "lamb"),
                                       _p(z1_extra, f32, "z1_extra"), _p(dz1, f32), int(bool(sample_major)), _p(loss, f32),
                                       _stream()))
    return dz1, loss


# Dense GEMMs of the MA agent critics' nature block (library GEMMs, not kernels of this package): fp32 accuracy at
# tensor-core speed by the same 3xTF32 split the update kernels use -- x = hi + lo with hi exactly representable in tf32,
# a.b ~ lo_a.hi_b + hi_a.lo_b + hi_a.hi_b, each product a TF32 library GEMM with fp32 accumulation.
def split_tf32(x):
    hi = (x.contiguous().view(torch.int32) & -8192).view(f32)
    return hi, x - hi


def matmul_3xtf32(a, b):
    """a @ b (torch.matmul broadcasting rules); a / b are tensors or (hi, lo) pairs from split_tf32."""
    ah, al = a if isinstance(a, tuple) else split_tf32(a)
    bh, bl = b if isinstance(b, tuple) else split_tf32(b)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        out = torch.matmul(al, bh)
        out += torch.matmul(ah, bl)
        out += torch.matmul(ah, bh)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


# The package's own dense product for the same job (csrc/gemm3h.cu): x = rs * (hi + lo) with fp16 hi / lo and a per-row
# power-of-two scale, three kind::f16 tcgen05 products per k-block with fp32 accumulation in TMEM, operands fed by TMA.
f16 = torch.float16


def _ldk(K):
    return (int(K) + 7) // 8 * 8


def split16_rows(x, out=None):
    """x `[R,K]` fp32 (row stride may exceed K) -> (hi, lo `[R,ldk]` fp16, rs `[R]`): the K-major operand form of gemm3h.
    `out` = a previous result to overwrite (static buffers for CUDA-graph replay)."""
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == f32 and x.is_cuda, "split16_rows: [R,K] fp32 CUDA rows"
    R, K = x.shape
    ldk = _ldk(K)
    if out is None:
        hi = torch.empty((R, ldk), dtype=f16, device=x.device)
        lo = torch.empty((R, ldk), dtype=f16, device=x.device)
        rs = torch.empty((R,), dtype=f32, device=x.device)
    else:
        hi, lo, rs = out
        assert hi.shape == (R, ldk) and lo.shape == (R, ldk) and rs.shape == (R,)
    check(lib().rmab_split16_rows(R, K, C.c_void_p(x.data_ptr()), x.stride(0), ldk, _p(hi, f16), _p(lo, f16), _p(rs, f32),
                                  _stream()))
    return hi, lo, rs


def split16_cols(x):
    """x `[Rk,Cn]` fp32 -> the split of x^T: (hi, lo `[Cn,ldk]` fp16 with k = the row index of x, rs `[Cn]`)."""
    assert x.dim() == 2 and x.stride(1) == 1 and x.dtype == f32 and x.is_cuda, "split16_cols: [Rk,Cn] fp32 CUDA rows"
    Rk, Cn = x.shape
    ldk = _ldk(Rk)
    hi = torch.empty((Cn, ldk), dtype=f16, device=x.device)
    lo = torch.empty((Cn, ldk), dtype=f16, device=x.device)
    rs = torch.empty((Cn,), dtype=f32, device=x.device)
    amax = torch.empty((Cn,), dtype=i32, device=x.device)
    check(lib().rmab_split16_cols(Rk, Cn, C.c_void_p(x.data_ptr()), x.stride(0), ldk, _p(hi, f16), _p(lo, f16), _p(rs, f32),
                                  _p(amax, i32), _stream()))
    return hi, lo, rs


def gemm3h(a, b, K, out=None, n_fast=True):
    """a = (Ah, Al, rsA) `[M,ldk]`, b = (Bh, Bl, rsB) `[N,ldk]` from split16_* -> a . b^T `[M,N]` fp32."""
    Ah, Al, rsA = a
    Bh, Bl, rsB = b
    M, ldk = Ah.shape
    N = Bh.shape[0]
    assert Bh.shape[1] == ldk and ldk >= K, "gemm3h: operands must share the padded K"
    out = torch.empty((M, N), dtype=f32, device=Ah.device) if out is None else out
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype == f32
    check(lib().rmab_gemm3h(M, N, int(K), ldk, _p(Ah, f16), _p(Al, f16), _p(rsA, f32), _p(Bh, f16), _p(Bl, f16), _p(rsB, f32),
                            C.c_void_p(out.data_ptr()), out.stride(0), int(bool(n_fast)), _stream()))
    return out


def gemm3h_adam(a, b, K, p, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, n_fast=True):
    """p `[M,N]` (with Adam moments m, v) <- Adam step with the gradient a . b^T, which is consumed in the product's epilogue."""
    Ah, Al, rsA = a
    Bh, Bl, rsB = b
    M, ldk = Ah.shape
    N = Bh.shape[0]
    assert p.shape == (M, N) and p.stride(1) == 1 and m.shape == p.shape and v.shape == p.shape and m.is_contiguous() and \
        v.is_contiguous() and p.stride(0) == N, "gemm3h_adam: parameter and moments must be contiguous [M,N]"
    check(lib().rmab_gemm3h_adam(M, N, int(K), ldk, _p(Ah, f16), _p(Al, f16), _p(rsA, f32), _p(Bh, f16), _p(Bl, f16), _p(rsB, f32),
                                 _p(p, f32), _p(m, f32), _p(v, f32), p.stride(0), float(lr), beta1, beta2, eps, int(step),
                                 int(bool(n_fast)), _stream()))


def adam_step(p, grad, m, v, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    check(lib().rmab_adam_step(p.numel(), _p(p, f32, "p"), _p(grad, f32, "grad"), _p(m, f32, "m"), _p(v, f32, "v"),
                               float(lr), float(beta1), float(beta2), float(eps), int(step), _stream()))


# -------------------------------------------------------------------------------------------- buffer
def gae(s_traj, a, val, last_val, lamb, Cc, R, gamma, lam, normalize=True, want_q=False, want_stats=False):
    N, T, E = a.shape
    S, A = R.shape[1], Cc.shape[0]
    dev = a.device
    adv = torch.empty((N, T, E), dtype=f32, device=dev)
    ret = torch.empty((N, T, E), dtype=f32, device=dev)
    q = torch.empty((N, T, E), dtype=f32, device=dev) if want_q else None
    stats = torch.empty((N, 2), dtype=f32, device=dev) if want_stats else None
    check(lib().rmab_gae(T, E, N, S, A, _p(s_traj, u8, "s_traj"), _p(a, u8, "a"), _p(val, f32, "val"),
                         _p(last_val, f32, "last_val"), _p(lamb, f32, "lamb"), _p(Cc, f32, "C"), _p(R, f32, "R"),
                         float(gamma), float(lam), int(bool(normalize)), _p(adv, f32), _p(ret, f32), _p(q, f32),
                         _p(stats, f32), _stream()))
    return adv, ret, q, stats


def gae_explicit(rew, cost, val, last_val, lamb, gamma, lam, normalize=True, want_q=False, want_stats=False):
    N, T, E = rew.shape
    dev = rew.device
    adv = torch.empty((N, T, E), dtype=f32, device=dev)
    ret = torch.empty((N, T, E), dtype=f32, device=dev)
    q = torch.empty((N, T, E), dtype=f32, device=dev) if want_q else None
    stats = torch.empty((N, 2), dtype=f32, device=dev) if want_stats else None
    check(lib().rmab_gae_explicit(T, E, N, _p(rew, f32, "rew"), _p(cost, f32, "cost"), _p(val, f32, "val"),
                                  _p(last_val, f32, "last_val"), _p(lamb, f32, "lamb"), float(gamma), float(lam),
                                  int(bool(normalize)), _p(adv, f32), _p(ret, f32), _p(q, f32), _p(stats, f32),
                                  _stream()))
    return adv, ret, q, stats


def cost_togo(a, Cc, gamma):
    N, T, E = a.shape
    ws = torch.empty((T, E), dtype=f64, device=a.device)
    out = torch.empty((T, E), dtype=f32, device=a.device)
    check(lib().rmab_cost_togo(T, E, N, Cc.shape[0], _p(a, u8, "a"), _p(Cc, f32, "C"), float(gamma), _p(ws, f64),
                               _p(out, f32), _stream()))
    return out


# ---------------------------------------------------------------------------------------- PPO update
def hyper(clip_ratio, entropy_coeff, pi_lr, vf_lr, pi_iters, v_iters, step0_pi=0, step0_v=0, beta1=0.9,
          beta2=0.999, eps=1e-8):
    return RmabHyper(float(clip_ratio), float(entropy_coeff), float(pi_lr), float(vf_lr), beta1, beta2, eps,
                     int(pi_iters), int(v_iters), int(step0_pi), int(step0_v))


def ppo_update(net, hp, Wpi, Wv, adam_pi, adam_v, s, a, adv, logp_old, ret, lamb, want_loss=False):
    """s may be the `[N,T+1,E]` trajectory (its first T rows are used) or `[N,T,E]`."""
    N, T, E = a.shape
    s_rows = s.shape[1]
    dev = a.device
    lp = torch.zeros((N, max(hp.pi_iters, 1)), dtype=f32, device=dev) if want_loss else None
    lv = torch.zeros((N, max(hp.v_iters, 1)), dtype=f32, device=dev) if want_loss else None
    check(lib().rmab_ppo_update(C.byref(net), C.byref(hp), T, s_rows, _p(Wpi, f32, "Wpi"), _p(Wv, f32, "Wv"),
                                _p(adam_pi, f32, "adam_pi"), _p(adam_v, f32, "adam_v"), _p(s, u8, "s"), _p(a, u8, "a"),
                                _p(adv, f32, "adv"), _p(logp_old, f32, "logp_old"), _p(ret, f32, "ret"),
                                _p(lamb, f32, "lamb"), _p(lp, f32), _p(lv, f32), _stream()))
    return lp, lv


# --------------------------------------------------------------------------- fused table-domain epoch
def stat_rows(domain):
    return int(lib().rmab_stat_rows(int(domain)))


def epoch_table(domain, net, T, Wpi, Wv, lamb, nature, ranges, R, Cc, gamma, lam, seed, t_act0, t_env0, s_traj, a,
                stats=None, arm_stats=None, env_sums=None, ws=None, adv=None, ret=None, last_val=None, env0=0, arm0=0):
    """Rollout + GAE + normalisation + loss statistics in one launch.  Returns (stats, arm_stats, env_sums)."""
    N, E = net.n_arms, net.n_envs
    dev = s_traj.device
    K = stat_rows(domain)
    stats = torch.empty((N, K, E), dtype=f32, device=dev) if stats is None else stats
    arm_stats = torch.empty((N, 4), dtype=f32, device=dev) if arm_stats is None else arm_stats
    env_sums = torch.empty((2, E), dtype=f32, device=dev) if env_sums is None else env_sums
    need = int(lib().rmab_epoch_ws_bytes(N, E))
    if ws is None or ws.numel() * ws.element_size() < need:
        ws = torch.empty((max(need, 8) + 7) // 8, dtype=f64, device=dev)
    check(lib().rmab_epoch_table(int(domain), C.byref(net), int(T), _p(Wpi, f32, "Wpi"), _p(Wv, f32, "Wv"),
                                 _p(lamb, f32, "lamb"), _p(nature, f32, "nature"), _p(ranges, f32, "ranges"),
                                 _p(R, f32, "R"), _p(Cc, f32, "C"), float(gamma), float(lam),
                                 int(seed) & 0xFFFFFFFFFFFFFFFF, int(t_act0), int(t_env0), int(env0), int(arm0),
                                 _p(s_traj, u8, "s_traj"), _p(a, u8, "a"), _p(stats, f32, "stats"),
                                 _p(arm_stats, f32, "arm_stats"), _p(env_sums, f32, "env_sums"), _p(ws, f64, "ws"),
                                 ws.numel() * ws.element_size(), _p(adv, f32, "adv"), _p(ret, f32, "ret"),
                                 _p(last_val, f32, "last_val"), _stream()))
    return stats, arm_stats, env_sums


def epoch_table_envshard(domain, net, T, Wpi, Wv, lamb, nature, ranges, R, Cc, gamma, lam, seed, t0, s_traj, stats, arm_stats,
                         env_sums, ws, last_val, env0, mom_out=None, mom_in=None):
    """rmab_epoch_table for env-sharded ranks: launch once with mom_out `[N,2]` fp64 (raw advantage sums of this rank's envs),
    all-reduce, launch again with mom_in `[N,2]` fp32 (global mean, std)."""
    check(lib().rmab_epoch_table_envshard(int(domain), C.byref(net), int(T), _p(Wpi, f32), _p(Wv, f32), _p(lamb, f32),
                                          _p(nature, f32), _p(ranges, f32), _p(R, f32), _p(Cc, f32), float(gamma), float(lam),
                                          int(seed) & 0xFFFFFFFFFFFFFFFF, int(t0), int(t0), int(env0), 0, _p(s_traj, u8), None,
                                          _p(stats, f32), _p(arm_stats, f32), _p(env_sums, f32), _p(ws, f64),
                                          ws.numel() * ws.element_size(), _p(last_val, f32), _p(mom_out, f64), _p(mom_in, f32),
                                          _stream()))


def ppo_update_table(net, hp, T, Wpi, Wv, adam_pi, adam_v, stats, arm_stats, lamb, want_loss=False):
    N = net.n_arms
    dev = stats.device
    lp = torch.zeros((N, max(hp.pi_iters, 1)), dtype=f32, device=dev) if want_loss else None
    lv = torch.zeros((N, max(hp.v_iters, 1)), dtype=f32, device=dev) if want_loss else None
    check(lib().rmab_ppo_update_table(C.byref(net), C.byref(hp), int(T), _p(Wpi, f32, "Wpi"), _p(Wv, f32, "Wv"),
                                      _p(adam_pi, f32, "adam_pi"), _p(adam_v, f32, "adam_v"), _p(stats, f32, "stats"),
                                      _p(arm_stats, f32, "arm_stats"), _p(lamb, f32, "lamb"), _p(lp, f32), _p(lv, f32),
                                      _stream()))
    return lp, lv


# ------------------------------------------------------------------------------------------ selection
_SELECT_WS = {}


def _select_ws(E, N, dev):
    need = int(lib().rmab_select_ws_bytes(E, N, 2))
    ws = _SELECT_WS.get((dev, need))
    if ws is None:
        ws = _SELECT_WS[(dev, need)] = torch.empty((need + 3) // 4, dtype=i32, device=dev)
    return ws, need


def budget_select_binary(p1, cost1, B, positive_only=False):
    N, E = p1.shape
    out = torch.empty((N, E), dtype=u8, device=p1.device)
    ws, need = _select_ws(E, N, p1.device) if N * E >= (1 << 21) else (None, 0)
    check(lib().rmab_budget_select_binary(E, N, _p(p1, f32, "p1"), float(cost1), float(B), int(bool(positive_only)),
                                          _p(out, u8), _p(ws, i32), need, _stream()))
    return out


def select_count(N, cost1, B):
    """Number of ones the while-loop of rmabppo_core.py:320-331 hands out."""
    k, spent = 0, 0.0
    while spent < float(B) and k < N:
        if spent + float(cost1) > float(B):
            break
        spent += float(cost1)
        k += 1
    return k


def budget_select_binary_sharded(p1, cost1, B, n_total, rank, world, group=None, positive_only=False):
    """Global top-k over arms that are sharded across ranks (contiguous arm ranges in rank order): p1 `[N_local,E]` of
    this rank -> its rows of the actions.  Per 8-bit radix pass one all-reduce of the per-env digit histograms
    `[ceil(E/32), 256, 32]`, then one all-gather of the per-env boundary-tie counts so that ties go to the
    higher GLOBAL arm index."""
    import torch.distributed as dist
    N, E = p1.shape
    dev = p1.device
    out = torch.empty((N, E), dtype=u8, device=dev)
    k = select_count(n_total, cost1, B)
    if k <= 0 or k >= n_total:
        out.fill_(1 if k >= n_total else 0)
        if positive_only and k >= n_total:
            out.copy_((p1 > 0).to(u8))
        return out
    G = (E + 31) // 32
    Nw = max(N, 1)
    ws, need = _select_ws(E, Nw, dev)
    hist = torch.zeros((G, 256, 32), dtype=i32, device=dev)
    local3 = None
    for ps in range(4):
        if N > 0:
            check(lib().rmab_select_hist(E, N, _p(p1, f32, "p1"), ps, _p(ws, i32), need, _p(hist, i32), _stream()))
        else:
            hist.zero_()
        if ps == 3:
            local3 = hist.clone()
        if world > 1:
            dist.all_reduce(hist, group=group)
        check(lib().rmab_select_scan(E, Nw, ps, k, _p(hist, i32), _p(ws, i32), need, _stream()))
    # keys equal to the threshold on ranks holding higher arm indices
    tau = ws[:G * 32]
    lane = torch.arange(32, device=dev).repeat(G)
    grp = torch.arange(G, device=dev).repeat_interleave(32)
    eq_local = local3[grp, (tau & 255).long(), lane].contiguous()            # [G*32]
    above = torch.zeros_like(eq_local)
    if world > 1:
        allc = torch.empty((world, G * 32), dtype=i32, device=dev)
        dist.all_gather_into_tensor(allc, eq_local, group=group)
        if rank + 1 < world:
            above = allc[rank + 1:].sum(dim=0).to(i32).contiguous()
    if N > 0:
        check(lib().rmab_select_mark(E, N, _p(p1, f32, "p1"), _p(above, i32), int(bool(positive_only)), _p(out, u8),
                                     _p(ws, i32), need, _stream()))
    return out


def budget_select_multi(probs, Cc, B):
    N, A, E = probs.shape
    out = torch.empty((N, E), dtype=u8, device=probs.device)
    check(lib().rmab_budget_select_multi(E, N, A, _p(probs, f32, "probs"), _p(Cc, f32, "C"), float(B), _p(out, u8),
                                         None, 0, _stream()))
    return out


def knapsack_multichoice(values, costs, B):
    """Exact multiple-choice knapsack per env (lp_methods.py:221-271): values `[N,A,E]` fp64, integer costs `[A]`, integer
    budget met with equality.  Returns (actions `[N,E]` u8, status `[E]` int32: 1 = infeasible)."""
    N, A, E = values.shape
    dev = values.device
    c = torch.as_tensor([int(round(float(x))) for x in costs], dtype=i32, device=dev)
    Bi = int(round(float(B)))
    need = int(lib().rmab_knapsack_ws_bytes(E, N, Bi))
    ws = torch.empty(need, dtype=u8, device=dev)
    actions = torch.empty((N, E), dtype=u8, device=dev)
    status = torch.empty((E,), dtype=i32, device=dev)
    check(lib().rmab_knapsack_multichoice(E, N, A, _p(values, f64, "values"), _p(c, i32), Bi, _p(actions, u8), _p(status, i32),
                                          _p(ws, u8), need, _stream()))
    return actions, status


# ------------------------------------------------------------------------------------ value iteration
def vi_batched(T, R, Cc, lambdas, gamma, eps, stop="full", want_q=False):
    N, S, A, _ = T.shape
    per_arm = lambdas.dim() == 2
    L = lambdas.shape[-1]
    dev = T.device
    V = torch.empty((N, L, S), dtype=f64, device=dev)
    pol = torch.empty((N, L, S), dtype=u8, device=dev)
    iters = torch.empty((N, L), dtype=i32, device=dev)
    mi = torch.empty((N, L), dtype=i32, device=dev)
    Q = torch.empty((N, L, S, A), dtype=f64, device=dev) if want_q else None
    check(lib().rmab_vi_batched(N, S, A, L, _p(T, f64, "T"), _p(R, f64, "R"), _p(Cc, f64, "C"),
                                _p(lambdas, f64, "lambdas"), int(per_arm), float(gamma), float(eps),
                                {"fast": _lib.STOP_FAST, "abs": _lib.STOP_ABS}.get(stop, _lib.STOP_FULL), _p(V, f64), _p(pol, u8),
                                _p(iters, i32), _p(mi, i32), _p(Q, f64), _stream()))
    return V, pol, iters, mi, Q


def vi_general(T, Rsa, gamma, eps, stop="full"):
    """ValueIteration for N MDPs with general rewards: T [N,S,A,S], Rsa [N,S,A] -> (V [N,S], policy, iters, max_iter)."""
    N, S, A, _ = T.shape
    dev = T.device
    V = torch.empty((N, S), dtype=f64, device=dev)
    pol = torch.empty((N, S), dtype=u8, device=dev)
    iters = torch.empty((N,), dtype=i32, device=dev)
    mi = torch.empty((N,), dtype=i32, device=dev)
    check(lib().rmab_vi_general(N, S, A, _p(T, f64, "T"), _p(Rsa, f64, "Rsa"), float(gamma), float(eps),
                                {"fast": _lib.STOP_FAST, "abs": _lib.STOP_ABS}.get(stop, _lib.STOP_FULL), _p(V, f64),
                                _p(pol, u8), _p(iters, i32), _p(mi, i32), _stream()))
    return V, pol, iters, mi


def whittle_bisect(T, R, Cc, gamma, eps, lam_lo, lam_hi, n_rounds, stop="full"):
    N, S, A, _ = T.shape
    idx = torch.empty((N, S), dtype=f64, device=T.device)
    check(lib().rmab_whittle_bisect(N, S, A, _p(T, f64, "T"), _p(R, f64, "R"), _p(Cc, f64, "C"), float(gamma),
                                    float(eps), _lib.STOP_FAST if stop == "fast" else _lib.STOP_FULL, float(lam_lo),
                                    float(lam_hi), int(n_rounds), _p(idx, f64), _stream()))
    return idx


# ------------------------------------------------------------------------ dense nature nets (MA-RMABPPO)
def nat_ws(M, H, n_loc, Dl, dev):
    need = int(lib().rmab_nat_ws_bytes(int(M), int(H), int(n_loc), int(Dl)))
    return torch.empty((need + 3) // 4, dtype=f32, device=dev), need


def _pv(t):
    """Raw pointer of a (possibly offset) contiguous CUDA view."""
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError("robustrmab_b200: contiguous CUDA view expected (no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def nat_l1_fwd(M, E, H, x0, sc0, W0, z1, ws, x1=None, W1=None):
    """x0 / x1: u8 `[n_loc, rows, E]` (or `[n_loc, E]`) whose first M / E rows are the samples; W0 / W1 `[n_loc, H]`."""
    n_loc = x0.shape[0]
    st0 = x0.stride(0)
    st1 = x1.stride(0) if x1 is not None else 0
    check(lib().rmab_nat_l1_fwd(M, E, H, n_loc, _pv(x0) if n_loc else None, st0, float(sc0), _pv(W0) if n_loc else None,
                                _pv(x1) if (x1 is not None and n_loc) else None, st1, _pv(W1) if (W1 is not None and n_loc) else None,
                                _p(z1, f32), _p(ws[0], f32), ws[1], _stream()))
    return z1


def nat_mid_fwd(M, H, z1, b1, W2, b2, a1, a2):
    check(lib().rmab_nat_mid_fwd(M, H, _p(z1, f32), _pv(b1), _pv(W2), _pv(b2), _p(a1, f32), _p(a2, f32), _stream()))


def nat_mu(M, H, Dl, a2, W3, b3, mu):
    check(lib().rmab_nat_mu(M, H, Dl, _p(a2, f32), _pv(W3), _pv(b3), _pv(mu), mu.stride(0), _stream()))
    return mu


def nat_logp(M, H, Dl, a2, W3, b3, log_std, act, logp, ws):
    """act: `[M, >= Dl]` view whose column 0 is this rank's first output row (row stride = act.stride(0))."""
    check(lib().rmab_nat_logp(M, H, Dl, _p(a2, f32), _pv(W3) if Dl else None, _pv(b3) if Dl else None,
                              _pv(log_std) if Dl else None, C.c_void_p(act.data_ptr()) if Dl else None, act.stride(0),
                              _p(logp, f32), _p(ws[0], f32), ws[1], _stream()))
    return logp


def nat_pi_coef(M, logp, logp_old, adv, clip_ratio, g):
    check(lib().rmab_nat_pi_coef(M, _p(logp, f32), _p(logp_old, f32), _p(adv, f32), float(clip_ratio), _p(g, f32), _stream()))
    return g


def nat_out_bwd(M, H, Dl, a2, W3, b3, log_std, act, g, dW3, db3, dls, da2, ws):
    check(lib().rmab_nat_out_bwd(M, H, Dl, _p(a2, f32), _pv(W3) if Dl else None, _pv(b3) if Dl else None,
                                 _pv(log_std) if Dl else None, C.c_void_p(act.data_ptr()) if Dl else None, act.stride(0),
                                 _p(g, f32), _pv(dW3) if Dl else None, _pv(db3) if Dl else None, _pv(dls) if Dl else None,
                                 _p(da2, f32), _p(ws[0], f32), ws[1], _stream()))


def nat_mid_bwd(M, H, da2, a1, a2, W2, d1, d2, dW2, db2, db1, ws):
    check(lib().rmab_nat_mid_bwd(M, H, _p(da2, f32), _p(a1, f32), _p(a2, f32), _pv(W2), _p(d1, f32), _p(d2, f32), _pv(dW2),
                                 _pv(db2), _pv(db1), _p(ws[0], f32), ws[1], _stream()))


def nat_l1_bwd(M, E, H, x0, sc0, d1, dW0, ws, x1=None, dW1=None):
    n_loc = x0.shape[0]
    if n_loc == 0:
        return
    check(lib().rmab_nat_l1_bwd(M, E, H, n_loc, _pv(x0), x0.stride(0), float(sc0), _pv(x1) if x1 is not None else None,
                                x1.stride(0) if x1 is not None else 0, _p(d1, f32), _pv(dW0), _pv(dW1) if dW1 is not None else None,
                                _p(ws[0], f32), ws[1], _stream()))


def nat_v_out(M, H, a2, W3, b3, v=None, ret=None, da2=None, dW3=None, db3=None, ws=None):
    check(lib().rmab_nat_v_out(M, H, _p(a2, f32), _pv(W3), _pv(b3), _p(ret, f32) if ret is not None else None,
                               _p(v, f32) if v is not None else None, _p(da2, f32) if da2 is not None else None, _pv(dW3), _pv(db3),
                               _p(ws[0], f32) if ws is not None else None, ws[1] if ws is not None else 0, _stream()))
    return v
