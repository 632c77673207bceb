"""Environment-step restatement (oracle, test-only): counterexample / ARMMAN / SIS robust envs.

Follows robust_rmab/environments/bandit_env_robust.py:
  CounterExampleRobustEnv.step  :860-893   (clip nature param to range :867-874, T[1,1]=[1-p,p] :876-877)
  ARMMANRobustEnv.step          :575-634   (clip :583-592, state-specific patch :595-616)
  SISRobustEnv.step/compute_distro/p_i_s   :1224-1243 / :1025-1084 / :1011-1022
  bound_nature_actions          :917-935, :684-710, :1260-1279
  reset                         :941-944, :715-718, :1285-1290
Reward is R[arm, s'] (reward on the NEXT state, :888/:629/:1238).

Vectorised over `[N, E]` (arm-major, env fastest).  Sampling uses the injected-uniform rule of
oracle/philox.py::categorical on fp32-rounded probabilities, categories in the reference's own
order (index 0..S-1 of T[i, s, a, :]).
"""
import numpy as np
from scipy.special import comb

from . import philox

DOMAIN_CE, DOMAIN_ARMMAN, DOMAIN_SIS = 0, 1, 2

F32 = np.float32


def ce_param_ranges(N):
    """bandit_env_robust.py:790-818 -- ranges cycle [0,1],[0.05,0.9],[0.1,0.95] by arm%3."""
    base = np.array([[0.0, 1.0], [0.05, 0.9], [0.1, 0.95]], dtype=np.float64)
    return base[np.arange(N) % 3]


def armman_param_ranges(N, percent_A=0.2, percent_B=0.2):
    """bandit_env_robust.py:395-465 -- `[N,3,2,2]` ranges for the A/B/C arm groups."""
    rA = [[[0.0, 1.0], [0.0, 1.0]], [[0.5, 1.0], [0.5, 1.0]], [[0.35, 0.85], [0.35, 0.85]]]
    rB = [[[0.0, 1.0], [0.0, 1.0]], [[0.35, 0.85], [0.15, 0.65]], [[0.35, 0.85], [0.35, 0.85]]]
    rC = [[[0.0, 1.0], [0.0, 1.0]], [[0.35, 0.85], [0.0, 0.5]], [[0.35, 0.85], [0.35, 0.85]]]
    nA, nB = int(N * percent_A), int(N * percent_B)
    out = np.empty((N, 3, 2, 2), dtype=np.float64)
    out[:nA] = rA
    out[nA:nA + nB] = rB
    out[nA + nB:] = rC
    return out


def sis_param_ranges(N):
    """bandit_env_robust.py:1088-1109 -- `[N,4,2]`: r_t in [0.5,0.99], lam/e1/e2 in [1,10]."""
    return np.tile(np.array([[0.5, 0.99], [1, 10], [1, 10], [1, 10]], dtype=np.float64), (N, 1, 1))


def sample_sub_ranges(ranges, rs):
    """sample_parameter_ranges (:468-481, :1112-1126): sorted U(0,1) pair scaled into the range."""
    draw = rs.rand(*ranges.shape)
    draw.sort(axis=-1)
    lo = ranges.min(axis=-1)[..., None]
    width = (ranges.max(axis=-1) - ranges.min(axis=-1))[..., None]
    return draw * width + lo


def rewards_table(domain, N, S):
    """R[N,S]: CE [0,1] (:843), ARMMAN [1,0.5,0] (:557), SIS linspace(0,1,S) (:1165)."""
    if domain == DOMAIN_CE:
        return np.tile(np.array([0.0, 1.0], dtype=F32), (N, 1))
    if domain == DOMAIN_ARMMAN:
        return np.tile(np.array([1.0, 0.5, 0.0], dtype=F32), (N, 1))
    return np.tile(np.linspace(0, 1, S).astype(F32), (N, 1))


def _clip(p, lo, hi):
    """Warn-and-clamp of :867-874 / :583-592, in fp32."""
    p = np.asarray(p, dtype=F32)
    return np.minimum(np.maximum(p, np.asarray(lo, F32)), np.asarray(hi, F32))


def ce_probs(s, a, p):
    """Transition row T[i,s,a,:] for the counterexample domain (:831-836, :876-877). fp32 `[...,2]`."""
    s = np.asarray(s)
    a = np.asarray(a)
    p = np.asarray(p, dtype=F32)
    out = np.empty(s.shape + (2,), dtype=F32)
    half = F32(0.5)
    one = F32(1.0)
    # s == 0: [0.5, 0.5]; s == 1, a == 0: [1, 0]; s == 1, a == 1: [1-p, p]
    out[..., 0] = np.where(s == 0, half, np.where(a == 0, one, one - p))
    out[..., 1] = np.where(s == 0, half, np.where(a == 0, F32(0.0), p))
    return out


def ce_step(s, a, nature, ranges, R, u):
    """One counterexample step.  s,a `[N,E]` ints; nature `[N]` or `[N,E]`; ranges `[N,2]`; u `[N,E]`.

    Returns (s_next `[N,E]` uint8, r `[N,E]` fp32).
    """
    s = np.asarray(s).astype(np.int64)
    a = np.asarray(a).astype(np.int64)
    nature = np.asarray(nature, dtype=F32)
    if nature.ndim == 1:
        nature = nature[:, None]
    p = _clip(nature, ranges[:, 0:1], ranges[:, 1:2])
    probs = ce_probs(s, a, np.broadcast_to(p, s.shape))
    s_next = philox.categorical(probs, u)
    r = np.take_along_axis(np.asarray(R, F32), s_next, axis=1)
    return s_next.astype(np.uint8), r


def armman_probs(s, a, p):
    """Transition row for ARMMAN after the state-specific patch (:595-616).  fp32 `[...,3]`.

    s=0: [p,1-p,0];  s=1,a=0: [0,1-p,p];  s=1,a=1: [p,1-p,0];  s=2: [0,1-p,p].
    """
    s = np.asarray(s)
    a = np.asarray(a)
    p = np.asarray(p, dtype=F32)
    q = (F32(1.0) - p).astype(F32)
    to0 = (s == 0) | ((s == 1) & (a == 1))
    out = np.zeros(s.shape + (3,), dtype=F32)
    out[..., 0] = np.where(to0, p, F32(0.0))
    out[..., 1] = q
    out[..., 2] = np.where(to0, F32(0.0), p)
    return out


def armman_step(s, a, nature, ranges, R, u, nature_is_table=True):
    """One ARMMAN step.  nature: `[N,S,A]` per-arm table (looked up at the current state, as
    nature_baselines_armman.py:224-232 does) or `[N,A,E]` per-env actions (what env.step receives,
    :575-580).  ranges `[N,S,A,2]`.
    """
    s = np.asarray(s).astype(np.int64)
    a = np.asarray(a).astype(np.int64)
    N, E = s.shape
    ar = np.arange(N)[:, None]
    nature = np.asarray(nature, dtype=F32)
    if nature_is_table:
        raw = nature[ar, s, a]
    else:
        raw = np.take_along_axis(nature, a[:, None, :], axis=1)[:, 0, :]
    lo = np.asarray(ranges, F32)[ar, s, a, 0]
    hi = np.asarray(ranges, F32)[ar, s, a, 1]
    p = _clip(raw, lo, hi)
    probs = armman_probs(s, a, p)
    s_next = philox.categorical(probs, u)
    r = np.take_along_axis(np.asarray(R, F32), s_next, axis=1)
    return s_next.astype(np.uint8), r


def sis_distro(pop, s, a, params):
    """compute_distro (:1025-1084) for scalar s,a and params (r_t, lam, e1, e2); float64 `[pop+1]`."""
    r_t, lam, e1, e2 = [float(x) for x in params]
    S = pop + 1
    distro = np.zeros(S, dtype=np.float64)
    beta_t = (pop - s) / pop
    if a == 0:
        q_t = 1 - np.e ** (-lam * beta_t * r_t)
    elif a == 1:
        q_t = 1 - np.e ** (-lam * beta_t / e1 * r_t)
    else:
        q_t = 1 - np.e ** (-lam * beta_t * r_t / e2)
    for sp in range(S):
        if pop - s <= sp <= pop:
            i = pop - sp
            if s == pop:
                prob = 1.0 if i == 0 else 0.0
            else:
                prob = comb(s, i) * q_t ** i * (1 - q_t) ** (s - i)
            distro[sp] = prob
    distro[distro < 1e-7] = 0
    return distro / distro.sum()


def sis_step(pop, s, a, params, R, u):
    """One SIS step.  params `[N,4]` (shared by envs) or `[N,4,E]`; float64 pmf, fp32 sampling."""
    s = np.asarray(s).astype(np.int64)
    a = np.asarray(a).astype(np.int64)
    N, E = s.shape
    params = np.asarray(params, dtype=np.float64)
    s_next = np.empty((N, E), dtype=np.int64)
    for n in range(N):
        for e in range(E):
            prm = params[n] if params.ndim == 2 else params[n, :, e]
            d = sis_distro(pop, int(s[n, e]), int(a[n, e]), prm)
            s_next[n, e] = philox.categorical(d.astype(F32), np.asarray(u)[n, e])
    r = np.take_along_axis(np.asarray(R, F32), s_next, axis=1)
    return s_next.astype(np.uint8), r


def bound_nature(raw, lo, hi):
    """bound_nature_actions (:703, :930, :1274): ((tanh(x)+1)/2)*(ub-lb)+lb, tanh in fp32 via torch
    and the affine part in float64 (python float arithmetic on a 0-d fp32 tensor promotes to fp32
    -- torch keeps the tensor dtype -- then the result is stored into a float64 array).
    """
    import torch
    t = torch.tanh(torch.as_tensor(np.asarray(raw), dtype=torch.float32))
    lo32 = torch.as_tensor(np.asarray(lo, dtype=np.float64))
    hi32 = torch.as_tensor(np.asarray(hi, dtype=np.float64))
    # elementwise replay of the scalar expression: ((t+1)/2) is fp32; *(ub-lb) multiplies an fp32
    # 0-d tensor by a python float (stays fp32); + lb likewise.
    width = (hi32 - lo32).to(torch.float32)
    out = ((t + 1) / 2) * width + lo32.to(torch.float32)
    return out.numpy().astype(np.float64)


def reset_states(seed, epoch, S, n_envs, n_arms, arm0=0, env0=0):
    """Uniform random start states (reset: :941-944 etc.) from Philox stream RESET, `[N,E]` uint8."""
    u = philox.uniform_grid(seed, philox.STREAM_RESET, epoch, n_envs, n_arms, arm0, env0)
    return np.minimum((u * F32(S)).astype(np.int64), S - 1).astype(np.uint8)
