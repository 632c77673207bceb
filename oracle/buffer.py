"""PPO buffer restatement: GAE / returns / cost-to-go and advantage normalisation (oracle, test-only).

Follows robust_rmab/algos/rmabppo/agent_oracle.py:
  RMABPPO_Buffer.finish_path  :90-139  (lambda-adjusted reward :120, q :125, delta :126,
                                         adv = discount_cumsum(delta, gamma*lam) :127,
                                         ret = discount_cumsum([r~, v_T], gamma)[:-1] :130,
                                         cdcost = discount_cumsum(sum_i cost, gamma)[:-1] :139)
  RMABPPO_Buffer.get          :144-161 (per-arm (adv-mean)/std)
  core.discount_cumsum        rmabppo_core.py:36-51 (scipy lfilter, float64, stored to fp32)
  mpi_statistics_scalar       utils/mpi_tools.py:76-97 (all float32)
The MA nature stream (nature_oracle.py:154-166) is the same recurrence on a `[T,E]` stream.
Layout `[N,T,E]` (arm-major); the reference's `[T,N]` is the E=1 transpose.
"""
import numpy as np

F32 = np.float32


def discount_cumsum(x, discount):
    """Reverse discounted cumulative sum along axis 0, float64 (rmabppo_core.py:36-51)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty_like(x)
    acc = np.zeros(x.shape[1:], dtype=np.float64)
    for t in range(x.shape[0] - 1, -1, -1):
        acc = x[t] + discount * acc
        out[t] = acc
    return out


def finish_path(rew, cost, val, last_val, lamb, gamma, lam_other):
    """rew, cost, val `[N,T,E]` fp32; last_val `[N,E]`; lamb `[T,E]` (or `[E]`).

    Returns adv, ret, q `[N,T,E]` fp32 (un-normalised) and cdcost `[T,E]` fp32.
    """
    rew = np.asarray(rew, F32)
    N, T, E = rew.shape
    lamb = np.asarray(lamb, F32)
    if lamb.ndim == 1:
        lamb = np.broadcast_to(lamb[None, :], (T, E))
    rt = np.moveaxis(rew, 1, 0).astype(np.float64)           # [T,N,E]
    ct = np.moveaxis(np.asarray(cost, F32), 1, 0).astype(np.float64)
    vt = np.moveaxis(np.asarray(val, F32), 1, 0).astype(np.float64)
    lv = np.asarray(last_val, np.float64)[None]                 # [1,N,E]; float64 as passed (:111)
    rews = np.concatenate([rt, lv], axis=0)                   # :111
    costs = np.concatenate([ct, np.zeros_like(lv)], axis=0)   # :113
    lambds = np.concatenate([lamb.astype(np.float64), np.zeros((1, E))], axis=0)[:, None, :]  # :115
    rews = rews - lambds * costs                              # :120
    vals = np.concatenate([vt, lv], axis=0)                   # :122
    qs = rews[:-1] + gamma * vals[1:]                         # :125
    deltas = qs - vals[:-1]                                   # :126
    adv = discount_cumsum(deltas, gamma * lam_other)          # :127
    ret = discount_cumsum(rews, gamma)[:-1]                   # :130
    cdcost = discount_cumsum(costs.sum(axis=1), gamma)[:-1]   # :117,:139
    back = lambda z: np.ascontiguousarray(np.moveaxis(z, 0, 1)).astype(F32)
    return back(adv), back(ret), back(qs), cdcost.astype(F32)


def statistics_scalar(x):
    """mpi_statistics_scalar (mpi_tools.py:76-97) at one process: float32 mean and population std."""
    x = np.array(x, dtype=F32).reshape(-1)
    s = F32(np.sum(x))
    n = F32(len(x))
    mean = F32(s / n)
    ssq = F32(np.sum((x - mean) ** 2))
    std = F32(np.sqrt(ssq / n))
    return mean, std


def normalize_adv(adv):
    """get() :152-155 -- per-arm normalisation over all T*E samples.  adv `[N,T,E]` fp32."""
    adv = np.asarray(adv, F32)
    out = np.empty_like(adv)
    for i in range(adv.shape[0]):
        mean, std = statistics_scalar(adv[i])
        with np.errstate(divide="ignore", invalid="ignore"):
            out[i] = (adv[i] - mean) / std
    return out


def finish_path_stream(rew, val, last_val, gamma, lam_other):
    """Nature stream of MA_RMABPPO_Buffer.finish_path (nature_oracle.py:154-166): `[T,E]` arrays."""
    r = np.concatenate([np.asarray(rew, np.float64), np.asarray(last_val, np.float64)[None]], axis=0)
    v = np.concatenate([np.asarray(val, np.float64), np.asarray(last_val, np.float64)[None]], axis=0)
    qs = r[:-1] + gamma * v[1:]
    deltas = qs - v[:-1]
    return (discount_cumsum(deltas, gamma * lam_other).astype(F32),
            discount_cumsum(r, gamma)[:-1].astype(F32), qs.astype(F32))
