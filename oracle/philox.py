"""Philox4x32-10 counter RNG and the uniform / categorical sampling rule (oracle, test-only).

The reference has no GPU-reproducible RNG (numpy RandomState.multinomial in
bandit_env_robust.py:886/627/1236 and torch's global generator in rmabppo_core.py:222), so the
B200 build fixes one: Philox4x32-10 (Salmon et al., SC'11 -- the published
algorithm, same constants as cuRAND/Random123), key = 64-bit run seed, counter =
(stream, t, env, global_arm).  One uniform per categorical draw: u = (word >> 8) * 2^-24 (see `uniform`).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)

STREAM_ENV, STREAM_ACT, STREAM_NATURE, STREAM_RESET, STREAM_CF, STREAM_INIT = 0, 1, 2, 3, 4, 7


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)])
    c0, c1, c2, c3 = c0.copy(), c1.copy(), c2.copy(), c3.copy()
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * M0
            p1 = c2.astype(np.uint64) * M1
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = p0.astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _u24(x):
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def uniform(seed, stream, t, env, arm):
    """fp32 uniform in [0,1) for (stream, t, env, arm) under 64-bit `seed` (t a scalar).

    Streams ENV and ACT share Philox blocks: the transition and action draws of time step t are words 2*(t&1) and
    2*(t&1)+1 of the block with counter (0, t>>1, env, arm), so a rollout needs one block per two steps.  Every
    other stream uses word 0 of the block with counter (stream, t, env, arm).
    """
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, seed >> 32
    t = int(t)
    if stream in (STREAM_ENV, STREAM_ACT):
        words = philox4x32_10(0, t >> 1, env, arm, k0, k1)
        return _u24(words[2 * (t & 1) + (1 if stream == STREAM_ACT else 0)])
    return _u24(philox4x32_10(stream, t, env, arm, k0, k1)[0])


def uniform_grid(seed, stream, t, n_envs, n_arms, arm0=0, env0=0):
    """`[N, E]` grid of uniforms for global arms arm0.. and envs env0.. (arm-major layout)."""
    arm = (np.arange(n_arms, dtype=np.uint32) + np.uint32(arm0))[:, None]
    env = (np.arange(n_envs, dtype=np.uint32) + np.uint32(env0))[None, :]
    return uniform(seed, stream, t, env, arm)


def normal_pair(seed, stream, t, env, idx):
    """Two fp32 standard normals (Box-Muller on x0,x1) for the Gaussian nature actor."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    x0, x1, _, _ = philox4x32_10(stream, t, env, idx, seed & 0xFFFFFFFF, seed >> 32)
    u1 = ((x0 >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -24)  # (0,1]
    u2 = (x1 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    r = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = (np.float32(2.0 * np.pi) * u2).astype(np.float32)
    return (r * np.cos(ang)).astype(np.float32), (r * np.sin(ang)).astype(np.float32)


def categorical(p, u):
    """Inverse-CDF categorical draw shared by oracle, reference shim and CUDA kernels.

    `p[..., K]` probabilities (rounded to fp32), `u[...]` fp32 uniforms.  cdf accumulated
    left-to-right in fp32; k = #{j : cdf_j <= u}, clamped to the last category with p > 0.
    """
    p = np.asarray(p, dtype=np.float32)
    u = np.asarray(u, dtype=np.float32)
    K = p.shape[-1]
    cdf = np.zeros(p.shape[:-1], dtype=np.float32)
    k = np.zeros(p.shape[:-1], dtype=np.int64)
    last = np.zeros(p.shape[:-1], dtype=np.int64)
    for j in range(K):
        cdf = (cdf + p[..., j]).astype(np.float32)
        k += (cdf <= u)
        last = np.where(p[..., j] > 0, j, last)
    return np.minimum(k, last)
