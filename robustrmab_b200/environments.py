"""Drop-in environments: the reference's robust RMAB envs (robust_rmab/environments/bandit_env_robust.py)
with the same constructor, attributes and gym-style calls, stepping on the GPU.

    CounterExampleRobustEnv(N, B, seed)            :740-958
    ARMMANRobustEnv(N, B, seed)                    :352-735
    SISRobustEnv(N, B, pop_size, seed)             :971-1307

Kept from the reference: attribute names (N, S, A, B, T, R, C, observation_space, action_space,
observation_dimension, action_dim_nature, PARAMETER_RANGES, sampled_parameter_ranges, current_full_state,
random_stream, init_seed), `reset() -> [N,1] int`, `step(a_agent, a_nature) -> (s'[N,1], r[N], False, None)`,
`seed`, `sample_parameter_ranges` (same numpy stream, so the same ranges for the same seed),
`bound_nature_actions(a_flat, state, reshape)`, `get_T_for_a_nature`, `random_agent_action`, and the
error behaviour (`assert N % 3 == 0`, ValueError for out-of-range SIS / get_T parameters, warn-and-clamp
in the table-domain step).

New: `n_envs=E` replicas stepped together (states `[N,E]`; with E = 1 shapes are the reference's), device
entry points `reset_device` / `step_device` that never leave HBM, and a Philox counter RNG for the
transitions (key = seed, counter = (stream, t, env, global arm); oracle/philox.py states the rule).  The
numpy `random_stream` still drives what the reference draws on the host (range sampling).
"""
import warnings

import numpy as np
import torch

from . import _lib, ops

_F32 = torch.float32


def _as_u8(x, dev, shape):
    if isinstance(x, torch.Tensor):
        t = x.to(device=dev, dtype=torch.uint8)
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(x).astype(np.uint8)), device=dev)
    return t.reshape(shape).contiguous()


class _RobustEnvBase:
    """Shared plumbing: device state `[N,E]`, Philox step counters, numpy <-> device conversion."""
    domain = None

    def _init_common(self, N, B, seed, S, A, n_envs, device, arm0):
        self.N, self.S, self.A, self.B = int(N), int(S), int(A), B
        self.observation_space = np.arange(S)
        self.action_space = np.arange(A)
        self.observation_dimension = 1
        self.action_dimension = 1
        self.init_seed = seed
        self.n_envs = int(n_envs)
        self.arm0 = int(arm0)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("robustrmab_b200 environments step on CUDA only (no CPU fallback)")
        self.random_stream = np.random.RandomState()
        self.sampled_parameter_ranges = None
        self._ranges_dev = None
        self._ranges_src = None
        self._state = torch.zeros((self.N, self.n_envs), dtype=torch.uint8, device=self.device)
        self._t = 0          # transitions taken (Philox time counter of stream ENV)
        self._resets = 0     # resets taken (Philox time counter of stream RESET)
        self.seed(seed=seed)

    def shard(self, arm0, arm1):
        """The env of the arm window [arm0, arm1) of this one: same per-arm tables, ranges and Philox counters (global
        arm index), so its transitions are exactly rows arm0..arm1 of this env's (shards by arm)."""
        import copy
        e = copy.copy(self)
        n, per = arm1 - arm0, self.action_dim_nature // self.N
        e.N, e.arm0, e.action_dim_nature = n, self.arm0 + arm0, n * per
        for name in ("PARAMETER_RANGES", "R", "sampled_parameter_ranges", "param_setting"):
            v = getattr(self, name, None)
            if v is not None:
                setattr(e, name, np.ascontiguousarray(np.asarray(v)[arm0:arm1]))
        e.T = FakeT((n,) + tuple(self.T.shape[1:])) if isinstance(self.T, FakeT) else np.ascontiguousarray(self.T[arm0:arm1])
        e._R_dev = self._R_dev[arm0:arm1].contiguous()
        e._state = self._state[arm0:arm1].contiguous()
        e._ranges_dev = e._ranges_src = None
        e.random_stream = np.random.RandomState()
        return e

    # -- state ---------------------------------------------------------------------------------
    @property
    def current_full_state(self):
        s = self._state.cpu().numpy().astype(np.int64)
        return s[:, 0] if self.n_envs == 1 else s

    @current_full_state.setter
    def current_full_state(self, value):
        self._state = _as_u8(value, self.device, (self.N, self.n_envs))

    def _obs(self, s):
        s = s.cpu().numpy().astype(np.int64)
        return s.reshape(self.N, self.observation_dimension) if self.n_envs == 1 else s

    def seed(self, seed=None):
        seed1 = seed
        if seed1 is None:
            seed1 = np.random.randint(1e9)
        self.random_stream.seed(seed=seed1)
        self._key = int(seed1)
        self._t = 0
        self._resets = 0
        return [seed1]

    def counter_base(self, dev_word=None):
        """Route the Philox time counter through a device word: with `dev_word` (int32 CUDA tensor of one element, holding the
        current `_t`) the kernels get `t - _t` as argument and add the word when they run, so a captured CUDA graph of a
        rollout can be replayed in later epochs after the caller refreshes the word.  None switches it off."""
        self._t_word, self._t_origin = dev_word, (self._t if dev_word is not None else 0)

    def _step_rng(self, stream=_lib.STREAM_ENV):
        word = getattr(self, "_t_word", None)
        if word is None:
            return ops.rng(self._key, stream, self._t, arm0=self.arm0)
        return ops.rng(self._key, stream, self._t - self._t_origin, arm0=self.arm0, t_base=word)

    def reset_device(self):
        """Uniform random start states (reference reset: random_stream.choice / randint)."""
        r = ops.rng(self._key, _lib.STREAM_RESET, self._resets, arm0=self.arm0)
        self._state = ops.reset_states(self.N, self.n_envs, self.S, r, device=self.device)
        self._resets += 1
        return self._state

    def reset(self):
        return self._obs(self.reset_device())

    def reset_random(self):
        return self.reset()

    def render(self):
        return None

    def close(self):
        pass

    # -- ranges --------------------------------------------------------------------------------
    def _ranges(self):
        r = self.sampled_parameter_ranges
        if r is None:
            raise ValueError("sampled_parameter_ranges must be set before stepping (the reference sets it "
                             "right after construction, agent_oracle.py:788)")
        if self._ranges_src is not r:
            self._ranges_dev = torch.as_tensor(np.ascontiguousarray(r), dtype=_F32).to(self.device)
            self._ranges_src = r
        return self._ranges_dev

    def _sample_sub_ranges(self):
        draw = self.random_stream.rand(*self.PARAMETER_RANGES.shape)
        lo = self.PARAMETER_RANGES.min(axis=-1)[..., None]
        width = (self.PARAMETER_RANGES.max(axis=-1) - self.PARAMETER_RANGES.min(axis=-1))[..., None]
        draw.sort(axis=-1)
        out = draw * width + lo
        assert (out[..., 0] >= self.PARAMETER_RANGES[..., 0]).all() and (out[..., 1] <= self.PARAMETER_RANGES[..., 1]).all()
        return out

    def _nature_dev(self, a_nature, per_arm_shape):
        """Nature action as a device fp32 tensor + whether it carries an env axis."""
        if isinstance(a_nature, torch.Tensor):
            t = a_nature.to(device=self.device, dtype=_F32)
        else:
            t = torch.as_tensor(np.ascontiguousarray(np.asarray(a_nature, dtype=np.float32)), device=self.device)
        n_per_arm = int(np.prod(per_arm_shape))
        if t.numel() == self.N * n_per_arm:
            return t.reshape((self.N,) + tuple(per_arm_shape)).contiguous(), False
        return t.reshape((self.N,) + tuple(per_arm_shape) + (self.n_envs,)).contiguous(), True

    def _finish_step(self, s_next, r):
        self._state = s_next
        self._t += 1
        if self.n_envs == 1:
            return self._obs(s_next), r.cpu().numpy().astype(np.float64)[:, 0], False, None
        return self._obs(s_next), r.cpu().numpy().astype(np.float64), False, None

    def _warn_clamped(self, nature, lo, hi):
        if bool(((nature < lo) | (nature > hi)).any()):
            warnings.warn("Warning! nature action outside allowed param range; setting to the bound of the range",
                          RuntimeWarning)

    def _binary_random_action(self):
        actions = np.zeros(self.N)
        choices = np.random.choice(np.arange(self.N), int(self.B), replace=False)
        actions[choices] = 1
        return actions


class CounterExampleRobustEnv(_RobustEnvBase):
    domain = _lib.DOMAIN_CE

    def __init__(self, N, B, seed, n_envs=1, device="cuda", arm0=0):
        assert N % 3 == 0
        self.action_dim_nature = N
        self.PARAMETER_RANGES = self.get_parameter_ranges(N)
        self._init_common(N, B, seed, 2, 2, n_envs, device, arm0)
        self.T, self.R, self.C = self.get_experiment(N)
        self._R_dev = torch.as_tensor(self.R, dtype=_F32).to(self.device).contiguous()

    def get_parameter_ranges(self, N):
        base = np.array([[0, 1], [0.05, 0.9], [0.1, 0.95]], dtype=np.float64)
        return base[np.arange(N) % 3]

    def sample_parameter_ranges(self):
        return np.copy(self.PARAMETER_RANGES)

    def get_experiment(self, N):
        t = np.array([[[0.5, 0.5], [0.5, 0.5]], [[1.0, 0.0], [0.0, -1.0]]])
        return np.tile(t, (N, 1, 1, 1)), np.array([[0, 1] for _ in range(N)]), np.array([0, 1])

    def random_agent_action(self):
        return self._binary_random_action()

    def step_device(self, a_agent, a_nature, want_reward=True):
        """a_agent `[N,E]` u8, a_nature `[N]` (shared by the envs) or `[N,E]`; returns device (s', r)."""
        ranges = self._ranges()
        nature, per_env = self._nature_dev(a_nature, ())
        r = self._step_rng()
        return ops.env_step_table(self.domain, self._state, a_agent, nature, ranges, self._R_dev, r=r,
                                  nature_per_env=per_env, want_reward=want_reward)

    def step(self, a_agent, a_nature):
        a = _as_u8(a_agent, self.device, (self.N, self.n_envs))
        nature, _ = self._nature_dev(a_nature, ())
        rg = self._ranges()
        lo, hi = (rg[:, 0], rg[:, 1]) if nature.dim() == 1 else (rg[:, 0:1], rg[:, 1:2])
        self._warn_clamped(nature, lo, hi)
        s_next, r = self.step_device(a, nature)
        if nature.dim() == 1:   # the reference leaves the clamped parameter in self.T (:876-877)
            p = torch.minimum(torch.maximum(nature, rg[:, 0]), rg[:, 1]).cpu().numpy().astype(np.float64)
            self.T[:, 1, 1, 1] = p
            self.T[:, 1, 1, 0] = 1 - p
        return self._finish_step(s_next, r)

    def get_T_for_a_nature(self, a_nature_expanded):
        p = np.asarray(a_nature_expanded, dtype=np.float64)
        rg = self.sampled_parameter_ranges
        bad = (p < rg[:, 0]) | (p > rg[:, 1])
        if bad.any():
            i = int(np.flatnonzero(bad)[0])
            raise ValueError("Nature setting outside allowed param range. Was %s but should be in %s" % (p[i], rg[i]))
        self.T[:, 1, 1, 1] = p
        self.T[:, 1, 1, 0] = 1 - p
        return np.copy(self.T)

    def bound_nature_actions(self, a_nature_flat, state=None, reshape=True):
        flat = np.asarray(a_nature_flat)
        x = torch.as_tensor(flat.reshape(self.N).astype(np.float32), device=self.device)
        rg = self._ranges()
        out = ops.bound_nature(x, rg[:, 0].contiguous(), rg[:, 1].contiguous()).cpu().numpy().astype(np.float64)
        return out if reshape else out.reshape(*flat.shape)


class ARMMANRobustEnv(_RobustEnvBase):
    domain = _lib.DOMAIN_ARMMAN

    def __init__(self, N, B, seed, n_envs=1, device="cuda", arm0=0):
        self.action_dim_nature = N * 2
        self.percent_A = 0.2
        self.percent_B = 0.2
        self.PARAMETER_RANGES = self.get_parameter_ranges(N)
        self._init_common(N, B, seed, 3, 2, n_envs, device, arm0)
        self.T, self.R, self.C = self.get_experiment(N)
        self._R_dev = torch.as_tensor(self.R, dtype=_F32).to(self.device).contiguous()

    def get_parameter_ranges(self, N):
        rA = [[[0.0, 1.0], [0.0, 1.0]], [[0.5, 1.0], [0.5, 1.0]], [[0.35, 0.85], [0.35, 0.85]]]
        rB = [[[0.0, 1.0], [0.0, 1.0]], [[0.35, 0.85], [0.15, 0.65]], [[0.35, 0.85], [0.35, 0.85]]]
        rC = [[[0.0, 1.0], [0.0, 1.0]], [[0.35, 0.85], [0.0, 0.5]], [[0.35, 0.85], [0.35, 0.85]]]
        nA, nB = int(N * self.percent_A), int(N * self.percent_B)
        out = np.empty((N, 3, 2, 2), dtype=np.float64)
        out[:nA], out[nA:nA + nB], out[nA + nB:] = rA, rB, rC
        return out

    def sample_parameter_ranges(self):
        return self._sample_sub_ranges()

    def get_experiment(self, N):
        tA = np.array([[[0.1, 0.9, 0.0], [0.1, 0.9, 0.0]], [[0, -1, -1], [-1, -1, 0]], [[0, 0.4, 0.6], [0.0, 0.4, 0.6]]])
        tB = np.array([[[0.9, 0.1, 0.0], [0.9, 0.1, 0.0]], [[0, -1, -1], [-1, -1, 0]], [[0, 0.4, 0.6], [0.0, 0.4, 0.6]]])
        nA, nB = int(N * 0.2), int(N * 0.2)
        T = np.array([tA] * nA + [tB] * nB + [tA] * (N - nA - nB), dtype=np.float64).reshape(N, 3, 2, 3)
        return T, np.array([[1, 0.5, 0] for _ in range(N)]), np.array([0, 1])

    def random_agent_action(self):
        return self._binary_random_action()

    def step_device(self, a_agent, a_nature, want_reward=True):
        """a_nature: `[N,A]` / `[N,A,E]` (what env.step receives, :575-580) or a `[N,S,A]` per-arm table that is
        looked up at the current state (what the baseline nature policies hold)."""
        ranges = self._ranges()
        if isinstance(a_nature, torch.Tensor) and a_nature.dim() == 3 and tuple(a_nature.shape) == (self.N, 3, 2):
            nature, per_env = a_nature.to(self.device, _F32).contiguous(), False
        else:
            nature, has_env = self._nature_dev(a_nature, (2,))
            if not has_env:
                nature = nature[:, :, None].expand(self.N, 2, self.n_envs).contiguous()
            per_env = True
        r = self._step_rng()
        return ops.env_step_table(self.domain, self._state, a_agent, nature, ranges, self._R_dev, r=r,
                                  nature_per_env=per_env, want_reward=want_reward)

    def step(self, a_agent, a_nature):
        a = _as_u8(a_agent, self.device, (self.N, self.n_envs))
        nature, has_env = self._nature_dev(a_nature, (2,))
        rg = self._ranges()                                              # [N,S,A,2]
        idx = self._state.long()                                          # [N,E]
        rs = rg[torch.arange(self.N, device=self.device)[:, None], idx]   # [N,E,A,2]
        nat_e = nature if has_env else nature[:, :, None].expand(self.N, 2, self.n_envs)
        chosen = torch.gather(nat_e, 1, a.long()[:, None, :])[:, 0]       # [N,E] the parameter actually used
        sel = torch.gather(rs, 2, a.long()[:, :, None, None].expand(-1, -1, 1, 2))[:, :, 0]
        self._warn_clamped(chosen, sel[..., 0], sel[..., 1])
        s_prev = self._state
        s_next, r = self.step_device(a, nature)
        if self.n_envs == 1:  # the reference leaves the patched row in self.T (:595-616)
            p = torch.minimum(torch.maximum(chosen, sel[..., 0]), sel[..., 1])[:, 0].cpu().numpy().astype(np.float64)
            sp = s_prev[:, 0].cpu().numpy()
            an = a[:, 0].cpu().numpy()
            for i in range(self.N):
                self._patch_T(i, int(sp[i]), int(an[i]), p[i])
        return self._finish_step(s_next, r)

    def _patch_T(self, i, s, a, p):
        if s == 0:
            self.T[i, s, a, 0], self.T[i, s, a, 1] = p, 1 - p
        elif s == 1 and a == 0:
            self.T[i, s, a, 2], self.T[i, s, a, 1] = p, 1 - p
        elif s == 1:
            self.T[i, s, a, 0], self.T[i, s, a, 1] = p, 1 - p
        else:
            self.T[i, s, a, 2], self.T[i, s, a, 1] = p, 1 - p

    def get_T_for_a_nature(self, a_nature_expanded):
        p = np.asarray(a_nature_expanded, dtype=np.float64)
        rg = self.sampled_parameter_ranges
        bad = (p < rg[..., 0]) | (p > rg[..., 1])
        if bad.any():
            i = tuple(int(v[0]) for v in np.nonzero(bad))
            raise ValueError("Nature setting outside allowed param range. Was %s but should be in %s" % (p[i], rg[i]))
        for i in range(self.N):
            for s in range(3):
                for a in range(2):
                    self._patch_T(i, s, a, p[i, s, a])
        return np.copy(self.T)

    def bound_nature_actions(self, a_nature_flat, state=None, reshape=True):
        flat = np.asarray(a_nature_flat)
        x = torch.as_tensor(flat.reshape(self.N, 2).astype(np.float32), device=self.device)
        st = torch.as_tensor(np.asarray(state).reshape(self.N).astype(np.int64), device=self.device)
        rg = self._ranges()[torch.arange(self.N, device=self.device), st]      # [N,A,2]
        out = ops.bound_nature(x, rg[..., 0].contiguous(), rg[..., 1].contiguous()).cpu().numpy().astype(np.float64)
        return out if reshape else out.reshape(*flat.shape)


class FakeT:
    """SISRobustEnv.T is only a shape carrier in the reference (:965-968)."""

    def __init__(self, shape):
        self.shape = shape


class SISRobustEnv(_RobustEnvBase):
    domain = _lib.DOMAIN_SIS

    def __init__(self, N, B, pop_size, seed, n_envs=1, device="cuda", arm0=0):
        if pop_size + 1 > 256:
            raise ValueError("pop_size must be <= 255 (states are stored as uint8)")
        self.pop_size = int(pop_size)
        self.action_dim_nature = N * 4
        self.PARAMETER_RANGES = self.get_parameter_ranges(N)
        self.param_setting = None
        self._init_common(N, B, seed, pop_size + 1, 3, n_envs, device, arm0)
        self.T, self.R, self.C = self.get_experiment(N)
        self._R_dev = torch.as_tensor(self.R, dtype=_F32).to(self.device).contiguous()

    def get_parameter_ranges(self, N):
        return np.tile(np.array([[0.5, 0.99], [1, 10], [1, 10], [1, 10]], dtype=np.float64), (N, 1, 1))

    def sample_parameter_ranges(self):
        return self._sample_sub_ranges()

    def get_experiment(self, N):
        return FakeT((N, self.S, self.A, self.S)), np.array([np.linspace(0, 1, self.S) for _ in range(N)]), \
            np.array([0, 1, 2])

    def random_agent_action(self):
        actions = np.zeros(self.N, dtype=int)
        cost = 0
        for arm in np.random.choice(np.arange(self.N), self.N, replace=False):
            n_valid = len(self.C[self.C <= self.B - cost])
            a = np.random.choice(np.arange(n_valid), 1, p=None)[0]
            cost += self.C[a]
            if cost > self.B:
                break
            actions[arm] = a
        return actions

    def set_params(self, a_nature):
        p = np.asarray(a_nature, dtype=np.float64)
        rg = self.sampled_parameter_ranges
        if ((p < rg[..., 0]) | (p > rg[..., 1])).any():
            raise ValueError("bad setting")
        self.param_setting = p.copy()

    def step_device(self, a_agent, a_nature, want_reward=True):
        nature, per_env = self._nature_dev(a_nature, (4,))
        r = self._step_rng()
        return ops.env_step_sis(self.pop_size, self._state, a_agent, nature, self._R_dev, r=r, nature_per_env=per_env,
                                want_reward=want_reward)

    def step(self, a_agent, a_nature):
        if not isinstance(a_nature, torch.Tensor) and np.asarray(a_nature).size == self.N * 4:
            self.set_params(np.asarray(a_nature).reshape(self.N, 4))
        a = _as_u8(a_agent, self.device, (self.N, self.n_envs))
        s_next, r = self.step_device(a, a_nature)
        return self._finish_step(s_next, r)

    def get_T_for_a_nature(self, a_nature):
        self.set_params(np.asarray(a_nature).reshape(self.N, 4))
        prm = torch.as_tensor(self.param_setting.astype(np.float32), device=self.device)
        return ops.sis_table(self.pop_size, prm).cpu().numpy().astype(np.float64)

    def bound_nature_actions(self, a_nature_flat, state=None, reshape=True):
        flat = np.asarray(a_nature_flat)
        x = torch.as_tensor(flat.reshape(self.N, 4).astype(np.float32), device=self.device)
        rg = self._ranges()
        out = ops.bound_nature(x, rg[..., 0].contiguous(), rg[..., 1].contiguous()).cpu().numpy().astype(np.float64)
        return out if reshape else out.reshape(*flat.shape)
