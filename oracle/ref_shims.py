"""Shims that let the UNMODIFIED reference run with injected randomness (test infrastructure only).

Used by oracle/gen_golden.py (and by tests that can see /root/reference) to pin the oracle against
the reference's own classes.  Nothing here is imported by the product package.

* `import_reference()` puts oracle/stubs (gym, mpi4py, matplotlib no-ops) and the reference root on
  sys.path -- the reference source is imported where it lies, never copied.
* `InjectedStream` replaces `env.random_stream` (a public attribute, bandit_env_robust.py:772): every
  `multinomial(1, p)` consumes ONE uniform from a Philox-keyed sequence and returns the one-hot of
  the fp32 left-to-right inverse CDF (oracle/philox.py::categorical).
* `patch_categorical_sample()` does the same for torch's `Categorical.sample`
  (rmabppo_core.py:222).
"""
import contextlib
import os
import sys

import numpy as np

from . import philox

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("RMAB_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "robust_rmab"))


def import_reference():
    for p in (REFERENCE_ROOT, os.path.join(HERE, "stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")


class UniformSource:
    """Uniforms keyed like the CUDA path: (stream, t, env=0, arm) under `seed`; one arm sweep per t."""

    def __init__(self, seed, stream, n_arms, env=0):
        self.seed, self.stream, self.n, self.env = seed, stream, n_arms, env
        self.t, self.i = 0, 0
        self.log = []

    def _row(self, t):
        return philox.uniform(self.seed, self.stream, t, self.env, np.arange(self.n, dtype=np.uint32))

    def next(self):
        u = self._row(self.t)[self.i]
        self.log.append(u)
        self.i += 1
        if self.i == self.n:
            self.i, self.t = 0, self.t + 1
        return np.float32(u)


class InjectedStream:
    """Drop-in for numpy RandomState on `env.random_stream`."""

    def __init__(self, step_source, reset_source, n_states):
        self.step_source, self.reset_source, self.S = step_source, reset_source, n_states

    def seed(self, seed=None):
        return None

    def multinomial(self, n, p):
        assert n == 1
        k = int(philox.categorical(np.asarray(p, dtype=np.float32), self.step_source.next()))
        out = np.zeros(len(p), dtype=np.int64)
        out[k] = 1
        return out

    def _reset_states(self, size):
        u = np.array([self.reset_source.next() for _ in range(size)], dtype=np.float32)
        return np.minimum((u * np.float32(self.S)).astype(np.int64), self.S - 1)

    def choice(self, space, size):
        return np.asarray(space)[self._reset_states(size)]

    def randint(self, low=0, high=None, size=None):
        return low + self._reset_states(size)


@contextlib.contextmanager
def patch_categorical_sample(source):
    from torch.distributions.categorical import Categorical
    import torch
    orig = Categorical.sample

    def sample(self, sample_shape=torch.Size()):
        p = self.probs.detach().numpy().astype(np.float32)
        assert p.ndim == 1
        return torch.as_tensor(int(philox.categorical(p, source.next())))

    Categorical.sample = sample
    try:
        yield
    finally:
        Categorical.sample = orig


@contextlib.contextmanager
def quiet():
    with open(os.devnull, "w") as f, contextlib.redirect_stdout(f):
        yield
