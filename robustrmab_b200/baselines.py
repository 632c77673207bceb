"""Baseline strategies as inputs of the device loops.

The reference's nature baselines (robust_rmab/baselines/nature_baselines_{counterexample,armman,sis}.py) "stay as they
are" (SURVEY.md section 2 row 6): they are host objects whose `get_nature_action(o)` returns a constant vector (CE
`[N]`, SIS `[N,4]`) or a per-(arm, state, action) lookup (ARMMAN `[N,A]` at the arms' states).  The kernels take that as
ONE parameter table per policy (`rmab_env_step_table` / `rmab_env_step_sis`, nature_per_env = 0), so

  * `nature_param_table(pol, env)` turns ANY deterministic, per-arm nature policy object -- the reference's classes
    included -- into that table: `pol.param_setting` when it has one, otherwise `get_nature_action` probed at the S
    constant state vectors;
  * the classes below are the same policies for all three domains (domain read off the shape of the sampled ranges:
    `[N,2]` CE, `[N,S,A,2]` ARMMAN, `[N,4,2]` SIS), so that the double-oracle drop-in (double_oracle.py:124-139) runs
    where the reference tree is absent.  Names / `__repr__` follow the reference classes (they become CSV columns,
    double_oracle.py:667).
"""
import numpy as np


def _domain_of(ranges):
    r = np.asarray(ranges)
    if r.ndim == 2:
        return "counterexample"
    if r.ndim == 4:
        return "armman"
    if r.ndim == 3 and r.shape[1] == 4:
        return "sis"
    raise ValueError(f"unrecognised sampled_parameter_ranges shape {r.shape}")


def nature_param_table(pol, env):
    """Per-arm parameter table of a deterministic nature policy: CE `[N]`, ARMMAN `[N,S,A]`, SIS `[N,4]` (float32).
    Returns None when the policy is not a per-arm table (it answers differently to the same observation, or depends on
    other arms' states) -- such policies go through the per-step path."""
    ps = getattr(pol, "param_setting", None)
    if ps is not None:
        return np.ascontiguousarray(np.asarray(ps, dtype=np.float32))
    if hasattr(pol, "get_nature_action_device") or not hasattr(pol, "get_nature_action"):
        return None
    N, S = int(env.N), int(env.S)
    dom = _domain_of(env.sampled_parameter_ranges)
    if dom != "armman":
        a = np.asarray(pol.get_nature_action(np.zeros(N, dtype=int)), dtype=np.float64)
        b = np.asarray(pol.get_nature_action(np.full(N, S - 1, dtype=int)), dtype=np.float64)
        c = np.asarray(pol.get_nature_action(np.zeros(N, dtype=int)), dtype=np.float64)
        if not (np.array_equal(a, b) and np.array_equal(a, c)):
            return None
        return np.ascontiguousarray(a.astype(np.float32))
    rows = [np.asarray(pol.get_nature_action(np.full(N, s, dtype=int)), dtype=np.float64) for s in range(S)]
    again = np.asarray(pol.get_nature_action(np.zeros(N, dtype=int)), dtype=np.float64)
    mixed = np.asarray(pol.get_nature_action(np.arange(N) % S), dtype=np.float64)       # arm i must only read o[i]
    tab = np.stack(rows, axis=1)                                                         # [N,S,A]
    if not np.array_equal(again, rows[0]) or not np.array_equal(mixed, tab[np.arange(N), np.arange(N) % S]):
        return None
    return np.ascontiguousarray(tab.astype(np.float32))


class _TableNaturePolicy:
    """A nature policy that plays one fixed parameter per arm (CE, SIS) or per (arm, state, action) (ARMMAN)."""
    base_name = "Table_Nature"

    def __init__(self, nature_params, ind):
        self.nature_params = np.asarray(nature_params, dtype=np.float64)
        self.ind = ind
        self.domain = _domain_of(self.nature_params)
        self.name = self.base_name + ("_c" if self.domain == "counterexample" else "")
        self.param_setting = self._table()

    def _table(self):
        raise NotImplementedError

    def __repr__(self):
        return "%s_%i" % (self.name, self.ind)

    def get_nature_action(self, o):
        if self.domain != "armman":
            return self.param_setting
        o = np.asarray(o).reshape(-1).astype(int)
        return self.param_setting[np.arange(o.shape[0]), o]                             # [N,A] at the arms' states

    def bound_nature_actions(self, actions, state=None, reshape=True):
        return actions


class PessimisticNaturePolicy(_TableNaturePolicy):
    """CE / ARMMAN: the lower end of every range (nature_baselines_counterexample.py:57-59, _armman.py:60-69);
    SIS: (r_t, lambda_c) at their upper ends, (e1, e2) at their lower ends (_sis.py:64-76)."""
    base_name = "Pessimist_Nature"

    def _table(self):
        p = self.nature_params
        if self.domain == "sis":
            return np.stack([p[:, 0, 1], p[:, 1, 1], p[:, 2, 0], p[:, 3, 0]], axis=1)
        return p[..., 0].copy()


class OptimisticNaturePolicy(_TableNaturePolicy):
    base_name = "Optimist_Nature"

    def _table(self):
        p = self.nature_params
        if self.domain == "sis":
            return np.stack([p[:, 0, 0], p[:, 1, 0], p[:, 2, 1], p[:, 3, 1]], axis=1)
        return p[..., 1].copy()


class MiddleNaturePolicy(_TableNaturePolicy):
    """Range midpoints, optionally perturbed by `perturbations` in [0,1) scaled to +- perturbation_size of the range
    width (_armman.py:147-171, _sis.py:167-187; the counterexample class computes the perturbation and then ignores it,
    _counterexample.py:127-136 -- kept)."""
    base_name = "Middle_Nature"

    def __init__(self, nature_params, ind, perturbations=None, perturbation_size=0.1):
        self.perturbations, self.perturbation_size = perturbations, perturbation_size
        super().__init__(nature_params, ind)

    def _table(self):
        p = self.nature_params
        mid = p.mean(axis=-1)
        if self.perturbations is not None and self.domain != "counterexample":
            width = np.ptp(p, axis=-1) * self.perturbation_size
            mid = mid + (np.asarray(self.perturbations) * width * 2 - width)
        return mid


class SampledRandomNaturePolicy(_TableNaturePolicy):
    """One uniform draw inside every range, fixed for the policy's lifetime (`sample_param_setting(seed)`,
    _armman.py:206-218)."""
    base_name = "Sampled_Random_Nature"

    def __init__(self, nature_params, ind):
        super().__init__(nature_params, ind)
        if self.domain == "sis":
            self.name = "Sampled_Random_Nature_sis"
        elif self.domain == "counterexample":
            self.name = "Sampled_Random_Nature"

    def _table(self):
        return None

    def sample_param_setting(self, seed):
        assert self.param_setting is None
        rs = np.random.RandomState()
        rs.seed(seed)
        lo, hi = self.nature_params[..., 0], self.nature_params[..., 1]
        self.param_setting = rs.rand(*lo.shape) * (hi - lo) + lo


class DetermNaturePolicy:
    """nature_baselines_counterexample.py:161-175: plays the table it was given."""

    def __init__(self, param_setting, tok):
        self.param_setting, self.tok, self.name = np.asarray(param_setting), tok, "Determ_Nature"

    def __repr__(self):
        return "%s_%s" % (self.name, self.tok)

    def get_nature_action(self, o):
        return self.param_setting

    def bound_nature_actions(self, actions, state=None, reshape=True):
        return actions


class PessimisticAgentPolicy:
    """baselines/agent_baselines.py:7-19: never acts."""

    def __init__(self, N, ind=0):
        self.N, self.ind, self.name = N, ind, "Pessimist_Agent"

    def __repr__(self):
        return "%s_%i" % (self.name, self.ind)

    def act_test(self, obs):
        return np.zeros(self.N)

    def act_test_device(self, s):
        import torch
        return torch.zeros_like(s)


class RandomAgentPolicy:
    """baselines/agent_baselines.py:22-32: `env.random_agent_action()` every step (host RNG of the env)."""

    def __init__(self, env, ind):
        self.env, self.ind, self.name = env, ind, "Random_Agent"

    def __repr__(self):
        return "%s_%i" % (self.name, self.ind)

    def act_test(self, o):
        return self.env.random_agent_action()

    def act_test_device(self, s):
        import torch
        E = s.shape[1]
        a = np.stack([np.asarray(self.env.random_agent_action()).reshape(-1) for _ in range(E)], axis=1)
        return torch.as_tensor(a.astype(np.uint8), device=s.device)
