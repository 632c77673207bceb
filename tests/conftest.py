import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
        return cache[name]
    return load


# ---- tolerance headroom report -------------------------------------------------------------------------------------
# RMAB_TOL_REPORT=<path> wraps numpy.testing.assert_allclose for the session and writes, per call site, the tolerance in
# force and the largest observed  |got - want| / (atol + rtol |want|)  (1.0 = at the limit), so that every tolerance
# in tests/ can be set from measured headroom rather than guessed (profiles/r2_tolerance_headroom.json is one such run).
_TOL_SITES = {}


def _install_tol_report():
    import inspect
    orig = np.testing.assert_allclose

    def wrapped(actual, desired, rtol=1e-7, atol=0, *a, **k):
        try:
            x, y = np.asarray(actual, dtype=np.float64), np.asarray(desired, dtype=np.float64)
            fr = inspect.stack()[1]
            site = f"{os.path.basename(fr.filename)}:{fr.lineno}"
            with np.errstate(all="ignore"):
                den = atol + rtol * np.abs(y)
                ratio = float(np.nanmax(np.where(den > 0, np.abs(x - y) / den, np.where(x == y, 0.0, np.inf)))) if x.size else 0.0
            rec = _TOL_SITES.setdefault(site, dict(rtol=rtol, atol=atol, worst=0.0, calls=0))
            rec["worst"], rec["calls"] = max(rec["worst"], ratio), rec["calls"] + 1
        except Exception:      # the report must never change a test's outcome
            pass
        return orig(actual, desired, rtol=rtol, atol=atol, *a, **k)
    np.testing.assert_allclose = wrapped


if os.environ.get("RMAB_TOL_REPORT"):
    _install_tol_report()


def pytest_sessionfinish(session, exitstatus):
    path = os.environ.get("RMAB_TOL_REPORT")
    if path and _TOL_SITES:
        import json
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w") as f:
            json.dump(dict(sorted(_TOL_SITES.items())), f, indent=1)
