"""The C-ABI library loads on a CPU-only box and exports every symbol include/rmab.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rmab.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rmab_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = _declared()
    assert "rmab_env_step_table" in names and "rmab_ppo_update" in names and len(names) >= 20


def test_library_exports_every_declared_symbol():
    from robustrmab_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(L, n)]
    assert not missing, missing
    assert set(_declared()) == set(_lib.SIGNATURES), set(_declared()) ^ set(_lib.SIGNATURES)
    assert _lib.lib().rmab_abi_version() == 1


def test_no_cpu_fallback():
    import torch
    from robustrmab_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.reset_states(2, 2, 2, ops.rng(0, 3, 0), out=torch.zeros((2, 2), dtype=torch.uint8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "robustrmab_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
