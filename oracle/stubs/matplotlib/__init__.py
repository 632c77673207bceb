"""Import-only stand-in for `matplotlib` (test infrastructure only)."""


def use(*a, **k):
    return None
