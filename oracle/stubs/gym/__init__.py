"""Import-only stand-in for `gym` so the reference modules load in the build container.

Test infrastructure only (used by oracle/gen_golden.py and the reference-pinning tests);
never imported by the product package. No arithmetic lives here.
"""


class Env:  # the reference only subclasses it
    pass


from . import spaces  # noqa: E402,F401
