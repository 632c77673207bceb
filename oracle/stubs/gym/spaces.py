"""Import-only stand-in for `gym.spaces` (the reference does `from gym.spaces import Box, Discrete`)."""


class _Space:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k


Box = _Space
Discrete = _Space
