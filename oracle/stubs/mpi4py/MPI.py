"""Size-1 COMM_WORLD: Allreduce copies, Bcast is a no-op (test infrastructure only)."""
import numpy as _np

SUM, MIN, MAX = "sum", "min", "max"


class _Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Allreduce(self, x, buf, op=SUM):
        _np.copyto(buf, _np.asarray(x, dtype=buf.dtype))

    def Bcast(self, x, root=0):
        return None


COMM_WORLD = _Comm()
