"""Import-only stand-in for `mpi4py`: a size-1 communicator (test infrastructure only)."""
from . import MPI  # noqa: F401
