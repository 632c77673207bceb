"""Drop-in for robust_rmab/utils/mpi_tools.py + mpi_pytorch.py on torch.distributed (NCCL on GPU tensors over
NVLink, gloo on CPU tensors): same function names, argument meaning and float32 arithmetic
(mpi_tools.py:62-67 casts everything to float32 before the all-reduce).

    mpi_fork        :6-42    -> no re-exec: launch with torchrun / torch.multiprocessing; returns immediately
    proc_id / num_procs      -> dist.get_rank / get_world_size (0 / 1 when no process group is initialised)
    mpi_op / mpi_sum / mpi_avg / mpi_statistics_scalar   :59-97
    mpi_avg_grads / sync_params / setup_pytorch_for_mpi  (mpi_pytorch.py) -- ONE flattened all-reduce / broadcast
        per module instead of one per parameter tensor.
"""
import numpy as np
import torch
import torch.distributed as dist


def _on():
    return dist.is_available() and dist.is_initialized()


def mpi_fork(n, bind_to_core=False, is_cannon=False):
    """The reference re-executes the script under mpirun; here ranks are created by the launcher (torchrun)."""
    return


def proc_id():
    return dist.get_rank() if _on() else 0


def num_procs():
    return dist.get_world_size() if _on() else 1


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if _on() and dist.get_backend() == "nccl" else torch.device("cpu")


def allreduce_(t, op="sum"):
    """In-place all-reduce of a tensor (any device the backend supports); no-op for a single process."""
    if num_procs() > 1:
        dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}[op])
    return t


def broadcast(x, root=0):
    if num_procs() > 1:
        t = torch.as_tensor(x).to(_device())
        dist.broadcast(t, src=root)
        if isinstance(x, np.ndarray):
            x[...] = t.cpu().numpy()
        elif isinstance(x, torch.Tensor):
            x.copy_(t)
    return x


def mpi_op(x, op):
    x, scalar = ([x], True) if np.isscalar(x) else (x, False)
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    buff = torch.as_tensor(np.asarray(x, dtype=np.float32)).clone().to(_device())
    allreduce_(buff, op)
    out = buff.cpu().numpy()
    return out[0] if scalar else out


def mpi_sum(x):
    return mpi_op(x, "sum")


def mpi_avg(x):
    return mpi_sum(x) / num_procs()


def mpi_statistics_scalar(x, with_min_and_max=False):
    x = np.array(x, dtype=np.float32)
    global_sum, global_n = mpi_sum([np.sum(x), len(x)])
    mean = global_sum / global_n
    global_sum_sq = mpi_sum(np.sum((x - mean) ** 2))
    std = np.sqrt(global_sum_sq / global_n)
    if with_min_and_max:
        global_min = mpi_op(np.min(x) if len(x) > 0 else np.inf, op="min")
        global_max = mpi_op(np.max(x) if len(x) > 0 else -np.inf, op="max")
        return mean, std, global_min, global_max
    return mean, std


def setup_pytorch_for_mpi():
    if torch.get_num_threads() == 1:
        return
    torch.set_num_threads(max(int(torch.get_num_threads() / num_procs()), 1))


def mpi_avg_grads(module):
    """Average gradient buffers across ranks with one flattened all-reduce."""
    if num_procs() == 1:
        return
    ps = [p for p in module.parameters() if p.grad is not None]
    if not ps:
        return
    flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in ps]).to(_device())
    allreduce_(flat)
    flat /= num_procs()
    o = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[o:o + n].reshape(p.grad.shape).to(p.grad.device))
        o += n


def sync_params(module):
    """Broadcast rank 0's parameters with one flattened broadcast."""
    if num_procs() == 1:
        return
    ps = list(module.parameters())
    flat = torch.cat([p.data.reshape(-1).to(torch.float32) for p in ps]).to(_device())
    dist.broadcast(flat, src=0)
    o = 0
    for p in ps:
        n = p.data.numel()
        p.data.copy_(flat[o:o + n].reshape(p.data.shape).to(p.data.device))
        o += n


def arm_shard(n_arms, rank=None, world=None):
    """Contiguous arm range [arm0, arm0 + n) of a rank."""
    rank = proc_id() if rank is None else rank
    world = num_procs() if world is None else world
    per = (n_arms + world - 1) // world
    arm0 = min(rank * per, n_arms)
    return arm0, min(per, n_arms - arm0)
