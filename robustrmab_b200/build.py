"""Build librmab_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python -m robustrmab_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librmab_b200.so")
SOURCES = ["common.cu", "env.cu", "ac.cu", "gae.cu", "epoch.cu", "ppo.cu", "ppo_table.cu", "ppo_table_mma.cu", "ppo_table_mma64.cu", "ppo_table_umma64.cu", "ppo_table_umma16.cu", "ma.cu", "select.cu", "vi.cu", "gemm3h.cu", "nature.cu", "table_fwd_umma64.cu", "ac_umma64.cu"]
HEADERS = ["common.cuh", "mlp.cuh", "umma.cuh", os.path.join("..", "..", "include", "rmab.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(f"--- {src} ---\n{out}")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"])
    build_probe(force)
    return LIB


def build_probe(force=False):
    """tools/umma_probe.cu -> tools/bin/umma_probe: the stand-alone check of the tcgen05 descriptor encodings
    ppo_table_umma64.cu uses (run by tests/test_gpu_kernels.py on the GPU box)."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tools", "umma_probe.cu")
    out = os.path.join(root, "tools", "bin", "umma_probe")
    if os.path.exists(src) and (force or _stale(out, [src])):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O2", "-o", out, src])
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
