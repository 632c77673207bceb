"""Rate of the package's dense product (csrc/gemm3h.cu) at the BASELINE configs[3] shapes, next to the library 3xTF32
arm it replaces (ops.matmul_3xtf32: three cuBLAS TF32 GEMMs).   python tools/gemm3h_bench.py [N_arms]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustrmab_b200 import ops  # noqa: E402

N_arms = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
h, T, E = 64, 20, 256
M, D, NH = T * E, 4 * N_arms, N_arms * h
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
nat = torch.rand((M, D), device=dev, generator=g) * 9 + 1
W = (torch.rand((NH, D), device=dev, generator=g) - 0.5) * 0.02
dz = torch.randn((M, NH), device=dev, generator=g) * 1e-4


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


flops = 2.0 * M * D * NH
# forward: z1 [M, NH] = nat [M, D] . W^T
a = ops.split16_rows(nat)
ms_split_w = timed(lambda: ops.split16_rows(W))
b = ops.split16_rows(W)
out = torch.empty((M, NH), device=dev)
for nf in (True, False):
    ms = timed(lambda: ops.gemm3h(a, b, D, out=out, n_fast=nf))
    print(f"z1 gemm3h n_fast={nf}: {ms:.2f} ms  {flops / ms / 1e9:.0f} TFLOP/s fp32-equivalent, {3 * flops / ms / 1e9:.0f} fp16 executed")
nh, nl = ops.split_tf32(nat)
wh, wl = ops.split_tf32(W)
ms_lib = timed(lambda: ops.matmul_3xtf32((nh, nl), (wh.t(), wl.t())))
ref = ops.matmul_3xtf32((nh, nl), (wh.t(), wl.t()))
want = nat[:64].double() @ W.double().t()
print(f"z1 library 3xTF32: {ms_lib:.2f} ms  {flops / ms_lib / 1e9:.0f} TFLOP/s fp32-equivalent; split16_rows(W) {ms_split_w:.2f} ms")
print("   max rel err vs fp64 (first 64 rows): gemm3h %.2e  library %.2e" % (
    float(((out[:64].double() - want).abs() / want.abs().clamp_min(1e-3)).max()),
    float(((ref[:64].double() - want).abs() / want.abs().clamp_min(1e-3)).max())))
del ref, nh, nl, wh, wl, out
# gradient: dW [NH, D] = dz^T [NH, M] . nat [M, D]
ms_split_dz = timed(lambda: ops.split16_cols(dz))
a2 = ops.split16_cols(dz)
b2 = ops.split16_cols(nat)
out2 = torch.empty((NH, D), device=dev)
for nf in (True, False):
    ms = timed(lambda: ops.gemm3h(a2, b2, M, out=out2, n_fast=nf))
    print(f"dW gemm3h n_fast={nf}: {ms:.2f} ms  {flops / ms / 1e9:.0f} TFLOP/s fp32-equivalent; split16_cols(dz) {ms_split_dz:.2f} ms")
dh, dl = ops.split_tf32(dz)
nh, nl = ops.split_tf32(nat)
ms_lib = timed(lambda: ops.matmul_3xtf32((dh.t(), dl.t()), (nh, nl)))
print(f"dW library 3xTF32: {ms_lib:.2f} ms  {flops / ms_lib / 1e9:.0f} TFLOP/s fp32-equivalent")
# rollout step: [E, D] x [D, NH]
a3 = ops.split16_rows(nat[:E].contiguous())
out3 = torch.empty((E, NH), device=dev)
ms = timed(lambda: ops.gemm3h(a3, b, D, out=out3, n_fast=False), reps=10)
wh, wl = ops.split_tf32(W)
ms_l = timed(lambda: ops.matmul_3xtf32((wh, wl), nat[:E].t()), reps=10)
print(f"rollout step [E={E}] gemm3h {ms:.3f} ms vs library {ms_l:.3f} ms")
