#!/usr/bin/env python
"""bench.py -- RMABPPO arm-steps/sec (rollout + update) on B200, with the CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--hidden 16|64]

A "step" is one RMABPPO epoch (agent_oracle.py:492-593) over one batch of synthetic arms: lambda-net
forward, T rollout steps (actor sample + env transition), bootstrap, GAE + advantage normalisation,
train_pi_iters + train_v_iters full-batch Adam steps per arm, lambda SGD step when due, env reset.
Workload at N=1: BASELINE.json configs[1] -- counterexample domain, N=4098 arms (4096 rounded to a
multiple of 3, bandit_env_robust.py:755) x 256 envs, T=100, hidden [16,16], 20+20 iterations.
Under torchrun every rank holds that many arms (weak scaling by arm).
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "RMABPPO arm-steps/sec (rollout+update)"
UNIT = "arm-steps/s"


def measure_tf32_gemm(dev, n=8192, reps=5):
    """Library TF32 GEMM rate on this GPU (torch.matmul fp32 with TF32 allowed, n^3, best of `reps`, CUDA events): the
    denominator SURVEY.md 8(d) asks for beside MEASURED_PEAKS.json, which has no tf32 line."""
    import torch
    a = torch.randn((n, n), device=dev)
    b = torch.randn((n, n), device=dev)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    best = 0.0
    try:
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return best


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--arms", type=int, default=4098)
    ap.add_argument("--envs", type=int, default=256)
    ap.add_argument("--horizon", type=int, default=100)
    ap.add_argument("--hidden", type=int, default=16)
    ap.add_argument("--domain", default="counterexample", choices=["counterexample", "armman"])
    ap.add_argument("--pi-iters", type=int, default=20)
    ap.add_argument("--v-iters", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-whittle", action="store_true")
    ap.add_argument("--no-armman", action="store_true", help="skip the configs[2] block (ARMMAN 100k x 1024, hidden 64)")
    ap.add_argument("--no-ma", action="store_true", help="skip the configs[3] block (MA-RMABPPO on SIS)")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-kernel HBM lines (env step, GAE, budget select)")
    ap.add_argument("--quick", action="store_true", help="headline only: no secondary blocks, no CPU legs")
    ap.add_argument("--vi-arms", type=int, default=99999)
    ap.add_argument("--vi-probes", type=int, default=64)
    ap.add_argument("--cpu-arms", type=int, default=192)
    ap.add_argument("--cpu-envs", type=int, default=16)
    a = ap.parse_args()
    if a.quick:
        a.no_cpu_baseline = a.no_whittle = a.no_armman = a.no_ma = a.no_kernels = True
    return a


def workload_name(a, n_ranks):
    return (f"RMABPPO {a.domain} domain, N={a.arms} arms/GPU x {a.envs} envs, T={a.horizon}, hidden [{a.hidden},{a.hidden}], "
            f"{a.pi_iters}+{a.v_iters} iters, binary action" + (f", {n_ranks} ranks sharded by arm" if n_ranks > 1 else ""))


# ------------------------------------------------------------------------------------------ CPU legs
def _cpu_problem(a, n_arms, n_envs, seed=0):
    """Synthetic problem of the bench's shape at a CPU-sized arm / env count (oracle objects)."""
    import numpy as np
    from oracle import envs as oenv, nets as onets
    rs = np.random.RandomState(seed)
    dom = oenv.DOMAIN_CE if a.domain == "counterexample" else oenv.DOMAIN_ARMMAN
    S = 2 if dom == oenv.DOMAIN_CE else 3
    base = oenv.ce_param_ranges(n_arms) if dom == oenv.DOMAIN_CE else oenv.armman_param_ranges(n_arms)
    ranges = oenv.sample_sub_ranges(base, rs).astype(np.float32)
    nature = (ranges[..., 0] + rs.rand(*ranges.shape[:-1]) * (ranges[..., 1] - ranges[..., 0])).astype(np.float32)
    Wpi = onets.init_linear_default(rs, n_arms, S + 1, a.hidden, 2)
    Wv = onets.init_linear_default(rs, n_arms, S + 1, a.hidden, 1)
    b = 1 / np.sqrt(n_arms)
    lam = [rs.uniform(-b, b, (8, n_arms)), rs.uniform(-b, b, 8), rs.uniform(-.35, .35, (8, 8)), rs.uniform(-.35, .35, 8),
           rs.uniform(-.35, .35, (1, 8)), rs.uniform(-.35, .35, 1)]
    return dom, S, Wpi, Wv, lam, nature, ranges, oenv.rewards_table(dom, n_arms, S), np.array([0, 1], np.float32)


def _cpu_worker(args):
    a, n_arms, n_envs, epochs, seed = args
    import torch
    torch.set_num_threads(1)
    from oracle.train import CpuRMABPPO
    dom, S, Wpi, Wv, lam, nature, ranges, R, C = _cpu_problem(a, n_arms, n_envs, seed)
    tr = CpuRMABPPO(dom, Wpi, Wv, lam, nature, ranges, R, C, n_envs, a.horizon, a.hidden, n_arms // 3,
                    train_pi_iters=a.pi_iters, train_v_iters=a.v_iters, seed=seed)
    tr.epoch()                                            # warm-up (imports, torch autograd setup)
    t0 = time.perf_counter()
    for _ in range(epochs):
        tr.epoch()
    return time.perf_counter() - t0


def cpu_port_rate(a, cores, epochs=1, pool=None):
    """arm-steps/s of the oracle port on `cores` host processes (one env shard each, like mpirun -np).  `pool`: a
    multiprocessing pool to reuse across calls (worker start-up -- importing torch -- is then paid once)."""
    import multiprocessing as mp
    jobs = [(a, a.cpu_arms, a.cpu_envs, epochs, 100 + i) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1 and pool is None:
        times = [_cpu_worker(jobs[0])]
    elif pool is not None:
        times = pool.map(_cpu_worker, jobs)
    else:
        with mp.get_context("spawn").Pool(cores) as p_:
            times = p_.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    per_job = a.cpu_arms * a.cpu_envs * a.horizon * epochs
    rate = sum(per_job / t for t in times)
    return rate, max(times), wall


def recorded_reference_classes():
    """The UNMODIFIED reference classes timed in the build container (tools/time_reference_here.py -> profiles/): the
    reference is pure Python that `pip install --target` cannot package (empty wheel), so it cannot run on the GPU box."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_reference_classes_cpu.json")))
    except (OSError, ValueError):
        return None


def run_reference(a):
    """--impl reference: the reference algorithm on the host cores.  The reference is pure Python and cannot
    travel to the GPU box (nor be compiled), so this times its CPU restatement (oracle/, pinned to the
    reference by tests/golden) with one process per host core, as `--cpu k` / mpirun would."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, 32))
    rates, t0 = [], time.perf_counter()
    step_times = []
    with mp.get_context("spawn").Pool(cores) as pool:       # one pool for the whole run: workers import torch once
        for _ in range(a.warmup):
            cpu_port_rate(a, cores, 1, pool)
            if time.perf_counter() - t0 > 90:
                break
        for _ in range(a.steps):
            rate, tmax, wall = cpu_port_rate(a, cores, 1, pool)
            rates.append(rate)
            step_times.append(tmax)
            if time.perf_counter() - t0 > 270:
                break
    value = sum(rates) / len(rates)
    sample = (f"{cores} processes x (N={a.cpu_arms} arms x {a.cpu_envs} envs x T={a.horizon}) per step, oracle port, "
              f"1 torch thread each")
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": len(rates), "warmup": a.warmup,
           "ms_per_step": 1e3 * sum(step_times) / len(step_times), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": {"workload": f"RMABPPO {a.domain} domain, {cores} host processes x (N={a.cpu_arms} arms x {a.cpu_envs} envs), "
                                  f"T={a.horizon}, hidden [{a.hidden},{a.hidden}], {a.pi_iters}+{a.v_iters} iters -- the CPU-sized sample of: "
                                  + workload_name(a, 1), "sample": sample},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                            "reference_classes_recorded": recorded_reference_classes()},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ ours
def run_ours(a):
    import torch
    import torch.distributed as dist
    from robustrmab_b200 import _lib
    from robustrmab_b200.trainer import RMABPPOTrainer
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    tr = RMABPPOTrainer(a.domain, a.arms * world, a.envs, a.horizon, hidden=a.hidden, train_pi_iters=a.pi_iters,
                        train_v_iters=a.v_iters, epochs=max(20, a.warmup + a.steps), seed=0, device=dev, rank=rank,
                        world_size=world)
    N, E, T = tr.N, tr.E, tr.T

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()          # samples cover the warm-up and the timed region (both under load)
    # the lambda-net SGD step runs every lamb_update_freq epochs only, i.e. possibly first inside the timed region: run that
    # branch once now with a zero learning rate (state unchanged) so its one-off kernel loads happen here
    tr.prewarm()
    for _ in range(max(a.warmup, 3)):
        tr.epoch()
    if world > 1:                      # NCCL picks its algorithm / channels lazily: settle it before the timed region
        t_w = torch.zeros(2 * E + 1, device=dev)
        for _ in range(20):
            dist.all_reduce(t_w)
    barrier()
    tr.timers = {}
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        tr.epoch()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    stage_ms = tr.stage_times()
    tr.timers = None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clk = clocks.stop() if rank == 0 else None
    total_arm_steps = float(a.arms * world) * E * T * a.steps
    value = total_arm_steps / (ms * 1e-3)

    # ---- end to end through the host-buffer API: H2D of the step's inputs, D2H of its results ----
    s0_host = torch.empty((N, E), dtype=torch.uint8).pin_memory()
    s0_host.copy_(tr.s_traj[:, 0])
    nat_host = tr.nature.cpu().pin_memory()
    for _ in range(2):
        tr.epoch_host(s0_host, nat_host)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        res = tr.epoch_host(s0_host, nat_host)
    e1.record()
    barrier()
    ems = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_value = total_arm_steps / (float(ems.item()) * 1e-3)
    h2d = s0_host.numel() + nat_host.numel() * 4
    d2h = sum(v.numel() * v.element_size() for v in res.values())

    # ---- secondary blocks that every rank takes part in (sharded by arm under torchrun) ----
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_blocks as bb
    S_dom = tr.S
    Ppi, Pv, buf_mib = tr.Ppi, tr.Pv, tr.buffer_bytes() / 2 ** 20
    alg_bytes = tr.algorithmic_bytes()
    del tr
    torch.cuda.empty_cache()
    blocks = {}
    if not a.no_ma:        # first: its rollout is launch-latency sensitive and measured slower right after the seconds-long ARMMAN epochs
        blocks["ma_sis"] = bb.guarded(bb.ma_sis_block, dev, peaks, rank, world)
    if not a.no_armman and a.hidden != 64:
        blocks["ce_h64"] = bb.guarded(bb.ce_h64_block, dev, peaks, rank, world)
        blocks["armman_full"] = bb.guarded(bb.armman_block, dev, peaks, rank, world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the dominant kernel against its binding roof ----
    dom_name = max(stage_ms, key=stage_ms.get) if stage_ms else "none"
    dom_ms = stage_ms.get(dom_name, 0.0)
    tf32_peak = peaks.get("bf16_tflops_sustained", 2250.0 * 0.6) / 2
    tf32_src = "measured bf16 sustained / 2 (tf32)" if "bf16_tflops_sustained" in peaks else "fallback 0.6 x 2250 / 2"
    traffic = None
    try:   # DRAM bytes per launch from the committed ncu capture of this same command (profiles/), default workload only
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        if (a.arms, a.envs, a.horizon, a.hidden, a.domain) == (4098, 256, 100, 16, "counterexample"):
            traffic = {"ppo_update": tj["ppo_update"], "rollout_gae_stats": tj["rollout_gae_stats"]}.get(dom_name)
    except (OSError, KeyError, ValueError):
        pass
    hbm_stage = {k: {"achieved_GBps": alg_bytes[k] / (stage_ms[k] * 1e-3) / 1e9,
                     "frac": alg_bytes[k] / (stage_ms[k] * 1e-3) / 1e9 / hbm_peak,
                     "algorithmic_bytes_per_launch": alg_bytes[k]} for k in alg_bytes if stage_ms.get(k)}
    if dom_name == "ppo_update":
        # The update's binding roof is the tensor / issue side (K5 is compute-bound; the statistics reformulation of
        # DESIGN.md 3.1 keeps the batch out of HBM).  achieved = EXECUTED tensor-core flops: 3 GEMMs x 3 split products of
        # [rows, h] x [h, h] per arm and Adam step, rows = S*E padded to the 128-row tile; peak = tf32 dense.
        h = a.hidden
        rows = ((S_dom * E + 127) // 128) * 128
        flops = float(N) * (a.pi_iters + a.v_iters) * rows * 3 * 3 * 2 * h * h
        ach = flops / (dom_ms * 1e-3) / 1e12
        per_sample = 6.0 * (a.pi_iters * Ppi + a.v_iters * Pv) * N * T * E
        roof = {"bound": "tensor", "kernel": f"ppo_update = ppo_table_umma{h}_kernel (pi + v launches of one epoch)",
                "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak, "traffic": traffic,
                "peak_source": tf32_src, "executed_tf32_flops_per_launch": flops, "ms_per_launch": dom_ms,
                "fp32_equivalent_TFLOPs": ach / 3, "rows_per_arm": S_dom * E, "samples_per_arm": T * E,
                "algorithmic_per_sample_TFLOPs": per_sample / (dom_ms * 1e-3) / 1e12,
                "note": "tensor-core flops actually executed (the statistics reformulation runs S*E rows per arm instead of T*E "
                        "samples); the CUDA-core work around the small GEMMs (tanh, splits, row sums, Adam) issue-bounds the kernel -- "
                        "see profiles/ for issue-active and tensor-pipe activity",
                "hbm_view": hbm_stage}
        if h == 64:
            tf32_measured = measure_tf32_gemm(dev)
            roof.update({"cublas_tf32_8192_TFLOPs": tf32_measured, "frac_of_cublas_tf32": ach / tf32_measured if tf32_measured else None})
    else:
        b_dom = alg_bytes.get(dom_name, 0)
        ach = b_dom / (dom_ms * 1e-3) / 1e9 if dom_ms else 0.0
        roof = {"bound": "hbm", "kernel": dom_name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": b_dom, "ms_per_launch": dom_ms,
                "hbm_view": hbm_stage}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
           "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(a, world), "l2": "per-step working set (statistics + packed nets + Adam "
                      f"moments, {buf_mib:.0f} MiB/GPU) exceeds the 126 MB L2; no explicit flush",
                      "parallelism": f"arm-sharded x{world}"},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": int(launches), "roofline": roof, "stage_ms_per_step": stage_ms, "clocks": clk}
    if not a.no_cpu_baseline:
        rate, tmax, _ = cpu_port_rate(a, 1, 1)
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"1 epoch of N={a.cpu_arms} arms x {a.cpu_envs} envs x T={a.horizon} "
                                         f"(oracle port, 1 thread, {tmax:.1f} s)",
                               "reference_classes_recorded": recorded_reference_classes()}
    out.update(blocks)
    if not a.no_whittle:
        out["whittle"] = bb.guarded(bb.whittle_block, dev, hbm_peak, peak_src, not a.no_cpu_baseline, a.vi_arms, a.vi_probes)
    if not a.no_kernels:
        out["kernels"] = bb.guarded(bb.kernel_lines, dev, hbm_peak, peak_src)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
