"""Secondary blocks of bench.py's JSON line: one function per BASELINE.json config that is not the headline, plus the
per-kernel HBM lines north_star asks for ("achieved HBM GB/s for the simulator, scan and value-iteration kernels").

Every block returns a dict with `config.workload`, the metric, `ms_per_step` and a `roofline` object
(`bound`, `achieved`, `peak`, `unit`, `frac`) whose algorithmic bytes / flops are stated next to it.  All timing is CUDA
events on the launching stream after warm-up; inputs are created once on the device and are larger than the 126 MB L2
unless the block says otherwise.  Blocks never raise: a failure is reported as {"error": "..."} so the headline survives.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from robustrmab_b200 import _lib, ops                                      # noqa: E402

UNIT = "arm-steps/s"
METRIC = "RMABPPO arm-steps/sec (rollout+update)"


def _timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _hbm(bytes_per_launch, ms, peak, src):
    ach = bytes_per_launch / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
            "peak_source": src, "algorithmic_bytes_per_launch": int(bytes_per_launch), "ms_per_launch": ms}


def _max_over_ranks(ms, dev, world):
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def guarded(fn, *a, **k):
    try:
        return fn(*a, **k)
    except Exception as e:   # noqa: BLE001 -- a secondary block must not take the headline line down
        torch.cuda.synchronize()
        return {"error": f"{type(e).__name__}: {e}"[:400]}


# ------------------------------------------------------------------------------------------ per-kernel HBM lines
def kernel_lines(dev, hbm_peak, peak_src, N=100_000, E=1024, T=10):
    """K1 env_step (ARMMAN table), K3 GAE + normalisation, K6 budget select (binary top-k) at the configs[2] shape
    (100 000 arms x 1024 envs); algorithmic bytes per unit are SURVEY.md 8(d)'s."""
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    rs = np.random.RandomState(0)
    from robustrmab_b200.trainer import RMABPPOTrainer
    # ---- K1: ARMMAN env step with the per-arm nature table, reward written (s 1 + a 1 + s' 1 + r 4 = 7 B per arm-step)
    ranges = RMABPPOTrainer.default_ranges(_lib.DOMAIN_ARMMAN, N, rs)
    nature = torch.as_tensor(ranges.mean(-1), dtype=torch.float32, device=dev).contiguous()
    rg = torch.as_tensor(ranges, dtype=torch.float32, device=dev).contiguous()
    R = torch.tensor([1.0, 0.5, 0.0], device=dev).repeat(N, 1).contiguous()
    s = torch.randint(0, 3, (N, E), dtype=torch.uint8, device=dev, generator=g)
    a = torch.randint(0, 2, (N, E), dtype=torch.uint8, device=dev, generator=g)
    s2, rew = torch.empty_like(s), torch.empty((N, E), dtype=torch.float32, device=dev)
    k = [0]

    def step():
        k[0] += 1
        ops.env_step_table(_lib.DOMAIN_ARMMAN, s, a, nature, rg, R, r=ops.rng(0, _lib.STREAM_ENV, k[0]), s_next=s2, rew=rew)
    ms = _timed(step)
    out["env_step_table"] = {"workload": f"ARMMAN env.step, {N} arms x {E} envs, per-arm nature table, Philox draws, reward stored",
                             "arm_steps_per_s": N * E / (ms * 1e-3), "bytes_per_arm_step": 7,
                             "roofline": _hbm(7 * N * E, ms, hbm_peak, peak_src)}

    def step_nr():
        k[0] += 1
        ops.env_step_table(_lib.DOMAIN_ARMMAN, s, a, nature, rg, R, r=ops.rng(0, _lib.STREAM_ENV, k[0]), s_next=s2,
                           want_reward=False)
    ms = _timed(step_nr)
    out["env_step_table_no_reward"] = {"workload": "same, reward derived downstream (s 1 + a 1 + s' 1 = 3 B per arm-step)",
                                       "arm_steps_per_s": N * E / (ms * 1e-3), "bytes_per_arm_step": 3,
                                       "roofline": _hbm(3 * N * E, ms, hbm_peak, peak_src)}
    del rew, s2
    # ---- K6: budget select, binary actions: top-B arms by p1 per env (read p1 4 + write a 1 = 5 B per arm-env)
    p1 = torch.rand((N, E), device=dev, generator=g)
    ms = _timed(lambda: ops.budget_select_binary(p1, 1.0, float(N // 5)))
    out["budget_select_binary"] = {"workload": f"act_test top-B (B = N/5) over {N} arms per env, {E} envs",
                                   "arm_envs_per_s": N * E / (ms * 1e-3), "bytes_per_arm_env": 5,
                                   "roofline": _hbm(5 * N * E, ms, hbm_peak, peak_src)}
    del p1
    # ---- K3: GAE + per-arm advantage normalisation (s' 1 + a 1 + v 4 read, adv 4 + ret 4 written = 14 B per arm-step)
    Ng = N // 2
    s_traj = torch.randint(0, 3, (Ng, T + 1, E), dtype=torch.uint8, device=dev, generator=g)
    at = torch.randint(0, 2, (Ng, T, E), dtype=torch.uint8, device=dev, generator=g)
    val = torch.randn((Ng, T, E), device=dev, generator=g)
    last = torch.randn((Ng, E), device=dev, generator=g)
    lamb = torch.rand((E,), device=dev, generator=g)
    Cc = torch.tensor([0.0, 1.0], device=dev)
    ms = _timed(lambda: ops.gae(s_traj, at, val, last, lamb, Cc, R[:Ng].contiguous(), 0.9, 0.97, normalize=True), reps=3, warm=1)
    out["gae"] = {"workload": f"finish_path + get: GAE-lambda, returns, per-arm advantage normalisation, {Ng} arms x T={T} x {E} envs",
                  "arm_steps_per_s": Ng * T * E / (ms * 1e-3), "bytes_per_arm_step": 14,
                  "roofline": _hbm(14 * Ng * T * E, ms, hbm_peak, peak_src)}
    del s_traj, at, val
    # ---- SIS env step (binomial next-state law evaluated on the fly, fp64) and the exact multiple-choice knapsack of the
    # Hawkins baseline (lp_methods.py:221-271) at the configs[3] shape: latency / ALU bound, rates only
    Ns, Es, pop = 2048, 256, 50
    ss = torch.randint(0, pop + 1, (Ns, Es), dtype=torch.uint8, device=dev, generator=g)
    sa = torch.randint(0, 3, (Ns, Es), dtype=torch.uint8, device=dev, generator=g)
    lo = torch.tensor([0.5, 1.0, 1.0, 1.0], device=dev)
    hi = torch.tensor([0.99, 10.0, 10.0, 10.0], device=dev)
    prm = (lo + torch.rand((Ns, 4), device=dev, generator=g) * (hi - lo)).contiguous()
    Rs = torch.linspace(0, 1, pop + 1, device=dev).repeat(Ns, 1).contiguous()

    def sis():
        k[0] += 1
        ops.env_step_sis(pop, ss, sa, prm, Rs, r=ops.rng(0, _lib.STREAM_ENV, k[0]))
    ms = _timed(sis)
    out["env_step_sis"] = {"workload": f"SIS env.step, pop={pop}, {Ns} arms x {Es} envs (binomial pmf per transition, fp64)",
                           "arm_steps_per_s": Ns * Es / (ms * 1e-3), "ms_per_launch": ms, "bound": "ALU (fp64 pmf recurrence)"}
    vals = torch.rand((Ns, 3, Es), device=dev, dtype=torch.float64, generator=g)
    Bk = int(0.2 * Ns)
    ms = _timed(lambda: ops.knapsack_multichoice(vals, [0, 1, 2], Bk), reps=2, warm=1)
    out["knapsack_multichoice"] = {"workload": f"exact multiple-choice knapsack, {Ns} arms x 3 actions (costs 0,1,2), budget {Bk} "
                                               f"met with equality, {Es} envs (one CTA per env, DP over arms)",
                                   "arms_per_s": Ns * Es / (ms * 1e-3), "ms_per_launch": ms,
                                   "bound": "latency (N dependent DP layers of B+1 fp64 cells)"}
    return out


# ------------------------------------------------------------------------------------------ configs[2]
def armman_block(dev, peaks, rank=0, world=1, n_total=100_000, E=1024, T=10, h=64, iters=20, steps=2, warm=1):
    """BASELINE configs[2] at FULL size: ARMMAN, 100 000 arms x 1024 envs, T = 10, hidden [64,64], 20+20 Adam steps.
    world == 1: all arms on one GPU; under torchrun the arms are split (strong scaling, n_total / world per rank)."""
    from robustrmab_b200.trainer import RMABPPOTrainer
    tr = RMABPPOTrainer("armman", n_total, E, T, hidden=h, train_pi_iters=iters, train_v_iters=iters, epochs=20, seed=0,
                        device=dev, rank=rank, world_size=world)
    tr.prewarm()
    for _ in range(warm):
        tr.epoch()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    tr.timers = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.epoch()
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)
    stage_ms = tr.stage_times()
    upd = stage_ms.get("ppo_update", 0.0)
    rows = ((3 * E + 127) // 128) * 128
    flops = float(tr.N) * 2 * iters * rows * 3 * 3 * 2 * h * h     # 3 GEMMs x 3 tf32 products per row and Adam step
    tf32_peak = peaks.get("bf16_tflops_sustained", 2250.0 * 0.6) / 2
    ach = flops / (upd * 1e-3) / 1e12 if upd else 0.0
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    n_local = tr.N
    del tr
    torch.cuda.empty_cache()
    return {"metric": METRIC, "unit": UNIT, "value": float(n_total) * E * T / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
            "n_gpus": world, "scaling": "strong" if world > 1 else "single GPU", "stage_ms_per_step": stage_ms,
            "dtype": "f32 (3xTF32 tcgen05 GEMMs, fp32 accumulation in TMEM)", "peak_mem_GiB": mem,
            "config": {"workload": f"RMABPPO armman domain, N={n_total} arms x {E} envs, T={T}, hidden [{h},{h}], "
                                   f"{iters}+{iters} iters, top-k budget B=N/5" +
                                   (f", arm-sharded over {world} GPUs ({n_local} arms on rank 0)" if world > 1 else ", one GPU")},
            "roofline": {"bound": "tensor", "kernel": "ppo_table_umma64_kernel (pi + v), rank 0", "achieved": ach,
                         "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak if tf32_peak else None, "traffic": None,
                         "executed_tf32_flops_per_launch": flops, "ms_per_launch": upd, "fp32_equivalent_TFLOPs": ach / 3,
                         "peak_source": "measured bf16 sustained / 2 (tf32)" if "bf16_tflops_sustained" in peaks else "fallback"}}


def ce_h64_block(dev, peaks, rank=0, world=1, n_per_gpu=4098, E=256, T=100, h=64, iters=20, steps=5, warm=2):
    """SURVEY 8(d) config 2, second width: the headline problem (CE, 4098 arms per GPU x 256 envs, T = 100, 20+20 Adam steps)
    with hidden [64,64] -- the hidden-64 tcgen05 kernels (table pre-pass of the rollout kernel, update) on S*E = 512 rows per
    arm.  Weak-scaled like the headline under torchrun."""
    from robustrmab_b200.trainer import RMABPPOTrainer
    n_total = n_per_gpu * world
    tr = RMABPPOTrainer("counterexample", n_total, E, T, hidden=h, train_pi_iters=iters, train_v_iters=iters, epochs=20, seed=0,
                        device=dev, rank=rank, world_size=world)
    tr.prewarm()
    for _ in range(warm):
        tr.epoch()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    tr.timers = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.epoch()
    e1.record()
    torch.cuda.synchronize()
    ms = _max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)
    stage_ms = tr.stage_times()
    upd = stage_ms.get("ppo_update", 0.0)
    rows = ((2 * E + 127) // 128) * 128
    flops = float(tr.N) * 2 * iters * rows * 3 * 3 * 2 * h * h     # 3 GEMMs x 3 tf32 products per row and Adam step
    tf32_peak = peaks.get("bf16_tflops_sustained", 2250.0 * 0.6) / 2
    ach = flops / (upd * 1e-3) / 1e12 if upd else 0.0
    del tr
    torch.cuda.empty_cache()
    return {"metric": METRIC, "unit": UNIT, "value": float(n_total) * E * T / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
            "n_gpus": world, "scaling": "weak" if world > 1 else "single GPU", "stage_ms_per_step": stage_ms,
            "dtype": "f32 (3xTF32 tcgen05 GEMMs, fp32 accumulation in TMEM)",
            "config": {"workload": f"RMABPPO counterexample domain, N={n_per_gpu} arms/GPU x {E} envs, T={T}, hidden [{h},{h}], "
                                   f"{iters}+{iters} iters, binary action"},
            "roofline": {"bound": "tensor", "kernel": "ppo_table_umma64_kernel (pi + v), rank 0", "achieved": ach,
                         "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak if tf32_peak else None, "traffic": None,
                         "executed_tf32_flops_per_launch": flops, "ms_per_launch": upd, "fp32_equivalent_TFLOPs": ach / 3,
                         "peak_source": "measured bf16 sustained / 2 (tf32)" if "bf16_tflops_sustained" in peaks else "fallback"}}


# ------------------------------------------------------------------------------------------ configs[3]
def ma_sis_block(dev, peaks, rank=0, world=1, N=2048, E=256, T=20, pop=50, h=64, pi_iters=4, v_iters=4, steps=3, warm=2):
    """BASELINE configs[3]: SIS epidemic domain (S = 51, A = 3, C = [0,1,2]) with the MA-RMABPPO nature oracle against
    the Pessimist agent strategy; N = 2048 arms (D = 4N nature parameters), E = 256, T = 20, hidden [64,64], 50
    counterfactual env steps per step, nature nets updated every epoch.  Arm-sharded under torchrun (strong scaling)."""
    from robustrmab_b200.baselines import PessimisticAgentPolicy
    from robustrmab_b200.environments import SISRobustEnv
    from robustrmab_b200.ma_core import RMABLambdaNatureOracle
    from robustrmab_b200.mpi_tools import arm_shard
    from robustrmab_b200.nature_oracle import MARolloutTrainer
    env = SISRobustEnv(N, 0.2 * N, pop, 0, n_envs=E, device=dev)
    env.random_stream.seed(0)
    env.sampled_parameter_ranges = env.sample_parameter_ranges()
    torch.manual_seed(0)
    a0, n_loc = arm_shard(N, rank, world)
    ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, env.action_dim_nature,
                                env, hidden_sizes=[h, h], C=env.C, N=N, B=env.B, one_hot_encode=False, non_ohe_obs_dim=1,
                                state_norm=pop, nature_state_norm=1, n_envs=E, device=dev,
                                arm_shard=(a0, a0 + n_loc) if world > 1 else None)
    tr = MARolloutTrainer(env, ac, T, 0.9, 0.97, 2.0, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, pi_iters, v_iters, 20, 1, 0, 0.5, 0.0, 0,
                          rank=rank, world_size=world)
    pess = PessimisticAgentPolicy(N)
    for _ in range(warm):
        tr.epoch(pess)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    tr.profile = True
    roll_ms = upd_ms = 0.0
    stages = {}
    for _ in range(steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        lamb, h1, s0, lv, lvn, _ = tr.rollout(pess)
        ev[1].record()
        tr.update(lamb, h1, s0, lv, lvn)
        tr.s = tr.env.reset_device()
        ev[2].record()
        torch.cuda.synchronize()
        roll_ms += ev[0].elapsed_time(ev[1]) / steps
        upd_ms += ev[1].elapsed_time(ev[2]) / steps
        for k_, v_ in tr.stages.report().items():
            stages[k_] = stages.get(k_, 0.0) + v_ / steps
    ms = _max_over_ranks(roll_ms + upd_ms, dev, world)
    ok = bool(torch.isfinite(ac.Wpi).all() and torch.isfinite(ac.Wv).all() and torch.isfinite(ac.W1n).all())
    D = 4 * N
    # dominant stage: the agent critics' nature-block products, 2 GEMMs of [T*E, D] x [D, n_loc*h] per Adam iteration
    gemm_flops = 2.0 * v_iters * 2.0 * (T * E) * D * (n_loc * h)
    gemm_ms = stages.get("critics.z1_gemm", 0.0) + stages.get("critics.grad_gemm", 0.0)
    bf16_peak = peaks.get("bf16_tflops_sustained", 2250.0 * 0.6)
    ach = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms else 0.0
    mem = torch.cuda.max_memory_allocated() / 2 ** 30
    info = "library 3xTF32" if tr.library_gemm else "rmab_gemm3h: tcgen05 kind::f16 x 3, TMA, CTA pairs"
    del tr, ac, env
    torch.cuda.empty_cache()
    return {"metric": "MA-RMABPPO arm-steps/sec (rollout incl. 50 counterfactual steps + update)", "unit": UNIT,
            "value": float(N) * E * T / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "n_gpus": world,
            "scaling": "strong" if world > 1 else "single GPU", "finite": ok, "peak_mem_GiB": mem,
            "stage_ms_per_step": {"rollout": roll_ms, "update": upd_ms, **{"update." + k_: v_ for k_, v_ in stages.items()}},
            "dtype": "f32",
            "config": {"workload": f"MA-RMABPPO nature oracle, SIS domain pop={pop} (S={pop + 1}, A=3), N={N} arms, D={D} nature "
                                   f"parameters, {E} envs, T={T}, hidden [{h},{h}], {pi_iters}+{v_iters} iters, nature nets every "
                                   f"epoch, Pessimist agent strategy" + (f", arm-sharded over {world} GPUs" if world > 1 else "")},
            "roofline": {"bound": "tensor", "kernel": f"agent-critic nature-block GEMMs ({info}), rank 0", "achieved": ach,
                         "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak if bf16_peak else None, "traffic": None,
                         "fp32_equivalent_flops_per_step": gemm_flops, "ms_per_step": gemm_ms,
                         "executed_f16_TFLOPs": 3 * ach,
                         "note": "achieved = fp32-equivalent flops (2 M N K per product) of the 2 x v_iters dense products over "
                                 "their stage time (the fp16 hi/lo split kernels of W1n and dz1 included); each product is evaluated "
                                 "as 3 fp16 tensor-core products with fp32 accumulation, so 1/3 of the 16-bit peak is the ceiling "
                                 "of `achieved` and executed_f16_TFLOPs / peak is the tensor-pipe view",
                         "peak_source": "measured bf16 sustained" if "bf16_tflops_sustained" in peaks else "fallback"}}


# ------------------------------------------------------------------------------------------ configs[4]
def _vi_bytes(T, R, Cc, lam, outs):
    return sum(x.numel() * x.element_size() for x in (T, R, Cc, lam)) + sum(x.numel() * x.element_size() for x in outs if x is not None)


def whittle_block(dev, hbm_peak, peak_src, want_cpu=True, N=99_999, L=64):
    """BASELINE configs[4]: batched value iteration over N arms x L lambda-probes (mdptoolbox ValueIteration per probe,
    simulator.py:504-521) and the Whittle index by bisection on the indifference condition (lp_methods.py:181).
    Headline = counterexample domain at eps = 1e-8; eps = 1e-1 (simulator.py:595 default), ARMMAN (S = 3) and SIS pop = 50
    (S = 51, general kernel) tables beside it."""
    from robustrmab_b200.trainer import RMABPPOTrainer
    gamma, rounds = 0.9, 20
    out = {"metric": "Whittle-index arms/sec", "unit": "arms/s", "dtype": "f64", "variants": {}}

    def ce_tables(n):
        p = RMABPPOTrainer.default_ranges(_lib.DOMAIN_CE, n, None).mean(-1)      # midpoints of PARAMETER_RANGES (:790-824)
        T = np.zeros((n, 2, 2, 2))
        T[:, 0] = 0.5
        T[:, 1, 0, 0] = 1.0
        T[:, 1, 1, 1], T[:, 1, 1, 0] = p, 1 - p
        return T, np.tile(np.array([0.0, 1.0]), (n, 1))

    def armman_tables(n):
        rs = np.random.RandomState(0)
        T = rs.rand(n, 3, 2, 3)
        T /= T.sum(-1, keepdims=True)
        return T, np.tile(np.array([1.0, 0.5, 0.0]), (n, 1))

    C = np.array([0.0, 1.0])

    def run(name, T, R, Cc, eps, label):
        n = T.shape[0]
        lam_lim = R.max() / (Cc[Cc > 0].min() * (1 - gamma))            # agent_baselines.py:61
        lambdas = np.linspace(0, lam_lim, L)
        dT, dR, dC, dL = (torch.as_tensor(x, device=dev) for x in (T, R, Cc, lambdas))
        res = ops.vi_batched(dT, dR, dC, dL, gamma, eps)
        ms_vi = _timed(lambda: ops.vi_batched(dT, dR, dC, dL, gamma, eps), reps=3, warm=1)
        b = _vi_bytes(dT, dR, dC, dL, res)
        block = {"workload": label, "ms_vi_batched": ms_vi, "vi_probe_solves_per_s": n * L / (ms_vi * 1e-3),
                 "mean_sweeps_per_probe": float(res[2].double().mean().item()),
                 "roofline": _hbm(b, ms_vi, hbm_peak, peak_src)}
        block["roofline"]["kernel"] = "vi_small_kernel (V, policy, iters, max_iter of every probe written)"
        if Cc.shape[0] == 2:
            ms_wh = _timed(lambda: ops.whittle_bisect(dT, dR, dC, gamma, eps, 0.0, lam_lim, rounds), reps=3, warm=1)
            block.update({"ms_whittle_bisect": ms_wh, "whittle_arms_per_s": n / (ms_wh * 1e-3)})
            bw = dT.numel() * 8 + dR.numel() * 8 + n * T.shape[1] * 8
            block["roofline_whittle"] = _hbm(bw, ms_wh, hbm_peak, peak_src)
            block["roofline_whittle"]["note"] = ("tables read once, one index per (arm, state) written; the solve is ALU / latency "
                                                 "bound (whole MDP in registers, SURVEY 8d), so the HBM fraction is small by construction")
        out["variants"][name] = block
        return block

    Tce, Rce = ce_tables(N)
    head = run("ce_eps1e-8", Tce, Rce, C, 1e-8, f"counterexample, N={N} arms x {L} probes, gamma={gamma}, eps=1e-8")
    run("ce_eps1e-1", Tce, Rce, C, 1e-1, f"counterexample, N={N} arms x {L} probes, eps=1e-1 (simulator.py:595 default)")
    Ta, Ra = armman_tables(N)
    run("armman_eps1e-8", Ta, Ra, C, 1e-8, f"ARMMAN-shaped tables (S=3, A=2), N={N} arms x {L} probes, eps=1e-8")
    # SIS pop = 50: S = 51, A = 3, tables from rmab_sis_table (31 KB per arm in f32, 62 KB in f64) -> the general VI kernel
    n_sis, pop = 4096, 50
    g = torch.Generator(device=dev).manual_seed(0)
    lo = torch.tensor([0.5, 1.0, 1.0, 1.0], device=dev)
    hi = torch.tensor([0.99, 10.0, 10.0, 10.0], device=dev)
    params = (lo + torch.rand((n_sis, 4), device=dev, generator=g) * (hi - lo)).contiguous()
    Tsis = ops.sis_table(pop, params).double().contiguous()
    Rs = torch.linspace(0, 1, pop + 1, device=dev, dtype=torch.float64).repeat(n_sis, 1).contiguous()
    Cs = torch.tensor([0.0, 1.0, 2.0], device=dev, dtype=torch.float64)
    Ls = 16
    lam_s = torch.linspace(0, 10.0, Ls, device=dev, dtype=torch.float64)
    res = ops.vi_batched(Tsis, Rs, Cs, lam_s, gamma, 1e-8)
    ms_s = _timed(lambda: ops.vi_batched(Tsis, Rs, Cs, lam_s, gamma, 1e-8), reps=2, warm=0)
    sweeps = float(res[2].double().mean().item())
    out["variants"]["sis_pop50_eps1e-8"] = {
        "workload": f"SIS pop={pop} (S={pop + 1}, A=3) tables from rmab_sis_table, N={n_sis} arms x {Ls} probes, eps=1e-8",
        "ms_vi_batched": ms_s, "vi_probe_solves_per_s": n_sis * Ls / (ms_s * 1e-3), "mean_sweeps_per_probe": sweeps,
        "roofline": _hbm(_vi_bytes(Tsis, Rs, Cs, lam_s, res), ms_s, hbm_peak, peak_src)}
    del Tsis, res
    out.update({"value": head["whittle_arms_per_s"], "vi_probe_solves_per_s": head["vi_probe_solves_per_s"],
                "ms_vi_batched": head["ms_vi_batched"], "ms_whittle_bisect": head["ms_whittle_bisect"],
                "config": {"workload": head["workload"] + f"; Whittle index by {rounds}-round bisection per (arm, state)"},
                "roofline": head["roofline"]})
    if want_cpu:
        from oracle import vi as ovi        # the CPU leg: the oracle port of mdptoolbox ValueIteration (checker code, timed here)
        n_cpu = 200
        lam_lim = 10.0
        t0 = time.perf_counter()
        ovi.batched_vi(Tce[:n_cpu], Rce[:n_cpu], C, np.linspace(0, lam_lim, L), gamma, 1e-8)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu * L / dt, "unit": "probe solves/s", "cores": 1, "kind": "port",
                               "sample": f"{n_cpu} arms x {L} probes at eps=1e-8 (oracle port of mdptoolbox ValueIteration, {dt:.1f} s)"}
    return out
