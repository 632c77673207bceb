"""CPU oracle for the RobustRMAB hot path -- TEST INFRASTRUCTURE ONLY.

A NumPy / torch-CPU restatement of the reference algorithm (killian-34/RobustRMAB) for the
rollout / GAE / PPO-update / budget-selection / value-iteration path, with injected randomness.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import this package, and only as the checker or the reported CPU baseline.  The
product package `robustrmab_b200` never imports it and has no CPU fallback.

Pinning: every function cites the reference file:line it follows and is checked against golden
vectors generated from the reference's own classes (oracle/gen_golden.py -> tests/golden/*.npz)
plus the mdptoolbox docstring known answers.  Parts whose reference needs Gurobi (Whittle-index
LP, Hawkins LP, knapsack ILP) are marked "parity unpinned" where they appear (oracle/vi.py).

Array layout (shared with the CUDA library): arm-major, env fastest --
states `[N, E]`, trajectories `[N, T, E]`, packed per-arm weights `[N, P]`.
"""
