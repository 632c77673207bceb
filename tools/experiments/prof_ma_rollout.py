import sys, json, torch, time
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
from robustrmab_b200.baselines import PessimisticAgentPolicy
from robustrmab_b200.environments import SISRobustEnv
from robustrmab_b200.ma_core import RMABLambdaNatureOracle
from robustrmab_b200.nature_oracle import MARolloutTrainer
dev = torch.device("cuda:0")
N, E, T, pop, h = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 256, 20, 50, 64
env = SISRobustEnv(N, 0.2 * N, pop, 0, n_envs=E, device=dev)
env.random_stream.seed(0)
env.sampled_parameter_ranges = env.sample_parameter_ranges()
torch.manual_seed(0)
ac = RMABLambdaNatureOracle(env.observation_space, env.action_space, env.sampled_parameter_ranges, env.action_dim_nature,
                            env, hidden_sizes=[h, h], C=env.C, N=N, B=env.B, one_hot_encode=False, non_ohe_obs_dim=1,
                            state_norm=pop, nature_state_norm=1, n_envs=E, device=dev)
tr = MARolloutTrainer(env, ac, T, 0.9, 0.97, 2.0, 2e-3, 2e-3, 2e-3, 2e-3, 2e-3, 4, 4, 20, 1, 0, 0.5, 0.0, 0)
pess = PessimisticAgentPolicy(N)
for _ in range(2):
    tr.epoch(pess)
torch.cuda.synchronize()
t0 = time.perf_counter()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    out = tr.rollout(pess)
    torch.cuda.synchronize()
print("rollout wall ms", (time.perf_counter() - t0) * 1e3)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
