"""Double-oracle bookkeeping over the device oracles: `DoubleOracle` of /double_oracle.py:51-278 (strategy sets, payoff
matrix, equilibrium of the restricted game) with the same constructor arguments, methods and loop order, so that
BASELINE configs[4] ("double_oracle.py counterexample ...") runs end to end on the drop-ins without nashpy.

Per iteration (`run`, :145-194): agent best response against the nature mixture (`AgentOracle.best_response` = the
RMABPPO loop), nature best response against the agent mixture (`NatureOracle.best_response` = MA-RMABPPO), one
`simulate_reward` per new (agent strategy, nature strategy) pair (:242-278; each call is ONE batched device rollout with
the `n_simu_epochs` evaluation rollouts as env replicas), then the minimax-regret equilibrium of the payoff matrix
(`nfg_solver.solve_minimax_regret`, a zero-sum LP).  The experiment CLI around it (argparse, CSV / plot output,
:317-815) stays in the reference.

New keyword arguments: `n_envs` (env replicas per training rollout; 1 = the reference's shapes) and `device`.
"""
import numpy as np

from .agent_oracle import AgentOracle
from .baselines import PessimisticAgentPolicy, PessimisticNaturePolicy
from .environments import ARMMANRobustEnv, CounterExampleRobustEnv, SISRobustEnv
from .nature_oracle import NatureOracle
from .nfg_solver import get_payoff, solve_minimax_regret, solve_minimax_regret_with_regret_array


class DoubleOracle:
    def __init__(self, data, N, budget, horizon, max_epochs_double_oracle, S=0, A=0, seed=0, reward_bound=1, n_cpu=1,
                 n_simu_epochs=50, home_dir="", exp_name="", gamma=0.99, pop_size=0, one_hot_encode=True, state_norm=1,
                 non_ohe_obs_dim=0, agent_kwargs=dict(), nature_kwargs=dict(), n_envs=1, device="cuda", verbose=False):
        self.data, self.N, self.budget, self.horizon, self.gamma = data, N, budget, horizon, gamma
        self.max_epochs_double_oracle, self.n_simu_epochs = max_epochs_double_oracle, n_simu_epochs
        self.S, self.A, self.seed, self.reward_bound = S, A, seed, reward_bound
        self.home_dir, self.exp_name, self.n_cpu = home_dir, exp_name, n_cpu
        self.pop_size, self.one_hot_encode = pop_size, one_hot_encode
        self.non_ohe_obs_dim, self.state_norm = non_ohe_obs_dim, state_norm
        self.nature_state_norm = 1
        self.verbose = verbose
        if data == 'armman':
            self.env_fn = lambda: ARMMANRobustEnv(N, budget, seed, n_envs=n_envs, device=device)
        elif data == 'counterexample':
            self.env_fn = lambda: CounterExampleRobustEnv(N, budget, seed, n_envs=n_envs, device=device)
        elif data == 'sis':
            self.env_fn = lambda: SISRobustEnv(N, budget, pop_size, seed, n_envs=n_envs, device=device)
        else:
            raise ValueError("data must be 'counterexample', 'armman' or 'sis'")
        self.env = self.env_fn()
        self.sampled_nature_parameter_ranges = self.env.sample_parameter_ranges()
        # the same ranges for every instantiation of the env (:103-105)
        self.env.sampled_parameter_ranges = self.sampled_nature_parameter_ranges
        self.agent_oracle = AgentOracle(data, N, S, A, budget, seed, reward_bound, agent_kwargs=agent_kwargs,
                                        home_dir=home_dir, exp_name=exp_name,
                                        sampled_nature_parameter_ranges=self.sampled_nature_parameter_ranges,
                                        pop_size=pop_size, one_hot_encode=one_hot_encode, state_norm=state_norm,
                                        non_ohe_obs_dim=non_ohe_obs_dim, n_envs=n_envs, device=device)
        self.nature_oracle = NatureOracle(data, N, S, A, budget, seed, reward_bound, nature_kwargs=nature_kwargs,
                                          home_dir=home_dir, exp_name=exp_name,
                                          sampled_nature_parameter_ranges=self.sampled_nature_parameter_ranges,
                                          pop_size=pop_size, one_hot_encode=one_hot_encode, state_norm=state_norm,
                                          non_ohe_obs_dim=non_ohe_obs_dim, nature_state_norm=self.nature_state_norm,
                                          n_envs=n_envs, device=device)
        pess_nature_pol = PessimisticNaturePolicy(self.sampled_nature_parameter_ranges, 0)
        self.agent_strategies = []                    # agent policies
        self.nature_strategies = [pess_nature_pol]    # nature policies
        self.payoffs = []                             # payoffs[i][j] = reward of agent strategy i vs nature strategy j
        self.equilibria = []
        self.update_payoffs_agent(PessimisticAgentPolicy(self.N, 0), 0)

    def _log(self, *a):
        if self.verbose:
            print(*a)

    def run(self):
        agent_eq = np.ones(len(self.agent_strategies)) / len(self.agent_strategies)
        nature_eq = np.ones(len(self.nature_strategies)) / len(self.nature_strategies)
        n_epochs = 0
        while True:
            self._log('epoch', n_epochs, 'n_agent_strategies', len(self.agent_strategies), 'n_nature_strategies',
                      len(self.nature_strategies))
            add_to_seed = 0
            agent_br = self.agent_oracle.best_response(self.nature_strategies, nature_eq, add_to_seed)
            nature_br = self.nature_oracle.best_response(self.agent_strategies, agent_eq, add_to_seed)
            self.update_payoffs(nature_br, agent_br, 0)    # same seeds for every pair's n_simu_epochs rollouts (:171)
            agent_eq, nature_eq = self.find_equilibrium()
            agent_eq, nature_eq = np.asarray(agent_eq, dtype=np.float64), np.asarray(nature_eq, dtype=np.float64)
            max_regret_game = np.array(self.payoffs) - np.array(self.payoffs).max(axis=0)
            self.equilibria.append((agent_eq, nature_eq, get_payoff(max_regret_game, agent_eq, nature_eq)))
            self._log('agent equilibrium', np.round(agent_eq, 3), 'nature equilibrium', np.round(nature_eq, 3),
                      'regret', self.equilibria[-1][2])
            if n_epochs >= self.max_epochs_double_oracle:
                break
            n_epochs += 1
            assert len(self.payoffs) == len(self.agent_strategies)
            assert len(self.payoffs[0]) == len(self.nature_strategies)
        return agent_eq, nature_eq

    def compute_regret(self, agent_s, nature_s, max_reward):
        reward = self.agent_oracle.simulate_reward(agent_s, nature_s, seed=self.seed, steps_per_epoch=self.horizon,
                                                   epochs=self.n_simu_epochs, gamma=self.gamma)
        return max_reward - reward

    def compute_payoff_regret(self, agent_eq):
        """Expected regret in the payoff matrix of a PURE agent strategy (:208-220)."""
        agent_eq = np.asarray(agent_eq)
        assert abs(agent_eq.sum() - 1) <= 1e-3
        regret = np.array(self.payoffs) - np.array(self.payoffs).max(axis=0)
        nz = np.where(agent_eq > 0)[0]
        if len(nz) == 1:
            return -np.min(regret[nz.item()])
        raise Exception('need to implement')

    def find_equilibrium(self, ignore_rows=0, ignore_cols=0):
        payoffs = np.array(self.payoffs)
        if ignore_rows > 0:
            payoffs = payoffs[:-ignore_rows]
        if ignore_cols > 0:
            payoffs = payoffs[:, :-ignore_cols]
        return solve_minimax_regret(payoffs)

    def find_equilibrium_with_regret_array(self, regret_array, ignore_rows=0, ignore_cols=0):
        payoffs = np.copy(regret_array)
        if ignore_rows > 0:
            payoffs = payoffs[:-ignore_rows]
        if ignore_cols > 0:
            payoffs = payoffs[:, :-ignore_cols]
        return solve_minimax_regret_with_regret_array(payoffs)

    def _reward(self, agent_s, nature_s, add_to_seed):
        self._log('Simulating rewards, %s vs. %s' % (agent_s, nature_s))
        return self.agent_oracle.simulate_reward(agent_s, nature_s, seed=self.seed + add_to_seed,
                                                 steps_per_epoch=self.horizon, epochs=self.n_simu_epochs, gamma=self.gamma)

    def update_payoffs(self, nature_br, agent_br, add_to_seed):
        self.update_payoffs_agent(agent_br, add_to_seed)
        self.update_payoffs_nature(nature_br, add_to_seed)

    def update_payoffs_agent(self, agent_br, add_to_seed=0):
        """Append an agent strategy: one reward per nature strategy (:242-259).  Returns its index."""
        self.agent_strategies.append(agent_br)
        self.payoffs.append([self._reward(agent_br, nature_s, add_to_seed) for nature_s in self.nature_strategies])
        return len(self.agent_strategies) - 1

    def update_payoffs_nature(self, nature_br, add_to_seed=0):
        """Append a nature strategy: one reward per agent strategy (:261-278).  Returns its index."""
        self.nature_strategies.append(nature_br)
        for i, agent_s in enumerate(self.agent_strategies):
            self.payoffs[i].append(self._reward(agent_s, nature_br, add_to_seed))
        return len(self.nature_strategies) - 1
