/*
 * rmab.h -- C ABI of librmab_b200.so: the B200 (sm_100a) kernels behind the RobustRMAB hot path.
 *
 * The reference (killian-34/RobustRMAB) is pure Python and has no FFI layer; each entry point
 * below replaces the BODY of one reference function and is what a ctypes binding inside that
 * function would call (see INTEGRATION.md for the stubs).  Reference file:line per entry point.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller (torch tensors); nothing is allocated
 *     or freed inside.  Workspaces are caller-provided and sized by the *_ws_bytes helpers.
 *   - Every call only enqueues work on `stream` (a cudaStream_t passed as void*) and returns.
 *   - Return 0 on success, <0 on error (RMAB_E_*); never throws.  rmab_last_error() gives the
 *     thread-local message of the last failing call.
 *   - Layout is arm-major, env fastest: states [N,E], trajectories [N,T,E], packed per-arm weights
 *     [N,P] with P = (in+1)h + (h+1)h + (h+1)out in torch.nn.Linear order W1 b1 W2 b2 W3 b3.
 *   - N is the number of arms held by THIS rank; `arm0` is the global index of its first arm so the
 *     Philox counters (stream, t, env, global_arm) give the same draws under any arm sharding.
 *   - RNG: Philox4x32-10, key = seed, u = (word >> 8) * 2^-24.  Streams ENV and ACT share blocks: the transition
 *     and action draws of time step t are words 2*(t&1) and 2*(t&1)+1 of the block with counter (0, t>>1, env,
 *     arm); every other stream uses word 0 of the block (stream, t, env, arm).  Categorical draw
 *     k = #{j: cdf_j <= u} with the cdf accumulated left to right in fp32, clamped to the last category with p > 0.
 *     If `u_inject` is non-NULL it supplies the uniforms ([N,E]) instead (parity tests).
 */
#ifndef RMAB_H_
#define RMAB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMAB_ABI_VERSION 1

typedef void* rmab_stream_t; /* cudaStream_t */

enum { RMAB_OK = 0, RMAB_E_BADARG = -1, RMAB_E_SHAPE = -2, RMAB_E_ARCH = -3, RMAB_E_LAUNCH = -4 };
enum { RMAB_DOMAIN_CE = 0, RMAB_DOMAIN_ARMMAN = 1, RMAB_DOMAIN_SIS = 2 };
enum { RMAB_STREAM_ENV = 0, RMAB_STREAM_ACT = 1, RMAB_STREAM_NATURE = 2, RMAB_STREAM_RESET = 3,
       RMAB_STREAM_CF = 4, RMAB_STREAM_INIT = 7 };
/* mdptoolbox stop_criterion 'full' / 'fast' (both also stop at the _boundIter iteration bound, which leaves V
 * span-converged only), and ABS: iterate until the sup-norm change is below thresh, ignoring the bound -- the
 * absolutely converged value function that the Hawkins objective (lp_methods.py:63) needs. */
enum { RMAB_STOP_FULL = 0, RMAB_STOP_FAST = 1, RMAB_STOP_ABS = 2 };

/* Philox counter block shared by the sampling kernels. */
typedef struct RmabRng {
  uint64_t seed;   /* Philox key */
  uint32_t stream; /* RMAB_STREAM_* */
  uint32_t t;      /* time counter */
  uint32_t env0;   /* global index of the first env in the call */
  uint32_t arm0;   /* global index of the first arm in the call */
  const uint32_t* t_base; /* optional DEVICE word added to t when the kernel runs (NULL = 0): lets a captured CUDA graph be
                             replayed with advancing counters */
} RmabRng;

/* Shape of the N per-arm actor / critic nets (rmabppo_core.py:131-174). Two hidden layers of equal
 * width `hidden` (8, 16, 32 or 64), tanh.  in_dim = (one_hot ? n_states : 1) + 1 (+ n_extra). */
typedef struct RmabNetDesc {
  int32_t n_arms;    /* N on this rank */
  int32_t n_envs;    /* E */
  int32_t n_states;  /* S */
  int32_t n_actions; /* A (actor outputs) */
  int32_t hidden;    /* h */
  int32_t one_hot;   /* 1: x = onehot(s) ++ [lambda]; 0: x = [s/state_norm, lambda] */
  float state_norm;  /* used when one_hot == 0 */
  int32_t n_extra;   /* extra critic inputs (MA-RMABPPO bounded nature vector); 0 for RMABPPO */
} RmabNetDesc;

/* PPO / Adam hyper-parameters of one update() call (agent_oracle.py:233-242, :383-388). */
typedef struct RmabHyper {
  double clip_ratio;
  double entropy_coeff;
  double pi_lr, vf_lr;
  double beta1, beta2, adam_eps; /* torch.optim.Adam defaults: 0.9, 0.999, 1e-8 */
  int32_t pi_iters, v_iters;
  int32_t adam_step0_pi, adam_step0_v; /* Adam steps already taken (optimisers persist across epochs) */
} RmabHyper;

int rmab_abi_version(void);
const char* rmab_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t rmab_launch_count(void);

/* ---- environment ------------------------------------------------------------------------- */

/* env.reset(): uniform random start states (bandit_env_robust.py:941-944, :715-718, :1285-1290).
 * s[N,E] = min(floor(u*S), S-1). */
int rmab_reset_states(int E, int N, int S, const RmabRng* rng, const float* u_inject, uint8_t* s,
                      rmab_stream_t stream);

/* CounterExampleRobustEnv.step (bandit_env_robust.py:860-893) / ARMMANRobustEnv.step (:575-634).
 *   nature_per_env == 0: per-arm table  CE [N]      ARMMAN [N,S,A] (looked up at the current state,
 *                        nature_baselines_armman.py:224-232)
 *   nature_per_env == 1: per-env action CE [N,E]    ARMMAN [N,A,E]
 *   ranges: CE [N,2], ARMMAN [N,S,A,2] (sampled_parameter_ranges; the clamp of :867-874 / :583-592)
 *   R [N,S]; r may be NULL.  Reward is R[arm, s_next]. */
int rmab_env_step_table(int domain, int E, int N, const uint8_t* s, const uint8_t* a,
                        const float* nature, int nature_per_env, const float* ranges, const float* R,
                        const RmabRng* rng, const float* u_inject, uint8_t* s_next, float* r,
                        rmab_stream_t stream);

/* SISRobustEnv.step (bandit_env_robust.py:1224-1243) with compute_distro (:1025-1084) evaluated on
 * the fly in fp64.  params [N,4] (nature_per_env == 0) or [N,4,E]; pop <= 255. */
int rmab_env_step_sis(int E, int N, int pop, const uint8_t* s, const uint8_t* a, const float* params,
                      int nature_per_env, const float* R, const RmabRng* rng, const float* u_inject,
                      uint8_t* s_next, float* r, rmab_stream_t stream);

/* SISRobustEnv.get_T_for_a_nature (:1247-1256): full table T[N,S,A,S] (fp32) for params [N,4]. */
int rmab_sis_table(int N, int pop, const float* params, float* T, rmab_stream_t stream);

/* env.bound_nature_actions (:703, :930, :1274): out = ((tanh(x)+1)/2)*(hi-lo)+lo, elementwise fp32.
 * lo/hi already gathered to the shape of x (n elements). */
int rmab_bound_nature(int64_t n, const float* x, const float* lo, const float* hi, float* out,
                      rmab_stream_t stream);

/* ---- actor-critic ------------------------------------------------------------------------ */

/* MLPActorCriticRMAB.step (rmabppo_core.py:198-244) for all arms x envs.
 * Wpi [N,Ppi], Wv [N,Pv]; s [N,E]; lamb [E].
 * extra (NULL unless net->n_extra > 0): the MA-RMABPPO agent critics also read the bounded nature action vector
 * (ma_rmabppo_core.py:312-316, critic input dim = in + n_extra).  Their first layer splits into the packed part
 * W1[h,in] and W1_nature[h,n_extra]; the caller computes the nature part for all arms and envs as ONE dense GEMM
 * ([N*h, n_extra] x [n_extra, E], a plain library GEMM) and passes the result extra[N,h,E], which the kernel adds to
 * the critic's first-layer pre-activation.
 * Outputs a [N,E] u8, v/logp/p1 [N,E] fp32, probs [N,A,E] fp32 (may be NULL). */
int rmab_ac_step(const RmabNetDesc* net, const float* Wpi, const float* Wv, const uint8_t* s,
                 const float* lamb, const float* extra, const RmabRng* rng, const float* u_inject,
                 uint8_t* a, float* v, float* logp, float* p1, float* probs, rmab_stream_t stream);

/* MLPLambdaNet forward (rmabppo_core.py:108-115; N -> 8 -> 8 -> 1): lamb[E] from states s[N,E]
 * (obs = s * obs_scale).  params packed W1[8,N] b1[8] W2[8,8] b2[8] W3[1,8] b3[1].
 * h1_partial (may be NULL) receives the first-layer pre-activation partial sums [E,8] WITHOUT bias so
 * that arm-sharded ranks can all-reduce them; when sharded, pass h1_in = the reduced sums instead of s. */
int rmab_lambda_forward(int E, int N, const float* params, int n_total, int arm0, const uint8_t* s,
                        float obs_scale, const float* h1_in, float* h1_partial, float* lamb, void* ws, size_t ws_bytes,
                        rmab_stream_t stream);
/* Optional workspace of rmab_lambda_forward (ws may be NULL): with it the first layer is computed by arm splits in
 * parallel and summed in split order. */
size_t rmab_lambda_ws_bytes(int E, int N);

/* SGD step on the lambda net for loss = mean_e lamb_e * coef_e  (compute_loss_lambda,
 * agent_oracle.py:360-376, :440-454; coef_e = B/(1-gamma) - cdcost_e[0]; the mean over envs is what
 * mpi_avg_grads does over ranks).  In place on params.  h1 [E,8] = the (all-reduced) first-layer sums
 * that rmab_lambda_forward wrote to h1_partial for the same states s; delta1_ws [E,8] scratch;
 * grad_out [P] (may be NULL) receives the gradient (W1 columns of this rank's arms only). */
int rmab_lambda_sgd(int E, int N, float* params, int n_total, int arm0, const uint8_t* s, float obs_scale,
                    const float* h1, const float* coef, float lr, float* delta1_ws, float* grad_out,
                    rmab_stream_t stream);

/* Fused rollout of T steps for the table domains with a per-arm nature table (the baseline nature
 * policies): actor tables are evaluated once per (arm, env, state), then each (arm, env) chain runs
 * T steps of  a ~ pi(.|s, lamb_e) ; s' ~ T(s, a, nature)  from Philox streams ACT / ENV with time
 * counters t_act0 + t and t_env0 + t.  s_traj [N,T+1,E] (row 0 = start state on input), a/logp/val
 * [N,T,E], last_val [N,E] = V(s_T). */
int rmab_rollout_table(int domain, const RmabNetDesc* net, int T, const float* Wpi, const float* Wv,
                       const float* lamb, const float* nature, const float* ranges, uint64_t seed,
                       uint32_t t_act0, uint32_t t_env0, uint32_t env0, uint32_t arm0, uint8_t* s_traj,
                       uint8_t* a, float* logp, float* val, float* last_val, rmab_stream_t stream);

/* ---- MA-RMABPPO rollout pieces ------------------------------------------------------------ */

/* Gaussian nature actor sample (ma_rmabppo_core.py:289-291, MLPGaussianActor :98-113): a = mu + exp(log_std) * eps,
 * logp[e] = sum_d Normal(mu, exp(log_std)).log_prob(a).  mu, a [E,D]; log_std [D]; eps from Philox stream NATURE
 * (Box-Muller on words 0/1 of the block (NATURE, t, env, d/2)) or eps_inject [E,D]. */
int rmab_gaussian_sample(int E, int D, const float* mu, const float* log_std, const RmabRng* rng,
                         const float* eps_inject, float* a, float* logp, rmab_stream_t stream);

/* Counterfactual reward estimate of the nature oracle (nature_oracle.py:685-701): from the SAME state s, take
 * n_trials independent env steps with action a and the given nature parameters, and return the mean over trials of
 * the reward R[arm, s'] per (arm, env) in r_mean [N,E] (the caller sums over arms).  Draws: Philox stream CF,
 * time counter rng->t * n_trials + trial.  domain CE / ARMMAN: nature, ranges as rmab_env_step_table;
 * domain SIS: nature = params as rmab_env_step_sis, ranges ignored, S = pop + 1. */
int rmab_env_counterfactual(int domain, int E, int N, int S, const uint8_t* s, const uint8_t* a, const float* nature,
                            int nature_per_env, const float* ranges, const float* R, const RmabRng* rng, int n_trials,
                            float* r_mean, rmab_stream_t stream);

/* One full-batch Adam iteration of the MA-RMABPPO agent critics' value loss (compute_loss_v_agent,
 * nature_oracle.py:389-416; update loop :520-527).  Critic i = MLP([x_i, bounded nature vector] -> h -> h -> 1); its
 * first layer splits into the packed block W1[h,in] (in Wv, as rmab_ppo_update) and W1_nature[h,D], which the caller
 * applies / differentiates with dense GEMMs:
 *   z1_extra[n, m, :] = W1_nature[n] . nature_bounded[m]      (input,  [N, T*E, h])
 *   dz1_out [n, m, :] = dLoss_n / d(first-layer pre-activation) (output, [N, T*E, h]) -> dW1_nature[n] = dz1^T . nature
 * The kernel does the forward / backward of everything else and the Adam step of the packed block (step number
 * hp->adam_step0_v + 1).  s [N,s_rows,E] u8 (first T rows), ret [N,T,E], lamb [E], loss_v [N] (may be NULL).
 * z1_sample_major != 0: both buffers are [T*E, N, h] -- the row-major result of ONE dense GEMM over all arms,
 * nature[T*E, D] . W1_nature[N*h, D]^T, instead of N batched ones. */
int rmab_critic_extra_iter(const RmabNetDesc* net, const RmabHyper* hp, int T, int s_rows, float* Wv, float* adam_v,
                           const uint8_t* s, const float* ret, const float* lamb, const float* z1_extra, float* dz1_out,
                           int z1_sample_major, float* loss_v, rmab_stream_t stream);

/* Elementwise Adam step (torch.optim.Adam defaults form) on a flat fp32 tensor; `step` is 1-based. */
int rmab_adam_step(int64_t n, float* p, const float* grad, float* m, float* v, double lr, double beta1, double beta2,
                   double eps, int step, rmab_stream_t stream);

/* ---- buffer ------------------------------------------------------------------------------ */

/* RMABPPO_Buffer.finish_path + get (agent_oracle.py:90-161): GAE-lambda advantages, rewards-to-go
 * and per-arm advantage normalisation, one CTA per arm, reverse scan over T per env.
 *   s_traj [N,T+1,E] (reward r_t = R[arm, s_{t+1}]), a [N,T,E], val [N,T,E], last_val [N,E], lamb [E],
 *   C [A], R [N,S].  Outputs adv (normalised when normalize != 0), ret, q (may be NULL) [N,T,E];
 *   adv_stats [N,2] = (mean, std) (may be NULL).  Recurrences run in fp64 like scipy's lfilter. */
int rmab_gae(int T, int E, int N, int S, int A, const uint8_t* s_traj, const uint8_t* a, const float* val,
             const float* last_val, const float* lamb, const float* C, const float* R, double gamma,
             double lam, int normalize, float* adv, float* ret, float* q, float* adv_stats,
             rmab_stream_t stream);

/* Same with explicit reward / cost arrays (the drop-in RMABPPO_Buffer.store path): rew, cost [N,T,E]. */
int rmab_gae_explicit(int T, int E, int N, const float* rew, const float* cost, const float* val,
                      const float* last_val, const float* lamb, double gamma, double lam, int normalize,
                      float* adv, float* ret, float* q, float* adv_stats, rmab_stream_t stream);

/* cdcost_buf (agent_oracle.py:117,:139): arm-summed, discounted cost-to-go [T,E] of this rank's arms.
 * ws: T*E doubles of scratch. */
int rmab_cost_togo(int T, int E, int N, int A, const uint8_t* a, const float* C, double gamma, double* ws,
                   float* cdcost, rmab_stream_t stream);

/* ---- PPO update -------------------------------------------------------------------------- */

/* update() (agent_oracle.py:399-436) for all arms: pi_iters full-batch Adam steps on the clipped
 * policy loss (compute_loss_pi :285-322) then v_iters on the value loss (compute_loss_v :325-339).
 * One CTA per arm; weights, gradients and Adam moments stay in shared memory across iterations.
 *   s, a [N,T,E] u8 (s = first T rows of s_traj); adv, logp_old, ret [N,T,E]; lamb [E];
 *   adam_pi [N,2,Ppi], adam_v [N,2,Pv] (m then v); loss_pi [N,pi_iters], loss_v [N,v_iters] (may be NULL)
 *   hold the loss evaluated before each step.  s_stride = row stride of s in elements of E per arm
 *   (T+1 when s aliases s_traj, T otherwise). */
int rmab_ppo_update(const RmabNetDesc* net, const RmabHyper* hp, int T, int s_rows, float* Wpi, float* Wv,
                    float* adam_pi, float* adam_v, const uint8_t* s, const uint8_t* a, const float* adv,
                    const float* logp_old, const float* ret, const float* lamb, float* loss_pi,
                    float* loss_v, rmab_stream_t stream);

/* ---- fused epoch for the one-hot table domains ------------------------------------------- */

/* Number of statistics rows K = 4*S*A + 3*S of rmab_epoch_table for a table domain (-1 otherwise).
 * Row order (c = s*A + a):  Aplus[c] | Aminus[c] | logp_old[c] | n[s] | mean_ret[s] | v_old[s] | n[c]. */
int rmab_stat_rows(int domain);
size_t rmab_epoch_ws_bytes(int N, int E);

/* One epoch's rollout + finish_path + get() for all arms x envs in one kernel (agent_oracle.py:506-570,
 * :90-161) and the per-(arm, env, state, action) sufficient statistics of compute_loss_pi / compute_loss_v
 * (:285-339) that rmab_ppo_update_table consumes:
 *   Aplus / Aminus = sums of the positive / negative normalised advantages of the samples that visited
 *   (s, a); logp_old = log pi_old(a | s, lambda_e); n = visit counts; mean_ret = mean reward-to-go per state.
 * Inputs as rmab_rollout_table plus R [N,S], C [A], gamma, lam (GAE lambda).  s_traj [N,T+1,E] (row 0 =
 * start state on input) and a [N,T,E] are written; stats [N,K,E]; arm_stats [N,4] = (adv mean, adv std,
 * sum (ret - mean_ret[s,e])^2, 0) (may be NULL); env_sums [2,E] = per-env sum over arms and steps of the
 * reward, and of the discounted cost-to-go at t = 0 (cdcost_buf[0], :117/:139) (may be NULL; needs ws of
 * rmab_epoch_ws_bytes); adv (normalised), ret [N,T,E] and last_val [N,E] are written when non-NULL.
 * a == NULL: the trajectory is not materialised -- it lives in shared memory as bit planes for the kernel's own two
 * backward passes -- and s_traj is then just the start states [N,E] (read only).  Needs (T/32 + 1) * (S == 2 ? 2 : 3)
 * * E * 4 bytes of shared memory beside the staged nets (fails with RMAB_E_BADARG otherwise). */
int rmab_epoch_table(int domain, const RmabNetDesc* net, int T, const float* Wpi, const float* Wv,
                     const float* lamb, const float* nature, const float* ranges, const float* R, const float* C,
                     double gamma, double lam, uint64_t seed, uint32_t t_act0, uint32_t t_env0, uint32_t env0,
                     uint32_t arm0, uint8_t* s_traj, uint8_t* a, float* stats, float* arm_stats, float* env_sums,
                     void* ws, size_t ws_bytes, float* adv, float* ret, float* last_val, rmab_stream_t stream);

/* rmab_epoch_table for ENV-sharded ranks (the reference's MPI layout, mpi_tools.py: every rank all arms, a slice of the
 * envs given by env0): get() normalises an arm's advantages with mean / std over the envs of ALL ranks
 * (mpi_statistics_scalar, agent_oracle.py:152-155).  First launch: adv_moments_out [N,2] fp64 receives this rank's raw sums
 * (sum adv, sum adv^2; the statistics written are to be ignored); the caller all-reduces them; second launch (same seeds =
 * same trajectory): adv_moments_in [N,2] = global (mean, std) used for the statistics.  Exactly one of the two is given. */
int rmab_epoch_table_envshard(int domain, const RmabNetDesc* net, int T, const float* Wpi, const float* Wv,
                              const float* lamb, const float* nature, const float* ranges, const float* R, const float* C,
                              double gamma, double lam, uint64_t seed, uint32_t t_act0, uint32_t t_env0, uint32_t env0,
                              uint32_t arm0, uint8_t* s_traj, uint8_t* a, float* stats, float* arm_stats, float* env_sums,
                              void* ws, size_t ws_bytes, float* last_val, double* adv_moments_out,
                              const float* adv_moments_in, rmab_stream_t stream);

/* update() (agent_oracle.py:399-436) on the statistics of rmab_epoch_table: the same pi_iters + v_iters
 * full-batch Adam steps as rmab_ppo_update, evaluated on S*E rows per arm instead of T*E samples (exact
 * regrouping of the loss sums; see csrc/epoch.cu).  T only enters as the 1/(T*E) mean. */
int rmab_ppo_update_table(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv,
                          float* adam_pi, float* adam_v, const float* stats, const float* arm_stats,
                          const float* lamb, float* loss_pi, float* loss_v, rmab_stream_t stream);

/* ---- dense products of the MA-RMABPPO agent critics' nature block ------------------------ */

/* Every agent critic reads the whole bounded nature vector (ma_rmabppo_core.py:203), so the first layers of all N
 * critics over the D nature inputs are ONE dense product per batch, z1 = nat . W1n^T (nature_oracle.py:388-416), and
 * its weight gradient is dW1n = dz1^T . nat (:505-513).  These entry points evaluate such products with fp32 accuracy
 * on the fp16 tensor-core path (csrc/gemm3h.cu): rmab_split16_* write x = rs * (hi + lo) with per-row power-of-two
 * scales (hi, lo fp16, 22 significant bits), rmab_gemm3h computes
 *     C[m, n] = rsA[m] * rsB[n] * sum_k (Ah Bh + Al Bh + Ah Bl)[m, n]       (fp32 accumulation)
 * for K-major operands Ah, Al [M, ldk] and Bh, Bl [N, ldk] (fp16, ldk % 8 == 0, 16-byte aligned), C [M, N] fp32 with
 * row stride ldc.  n_fast = 1 walks consecutive tiles along N (A block reused), 0 along M (B block reused).
 * rmab_split16_rows: X [R, K] (row stride ldx) -> Xh, Xl [R, ldk], rs [R].
 * rmab_split16_cols: X [Rk, Cn] -> the TRANSPOSED split Xh, Xl [Cn, ldk] (k = row of X), rs [Cn]; amax_ws: Cn words. */
int rmab_split16_rows(int R, int K, const float* X, int64_t ldx, int ldk, void* Xh, void* Xl, float* rs, rmab_stream_t stream);
int rmab_split16_cols(int Rk, int Cn, const float* X, int64_t ldx, int ldk, void* Xh, void* Xl, float* rs, uint32_t* amax_ws,
                      rmab_stream_t stream);
int rmab_gemm3h(int M, int N, int K, int ldk, const void* Ah, const void* Al, const float* rsA, const void* Bh, const void* Bl,
                const float* rsB, float* C, int64_t ldc, int n_fast, rmab_stream_t stream);
/* The same product as a GRADIENT consumed in the epilogue: P [M, N] (row stride ldc) is the parameter tensor, Mo / Vo its
 * Adam moments; every element gets exactly rmab_adam_step's update with g = the product, under the next tile's main loop
 * (dW1n = dz1^T . nat followed by the Adam step of nature_oracle.py:520-527, without materialising dW1n). */
int rmab_gemm3h_adam(int M, int N, int K, int ldk, const void* Ah, const void* Al, const float* rsA, const void* Bh, const void* Bl,
                     const float* rsB, float* P, float* Mo, float* Vo, int64_t ldc, double lr, double beta1, double beta2,
                     double eps, int step, int n_fast, rmab_stream_t stream);

/* ---- the dense nature nets of MA-RMABPPO ------------------------------------------------- */

/* Gaussian nature actor N -> h -> h -> D (ma_rmabppo_core.py:84-113) and nature critic 2N -> h -> h -> 1 (:116-123) as
 * tiled fp32 kernels (csrc/nature.cu); losses and updates follow nature_oracle.py:359-386, :421-428, :558-576.
 * Packed parameters: [W1T (Nin x H) | b1 | W2 (H x H, [out][in]) | b2 | W3 (OUT x H) | b3 | log_std]; W1 is TRANSPOSED so
 * that an arm's rows are contiguous.  Layer 1 reads the u8 trajectories directly: x [n][t][e] with arm stride `stride`
 * (elements), sample m = t * E + e, value sc * x; the optional second source (the agent actions of the critic) has its own
 * weight block.  Arm-sharded use: pass the rank's arms and weight rows; l1_fwd / logp / out_bwd.da2 are
 * then partial sums the caller all-reduces.  All reductions are two-stage in a fixed order (deterministic).
 * ws: rmab_nat_ws_bytes(M, H, n_loc, Dl) bytes.  H in {8, 16, 32, 64}. */
size_t rmab_nat_ws_bytes(int M, int H, int n_loc, int Dl);
/* z1 [M,H] = sum_n sc0 x0[n,m] W0[n,:] (+ x1[n,m] W1[n,:]) */
int rmab_nat_l1_fwd(int M, int E, int H, int n_loc, const uint8_t* x0, int64_t stride0, float sc0, const float* W0,
                    const uint8_t* x1, int64_t stride1, const float* W1, float* z1, void* ws, size_t ws_bytes, rmab_stream_t stream);
/* a1 = tanh(z1 + b1), a2 = tanh(W2 a1 + b2) */
int rmab_nat_mid_fwd(int M, int H, const float* z1, const float* b1, const float* W2, const float* b2, float* a1, float* a2,
                     rmab_stream_t stream);
/* mu [M, Dl] (row stride ld_mu) = a2 W3^T + b3 */
int rmab_nat_mu(int M, int H, int Dl, const float* a2, const float* W3, const float* b3, float* mu, int64_t ld_mu, rmab_stream_t stream);
/* logp [M] = sum_d Normal(mu_d, exp(log_std_d)).log_prob(act[m, d]) over the Dl given rows */
int rmab_nat_logp(int M, int H, int Dl, const float* a2, const float* W3, const float* b3, const float* log_std, const float* act,
                  int64_t ld_act, float* logp, void* ws, size_t ws_bytes, rmab_stream_t stream);
/* g [M] = d/dlogp of -mean(min(ratio adv, clip(ratio) adv)), ratio = exp(logp - logp_old) */
int rmab_nat_pi_coef(int M, const float* logp, const float* logp_old, const float* adv, float clip_ratio, float* g, rmab_stream_t stream);
/* backward of the actor's output layer given g: dW3 [Dl,H], db3 [Dl], dlog_std [Dl], da2 [M,H] */
int rmab_nat_out_bwd(int M, int H, int Dl, const float* a2, const float* W3, const float* b3, const float* log_std, const float* act,
                     int64_t ld_act, const float* g, float* dW3, float* db3, float* dlog_std, float* da2, void* ws, size_t ws_bytes,
                     rmab_stream_t stream);
/* d2 = da2 (1 - a2^2), d1 = (W2^T d2)(1 - a1^2); dW2 = d2^T a1, db2 = sum d2, db1 = sum d1 */
int rmab_nat_mid_bwd(int M, int H, const float* da2, const float* a1, const float* a2, const float* W2, float* d1, float* d2,
                     float* dW2, float* db2, float* db1, void* ws, size_t ws_bytes, rmab_stream_t stream);
/* dW0 [n_loc,H] = sum_m sc0 x0[n,m] d1[m,:] (and dW1 for the second source) */
int rmab_nat_l1_bwd(int M, int E, int H, int n_loc, const uint8_t* x0, int64_t stride0, float sc0, const uint8_t* x1, int64_t stride1,
                    const float* d1, float* dW0, float* dW1, void* ws, size_t ws_bytes, rmab_stream_t stream);
/* critic output: v [M] = a2 W3 + b3; with ret: backward of mean((v - ret)^2): da2 [M,H], dW3 [H], db3 [1] */
int rmab_nat_v_out(int M, int H, const float* a2, const float* W3, const float* b3, const float* ret, float* v, float* da2,
                   float* dW3, float* db3, void* ws, size_t ws_bytes, rmab_stream_t stream);

/* ---- budget-constrained selection -------------------------------------------------------- */

size_t rmab_select_ws_bytes(int E, int N, int A);

/* Top-k selection for binary actions: action 1 for the top-k arms by key p1 per env, k = number of C[1]-cost actions
 * that fit in B; ties -> higher arm index.  p1 [N,E] -> actions [N,E].
 *   positive_only = 0: act_test_deterministic (rmabppo_core.py:289-334) and top-B by index (simulator.py:309-325;
 *                      keys may be <= 0).
 *   positive_only = 1: act_test_deterministic_multiaction (:338-445) for C = [0,1], which never assigns an action whose
 *                      probability is exactly 0 (the sweep ends when no positive entry is left, :434), so fewer than k
 *                      arms act if fewer than k have p1 > 0. */
int rmab_budget_select_binary(int E, int N, const float* p1, float cost1, float B, int positive_only, uint8_t* actions,
                              void* ws, size_t ws_bytes, rmab_stream_t stream);

/* The phases of the large-problem path of rmab_budget_select_binary, exposed so that ARM-SHARDED ranks can select a
 * global top-k: per pass (0..3, 8 key bits each, from the top) every rank computes its per-env digit
 * histogram hist [ceil(E/32)][256][32] uint32 (lane = env within the group), the caller all-reduces it, and the scan
 * fixes the next 8 bits of every env's threshold key.  rmab_select_mark then writes the actions; above_ranks [E] =
 * number of keys equal to the threshold held by ranks that own HIGHER arm indices (NULL on one rank), so that ties
 * go to the higher global arm index as np.argsort does in the reference.  ws: rmab_select_ws_bytes(E, N, 2) bytes;
 * its first 3*32*ceil(E/32) words are prefix | remain | eqtot per env.  hist == NULL uses the workspace's own array. */
int rmab_select_hist(int E, int N, const float* p1, int pass, void* ws, size_t ws_bytes, uint32_t* hist, rmab_stream_t stream);
int rmab_select_scan(int E, int N, int pass, int k, const uint32_t* hist, void* ws, size_t ws_bytes, rmab_stream_t stream);
int rmab_select_mark(int E, int N, const float* p1, const uint32_t* above_ranks, int positive_only, uint8_t* actions,
                     void* ws, size_t ws_bytes, rmab_stream_t stream);

/* act_test_deterministic_multiaction (rmabppo_core.py:338-445), one CTA per env, stepwise argmax
 * formulation (oracle/select.py::greedy_multiaction_stepwise).  probs [N,A,E], C [A] strictly
 * increasing. */
int rmab_budget_select_multi(int E, int N, int A, const float* probs, const float* C, float B,
                             uint8_t* actions, void* ws, size_t ws_bytes, rmab_stream_t stream);

/* Exact multiple-choice knapsack = the ILP of lp_methods.py:221-271 behind HawkinsAgentPolicy.act_test
 * (agent_baselines.py:89): maximise sum_i values[i, a_i] subject to sum_i costs[a_i] == B with one action per arm.
 * Integer costs; dynamic programme over arms with all budgets in parallel, one CTA per env; ties -> lowest action.
 * values [N,A,E] fp64, costs [A] int32, actions [N,E]; status [E] = 0, or 1 when the budget cannot be met exactly
 * (the reference's model is infeasible there; actions are then all 0).  ws: rmab_knapsack_ws_bytes(E, N, B) bytes. */
size_t rmab_knapsack_ws_bytes(int E, int N, int B);
int rmab_knapsack_multichoice(int E, int N, int A, const double* values, const int32_t* costs, int B, uint8_t* actions,
                              int32_t* status, void* ws, size_t ws_bytes, rmab_stream_t stream);

/* ---- value iteration / Whittle index ----------------------------------------------------- */

/* mdptoolbox.mdp.ValueIteration (mdp.py:1293-1410 incl. _boundIter) for all arms x lambda-probes, on the
 * per-probe MDP of simulator.py:504-510 with reward R[n,s] - lambda*C[a].  All fp64 like the reference.
 *   T [N,S,A,S], R [N,S], C [A], lambdas [L] (lambdas_per_arm == 0) or [N,L].
 *   V [N,L,S], policy [N,L,S] u8 (may be NULL), iters [N,L] i32 (-1 and V = NaN where the reference raises;
 *   may be NULL), max_iter [N,L] i32 (may be NULL), Q [N,L,S,A] (may be NULL; Q from the final V as
 *   simulator.py:518-521). */
int rmab_vi_batched(int N, int S, int A, int L, const double* T, const double* R, const double* C,
                    const double* lambdas, int lambdas_per_arm, double gamma, double eps, int stop_mode,
                    double* V, uint8_t* policy, int32_t* iters, int32_t* max_iter, double* Q,
                    rmab_stream_t stream);

/* mdptoolbox.mdp.ValueIteration with a general reward: T [N,S,A,S], Rsa [N,S,A] (the matrix MDP._computeReward
 * collapses any accepted reward shape to, mdp.py:292-302), one solve per arm.  V [N,S], policy [N,S] (may be NULL),
 * iters / max_iter [N] (may be NULL).  Used by the drop-in `ValueIteration` class (robustrmab_b200/mdp.py). */
int rmab_vi_general(int N, int S, int A, const double* T, const double* Rsa, double gamma, double eps, int stop_mode,
                    double* V, uint8_t* policy, int32_t* iters, int32_t* max_iter, rmab_stream_t stream);

/* Whittle index by bisection on the indifference condition Q(s,1;l) = Q(s,0;l) (lp_methods.py:181),
 * n_rounds halvings of [lam_lo, lam_hi] per (arm, state), each probe one ValueIteration as above.
 * index [N,S] fp64. */
int rmab_whittle_bisect(int N, int S, int A, const double* T, const double* R, const double* C, double gamma,
                        double eps, int stop_mode, double lam_lo, double lam_hi, int n_rounds, double* index,
                        rmab_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RMAB_H_ */
