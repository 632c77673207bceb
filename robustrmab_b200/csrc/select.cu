// Budget-constrained action selection (act_test).
//
// Reference behaviour: robust_rmab/algos/rmabppo/rmabppo_core.py act_test_deterministic :289-334 (binary:
// action 1 for the top-k arms by p1) and act_test_deterministic_multiaction :338-445 (greedy with hiding
// and sweep restarts).  Tie rule: descending value, ties -> higher arm index (np.argsort stable, reversed).
//   binary: per-env radix select (4 x 8-bit passes over the sortable fp32 key) + one marking pass; one CTA
//           per 32 envs, lane = env, so every pass streams p1[N,E] coalesced.  HBM-bound, ~6 reads of p1.
//   multi : one CTA per env runs the stepwise formulation of the greedy (oracle/select.py::
//           greedy_multiaction_stepwise): each step is a block-wide argmax over the snapshot keys.
#include "common.cuh"

namespace rmab {

__device__ __forceinline__ uint32_t sortable(float x) {
  const uint32_t b = __float_as_uint(x);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ------------------------------------------------------------------------------------------------
constexpr int kSelEnvs = 32;
constexpr int kSelWarps = 8;

__global__ void __launch_bounds__(kSelEnvs* kSelWarps) select_binary_kernel(int E, int N, const float* __restrict__ p1,
                                                                           int k, int positive_only,
                                                                           uint8_t* __restrict__ actions) {
  __shared__ uint32_t hist[kSelEnvs][257];
  __shared__ uint32_t prefix[kSelEnvs];   // key bits fixed so far
  __shared__ uint32_t remain[kSelEnvs];   // how many still to take among keys matching the prefix
  __shared__ uint32_t eqcnt[kSelWarps][kSelEnvs];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int e = blockIdx.x * kSelEnvs + lane;
  const bool live = e < E;
  const int chunk = (N + kSelWarps - 1) / kSelWarps;
  const int n0 = min(N, w * chunk), n1 = min(N, n0 + chunk);
  if (k <= 0 || k >= N) {
    if (live)
      for (int n = n0; n < n1; ++n)
        actions[(int64_t)n * E + e] = (k >= N && (!positive_only || p1[(int64_t)n * E + e] > 0.f)) ? 1 : 0;
    return;
  }
  if (w == 0) { prefix[lane] = 0u; remain[lane] = (uint32_t)k; }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = threadIdx.x; i < kSelEnvs * 257; i += blockDim.x) (&hist[0][0])[i] = 0u;
    __syncthreads();
    const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    const uint32_t pf = prefix[lane];
    if (live)
      for (int n = n0; n < n1; ++n) {
        const uint32_t key = sortable(p1[(int64_t)n * E + e]);
        if ((key & himask) == pf) atomicAdd(&hist[lane][(key >> shift) & 255u], 1u);
      }
    __syncthreads();
    if (w == 0) {
      // walk digits from the top: the k-th largest key has the digit where the running count reaches `remain`
      uint32_t need = remain[lane], run = 0u;
      int dsel = 0;
      for (int dgt = 255; dgt >= 0; --dgt) {
        const uint32_t c = hist[lane][dgt];
        if (run + c >= need) { dsel = dgt; break; }
        run += c;
      }
      prefix[lane] = pf | ((uint32_t)dsel << shift);
      remain[lane] = need - run;
    }
    __syncthreads();
  }
  // prefix = exact key of the k-th largest element; remain = how many of the keys equal to it are taken,
  // highest arm index first.
  const uint32_t tau = prefix[lane], r = remain[lane];
  uint32_t eq = 0u;
  if (live)
    for (int n = n0; n < n1; ++n) eq += (sortable(p1[(int64_t)n * E + e]) == tau) ? 1u : 0u;
  eqcnt[w][lane] = eq;
  __syncthreads();
  uint32_t above = 0u;  // equal keys in higher-index chunks
  for (int ww = w + 1; ww < kSelWarps; ++ww) above += eqcnt[ww][lane];
  if (live)
    for (int n = n1 - 1; n >= n0; --n) {
      const float pv = p1[(int64_t)n * E + e];
      const uint32_t key = sortable(pv);
      uint8_t a = 0;
      if (key > tau) a = 1;
      else if (key == tau) { a = (above < r) ? 1 : 0; ++above; }
      // the multi-action greedy never assigns an action whose probability is exactly 0 (rmabppo_core.py:434)
      if (positive_only && !(pv > 0.f)) a = 0;
      actions[(int64_t)n * E + e] = a;
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int kMultiThreads = 256;
constexpr int kMultiMaxA = 8;

__global__ void __launch_bounds__(kMultiThreads) select_multi_kernel(int E, int N, int A,
                                                                     const float* __restrict__ probs,
                                                                     const float* __restrict__ Cg, double B,
                                                                     uint8_t* __restrict__ actions) {
  extern __shared__ __align__(16) unsigned char sraw[];
  float* pi = reinterpret_cast<float*>(sraw);   // [N][A]
  float* key = pi + (size_t)N * A;               // [N] snapshot of the row max; -1 = visited in this sweep
  uint8_t* act = reinterpret_cast<uint8_t*>(key + N);
  __shared__ unsigned long long wbest[kMultiThreads / 32];
  __shared__ int s_i[4];
  __shared__ double s_spent;
  __shared__ int s_cnt[kMultiMaxA];  // arms currently at each action
  __shared__ float C[kMultiMaxA];
  const int e = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < kMultiMaxA) { C[tid] = tid < A ? Cg[tid] : 0.f; s_cnt[tid] = (tid == 0) ? N : 0; }
  int pos_local = 0;
  for (int i = tid; i < N * A; i += blockDim.x) {
    const float v = probs[(int64_t)i * E + e];  // probs[n, a, e]
    pi[i] = v;
    pos_local += v > 0.f;
  }
  for (int n = tid; n < N; n += blockDim.x) act[n] = 0;
  if (tid == 0) { s_spent = 0.0; s_i[0] = 0; s_i[1] = 0; s_i[2] = 0; }
  __syncthreads();
  // s_i[0] = done flag, s_i[1] = positive entries left, s_i[2] = restart flag
  atomicAdd(&s_i[1], pos_local);
  __syncthreads();

  while (true) {
    // ---- (re)start of a sweep: snapshot the row maxima ----
    if (s_i[0] || !(s_spent < B)) break;
    for (int n = tid; n < N; n += blockDim.x) {
      float m = pi[(size_t)n * A];
      for (int a = 1; a < A; ++a) m = fmaxf(m, pi[(size_t)n * A + a]);
      key[n] = m;
    }
    if (tid == 0) s_i[2] = 0;
    __syncthreads();
    for (int n_vis = 0; n_vis < N; ++n_vis) {
      // block-wide argmax over (key, index), ties -> higher index; visited arms carry key -1
      unsigned long long best = 0ull;
      for (int n = tid; n < N; n += blockDim.x) {
        const unsigned long long c = ((unsigned long long)sortable(key[n]) << 32) | (unsigned)n;
        best = c > best ? c : best;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
      if (lane == 0) wbest[w] = best;
      __syncthreads();
      if (tid == 0) {
        unsigned long long b = wbest[0];
        for (int i = 1; i < kMultiThreads / 32; ++i) b = wbest[i] > b ? wbest[i] : b;
        const int arm = (int)(b & 0xFFFFFFFFull);
        float* row = pi + (size_t)arm * A;
        int arm_a = 0;
        float mx = row[0];
        for (int a = 1; a < A; ++a)
          if (row[a] >= mx) { mx = row[a]; arm_a = a; }  // stable argsort -> last maximum
        const int old = act[arm];
        const double spent = s_spent;
        if (spent + (double)C[arm_a] - (double)C[old] > B) {
          s_i[0] = 1;  // the visited arm's choice is unaffordable: stop everything (:396-399)
        } else {
          key[arm] = -1.f;
          act[arm] = (uint8_t)arm_a;
          int hidden = 0;
          for (int a = 0; a <= arm_a; ++a) { hidden += row[a] > 0.f; row[a] = 0.f; }
          s_i[1] -= hidden;
          s_cnt[old] -= 1;
          s_cnt[arm_a] += 1;
          const double ns = spent + (double)C[arm_a] - (double)C[old];
          s_spent = ns;
          int amin = 0;
          while (amin < A - 1 && s_cnt[amin] == 0) ++amin;
          s_i[2] = ((double)C[A - 1] - (double)C[amin] > B - ns) ? 1 : 0;
        }
      }
      __syncthreads();
      if (s_i[0]) break;
      if (s_i[2]) {
        // hide every action that no longer fits (:410-426) and restart the sweep
        const double room = B - s_spent;
        int hid = 0;
        for (int i = tid; i < N * A; i += blockDim.x) {
          const int n = i / A, a = i - n * A;
          if ((double)C[a] - (double)C[act[n]] > room) { hid += pi[i] > 0.f; pi[i] = 0.f; }
        }
        if (hid) atomicSub(&s_i[1], hid);
        __syncthreads();
        if (tid == 0 && s_i[1] <= 0) s_i[0] = 1;  // `if not pi_list.sum() > 0` (:434)
        __syncthreads();
        break;  // restart: new snapshot, visited flags cleared
      }
      if (s_i[1] <= 0) {
        __syncthreads();
        if (tid == 0) s_i[0] = 1;
        __syncthreads();
        break;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int n = tid; n < N; n += blockDim.x) actions[(int64_t)n * E + e] = act[n];
}

}  // namespace rmab

using namespace rmab;

// ---- exact multiple-choice knapsack (lp_methods.py:221-271, the ILP behind HawkinsAgentPolicy.act_test) ----
// maximise sum_i values[i, a_i]  s.t.  sum_i C[a_i] == B, one action per arm; integer costs.  One CTA per env runs the
// textbook DP over arms, f_i[b] = max_a f_{i-1}[b - C[a]] + values[i, a] (ties -> lowest action), all budgets b in
// parallel; the argmax table [N, B+1] (caller workspace) is walked back by one thread.  fp64 like the values it gets.
constexpr int kKnapThreads = 256;
constexpr int kKnapMaxA = 8;
constexpr uint8_t kKnapNone = 255;

__global__ void __launch_bounds__(kKnapThreads) knapsack_mc_kernel(int E, int N, int A, const double* __restrict__ values,
                                                                    const int* __restrict__ costs, int B,
                                                                    uint8_t* __restrict__ actions, int* __restrict__ status,
                                                                    uint8_t* __restrict__ choice_ws) {
  extern __shared__ __align__(16) double knap_smem[];
  __shared__ int c_s[kKnapMaxA];
  __shared__ double v_s[kKnapMaxA];
  const int e = blockIdx.x, tid = threadIdx.x;
  const int W = B + 1;
  double* f_prev = knap_smem;
  double* f_cur = knap_smem + W;
  uint8_t* choice = choice_ws + (size_t)e * N * W;
  const double NEG = -1.0e300;
  if (tid < A) c_s[tid] = costs[tid];
  for (int b = tid; b < W; b += kKnapThreads) f_prev[b] = b == 0 ? 0.0 : NEG;
  __syncthreads();
  for (int i = 0; i < N; ++i) {
    if (tid < A) v_s[tid] = values[((size_t)i * A + tid) * E + e];
    __syncthreads();
    for (int b = tid; b < W; b += kKnapThreads) {
      double best = NEG;
      uint8_t arg = kKnapNone;
      for (int a = 0; a < A; ++a) {
        const int c = c_s[a];
        if (b >= c) {
          const double prev = f_prev[b - c];
          if (prev > NEG) {
            const double cand = prev + v_s[a];
            if (cand > best) { best = cand; arg = (uint8_t)a; }
          }
        }
      }
      f_cur[b] = best;
      choice[(size_t)i * W + b] = arg;
    }
    __syncthreads();
    double* t = f_prev; f_prev = f_cur; f_cur = t;
  }
  if (tid == 0) {
    const bool feasible = N == 0 ? B == 0 : f_prev[B] > NEG;
    status[e] = feasible ? 0 : 1;
    int b = B;
    for (int i = N - 1; i >= 0; --i) {
      const uint8_t a = feasible ? choice[(size_t)i * W + b] : 0;
      actions[(size_t)i * E + e] = a;
      if (feasible) b -= c_s[a];
    }
  }
}

extern "C" size_t rmab_select_ws_bytes(int E, int N, int A) {
  (void)E; (void)N; (void)A;
  return 16;  // both kernels keep their working set in shared memory; a token workspace keeps the ABI stable
}

extern "C" int rmab_budget_select_binary(int E, int N, const float* p1, float cost1, float B, int positive_only,
                                         uint8_t* actions, void* ws, size_t ws_bytes, rmab_stream_t stream) {
  (void)ws; (void)ws_bytes;
  RMAB_REQUIRE(E > 0 && N >= 0 && p1 && actions, RMAB_E_BADARG, "rmab_budget_select_binary: bad argument");
  if (N == 0) return RMAB_OK;
  // number of ones exactly as the while-loop of rmabppo_core.py:320-331 counts them
  int k = 0;
  double spent = 0.0;
  while (spent < (double)B && k < N) {
    if (spent + (double)cost1 > (double)B) break;
    spent += (double)cost1;
    ++k;
  }
  select_binary_kernel<<<(E + kSelEnvs - 1) / kSelEnvs, kSelEnvs * kSelWarps, 0, (cudaStream_t)stream>>>(
      E, N, p1, k, positive_only, actions);
  return check_launch("rmab_budget_select_binary");
}

extern "C" int rmab_budget_select_multi(int E, int N, int A, const float* probs, const float* C, float B,
                                        uint8_t* actions, void* ws, size_t ws_bytes, rmab_stream_t stream) {
  (void)ws; (void)ws_bytes;
  RMAB_REQUIRE(E > 0 && N >= 0 && A >= 2 && A <= kMultiMaxA && probs && C && actions, RMAB_E_BADARG,
               "rmab_budget_select_multi: bad argument");
  if (N == 0) return RMAB_OK;
  const size_t smem = (size_t)N * A * 4 + (size_t)N * 4 + (((size_t)N + 15) & ~(size_t)15);
  RMAB_REQUIRE(smem <= 200 * 1024, RMAB_E_SHAPE,
               "rmab_budget_select_multi: N=%d x A=%d does not fit the per-env shared-memory working set", N, A);
  if (smem > 48 * 1024)
    RMAB_REQUIRE(cudaFuncSetAttribute(select_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) ==
                     cudaSuccess,
                 RMAB_E_LAUNCH, "rmab_budget_select_multi: cannot reserve %zu B smem", smem);
  select_multi_kernel<<<E, kMultiThreads, smem, (cudaStream_t)stream>>>(E, N, A, probs, C, (double)B, actions);
  return check_launch("rmab_budget_select_multi");
}

extern "C" size_t rmab_knapsack_ws_bytes(int E, int N, int B) {
  if (E <= 0 || N <= 0 || B < 0) return 16;
  return (size_t)E * (size_t)N * (size_t)(B + 1);
}

extern "C" int rmab_knapsack_multichoice(int E, int N, int A, const double* values, const int32_t* costs, int B,
                                         uint8_t* actions, int32_t* status, void* ws, size_t ws_bytes,
                                         rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N >= 0 && A >= 1 && A <= kKnapMaxA && B >= 0 && values && costs && actions && status, RMAB_E_BADARG,
               "rmab_knapsack_multichoice: bad argument");
  if (N == 0) return RMAB_OK;
  RMAB_REQUIRE(ws && ws_bytes >= rmab_knapsack_ws_bytes(E, N, B), RMAB_E_BADARG,
               "rmab_knapsack_multichoice: workspace of %zu B needed (rmab_knapsack_ws_bytes)", rmab_knapsack_ws_bytes(E, N, B));
  const size_t smem = 2 * (size_t)(B + 1) * sizeof(double);
  RMAB_REQUIRE(smem <= 200 * 1024, RMAB_E_SHAPE, "rmab_knapsack_multichoice: budget %d exceeds the shared-memory DP rows", B);
  if (smem > 48 * 1024)
    RMAB_REQUIRE(cudaFuncSetAttribute(knapsack_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess,
                 RMAB_E_LAUNCH, "rmab_knapsack_multichoice: cannot reserve %zu B smem", smem);
  knapsack_mc_kernel<<<E, kKnapThreads, smem, (cudaStream_t)stream>>>(E, N, A, values, costs, B, actions, status,
                                                                      static_cast<uint8_t*>(ws));
  return check_launch("rmab_knapsack_multichoice");
}
