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
// Large problems (act_test over 10^5 arms x 10^3 envs): the same 4 x 8-bit radix select spread over the whole GPU.
// Grid = env groups (32 envs, lane = env) x arm chunks; every pass streams p1 once, coalesced, into a per-CTA shared
// histogram [digit][lane] (bank = lane: conflict-free), written out as a partial; a reduce kernel sums the partials of
// an env group, a scan kernel walks the 256 digits per env.  The reduce output is the array ranks all-reduce when the
// arms are sharded.  Marking is one more pass: key > tau -> 1, key == tau -> 1 for the `remain` highest
// arm indices (decided without ordering unless an env really has a tie at the boundary).
// ws (u32 words): prefix[E] remain[E] eqtot[E] | tot[G][256][32] | part[G*Cn][256][32].
constexpr int kSelThreads = 256;
constexpr int kSelTargetCtas = 4 * kNumSMs;

struct SelPlan {
  int G, Cn, chunk;
  size_t tot_off, part_off, words;
};

static SelPlan select_plan(int E, int N) {
  SelPlan p;
  p.G = (E + 31) / 32;
  int cn = kSelTargetCtas / p.G;
  const int max_cn = (N + 255) / 256;
  cn = cn < 1 ? 1 : cn;
  p.Cn = cn > max_cn ? (max_cn < 1 ? 1 : max_cn) : cn;
  p.chunk = (N + p.Cn - 1) / p.Cn;
  p.tot_off = 3 * (size_t)p.G * 32;
  p.part_off = p.tot_off + (size_t)p.G * 8192;
  p.words = p.part_off + (size_t)p.G * p.Cn * 8192;
  return p;
}

__global__ void __launch_bounds__(kSelThreads) select_hist_kernel(int E, int N, const float* __restrict__ p1, int chunk, int pass,
                                                                  const uint32_t* __restrict__ prefix, uint32_t* __restrict__ part) {
  __shared__ uint32_t hist[256][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = blockIdx.x, c = blockIdx.y;
  const int e = g * 32 + lane;
  const bool live = e < E;
  for (int i = threadIdx.x; i < 8192; i += kSelThreads) (&hist[0][0])[i] = 0u;
  __syncthreads();
  const int n0 = min(N, c * chunk), n1 = min(N, n0 + chunk);
  const int shift = 24 - 8 * pass;
  const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
  const uint32_t pf = (live && pass > 0) ? prefix[e] : 0u;
  if (live) {
    const float* col = p1 + e;
    int n = n0 + w;
    for (; n + 3 * 8 < n1; n += 4 * 8) {   // four independent loads in flight per thread
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldg(col + (int64_t)(n + 8 * j) * E);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t key = sortable(v[j]);
        if ((key & himask) == pf) atomicAdd(&hist[(key >> shift) & 255u][lane], 1u);
      }
    }
    for (; n < n1; n += 8) {
      const uint32_t key = sortable(__ldg(col + (int64_t)n * E));
      if ((key & himask) == pf) atomicAdd(&hist[(key >> shift) & 255u][lane], 1u);
    }
  }
  __syncthreads();
  uint32_t* out = part + ((size_t)g * gridDim.y + c) * 8192;
  for (int i = threadIdx.x; i < 8192; i += kSelThreads) out[i] = (&hist[0][0])[i];
}

__global__ void __launch_bounds__(kSelThreads) select_reduce_kernel(int Cn, const uint32_t* __restrict__ part,
                                                                    uint32_t* __restrict__ tot) {
  const int g = blockIdx.x;
  const int i = blockIdx.y * kSelThreads + threadIdx.x;   // 0 .. 8191
  uint32_t s = 0u;
  for (int c = 0; c < Cn; ++c) s += part[((size_t)g * Cn + c) * 8192 + i];
  tot[(size_t)g * 8192 + i] = s;
}

// one warp per env group: lane = env walks the digits from the top
__global__ void __launch_bounds__(32) select_scan_kernel(int E, int pass, uint32_t k, const uint32_t* __restrict__ tot,
                                                         uint32_t* __restrict__ prefix, uint32_t* __restrict__ remain,
                                                         uint32_t* __restrict__ eqtot) {
  const int g = blockIdx.x, lane = threadIdx.x, e = g * 32 + lane;
  if (e >= E) return;
  const uint32_t* h = tot + (size_t)g * 8192 + lane;
  const uint32_t need = pass == 0 ? k : remain[e];
  const uint32_t pf = pass == 0 ? 0u : prefix[e];
  uint32_t run = 0u, cnt = 0u;
  int dsel = 0;
  for (int dgt = 255; dgt >= 0; --dgt) {
    cnt = h[dgt * 32];
    if (run + cnt >= need) { dsel = dgt; break; }
    run += cnt;
  }
  prefix[e] = pf | ((uint32_t)dsel << (24 - 8 * pass));
  remain[e] = need - run;
  if (pass == 3) eqtot[e] = cnt;
}

__global__ void __launch_bounds__(kSelThreads) select_mark_kernel(int E, int N, const float* __restrict__ p1, int chunk,
                                                                  const uint32_t* __restrict__ prefix,
                                                                  const uint32_t* __restrict__ remain,
                                                                  const uint32_t* __restrict__ eqtot,
                                                                  const uint32_t* __restrict__ part,
                                                                  const uint32_t* __restrict__ above_ranks, int positive_only,
                                                                  int only_ties, uint8_t* __restrict__ actions) {
  __shared__ uint32_t eqcnt[kSelThreads / 32][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = blockIdx.x, c = blockIdx.y, Cn = gridDim.y;
  const int e = g * 32 + lane;
  const bool live = e < E;
  const int n0 = min(N, c * chunk), n1 = min(N, n0 + chunk);
  const uint32_t tau = live ? prefix[e] : 0u, r = live ? remain[e] : 0u;
  const bool tie = live && eqtot[e] > r;
  if (!__syncthreads_or(tie ? 1 : 0)) {
    // every key equal to the threshold is taken: no ordering needed
    if (only_ties) return;   // select_mark4_kernel has written this env group
    if (live) {
      const float* col = p1 + e;
      uint8_t* ac = actions + e;
      int n = n0 + w;
      for (; n + 3 * 8 < n1; n += 4 * 8) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldg(col + (int64_t)(n + 8 * j) * E);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint8_t a = sortable(v[j]) >= tau ? 1 : 0;
          if (positive_only && !(v[j] > 0.f)) a = 0;
          ac[(int64_t)(n + 8 * j) * E] = a;
        }
      }
      for (; n < n1; n += 8) {
        const float pv = __ldg(col + (int64_t)n * E);
        uint8_t a = sortable(pv) >= tau ? 1 : 0;
        if (positive_only && !(pv > 0.f)) a = 0;
        ac[(int64_t)n * E] = a;
      }
    }
    return;
  }
  // ordered path: equal keys are taken from the highest (global) arm index down
  const int sub = (n1 - n0 + (kSelThreads / 32) - 1) / (kSelThreads / 32);
  const int s0 = min(n1, n0 + w * sub), s1 = min(n1, s0 + sub);
  uint32_t eq = 0u;
  if (live)
    for (int n = s0; n < s1; ++n) eq += (sortable(p1[(int64_t)n * E + e]) == tau) ? 1u : 0u;
  eqcnt[w][lane] = eq;
  __syncthreads();
  if (!live) return;
  uint32_t above = above_ranks ? above_ranks[e] : 0u;
  for (int cc = c + 1; cc < Cn; ++cc) above += part[((size_t)g * Cn + cc) * 8192 + (tau & 255u) * 32 + lane];
  for (int ww = w + 1; ww < kSelThreads / 32; ++ww) above += eqcnt[ww][lane];
  for (int n = s1 - 1; n >= s0; --n) {
    const float pv = p1[(int64_t)n * E + e];
    const uint32_t key = sortable(pv);
    uint8_t a = 0;
    if (key > tau) a = 1;
    else if (key == tau) { a = (above < r) ? 1 : 0; ++above; }
    if (positive_only && !(pv > 0.f)) a = 0;
    actions[(int64_t)n * E + e] = a;
  }
}

// Marking, vectorised: lane = 4 consecutive envs (float4 load, uchar4 store), used when E % 4 == 0.  Env groups of 32 that
// have a tie at the threshold are skipped here and written by select_mark_kernel (launched right after with only_ties = 1).
__global__ void __launch_bounds__(kSelThreads) select_mark4_kernel(int E, int N, const float* __restrict__ p1, int chunk,
                                                                   const uint32_t* __restrict__ prefix,
                                                                   const uint32_t* __restrict__ remain,
                                                                   const uint32_t* __restrict__ eqtot, int positive_only,
                                                                   uint8_t* __restrict__ actions) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int e0 = (blockIdx.x * 32 + lane) * 4;
  const bool live = e0 < E;
  const int c = blockIdx.y;
  const int n0 = min(N, c * chunk), n1 = min(N, n0 + chunk);
  uint32_t tau[4] = {0u, 0u, 0u, 0u};
  bool tie = false;
  if (live) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tau[j] = prefix[e0 + j];
      tie |= eqtot[e0 + j] > remain[e0 + j];
    }
  }
  // the scalar kernel's env groups are 32 wide = 8 lanes here
  const uint32_t tb = __ballot_sync(0xffffffffu, tie);
  const bool skip = !live || ((tb >> (lane & ~7)) & 0xFFu) != 0u;
  if (skip) return;
  const float* col = p1 + e0;
  uint8_t* ac = actions + e0;
  for (int n = n0 + w; n < n1; n += 2 * (kSelThreads / 32)) {
    const int nb = n + kSelThreads / 32;
    const float4 v0 = *reinterpret_cast<const float4*>(col + (int64_t)n * E);
    float4 v1 = v0;
    if (nb < n1) v1 = *reinterpret_cast<const float4*>(col + (int64_t)nb * E);
    const float a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w};
    uchar4 o0, o1;
    uint8_t* q0 = reinterpret_cast<uint8_t*>(&o0);
    uint8_t* q1 = reinterpret_cast<uint8_t*>(&o1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      q0[j] = (sortable(a0[j]) >= tau[j] && (!positive_only || a0[j] > 0.f)) ? 1 : 0;
      q1[j] = (sortable(a1[j]) >= tau[j] && (!positive_only || a1[j] > 0.f)) ? 1 : 0;
    }
    *reinterpret_cast<uchar4*>(ac + (int64_t)n * E) = o0;
    if (nb < n1) *reinterpret_cast<uchar4*>(ac + (int64_t)nb * E) = o1;
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
  (void)A;
  if (E <= 0 || N <= 0) return 16;
  return select_plan(E, N).words * sizeof(uint32_t);
}

// number of ones exactly as the while-loop of rmabppo_core.py:320-331 counts them
static int select_count(int N, float cost1, float B) {
  int k = 0;
  double spent = 0.0;
  while (spent < (double)B && k < N) {
    if (spent + (double)cost1 > (double)B) break;
    spent += (double)cost1;
    ++k;
  }
  return k;
}

extern "C" int rmab_select_hist(int E, int N, const float* p1, int pass, void* ws, size_t ws_bytes, uint32_t* hist,
                                rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N > 0 && p1 && ws && pass >= 0 && pass < 4, RMAB_E_BADARG, "rmab_select_hist: bad argument");
  const SelPlan p = select_plan(E, N);
  RMAB_REQUIRE(ws_bytes >= p.words * sizeof(uint32_t), RMAB_E_BADARG, "rmab_select_hist: workspace of %zu B needed",
               p.words * sizeof(uint32_t));
  uint32_t* w = static_cast<uint32_t*>(ws);
  cudaStream_t st = (cudaStream_t)stream;
  select_hist_kernel<<<dim3(p.G, p.Cn), kSelThreads, 0, st>>>(E, N, p1, p.chunk, pass, w, w + p.part_off);
  select_reduce_kernel<<<dim3(p.G, 8192 / kSelThreads), kSelThreads, 0, st>>>(p.Cn, w + p.part_off, hist ? hist : w + p.tot_off);
  return check_launch("rmab_select_hist");
}

extern "C" int rmab_select_scan(int E, int N, int pass, int k, const uint32_t* hist, void* ws, size_t ws_bytes,
                                rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N > 0 && ws && pass >= 0 && pass < 4 && k > 0, RMAB_E_BADARG, "rmab_select_scan: bad argument");
  const SelPlan p = select_plan(E, N);
  RMAB_REQUIRE(ws_bytes >= p.words * sizeof(uint32_t), RMAB_E_BADARG, "rmab_select_scan: workspace too small");
  uint32_t* w = static_cast<uint32_t*>(ws);
  const size_t GE = (size_t)p.G * 32;
  select_scan_kernel<<<p.G, 32, 0, (cudaStream_t)stream>>>(E, pass, (uint32_t)k, hist ? hist : w + p.tot_off, w, w + GE,
                                                           w + 2 * GE);
  return check_launch("rmab_select_scan");
}

extern "C" int rmab_select_mark(int E, int N, const float* p1, const uint32_t* above_ranks, int positive_only,
                                uint8_t* actions, void* ws, size_t ws_bytes, rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N > 0 && p1 && ws && actions, RMAB_E_BADARG, "rmab_select_mark: bad argument");
  const SelPlan p = select_plan(E, N);
  RMAB_REQUIRE(ws_bytes >= p.words * sizeof(uint32_t), RMAB_E_BADARG, "rmab_select_mark: workspace too small");
  uint32_t* w = static_cast<uint32_t*>(ws);
  const size_t GE = (size_t)p.G * 32;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (E % 4) == 0 && ((uintptr_t)p1 % 16) == 0 && ((uintptr_t)actions % 4) == 0;
  if (vec) {
    // its env groups are 128 wide: own arm chunking so that the grid still fills the GPU
    const int g4 = (E / 4 + 31) / 32;
    int cn4 = (2 * kSelTargetCtas) / g4;
    const int max_cn4 = (N + 63) / 64;
    cn4 = cn4 < 1 ? 1 : (cn4 > max_cn4 ? max_cn4 : cn4);
    const int chunk4 = (N + cn4 - 1) / cn4;
    select_mark4_kernel<<<dim3(g4, cn4), kSelThreads, 0, st>>>(E, N, p1, chunk4, w, w + GE, w + 2 * GE, positive_only, actions);
    g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  select_mark_kernel<<<dim3(p.G, p.Cn), kSelThreads, 0, st>>>(E, N, p1, p.chunk, w, w + GE, w + 2 * GE, w + p.part_off,
                                                              above_ranks, positive_only, vec ? 1 : 0, actions);
  return check_launch("rmab_select_mark");
}

extern "C" int rmab_budget_select_binary(int E, int N, const float* p1, float cost1, float B, int positive_only,
                                         uint8_t* actions, void* ws, size_t ws_bytes, rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N >= 0 && p1 && actions, RMAB_E_BADARG, "rmab_budget_select_binary: bad argument");
  if (N == 0) return RMAB_OK;
  const int k = select_count(N, cost1, B);
  // the one-CTA-per-32-envs kernel wins while the whole problem is a few waves of it; beyond that (or whenever the caller
  // provides the workspace and the problem is large) the multi-CTA radix select streams p1 at HBM speed
  const bool big = ws && ws_bytes >= rmab_select_ws_bytes(E, N, 2) && (int64_t)N * E >= (1 << 21) && k > 0 && k < N;
  if (!big) {
    select_binary_kernel<<<(E + kSelEnvs - 1) / kSelEnvs, kSelEnvs * kSelWarps, 0, (cudaStream_t)stream>>>(
        E, N, p1, k, positive_only, actions);
    return check_launch("rmab_budget_select_binary");
  }
  for (int pass = 0; pass < 4; ++pass) {
    int rc = rmab_select_hist(E, N, p1, pass, ws, ws_bytes, nullptr, stream);
    if (rc != RMAB_OK) return rc;
    rc = rmab_select_scan(E, N, pass, k, nullptr, ws, ws_bytes, stream);
    if (rc != RMAB_OK) return rc;
  }
  return rmab_select_mark(E, N, p1, nullptr, positive_only, actions, ws, ws_bytes, stream);
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
