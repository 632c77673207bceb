// PPO-buffer kernels: GAE-lambda advantages, rewards-to-go, q, per-arm advantage normalisation and the
// arm-summed discounted cost-to-go.
//
// Reference behaviour: robust_rmab/algos/rmabppo/agent_oracle.py RMABPPO_Buffer.finish_path :90-139,
// get :144-161; core.discount_cumsum rmabppo_core.py:36-51 (scipy lfilter in float64, stored as fp32);
// mpi_statistics_scalar utils/mpi_tools.py:76-97 (float32 mean / population std).
// Segmented reverse scan: one CTA per arm, one thread per env walks t = T-1..0 (coalesced across envs),
// in fp64 with explicit round-to-nearest mul/add so the recurrence equals the reference's bit for bit.
// HBM-bound: algorithmic 14 B per arm-step (s' 1 + a 1 + val 4 in, adv 4 + ret 4 out; +4 if q is kept).
#include "common.cuh"

namespace rmab {

template <bool EXPLICIT>
__global__ void __launch_bounds__(256) gae_kernel(int T, int E, int S, int A, const uint8_t* __restrict__ s_traj,
                                                  const uint8_t* __restrict__ act, const float* __restrict__ rew,
                                                  const float* __restrict__ cost, const float* __restrict__ val,
                                                  const float* __restrict__ last_val, const float* __restrict__ lamb,
                                                  const float* __restrict__ C, const float* __restrict__ R, double gamma,
                                                  double lam, int normalize, float* __restrict__ adv,
                                                  float* __restrict__ ret, float* __restrict__ q,
                                                  float* __restrict__ adv_stats) {
  __shared__ float red[32];
  __shared__ float sR[256];
  __shared__ float sC[16];
  const int n = blockIdx.x;
  if (!EXPLICIT) {
    for (int i = threadIdx.x; i < S; i += blockDim.x) sR[i] = R[(int64_t)n * S + i];
    for (int i = threadIdx.x; i < A && i < 16; i += blockDim.x) sC[i] = C[i];
    __syncthreads();
  }
  const double gl = __dmul_rn(gamma, lam);
  const int64_t base = (int64_t)n * T * E;
  const int64_t tbase = (int64_t)n * (T + 1) * E;
  float fsum = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const double lv = (double)last_val[(int64_t)n * E + e];
    const double lm = (double)lamb[e];
    double next_v = lv, adv_acc = 0.0, ret_acc = lv;
    for (int t = T - 1; t >= 0; --t) {
      const int64_t o = base + (int64_t)t * E + e;
      double r, c;
      if (EXPLICIT) {
        r = (double)rew[o];
        c = (double)cost[o];
      } else {
        int sn = s_traj[tbase + (int64_t)(t + 1) * E + e];
        int a = act[o];
        r = (double)sR[sn < S ? sn : S - 1];
        c = (double)sC[a < A ? a : A - 1];
      }
      const double v = (double)val[o];
      const double rt = __dsub_rn(r, __dmul_rn(lm, c));                 // :120
      const double qv = __dadd_rn(rt, __dmul_rn(gamma, next_v));         // :125
      const double delta = __dsub_rn(qv, v);                             // :126
      adv_acc = __dadd_rn(delta, __dmul_rn(gl, adv_acc));                // :127
      ret_acc = __dadd_rn(rt, __dmul_rn(gamma, ret_acc));                // :130
      const float af = (float)adv_acc;
      adv[o] = af;
      ret[o] = (float)ret_acc;
      if (q) q[o] = (float)qv;
      fsum += af;
      next_v = v;
    }
  }
  if (!normalize && !adv_stats) return;
  const float cnt = (float)((int64_t)T * E);
  const float mean = block_sum(fsum, red) / cnt;
  float fss = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    for (int t = 0; t < T; ++t) {
      const float d = adv[base + (int64_t)t * E + e] - mean;
      fss = fmaf(d, d, fss);
    }
  const float stdv = sqrtf(block_sum(fss, red) / cnt);
  if (adv_stats && threadIdx.x == 0) {
    adv_stats[2 * (int64_t)n] = mean;
    adv_stats[2 * (int64_t)n + 1] = stdv;
  }
  if (!normalize) return;
  for (int e = threadIdx.x; e < E; e += blockDim.x)
    for (int t = 0; t < T; ++t) {
      const int64_t o = base + (int64_t)t * E + e;
      adv[o] = __fdiv_rn(adv[o] - mean, stdv);                           // :155
    }
}

// arm_summed_costs (:117): csum[t,e] = sum_n C[a[n,t,e]] in fp64, arms in ascending order within 8
// contiguous chunks whose partials are added in chunk order.
__global__ void __launch_bounds__(256) cost_sum_kernel(int T, int E, int N, int A, const uint8_t* __restrict__ act,
                                                       const float* __restrict__ C, double* __restrict__ csum) {
  __shared__ double part[8][32];
  __shared__ float sC[16];
  if (threadIdx.x < 16) sC[threadIdx.x] = threadIdx.x < A ? C[threadIdx.x] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t = blockIdx.y;
  const int e = blockIdx.x * 32 + lane;
  const int chunk = (N + 7) / 8;
  const int n0 = w * chunk, n1 = min(N, n0 + chunk);
  double acc = 0.0;
  if (e < E)
    for (int n = n0; n < n1; ++n) {
      const int a = act[((int64_t)n * T + t) * E + e];
      acc += (double)sC[a < A ? a : A - 1];
    }
  part[w][lane] = acc;
  __syncthreads();
  if (w == 0 && e < E) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += part[k][lane];
    csum[(int64_t)t * E + e] = tot;
  }
}

// cdcost = discount_cumsum([csum, 0], gamma)[:-1]  (:139)
__global__ void cost_scan_kernel(int T, int E, const double* __restrict__ csum, double gamma,
                                 float* __restrict__ cdcost) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double acc = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    acc = __dadd_rn(csum[(int64_t)t * E + e], __dmul_rn(gamma, acc));
    cdcost[(int64_t)t * E + e] = (float)acc;
  }
}

}  // namespace rmab

using namespace rmab;

extern "C" int rmab_gae(int T, int E, int N, int S, int A, const uint8_t* s_traj, const uint8_t* a, const float* val,
                        const float* last_val, const float* lamb, const float* C, const float* R, double gamma,
                        double lam, int normalize, float* adv, float* ret, float* q, float* adv_stats,
                        rmab_stream_t stream) {
  RMAB_REQUIRE(T > 0 && E > 0 && N >= 0 && S >= 1 && S <= 256 && A >= 1 && A <= 16, RMAB_E_SHAPE, "rmab_gae: bad shape");
  RMAB_REQUIRE(s_traj && a && val && last_val && lamb && C && R && adv && ret, RMAB_E_BADARG, "rmab_gae: null pointer");
  if (N == 0) return RMAB_OK;
  gae_kernel<false><<<N, 256, 0, (cudaStream_t)stream>>>(T, E, S, A, s_traj, a, nullptr, nullptr, val, last_val, lamb, C,
                                                        R, gamma, lam, normalize, adv, ret, q, adv_stats);
  return check_launch("rmab_gae");
}

extern "C" int rmab_gae_explicit(int T, int E, int N, const float* rew, const float* cost, const float* val,
                                 const float* last_val, const float* lamb, double gamma, double lam, int normalize,
                                 float* adv, float* ret, float* q, float* adv_stats, rmab_stream_t stream) {
  RMAB_REQUIRE(T > 0 && E > 0 && N >= 0, RMAB_E_SHAPE, "rmab_gae_explicit: bad shape");
  RMAB_REQUIRE(rew && cost && val && last_val && lamb && adv && ret, RMAB_E_BADARG, "rmab_gae_explicit: null pointer");
  if (N == 0) return RMAB_OK;
  gae_kernel<true><<<N, 256, 0, (cudaStream_t)stream>>>(T, E, 1, 1, nullptr, nullptr, rew, cost, val, last_val, lamb,
                                                       nullptr, nullptr, gamma, lam, normalize, adv, ret, q, adv_stats);
  return check_launch("rmab_gae_explicit");
}

extern "C" int rmab_cost_togo(int T, int E, int N, int A, const uint8_t* a, const float* C, double gamma, double* ws,
                              float* cdcost, rmab_stream_t stream) {
  RMAB_REQUIRE(T > 0 && E > 0 && N >= 0 && A >= 1 && A <= 16 && a && C && ws && cdcost, RMAB_E_BADARG,
               "rmab_cost_togo: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  cost_sum_kernel<<<dim3((E + 31) / 32, T), 256, 0, st>>>(T, E, N, A, a, C, ws);
  int rc = check_launch("rmab_cost_togo(sum)");
  if (rc) return rc;
  cost_scan_kernel<<<(E + 127) / 128, 128, 0, st>>>(T, E, ws, gamma, cdcost);
  return check_launch("rmab_cost_togo(scan)");
}
