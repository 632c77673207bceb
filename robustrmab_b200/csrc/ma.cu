// MA-RMABPPO rollout pieces: Gaussian nature-actor sampling and the counterfactual reward estimate.
//
// Reference behaviour: robust_rmab/algos/ma_rmabppo/ma_rmabppo_core.py MLPGaussianActor :98-113, step :289-291
// (a ~ Normal(mu_net(obs), exp(log_std)), logp = sum over action dims); nature_oracle.py:685-701 (50 env steps
// from the same state with the test agent action, mean of the summed rewards).
#include <math.h>
#include "common.cuh"

namespace rmab {

// Two standard normals from one Philox block (Box-Muller on words 0 / 1), as oracle/philox.py::normal_pair.
__device__ __forceinline__ void philox_normal_pair(uint32_t k0, uint32_t k1, uint32_t stream, uint32_t t, uint32_t env,
                                                   uint32_t idx, float& z0, float& z1) {
  const uint4 x = philox4x32_10(stream, t, env, idx, k0, k1);
  const float u1 = __fmul_rn(__fadd_rn((float)(x.x >> 8), 1.f), 5.9604644775390625e-08f);  // (0,1]
  const float u2 = (float)(x.y >> 8) * 5.9604644775390625e-08f;
  const float r = sqrtf(__fmul_rn(-2.f, logf(u1)));
  const float ang = __fmul_rn(6.2831854820251465f, u2);
  z0 = __fmul_rn(r, cosf(ang));
  z1 = __fmul_rn(r, sinf(ang));
}

__global__ void __launch_bounds__(256) gaussian_sample_kernel(int E, int D, const float* __restrict__ mu,
                                                              const float* __restrict__ log_std, Rng g,
                                                              const float* __restrict__ eps_inject,
                                                              float* __restrict__ a, float* __restrict__ logp) {
  __shared__ float red[32];
  const int e = blockIdx.x;
  float acc = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float eps;
    if (eps_inject) eps = eps_inject[(int64_t)e * D + d];
    else {
      float z0, z1;
      philox_normal_pair(g.k0, g.k1, g.stream, rng_t(g), g.env0 + (uint32_t)e, (uint32_t)(d >> 1), z0, z1);
      eps = (d & 1) ? z1 : z0;
    }
    const float ls = log_std[d];
    const float sd = expf(ls);
    const float m = mu[(int64_t)e * D + d];
    const float v = __fadd_rn(m, __fmul_rn(sd, eps));
    a[(int64_t)e * D + d] = v;
    // torch Normal.log_prob: -((value - loc)^2) / (2 var) - log(scale) - log(sqrt(2 pi))
    const float diff = __fsub_rn(v, m);
    const float var = __fmul_rn(sd, sd);
    acc += __fsub_rn(__fsub_rn(-__fdiv_rn(__fmul_rn(diff, diff), __fmul_rn(2.f, var)), ls), 0.9189385332046727f);
  }
  const float tot = block_sum(acc, red);
  if (threadIdx.x == 0) logp[e] = tot;
}

// ---- counterfactual steps ------------------------------------------------------------------------
__device__ __forceinline__ int cf_table_next(int domain, int s, int a, float p, float u) {
  if (domain == RMAB_DOMAIN_CE) {
    float pr[2];
    if (s == 0) { pr[0] = 0.5f; pr[1] = 0.5f; }
    else if (a == 0) { pr[0] = 1.f; pr[1] = 0.f; }
    else { pr[0] = __fsub_rn(1.f, p); pr[1] = p; }
    return categorical<2>(pr, u);
  }
  const bool to0 = (s == 0) || (s == 1 && a == 1);
  float pr[3];
  pr[0] = to0 ? p : 0.f;
  pr[1] = __fsub_rn(1.f, p);
  pr[2] = to0 ? 0.f : p;
  return categorical<3>(pr, u);
}

__global__ void cf_table_kernel(int domain, int E, int N, const uint8_t* __restrict__ s, const uint8_t* __restrict__ a,
                                const float* __restrict__ nature, int nature_per_env, const float* __restrict__ ranges,
                                const float* __restrict__ R, Rng g, int n_trials, float* __restrict__ r_mean) {
  const int S = domain == RMAB_DOMAIN_CE ? 2 : 3, A = 2;
  const int64_t total = (int64_t)N * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / E), e = (int)(i - (int64_t)n * E);
    const int si = min((int)s[i], S - 1), ai = min((int)a[i], A - 1);
    float p, lo, hi;
    if (domain == RMAB_DOMAIN_CE) {
      p = nature_per_env ? nature[i] : __ldg(nature + n);
      lo = __ldg(ranges + 2 * (int64_t)n);
      hi = __ldg(ranges + 2 * (int64_t)n + 1);
    } else {
      const int64_t cell = ((int64_t)n * S + si) * A + ai;
      p = nature_per_env ? nature[((int64_t)n * A + ai) * E + e] : __ldg(nature + cell);
      lo = __ldg(ranges + 2 * cell);
      hi = __ldg(ranges + 2 * cell + 1);
    }
    p = fminf(fmaxf(p, lo), hi);
    float acc = 0.f;
    for (int tr = 0; tr < n_trials; ++tr) {
      const uint4 x = philox4x32_10(RMAB_STREAM_CF, rng_t(g) * (uint32_t)n_trials + (uint32_t)tr, g.env0 + (uint32_t)e,
                                    g.arm0 + (uint32_t)n, g.k0, g.k1);
      const int sn = cf_table_next(domain, si, ai, p, u24(x.x));
      acc += __ldg(R + (int64_t)n * S + sn);
    }
    r_mean[i] = acc / (float)n_trials;
  }
}

// SIS: the binomial pmf of (s, a, params) is built once per (arm, env) and sampled n_trials times.
constexpr int kCfMaxS = 256;
__device__ double cf_sis_q(int pop, int s, int a, double r_t, double lam, double e1, double e2) {
  const double beta = (double)(pop - s) / (double)pop;
  double x;
  if (a == 0) x = __dmul_rn(__dmul_rn(-lam, beta), r_t);
  else if (a == 1) x = __dmul_rn(__ddiv_rn(__dmul_rn(-lam, beta), e1), r_t);
  else x = __ddiv_rn(__dmul_rn(__dmul_rn(-lam, beta), r_t), e2);
  return 1.0 - pow(2.718281828459045, x);
}

__global__ void __launch_bounds__(128) cf_sis_kernel(int E, int N, int pop, const uint8_t* __restrict__ s,
                                                      const uint8_t* __restrict__ a, const float* __restrict__ params,
                                                      int nature_per_env, const float* __restrict__ R, Rng g,
                                                      int n_trials, float* __restrict__ r_mean) {
  const int S = pop + 1;
  const int64_t total = (int64_t)N * E;
  float cdf[kCfMaxS];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / E), e = (int)(i - (int64_t)n * E);
    const int si = min((int)s[i], S - 1), ai = min((int)a[i], 2);
    if (si == pop) {  // absorbing (:1013-1017)
      r_mean[i] = __ldg(R + (int64_t)n * S + pop);
      continue;
    }
    double prm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      prm[k] = (double)(nature_per_env ? params[((int64_t)n * 4 + k) * E + e] : __ldg(params + (int64_t)n * 4 + k));
    const double q = cf_sis_q(pop, si, ai, prm[0], prm[1], prm[2], prm[3]);
    // pmf over i = 0..si infected-to-recover draws, categories sp = pop - i in ascending sp order (k = si .. 0)
    const double omq = 1.0 - q;
    double prob[kCfMaxS];
    if (q <= 0.0) { prob[0] = 1.0; for (int k = 1; k <= si; ++k) prob[k] = 0.0; }
    else if (omq <= 0.0) { for (int k = 0; k < si; ++k) prob[k] = 0.0; prob[si] = 1.0; }
    else {
      int m = (int)((double)(si + 1) * q);
      if (m > si) m = si;
      const double lg = lgamma((double)si + 1.0) - lgamma((double)m + 1.0) - lgamma((double)(si - m) + 1.0);
      prob[m] = exp(lg + (double)m * log(q) + (double)(si - m) * log(omq));
      const double ratio = q / omq, iratio = omq / q;
      for (int k = m; k < si; ++k) prob[k + 1] = prob[k] * ((double)(si - k) / (double)(k + 1)) * ratio;
      for (int k = m; k > 0; --k) prob[k - 1] = prob[k] * ((double)k / (double)(si - k + 1)) * iratio;
    }
    double tot = 0.0;
    for (int k = 0; k <= si; ++k) { if (prob[k] < 1e-7) prob[k] = 0.0; tot += prob[k]; }
    float c = 0.f;
    int last = 0;
    for (int k = si, j = 0; k >= 0; --k, ++j) {
      const float pf = (float)(prob[k] / tot);
      c = __fadd_rn(c, pf);
      cdf[j] = c;               // cdf over categories sp = pop - si + j
      if (pf > 0.f) last = j;
    }
    float acc = 0.f;
    for (int tr = 0; tr < n_trials; ++tr) {
      const uint4 x = philox4x32_10(RMAB_STREAM_CF, rng_t(g) * (uint32_t)n_trials + (uint32_t)tr, g.env0 + (uint32_t)e,
                                    g.arm0 + (uint32_t)n, g.k0, g.k1);
      const float u = u24(x.x);
      // cnt = #{j <= si : cdf[j] <= u}; the cdf is non-decreasing (round-to-nearest sums of non-negative terms), so that is
      // the first index with cdf > u: a binary search instead of si + 1 comparisons per trial
      int lo = 0, hi = si + 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
      }
      const int cnt = lo;
      const int sn = pop - si + (cnt < last ? cnt : last);
      acc += __ldg(R + (int64_t)n * S + sn);
    }
    r_mean[i] = acc / (float)n_trials;
  }
}

// Elementwise Adam step (torch.optim.Adam defaults form, _single_tensor_adam) on a flat tensor: used for the
// agent critics' first-layer block over the nature inputs, whose gradient comes from a dense GEMM.
__global__ void adam_step_kernel(int64_t n, float* __restrict__ p, const float* __restrict__ grad, float* __restrict__ m,
                                 float* __restrict__ v, float lerp_w, float beta2, float omb2, float step_size, float bc2s,
                                 float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float mi = m[i], vi = v[i];
    p[i] = adam_update(p[i], grad[i], mi, vi, lerp_w, beta2, omb2, step_size, bc2s, eps);
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace rmab

using namespace rmab;

extern "C" int rmab_adam_step(int64_t n, float* p, const float* grad, float* m, float* v, double lr, double beta1,
                              double beta2, double eps, int step, rmab_stream_t stream) {
  RMAB_REQUIRE(n >= 0 && p && grad && m && v && step >= 1, RMAB_E_BADARG, "rmab_adam_step: bad argument");
  if (n == 0) return RMAB_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const int64_t b = (n + 255) / 256;
  const int grid = (int)(b < (int64_t)kNumSMs * 16 ? b : (int64_t)kNumSMs * 16);
  adam_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, p, grad, m, v, (float)(1.0 - beta1), (float)beta2,
                                                          (float)(1.0 - beta2), (float)(lr / bc1), (float)sqrt(bc2),
                                                          (float)eps);
  return check_launch("rmab_adam_step");
}

extern "C" int rmab_gaussian_sample(int E, int D, const float* mu, const float* log_std, const RmabRng* rng,
                                    const float* eps_inject, float* a, float* logp, rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && D > 0 && mu && log_std && a && logp && (rng || eps_inject), RMAB_E_BADARG,
               "rmab_gaussian_sample: bad argument");
  RmabRng z = {0, RMAB_STREAM_NATURE, 0, 0, 0, nullptr};
  gaussian_sample_kernel<<<E, 256, 0, (cudaStream_t)stream>>>(E, D, mu, log_std, make_rng(rng ? rng : &z), eps_inject, a,
                                                             logp);
  return check_launch("rmab_gaussian_sample");
}

extern "C" int rmab_env_counterfactual(int domain, int E, int N, int S, const uint8_t* s, const uint8_t* a,
                                       const float* nature, int nature_per_env, const float* ranges, const float* R,
                                       const RmabRng* rng, int n_trials, float* r_mean, rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N >= 0 && n_trials > 0 && s && a && nature && R && rng && r_mean, RMAB_E_BADARG,
               "rmab_env_counterfactual: bad argument");
  if (N == 0) return RMAB_OK;
  const Rng g = make_rng(rng);
  const int64_t total = (int64_t)N * E;
  cudaStream_t st = (cudaStream_t)stream;
  if (domain == RMAB_DOMAIN_SIS) {
    RMAB_REQUIRE(S >= 2 && S <= kCfMaxS, RMAB_E_SHAPE, "rmab_env_counterfactual: SIS needs 2 <= S <= 256");
    const int grid = (int)((total + 127) / 128 < (int64_t)kNumSMs * 16 ? (total + 127) / 128 : (int64_t)kNumSMs * 16);
    cf_sis_kernel<<<grid, 128, 0, st>>>(E, N, S - 1, s, a, nature, nature_per_env, R, g, n_trials, r_mean);
  } else {
    RMAB_REQUIRE(domain == RMAB_DOMAIN_CE || domain == RMAB_DOMAIN_ARMMAN, RMAB_E_BADARG,
                 "rmab_env_counterfactual: unknown domain %d", domain);
    RMAB_REQUIRE(ranges, RMAB_E_BADARG, "rmab_env_counterfactual: table domains need ranges");
    const int grid = (int)((total + 255) / 256 < (int64_t)kNumSMs * 16 ? (total + 255) / 256 : (int64_t)kNumSMs * 16);
    cf_table_kernel<<<grid, 256, 0, st>>>(domain, E, N, s, a, nature, nature_per_env, ranges, R, g, n_trials, r_mean);
  }
  return check_launch("rmab_env_counterfactual");
}
