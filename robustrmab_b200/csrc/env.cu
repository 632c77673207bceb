// Environment-step kernels: reset, counterexample / ARMMAN table domains, SIS (binomial pmf on the fly),
// SIS full table, tanh bounding of nature actions.
//
// Reference behaviour (robust_rmab/environments/bandit_env_robust.py):
//   CounterExampleRobustEnv.step :860-893, ARMMANRobustEnv.step :575-634, SISRobustEnv.step :1224-1243,
//   compute_distro :1025-1084, p_i_s :1011-1022, bound_nature_actions :703/:930/:1274, reset :941-944.
// All of these are HBM-streaming, one (arm, env) element per lane slot; arm-major layout [N,E] makes the
// env index the coalesced one and lets the per-arm parameters be warp-uniform broadcasts.
#include "common.cuh"

namespace rmab {

// ------------------------------------------------------------------------------------------------
__global__ void reset_states_kernel(int E, int64_t total, int S, Rng g, const float* __restrict__ u_inject,
                                    uint8_t* __restrict__ s) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int n = (int)(i / E), e = (int)(i - (int64_t)n * E);
    const float u = draw_uniform(g, u_inject, i, e, n);
    int v = (int)(__fmul_rn(u, (float)S));
    s[i] = (uint8_t)(v < S - 1 ? v : S - 1);
  }
}

// ------------------------------------------------------------------------------------------------
// Table domains.  One thread handles VEC consecutive envs of one arm (uchar4 loads when VEC == 4).
__device__ __forceinline__ int ce_next(int s, int a, float p, float u) {
  // T[s=0] = [.5,.5]; T[1,0] = [1,0]; T[1,1] = [1-p, p]   (:831-836, :876-877)
  float pr[2];
  if (s == 0) { pr[0] = 0.5f; pr[1] = 0.5f; }
  else if (a == 0) { pr[0] = 1.f; pr[1] = 0.f; }
  else { pr[0] = __fsub_rn(1.f, p); pr[1] = p; }
  return categorical<2>(pr, u);
}

__device__ __forceinline__ int armman_next(int s, int a, float p, float u) {
  // s=0: [p,1-p,0]; s=1,a=0: [0,1-p,p]; s=1,a=1: [p,1-p,0]; s=2: [0,1-p,p]   (:595-616)
  const bool to0 = (s == 0) || (s == 1 && a == 1);
  float pr[3];
  pr[0] = to0 ? p : 0.f;
  pr[1] = __fsub_rn(1.f, p);
  pr[2] = to0 ? 0.f : p;
  return categorical<3>(pr, u);
}

template <int DOMAIN, int VEC>
__global__ void env_step_table_kernel(int E, int N, const uint8_t* __restrict__ s, const uint8_t* __restrict__ a,
                                      const float* __restrict__ nature, int nature_per_env,
                                      const float* __restrict__ ranges, const float* __restrict__ R, Rng g,
                                      const float* __restrict__ u_inject, uint8_t* __restrict__ s_next,
                                      float* __restrict__ r) {
  constexpr int S = DOMAIN == RMAB_DOMAIN_CE ? 2 : 3;
  constexpr int A = 2;
  const int EV = E / VEC;  // vec groups per arm
  const int64_t total = (int64_t)N * EV;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < total; gi += stride) {
    const int n = (int)(gi / EV);
    const int e0 = (int)(gi - (int64_t)n * EV) * VEC;
    const int64_t base = (int64_t)n * E + e0;
    uint8_t sv[VEC], av[VEC], outv[VEC];
    if (VEC == 4) {
      *reinterpret_cast<uchar4*>(sv) = *reinterpret_cast<const uchar4*>(s + base);
      *reinterpret_cast<uchar4*>(av) = *reinterpret_cast<const uchar4*>(a + base);
    } else {
      sv[0] = s[base];
      av[0] = a[base];
    }
    float Rn[S];
#pragma unroll
    for (int k = 0; k < S; ++k) Rn[k] = __ldg(R + (int64_t)n * S + k);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const int e = e0 + v;
      const int si = sv[v] < S ? sv[v] : S - 1, ai = av[v] < A ? av[v] : A - 1;
      float p, lo, hi;
      if (DOMAIN == RMAB_DOMAIN_CE) {
        p = nature_per_env ? nature[base + v] : __ldg(nature + n);
        lo = __ldg(ranges + 2 * (int64_t)n);
        hi = __ldg(ranges + 2 * (int64_t)n + 1);
      } else {
        const int64_t cell = ((int64_t)n * S + si) * A + ai;
        p = nature_per_env ? nature[((int64_t)n * A + ai) * E + e] : __ldg(nature + cell);
        lo = __ldg(ranges + 2 * cell);
        hi = __ldg(ranges + 2 * cell + 1);
      }
      p = fminf(fmaxf(p, lo), hi);  // warn-and-clamp of :867-874 / :583-592
      const float u = draw_uniform(g, u_inject, base + v, e, n);
      const int sn = DOMAIN == RMAB_DOMAIN_CE ? ce_next(si, ai, p, u) : armman_next(si, ai, p, u);
      outv[v] = (uint8_t)sn;
      if (r) {
        float rv = Rn[0];
#pragma unroll
        for (int k = 1; k < S; ++k) rv = (sn == k) ? Rn[k] : rv;
        r[base + v] = rv;
      }
    }
    if (VEC == 4) *reinterpret_cast<uchar4*>(s_next + base) = *reinterpret_cast<uchar4*>(outv);
    else s_next[base] = outv[0];
  }
}

// ------------------------------------------------------------------------------------------------
// SIS: next state pop - i, i ~ Binomial(s, q) with entries < 1e-7 zeroed and the row renormalised
// (compute_distro :1049-1082), q from (r_t, lam, e1, e2) in fp64 with the reference's operation order.
constexpr int kSisMaxS = 256;

__device__ __forceinline__ double sis_q(int pop, int s, int a, double r_t, double lam, double e1, double e2) {
  const double beta = (double)(pop - s) / (double)pop;
  double x;
  if (a == 0) x = __dmul_rn(__dmul_rn(-lam, beta), r_t);
  else if (a == 1) x = __dmul_rn(__ddiv_rn(__dmul_rn(-lam, beta), e1), r_t);
  else x = __ddiv_rn(__dmul_rn(__dmul_rn(-lam, beta), r_t), e2);
  return 1.0 - pow(2.718281828459045, x);  // np.e ** x  (:1057-1061)
}

// Raw binomial pmf prob[i], i = 0..s, by the upward recurrence from (1-q)^s (fp64).  s < pop here.
__device__ __forceinline__ void sis_pmf(int s, double q, double* prob) {
  const double omq = 1.0 - q;
  if (q <= 0.0) {
    prob[0] = 1.0;
    for (int i = 1; i <= s; ++i) prob[i] = 0.0;
    return;
  }
  if (omq <= 0.0) {
    for (int i = 0; i < s; ++i) prob[i] = 0.0;
    prob[s] = 1.0;
    return;
  }
  // start at the mode to avoid underflow of (1-q)^s or q^s, then recur both ways
  int m = (int)((double)(s + 1) * q);
  if (m > s) m = s;
  const double lg = lgamma((double)s + 1.0) - lgamma((double)m + 1.0) - lgamma((double)(s - m) + 1.0);
  prob[m] = exp(lg + (double)m * log(q) + (double)(s - m) * log(omq));
  const double ratio = q / omq;
  for (int i = m; i < s; ++i) prob[i + 1] = prob[i] * ((double)(s - i) / (double)(i + 1)) * ratio;
  const double iratio = omq / q;
  for (int i = m; i > 0; --i) prob[i - 1] = prob[i] * ((double)i / (double)(s - i + 1)) * iratio;
}

// Normalised fp32 row distro[sp], sp = 0..pop into `out` (stride `ostride`), returns nothing.
__device__ __forceinline__ void sis_row(int pop, int s, double q, double* prob, float* out, int64_t ostride) {
  const int S = pop + 1;
  if (s == pop) {  // p_i_s: i == 0 w.p. 1  (:1013-1017)
    for (int sp = 0; sp < S; ++sp) out[sp * ostride] = (sp == pop) ? 1.f : 0.f;
    return;
  }
  sis_pmf(s, q, prob);
  double tot = 0.0;
  for (int sp = pop - s; sp <= pop; ++sp) {
    double p = prob[pop - sp];
    if (p < 1e-7) { p = 0.0; prob[pop - sp] = 0.0; }
    tot += p;
  }
  for (int sp = 0; sp < S; ++sp) out[sp * ostride] = (sp < pop - s) ? 0.f : (float)(prob[pop - sp] / tot);
}

__global__ void env_step_sis_kernel(int E, int N, int pop, const uint8_t* __restrict__ s,
                                    const uint8_t* __restrict__ a, const float* __restrict__ params,
                                    int nature_per_env, const float* __restrict__ R, Rng g,
                                    const float* __restrict__ u_inject, uint8_t* __restrict__ s_next,
                                    float* __restrict__ r) {
  const int S = pop + 1;
  const int64_t total = (int64_t)N * E;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double prob[kSisMaxS];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int n = (int)(i / E), e = (int)(i - (int64_t)n * E);
    int si = s[i];
    si = si < S ? si : S - 1;
    int ai = a[i];
    ai = ai < 3 ? ai : 2;
    double prm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      prm[k] = (double)(nature_per_env ? params[((int64_t)n * 4 + k) * E + e] : __ldg(params + (int64_t)n * 4 + k));
    const float u = draw_uniform(g, u_inject, i, e, n);
    int sn;
    if (si == pop) {
      sn = pop;
    } else {
      const double q = sis_q(pop, si, ai, prm[0], prm[1], prm[2], prm[3]);
      sis_pmf(si, q, prob);
      double tot = 0.0;
      for (int k = 0; k <= si; ++k) {
        if (prob[k] < 1e-7) prob[k] = 0.0;
        tot += prob[k];
      }
      // categories sp = pop-si .. pop  <->  k = si .. 0 ; leading categories have p = 0 (cdf stays 0 <= u)
      float cdf = 0.f;
      int cnt = pop - si, last = 0;
      for (int k = si; k >= 0; --k) {
        const float pf = (float)(prob[k] / tot);
        cdf = __fadd_rn(cdf, pf);
        cnt += (cdf <= u) ? 1 : 0;
        if (pf > 0.f) last = pop - k;
      }
      sn = cnt < last ? cnt : last;
    }
    s_next[i] = (uint8_t)sn;
    if (r) r[i] = __ldg(R + (int64_t)n * S + sn);
  }
}

__global__ void sis_table_kernel(int N, int pop, const float* __restrict__ params, float* __restrict__ T) {
  const int S = pop + 1;
  const int64_t total = (int64_t)N * S * 3;
  double prob[kSisMaxS];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int a = (int)(i % 3);
    const int s = (int)((i / 3) % S);
    const int n = (int)(i / (3 * S));
    const float* p = params + (int64_t)n * 4;
    const double q = sis_q(pop, s, a, (double)p[0], (double)p[1], (double)p[2], (double)p[3]);
    sis_row(pop, s, q, prob, T + i * S, 1);
  }
}

__global__ void bound_nature_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ lo,
                                    const float* __restrict__ hi, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = tanhf(x[i]);
    const float half = __fdiv_rn(__fadd_rn(t, 1.f), 2.f);
    out[i] = __fadd_rn(__fmul_rn(half, __fsub_rn(hi[i], lo[i])), lo[i]);
  }
}

static inline int grid_for(int64_t work, int block) {
  int64_t b = (work + block - 1) / block;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace rmab

using namespace rmab;

extern "C" int rmab_reset_states(int E, int N, int S, const RmabRng* rng, const float* u_inject, uint8_t* s,
                                 rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N >= 0 && S > 0 && S <= 256 && s && (rng || u_inject), RMAB_E_BADARG, "rmab_reset_states: bad argument");
  if (N == 0) return RMAB_OK;
  RmabRng z = {0, 0, 0, 0, 0};
  const int64_t total = (int64_t)N * E;
  reset_states_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(E, total, S, make_rng(rng ? rng : &z),
                                                                             u_inject, s);
  return check_launch("rmab_reset_states");
}

extern "C" int rmab_env_step_table(int domain, int E, int N, const uint8_t* s, const uint8_t* a, const float* nature,
                                   int nature_per_env, const float* ranges, const float* R, const RmabRng* rng,
                                   const float* u_inject, uint8_t* s_next, float* r, rmab_stream_t stream) {
  RMAB_REQUIRE(domain == RMAB_DOMAIN_CE || domain == RMAB_DOMAIN_ARMMAN, RMAB_E_BADARG,
               "rmab_env_step_table: domain %d is not a table domain", domain);
  RMAB_REQUIRE(E > 0 && N >= 0 && s && a && nature && ranges && R && s_next && (rng || u_inject), RMAB_E_BADARG,
               "rmab_env_step_table: null pointer or bad shape");
  if (N == 0) return RMAB_OK;
  RmabRng z = {0, 0, 0, 0, 0};
  const Rng g = make_rng(rng ? rng : &z);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec4 = (E % 4 == 0) && ((((uintptr_t)s | (uintptr_t)a | (uintptr_t)s_next) & 3) == 0);
  const int64_t work = (int64_t)N * (vec4 ? E / 4 : E);
  const int grid = grid_for(work, 256);
#define LAUNCH(D, V) \
  env_step_table_kernel<D, V><<<grid, 256, 0, st>>>(E, N, s, a, nature, nature_per_env, ranges, R, g, u_inject, s_next, r)
  if (domain == RMAB_DOMAIN_CE) { if (vec4) LAUNCH(RMAB_DOMAIN_CE, 4); else LAUNCH(RMAB_DOMAIN_CE, 1); }
  else { if (vec4) LAUNCH(RMAB_DOMAIN_ARMMAN, 4); else LAUNCH(RMAB_DOMAIN_ARMMAN, 1); }
#undef LAUNCH
  return check_launch("rmab_env_step_table");
}

extern "C" int rmab_env_step_sis(int E, int N, int pop, const uint8_t* s, const uint8_t* a, const float* params,
                                 int nature_per_env, const float* R, const RmabRng* rng, const float* u_inject,
                                 uint8_t* s_next, float* r, rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N >= 0 && pop >= 1 && pop < kSisMaxS, RMAB_E_SHAPE, "rmab_env_step_sis: pop must be in [1,255]");
  RMAB_REQUIRE(s && a && params && R && s_next && (rng || u_inject), RMAB_E_BADARG, "rmab_env_step_sis: null pointer");
  if (N == 0) return RMAB_OK;
  RmabRng z = {0, 0, 0, 0, 0};
  env_step_sis_kernel<<<grid_for((int64_t)N * E, 128), 128, 0, (cudaStream_t)stream>>>(
      E, N, pop, s, a, params, nature_per_env, R, make_rng(rng ? rng : &z), u_inject, s_next, r);
  return check_launch("rmab_env_step_sis");
}

extern "C" int rmab_sis_table(int N, int pop, const float* params, float* T, rmab_stream_t stream) {
  RMAB_REQUIRE(N >= 0 && pop >= 1 && pop < kSisMaxS && params && T, RMAB_E_BADARG, "rmab_sis_table: bad argument");
  if (N == 0) return RMAB_OK;
  sis_table_kernel<<<grid_for((int64_t)N * (pop + 1) * 3, 128), 128, 0, (cudaStream_t)stream>>>(N, pop, params, T);
  return check_launch("rmab_sis_table");
}

extern "C" int rmab_bound_nature(int64_t n, const float* x, const float* lo, const float* hi, float* out,
                                 rmab_stream_t stream) {
  RMAB_REQUIRE(n >= 0 && x && lo && hi && out, RMAB_E_BADARG, "rmab_bound_nature: bad argument");
  if (n == 0) return RMAB_OK;
  bound_nature_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(n, x, lo, hi, out);
  return check_launch("rmab_bound_nature");
}
