// Actor-critic kernels: per-arm MLP forward + categorical sample (ac_step), fused table-domain rollout,
// lambda-net forward / SGD.
//
// Reference behaviour: robust_rmab/algos/rmabppo/rmabppo_core.py  MLPActorCriticRMAB.step :198-244,
// mlp :24-29, MLPLambdaNet :108-115; agent_oracle.py compute_loss_lambda :360-376, rollout loop :506-570.
// One CTA per arm: the arm's packed actor + critic weights are staged once in shared memory and every
// thread (one env each) reads them as warp-uniform float4 broadcasts.
#include <stdlib.h>
#include "common.cuh"
#include "mlp.cuh"

namespace rmab {

// ------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(128) ac_step_kernel(RmabNetDesc d, const float* __restrict__ Wpi,
                                                      const float* __restrict__ Wv, const uint8_t* __restrict__ s,
                                                      const float* __restrict__ lamb,
                                                      const float* __restrict__ z1_extra, Rng g,
                                                      const float* __restrict__ u_inject, uint8_t* __restrict__ a_out,
                                                      float* __restrict__ v_out, float* __restrict__ logp_out,
                                                      float* __restrict__ p1_out, float* __restrict__ probs_out) {
  extern __shared__ __align__(16) float smem[];
  const int E = d.n_envs, S = d.n_states, A = d.n_actions;
  const int in_dim = (d.one_hot ? S : 1) + 1;
  const SmemNet Lp(in_dim, H, A), Lv(in_dim, H, 1);
  const MlpLayout Gp(in_dim, H, A), Gv(in_dim, H, 1);
  float* sp = smem;
  float* sv = smem + Lp.size;
  const int n = blockIdx.x;
  stage_net(sp, Lp, Wpi + (int64_t)n * Gp.P);
  stage_net(sv, Lv, Wv + (int64_t)n * Gv.P);
  __syncthreads();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const int64_t idx = (int64_t)n * E + e;
    int si = s[idx];
    si = si < S ? si : S - 1;
    const float lam = lamb[e];
    float logit[kMaxA], vout[kMaxA], a1[H];
    mlp_layer1<H>(sp, Lp, d.one_hot, si, d.state_norm, lam, a1);
    mlp_tail<H>(sp, Lp, a1, A, logit);
    // MA agent critic: first-layer addend z1_extra[n][j][e] = (W1_nature[n] . bounded nature vector of env e)[j]
    mlp_layer1<H>(sv, Lv, d.one_hot, si, d.state_norm, lam, a1, z1_extra ? z1_extra + (int64_t)n * H * E + e : nullptr, E);
    mlp_tail<H>(sv, Lv, a1, 1, vout);
    float p[kMaxA], lp[kMaxA];
    softmax_small(logit, A, p, lp);
    const float u = draw_uniform(g, u_inject, idx, e, n);
    const int act = categorical_rt(p, A, u);
    a_out[idx] = (uint8_t)act;
    v_out[idx] = vout[0];
    float lpa = lp[0];
#pragma unroll
    for (int k = 1; k < kMaxA; ++k) lpa = (act == k) ? lp[k] : lpa;
    logp_out[idx] = lpa;
    if (p1_out) p1_out[idx] = p[1];
    if (probs_out)
      for (int k = 0; k < A; ++k) probs_out[((int64_t)n * A + k) * E + e] = p[k];
  }
}

// ------------------------------------------------------------------------------------------------
// Fused rollout for the one-hot table domains with a per-arm nature table: the nets only ever see
// (state, lambda_e), so each (arm, env) thread evaluates them once per state and then walks T steps of
// table lookups + two Philox draws per step.  HBM traffic is the trajectory write (10 B per arm-step).
__device__ __forceinline__ int ce_next_d(int s, int a, float p, float u) {
  float pr[2];
  if (s == 0) { pr[0] = 0.5f; pr[1] = 0.5f; }
  else if (a == 0) { pr[0] = 1.f; pr[1] = 0.f; }
  else { pr[0] = __fsub_rn(1.f, p); pr[1] = p; }
  return categorical<2>(pr, u);
}

__device__ __forceinline__ int armman_next_d(int s, int a, float p, float u) {
  const bool to0 = (s == 0) || (s == 1 && a == 1);
  float pr[3];
  pr[0] = to0 ? p : 0.f;
  pr[1] = __fsub_rn(1.f, p);
  pr[2] = to0 ? 0.f : p;
  return categorical<3>(pr, u);
}

template <int S>
__device__ __forceinline__ float pick(const float (&t)[S], int s) {
  float v = t[0];
#pragma unroll
  for (int k = 1; k < S; ++k) v = (s == k) ? t[k] : v;
  return v;
}

template <int H, int DOMAIN>
__global__ void __launch_bounds__(128) rollout_table_kernel(RmabNetDesc d, int T, const float* __restrict__ Wpi,
                                                            const float* __restrict__ Wv,
                                                            const float* __restrict__ lamb,
                                                            const float* __restrict__ nature,
                                                            const float* __restrict__ ranges, uint32_t k0, uint32_t k1,
                                                            uint32_t t_act0, uint32_t t_env0, uint32_t env0,
                                                            uint32_t arm0, uint8_t* __restrict__ s_traj,
                                                            uint8_t* __restrict__ a_out, float* __restrict__ logp_out,
                                                            float* __restrict__ val_out,
                                                            float* __restrict__ last_val) {
  constexpr int S = DOMAIN == RMAB_DOMAIN_CE ? 2 : 3;
  constexpr int A = 2;
  extern __shared__ __align__(16) float smem[];
  const int E = d.n_envs;
  const SmemNet Lp(S + 1, H, A), Lv(S + 1, H, 1);
  const MlpLayout Gp(S + 1, H, A), Gv(S + 1, H, 1);
  float* sp = smem;
  float* sv = smem + Lp.size;
  const int n = blockIdx.x;
  stage_net(sp, Lp, Wpi + (int64_t)n * Gp.P);
  stage_net(sv, Lv, Wv + (int64_t)n * Gv.P);
  __syncthreads();
  // clamped nature parameter per (state, action) of this arm (:867-874 / :583-592)
  float pn0[S], pn1[S];
#pragma unroll
  for (int si = 0; si < S; ++si) {
    if (DOMAIN == RMAB_DOMAIN_CE) {
      const float p = fminf(fmaxf(__ldg(nature + n), __ldg(ranges + 2 * (int64_t)n)), __ldg(ranges + 2 * (int64_t)n + 1));
      pn0[si] = p;
      pn1[si] = p;
    } else {
      const int64_t c0 = ((int64_t)n * S + si) * A;
      pn0[si] = fminf(fmaxf(__ldg(nature + c0), __ldg(ranges + 2 * c0)), __ldg(ranges + 2 * c0 + 1));
      pn1[si] = fminf(fmaxf(__ldg(nature + c0 + 1), __ldg(ranges + 2 * c0 + 2)), __ldg(ranges + 2 * c0 + 3));
    }
  }
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float lam = lamb[e];
    float p0t[S], p1t[S], lp0t[S], lp1t[S], vt[S];
#pragma unroll
    for (int si = 0; si < S; ++si) {
      float a1[H], logit[kMaxA], vout[kMaxA], p[kMaxA], lp[kMaxA];
      mlp_layer1<H>(sp, Lp, 1, si, 1.f, lam, a1);
      mlp_tail<H>(sp, Lp, a1, A, logit);
      mlp_layer1<H>(sv, Lv, 1, si, 1.f, lam, a1);
      mlp_tail<H>(sv, Lv, a1, 1, vout);
      softmax_small(logit, A, p, lp);
      p0t[si] = p[0]; p1t[si] = p[1]; lp0t[si] = lp[0]; lp1t[si] = lp[1]; vt[si] = vout[0];
    }
    const int64_t traj = (int64_t)n * (T + 1) * E + e;  // s_traj[n, 0, e]
    const int64_t step = (int64_t)n * T * E + e;        // x[n, 0, e]
    int si = s_traj[traj];
    si = si < S ? si : S - 1;
    const uint32_t ge = env0 + (uint32_t)e, gn = arm0 + (uint32_t)n;
    for (int t = 0; t < T; ++t) {
      float pr[2] = {pick<S>(p0t, si), pick<S>(p1t, si)};
      const float ua = philox_uniform(k0, k1, RMAB_STREAM_ACT, t_act0 + (uint32_t)t, ge, gn);
      const int act = categorical<2>(pr, ua);
      const int64_t o = step + (int64_t)t * E;
      a_out[o] = (uint8_t)act;
      logp_out[o] = act ? pick<S>(lp1t, si) : pick<S>(lp0t, si);
      val_out[o] = pick<S>(vt, si);
      const float p = act ? pick<S>(pn1, si) : pick<S>(pn0, si);
      const float ue = philox_uniform(k0, k1, RMAB_STREAM_ENV, t_env0 + (uint32_t)t, ge, gn);
      si = DOMAIN == RMAB_DOMAIN_CE ? ce_next_d(si, act, p, ue) : armman_next_d(si, act, p, ue);
      s_traj[traj + (int64_t)(t + 1) * E] = (uint8_t)si;
    }
    last_val[(int64_t)n * E + e] = pick<S>(vt, si);
  }
}

// ------------------------------------------------------------------------------------------------
// Lambda net N -> 8 -> 8 -> 1.  One CTA per 32 envs; warps stride over the arms, lane = env, so the
// state reads s[n, e0..e0+31] are coalesced and W1[j, n] is a warp-uniform broadcast.
constexpr int kLamH = 8;

__global__ void __launch_bounds__(1024) lambda_forward_kernel(int E, int N, const float* __restrict__ params,
                                                              int n_total, int arm0, const uint8_t* __restrict__ s,
                                                              float obs_scale, const float* __restrict__ h1_in,
                                                              int n_split, float* __restrict__ h1_partial,
                                                              float* __restrict__ lamb) {
  __shared__ float red[32][kLamH][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const float* W1 = params;
  const float* b1 = W1 + (int64_t)kLamH * n_total;
  const float* W2 = b1 + kLamH;
  const float* b2 = W2 + kLamH * kLamH;
  const float* W3 = b2 + kLamH;
  const float* b3 = W3 + kLamH;
  float h[kLamH];
#pragma unroll
  for (int j = 0; j < kLamH; ++j) h[j] = 0.f;
  if (h1_in) {
    if (blockDim.x >= 256) {
      // split sums spread over the block: thread t adds (env t >> 3 of this block, feature t & 7) over the splits in split
      // order (deterministic, and the same order as the one-warp form below); consecutive threads read consecutive words
      const int le = threadIdx.x >> 3, j = threadIdx.x & 7, ee = blockIdx.x * 32 + le;
      float acc = 0.f;
      if (threadIdx.x < 256 && ee < E)
#pragma unroll 8
        for (int k = 0; k < n_split; ++k) acc += __ldg(h1_in + ((int64_t)k * E + ee) * kLamH + j);
      if (threadIdx.x < 256) red[0][j][le] = acc;
      __syncthreads();
      if (w == 0 && e < E)
#pragma unroll
        for (int jj = 0; jj < kLamH; ++jj) h[jj] = red[0][jj][lane];
    } else if (w == 0 && e < E)
#pragma unroll 4
      for (int k = 0; k < n_split; ++k) {  // partial sums of the arm splits, added in split order (deterministic)
        static_assert(kLamH == 8, "two float4 per (split, env)");
        const float4* p4 = reinterpret_cast<const float4*>(h1_in + ((int64_t)k * E + e) * kLamH);
        const float4 x = __ldg(p4), y = __ldg(p4 + 1);
        h[0] += x.x; h[1] += x.y; h[2] += x.z; h[3] += x.w;
        h[4] += y.x; h[5] += y.y; h[6] += y.z; h[7] += y.w;
      }
  } else {
    for (int n = w; n < N; n += nw) {
      const float x = (e < E) ? (float)s[(int64_t)n * E + e] * obs_scale : 0.f;
#pragma unroll
      for (int j = 0; j < kLamH; ++j) h[j] = fmaf(__ldg(W1 + (int64_t)j * n_total + arm0 + n), x, h[j]);
    }
#pragma unroll
    for (int j = 0; j < kLamH; ++j) red[w][j][lane] = h[j];
    __syncthreads();
    if (w == 0) {
#pragma unroll
      for (int j = 0; j < kLamH; ++j) {
        float acc = 0.f;
        for (int ww = 0; ww < nw; ++ww) acc += red[ww][j][lane];  // fixed order: deterministic
        h[j] = acc;
      }
    }
  }
  if (w != 0 || e >= E) return;
  if (h1_partial)
#pragma unroll
    for (int j = 0; j < kLamH; ++j) h1_partial[(int64_t)e * kLamH + j] = h[j];
  if (!lamb) return;
  float a1[kLamH], out = __ldg(b3);
#pragma unroll
  for (int j = 0; j < kLamH; ++j) a1[j] = tanhf(h[j] + __ldg(b1 + j));
#pragma unroll
  for (int j = 0; j < kLamH; ++j) {
    float z = __ldg(b2 + j);
#pragma unroll
    for (int k = 0; k < kLamH; ++k) z = fmaf(__ldg(W2 + j * kLamH + k), a1[k], z);
    out = fmaf(__ldg(W3 + j), tanhf(z), out);
  }
  lamb[e] = out;
}

// First-layer partial sums of one arm split: grid (ceil(E/32), n_split), 8 warps stride the split's arms, lane = env.
// part[split][e][j] = sum_{n in split} W1[j, arm0 + n] * s[n, e] * scale, warps merged in a fixed order.
__global__ void __launch_bounds__(256) lambda_partial_kernel(int E, int N, const float* __restrict__ W1, int n_total,
                                                             int arm0, const uint8_t* __restrict__ s, float obs_scale,
                                                             int chunk, float* __restrict__ part) {
  __shared__ float red[8][kLamH][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const int n0 = blockIdx.y * chunk, n1 = min(N, n0 + chunk);
  float h[kLamH];
#pragma unroll
  for (int j = 0; j < kLamH; ++j) h[j] = 0.f;
  for (int n = n0 + w; n < n1; n += 8) {
    const float x = (e < E) ? (float)s[(int64_t)n * E + e] * obs_scale : 0.f;
#pragma unroll
    for (int j = 0; j < kLamH; ++j) h[j] = fmaf(__ldg(W1 + (int64_t)j * n_total + arm0 + n), x, h[j]);
  }
#pragma unroll
  for (int j = 0; j < kLamH; ++j) red[w][j][lane] = h[j];
  __syncthreads();
  if (w == 0 && e < E) {
#pragma unroll
    for (int j = 0; j < kLamH; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) acc += red[ww][j][lane];
      part[((int64_t)blockIdx.y * E + e) * kLamH + j] = acc;
    }
  }
}

// Arm splits of the first layer: 64 arms per CTA (8 per warp) keeps the dependent-load chain short -- the kernel is latency
// bound (256 arms per CTA: 40 us at 4098 arms x 256 envs) -- up to 128 splits (workspace 4 KB x E).
static int lambda_splits(int N) {
  const int k = N / 64;
  return k < 1 ? 1 : (k > 128 ? 128 : k);
}

// Backward of mean_e lamb_e * coef_e through the two small layers (single CTA, thread per env, fixed-order
// reductions), writes delta1[E,8] for the W1 kernel and applies SGD to b1, W2, b2, W3, b3.
__global__ void __launch_bounds__(256) lambda_sgd_small_kernel(int E, float* __restrict__ params, int n_total,
                                                               const float* __restrict__ h1,
                                                               const float* __restrict__ coef, float lr,
                                                               float* __restrict__ delta1,
                                                               float* __restrict__ grad_out) {
  constexpr int NS = kLamH + kLamH * kLamH + kLamH + kLamH + 1;  // b1 W2 b2 W3 b3
  __shared__ float g[NS];
  __shared__ float wl[NS];
  float* small = params + (int64_t)kLamH * n_total;
  for (int i = threadIdx.x; i < NS; i += blockDim.x) { g[i] = 0.f; wl[i] = small[i]; }
  __syncthreads();
  const float* b1 = wl;
  const float* W2 = b1 + kLamH;
  const float* b2 = W2 + kLamH * kLamH;
  const float* W3 = b2 + kLamH;
  // Envs are processed in ascending order by ONE warp-strided pass with a fixed-order accumulation:
  // thread t handles envs t, t+blockDim, ...; per-thread partials are then summed in thread order.
  float acc[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) acc[i] = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float a1[kLamH], a2[kLamH], d2[kLamH];
#pragma unroll
    for (int j = 0; j < kLamH; ++j) a1[j] = tanhf(h1[(int64_t)e * kLamH + j] + b1[j]);
#pragma unroll
    for (int j = 0; j < kLamH; ++j) {
      float z = b2[j];
#pragma unroll
      for (int k = 0; k < kLamH; ++k) z = fmaf(W2[j * kLamH + k], a1[k], z);
      a2[j] = tanhf(z);
    }
    const float d3 = coef[e] / (float)E;
    acc[NS - 1] += d3;  // b3
#pragma unroll
    for (int j = 0; j < kLamH; ++j) {
      acc[kLamH + kLamH * kLamH + kLamH + j] += d3 * a2[j];  // W3
      d2[j] = d3 * W3[j] * (1.f - a2[j] * a2[j]);
      acc[kLamH + kLamH * kLamH + j] += d2[j];  // b2
#pragma unroll
      for (int k = 0; k < kLamH; ++k) acc[kLamH + j * kLamH + k] += d2[j] * a1[k];  // W2
    }
#pragma unroll
    for (int k = 0; k < kLamH; ++k) {
      float d = 0.f;
#pragma unroll
      for (int j = 0; j < kLamH; ++j) d = fmaf(d2[j], W2[j * kLamH + k], d);
      d *= (1.f - a1[k] * a1[k]);
      delta1[(int64_t)e * kLamH + k] = d;
      acc[k] += d;  // b1
    }
  }
  // deterministic reduction: a fixed butterfly inside every warp, then the warps in order (threads without an env hold
  // zeros).  Round 1 added the threads into smem one after another: blockDim.x block barriers, 125 us.
  __shared__ float gw[8][NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) {
    const float x = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) gw[threadIdx.x >> 5][i] = x;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS; i += blockDim.x) {
    float x = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += gw[w][i];
    g[i] = x;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS; i += blockDim.x) {
    small[i] = wl[i] - lr * g[i];
    if (grad_out) grad_out[(int64_t)kLamH * n_total + i] = g[i];
  }
}

// dW1[k, n] = sum_e delta1[e, k] * s[n, e] * scale ; one warp per arm.
__global__ void __launch_bounds__(256) lambda_sgd_w1_kernel(int E, int N, float* __restrict__ params, int n_total,
                                                            int arm0, const uint8_t* __restrict__ s, float obs_scale,
                                                            const float* __restrict__ delta1, float lr,
                                                            float* __restrict__ grad_out) {
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  float acc[kLamH];
#pragma unroll
  for (int k = 0; k < kLamH; ++k) acc[k] = 0.f;
  for (int e = lane; e < E; e += 32) {
    const float x = (float)s[(int64_t)n * E + e] * obs_scale;
#pragma unroll
    for (int k = 0; k < kLamH; ++k) acc[k] = fmaf(delta1[(int64_t)e * kLamH + k], x, acc[k]);
  }
#pragma unroll
  for (int k = 0; k < kLamH; ++k) {
    const float gsum = warp_sum(acc[k]);
    if (lane == 0) {
      const int64_t o = (int64_t)k * n_total + arm0 + n;
      params[o] -= lr * gsum;
      if (grad_out) grad_out[o] = gsum;
    }
  }
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : -1;
}

}  // namespace rmab

using namespace rmab;

static int check_net(const RmabNetDesc* d, const char* who) {
  RMAB_REQUIRE(d, RMAB_E_BADARG, "%s: null net descriptor", who);
  RMAB_REQUIRE(d->hidden == 8 || d->hidden == 16 || d->hidden == 32 || d->hidden == 64, RMAB_E_SHAPE,
               "%s: hidden width %d unsupported (8, 16, 32 or 64; two hidden layers)", who, d->hidden);
  RMAB_REQUIRE(d->n_actions >= 2 && d->n_actions <= kMaxA, RMAB_E_SHAPE, "%s: n_actions %d not in [2,%d]", who,
               d->n_actions, kMaxA);
  RMAB_REQUIRE(d->n_states >= 1 && d->n_states <= 256 && d->n_envs > 0 && d->n_arms >= 0, RMAB_E_SHAPE,
               "%s: bad shape N=%d E=%d S=%d", who, d->n_arms, d->n_envs, d->n_states);
  RMAB_REQUIRE(!d->one_hot || d->n_states + 1 <= kMaxIn, RMAB_E_SHAPE, "%s: one-hot input with S=%d exceeds %d inputs",
               who, d->n_states, kMaxIn);
  RMAB_REQUIRE(d->one_hot || d->state_norm != 0.f, RMAB_E_BADARG, "%s: state_norm must be non-zero", who);
  return RMAB_OK;
}

namespace rmab {
int ac_step_umma64_run(const RmabNetDesc* net, const float* Wpi, const float* Wv, const uint8_t* s, const float* lamb,
                       const float* extra, const Rng& rng, const float* u_inject, uint8_t* a, float* v, float* logp, float* p1,
                       float* probs, cudaStream_t st);   // ac_umma64.cu
}

extern "C" int rmab_ac_step(const RmabNetDesc* net, const float* Wpi, const float* Wv, const uint8_t* s,
                            const float* lamb, const float* extra, const RmabRng* rng, const float* u_inject,
                            uint8_t* a, float* v, float* logp, float* p1, float* probs, rmab_stream_t stream) {
  int rc = check_net(net, "rmab_ac_step");
  if (rc) return rc;
  RMAB_REQUIRE((net->n_extra == 0) == (extra == nullptr), RMAB_E_SHAPE,
               "rmab_ac_step: the critic addend `extra` [N,h,E] must be given exactly when net->n_extra > 0");
  RMAB_REQUIRE(Wpi && Wv && s && lamb && a && v && logp && (rng || u_inject), RMAB_E_BADARG, "rmab_ac_step: null pointer");
  if (net->n_arms == 0) return RMAB_OK;
  RmabRng z = {0, RMAB_STREAM_ACT, 0, 0, 0};
  const Rng g = make_rng(rng ? rng : &z);
  const int in_dim = (net->one_hot ? net->n_states : 1) + 1;
  const size_t smem = sizeof(float) * (SmemNet(in_dim, net->hidden, net->n_actions).size + SmemNet(in_dim, net->hidden, 1).size);
  cudaStream_t st = (cudaStream_t)stream;
  // hidden 64 with at least half a 128-row tile of envs per arm: the tcgen05 forward (ac_umma64.cu); RMAB_AC_FFMA=1 = A/B arm
  if (net->hidden == 64 && net->n_envs >= 64 && getenv("RMAB_AC_FFMA") == nullptr)
    return ac_step_umma64_run(net, Wpi, Wv, s, lamb, extra, g, u_inject, a, v, logp, p1, probs, st);
#define LAUNCH(H)                                                                                              \
  {                                                                                                            \
    RMAB_REQUIRE(set_smem(ac_step_kernel<H>, smem) == 0, RMAB_E_LAUNCH, "rmab_ac_step: cannot reserve %zu B smem", smem); \
    ac_step_kernel<H><<<net->n_arms, 128, smem, st>>>(*net, Wpi, Wv, s, lamb, extra, g, u_inject, a, v, logp, p1, probs); \
  }
  switch (net->hidden) {
    case 8: LAUNCH(8) break;
    case 16: LAUNCH(16) break;
    case 32: LAUNCH(32) break;
    default: LAUNCH(64) break;
  }
#undef LAUNCH
  return check_launch("rmab_ac_step");
}

extern "C" int rmab_rollout_table(int domain, const RmabNetDesc* net, int T, const float* Wpi, const float* Wv,
                                  const float* lamb, const float* nature, const float* ranges, uint64_t seed,
                                  uint32_t t_act0, uint32_t t_env0, uint32_t env0, uint32_t arm0, uint8_t* s_traj,
                                  uint8_t* a, float* logp, float* val, float* last_val, rmab_stream_t stream) {
  int rc = check_net(net, "rmab_rollout_table");
  if (rc) return rc;
  RMAB_REQUIRE(domain == RMAB_DOMAIN_CE || domain == RMAB_DOMAIN_ARMMAN, RMAB_E_BADARG,
               "rmab_rollout_table: domain %d is not a table domain", domain);
  const int S = domain == RMAB_DOMAIN_CE ? 2 : 3;
  RMAB_REQUIRE(net->one_hot && net->n_states == S && net->n_actions == 2 && net->n_extra == 0, RMAB_E_SHAPE,
               "rmab_rollout_table: needs one-hot S=%d, A=2 nets", S);
  RMAB_REQUIRE(T > 0 && Wpi && Wv && lamb && nature && ranges && s_traj && a && logp && val && last_val, RMAB_E_BADARG,
               "rmab_rollout_table: null pointer");
  if (net->n_arms == 0) return RMAB_OK;
  const size_t smem = sizeof(float) * (SmemNet(S + 1, net->hidden, 2).size + SmemNet(S + 1, net->hidden, 1).size);
  const uint32_t k0 = (uint32_t)(seed & 0xFFFFFFFFull), k1 = (uint32_t)(seed >> 32);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(H, D)                                                                                               \
  {                                                                                                                \
    RMAB_REQUIRE(set_smem(rollout_table_kernel<H, D>, smem) == 0, RMAB_E_LAUNCH, "rmab_rollout_table: smem %zu B", smem); \
    rollout_table_kernel<H, D><<<net->n_arms, 128, smem, st>>>(*net, T, Wpi, Wv, lamb, nature, ranges, k0, k1, t_act0, \
                                                               t_env0, env0, arm0, s_traj, a, logp, val, last_val); \
  }
#define LAUNCH_H(D)              \
  switch (net->hidden) {         \
    case 8: LAUNCH(8, D) break;  \
    case 16: LAUNCH(16, D) break; \
    case 32: LAUNCH(32, D) break; \
    default: LAUNCH(64, D) break; \
  }
  if (domain == RMAB_DOMAIN_CE) { LAUNCH_H(RMAB_DOMAIN_CE) } else { LAUNCH_H(RMAB_DOMAIN_ARMMAN) }
#undef LAUNCH_H
#undef LAUNCH
  return check_launch("rmab_rollout_table");
}

extern "C" size_t rmab_lambda_ws_bytes(int E, int N) { return (size_t)lambda_splits(N) * (size_t)E * kLamH * sizeof(float); }

extern "C" int rmab_lambda_forward(int E, int N, const float* params, int n_total, int arm0, const uint8_t* s,
                                   float obs_scale, const float* h1_in, float* h1_partial, float* lamb, void* ws,
                                   size_t ws_bytes, rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N >= 0 && params && n_total >= N && arm0 >= 0 && arm0 + N <= n_total && (s || h1_in) &&
                   (lamb || h1_partial),
               RMAB_E_BADARG, "rmab_lambda_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int nsp = lambda_splits(N);
  if (!h1_in && ws && nsp > 1 && ws_bytes >= rmab_lambda_ws_bytes(E, N)) {
    // arm-split path: the first layer is spread over (E/32) x nsp CTAs, then one small kernel adds the splits in
    // order and runs the 8 -> 8 -> 1 tail (same summation order for a given N, so results are reproducible)
    const int chunk = (N + nsp - 1) / nsp;
    lambda_partial_kernel<<<dim3((E + 31) / 32, nsp), 256, 0, st>>>(E, N, params, n_total, arm0, s, obs_scale, chunk,
                                                                   (float*)ws);
    int rc = check_launch("rmab_lambda_forward(partial)");
    if (rc) return rc;
    lambda_forward_kernel<<<(E + 31) / 32, 256, 0, st>>>(E, N, params, n_total, arm0, s, obs_scale, (const float*)ws, nsp,
                                                       h1_partial, lamb);
    return check_launch("rmab_lambda_forward(tail)");
  }
  const int threads = h1_in ? 32 : 1024;
  lambda_forward_kernel<<<(E + 31) / 32, threads, 0, st>>>(E, N, params, n_total, arm0, s, obs_scale, h1_in, 1, h1_partial,
                                                          lamb);
  return check_launch("rmab_lambda_forward");
}

extern "C" int rmab_lambda_sgd(int E, int N, float* params, int n_total, int arm0, const uint8_t* s, float obs_scale,
                               const float* h1, const float* coef, float lr, float* delta1_ws, float* grad_out,
                               rmab_stream_t stream) {
  RMAB_REQUIRE(E > 0 && N >= 0 && params && s && h1 && coef && delta1_ws && n_total >= N && arm0 >= 0 &&
                   arm0 + N <= n_total,
               RMAB_E_BADARG, "rmab_lambda_sgd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  lambda_sgd_small_kernel<<<1, 256, 0, st>>>(E, params, n_total, h1, coef, lr, delta1_ws, grad_out);
  int rc = check_launch("rmab_lambda_sgd(small)");
  if (rc || N == 0) return rc;
  lambda_sgd_w1_kernel<<<(N + 7) / 8, 256, 0, st>>>(E, N, params, n_total, arm0, s, obs_scale, delta1_ws, lr, grad_out);
  return check_launch("rmab_lambda_sgd(W1)");
}
