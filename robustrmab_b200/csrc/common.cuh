// Shared device/host helpers for librmab_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/rmab.h"

namespace rmab {

extern thread_local char g_err[256];
extern std::atomic<uint64_t> g_launches;

int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define RMAB_REQUIRE(cond, code, ...) \
  do {                                \
    if (!(cond)) return ::rmab::fail(code, __VA_ARGS__); \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- Philox4x32-10 (Salmon et al. SC'11; same constants as cuRAND / Random123) -----------------
struct Rng {
  uint32_t k0, k1, stream, t, env0, arm0;
  const uint32_t* t_base;   // optional device word added to t (graph replay with advancing counters)
};

__device__ __forceinline__ uint32_t rng_t(const Rng& g) { return g.t + (g.t_base ? __ldg(g.t_base) : 0u); }

__host__ inline Rng make_rng(const RmabRng* r) {
  Rng g;
  g.k0 = (uint32_t)(r->seed & 0xFFFFFFFFull);
  g.k1 = (uint32_t)(r->seed >> 32);
  g.stream = r->stream;
  g.t = r->t;
  g.env0 = r->env0;
  g.arm0 = r->arm0;
  g.t_base = r->t_base;
  return g;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// One Philox block serves the two draws of two consecutive time steps: the env-transition draw (stream ENV) and the
// action draw (stream ACT) of step t are words 2*(t&1) and 2*(t&1)+1 of the block with counter (0, t>>1, env, arm).
// Every other stream uses word 0 of the block with counter (stream, t, env, arm).  u = (word >> 8) * 2^-24.
__device__ __forceinline__ float u24(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

__device__ __forceinline__ uint4 philox_step_block(uint32_t k0, uint32_t k1, uint32_t t, uint32_t env, uint32_t arm) {
  return philox4x32_10(0u, t >> 1, env, arm, k0, k1);
}

__device__ __forceinline__ uint32_t step_word(const uint4& b, uint32_t t, bool act) {
  return (t & 1u) ? (act ? b.w : b.z) : (act ? b.y : b.x);
}

__device__ __forceinline__ float philox_uniform(uint32_t k0, uint32_t k1, uint32_t stream, uint32_t t, uint32_t env,
                                                uint32_t arm) {
  if (stream == RMAB_STREAM_ENV || stream == RMAB_STREAM_ACT)
    return u24(step_word(philox_step_block(k0, k1, t, env, arm), t, stream == RMAB_STREAM_ACT));
  return u24(philox4x32_10(stream, t, env, arm, k0, k1).x);
}

__device__ __forceinline__ float draw_uniform(const Rng& g, const float* __restrict__ u_inject, int64_t idx, int e,
                                              int n) {
  if (u_inject) return u_inject[idx];
  return philox_uniform(g.k0, g.k1, g.stream, rng_t(g), g.env0 + (uint32_t)e, g.arm0 + (uint32_t)n);
}

// Inverse-CDF categorical draw: k = #{j : cdf_j <= u}, cdf accumulated left to right in fp32 with
// round-to-nearest adds (no fma contraction), clamped to the last category with p > 0.
template <int K>
__device__ __forceinline__ int categorical(const float (&p)[K], float u) {
  float cdf = 0.f;
  int k = 0, last = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    cdf = __fadd_rn(cdf, p[j]);
    k += (cdf <= u) ? 1 : 0;
    if (p[j] > 0.f) last = j;
  }
  return k < last ? k : last;
}

// Two categories: the general rule above reduces to one conjunction -- with p[1] > 0 the clamp is min(k, 1) and
// p[0] + p[1] <= u implies p[0] <= u (round-to-nearest addition of a non-negative term is monotone), so k >= 1 exactly when
// p[0] <= u; with p[1] <= 0 (or NaN) the last admissible category is 0.  Same result for every input, NaNs included.
template <>
__device__ __forceinline__ int categorical<2>(const float (&p)[2], float u) {
  return (p[1] > 0.f && p[0] <= u) ? 1 : 0;
}

// One Adam update (torch.optim.Adam defaults form) with every rounding spelled out, so that the stand-alone step and the
// fused GEMM epilogue give the same bits: returns the new parameter, updates m and v.
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, float lerp_w, float beta2, float omb2,
                                             float step_size, float bc2s, float eps) {
  const float mi = __fmaf_rn(lerp_w, __fsub_rn(g, m), m);
  const float vi = __fmaf_rn(__fmul_rn(omb2, g), g, __fmul_rn(v, beta2));
  m = mi;
  v = vi;
  const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2s), eps);
  return __fsub_rn(p, __fmul_rn(step_size, __fdiv_rn(mi, denom)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum (all threads get the result); `red` holds >= 32 floats of smem.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;
}

}  // namespace rmab
