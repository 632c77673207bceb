// Per-arm MLP helpers: packed-weight layout, shared-memory staging, thread-per-sample forward.
// Packed layout (torch.nn.Linear order, rmabppo_core.py:24-29): W1[h,in] b1[h] W2[h,h] b2[h] W3[out,h] b3[out].
#pragma once
#include "common.cuh"

namespace rmab {

// tanh(x) = 1 - 2 / (2^(2 log2(e) x) + 1): FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA.  Absolute error < 2e-7 (the level of fp32
// rounding of the pre-activation itself); saturates correctly (ex2 -> inf gives 1, ex2 -> 0 gives -1).  The same formula as
// in the update kernels, so rollout and update evaluate the nets identically.
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

constexpr int kMaxA = 4;   // actor outputs supported (reference domains: 2 or 3)
constexpr int kMaxIn = 8;  // input features supported by the thread-per-sample kernels

struct MlpLayout {
  int in, h, out, oW1, ob1, oW2, ob2, oW3, ob3, P;
  __host__ __device__ MlpLayout(int in_, int h_, int out_) : in(in_), h(h_), out(out_) {
    oW1 = 0;
    ob1 = oW1 + h * in;
    oW2 = ob1 + h;
    ob2 = oW2 + h * h;
    oW3 = ob2 + h;
    ob3 = oW3 + out * h;
    P = ob3 + out;
  }
  __host__ __device__ int padded() const { return (P + 3) & ~3; }
};

// Shared-memory image of one net used by the forward helpers.  W2 and W3 rows must start 16-byte
// aligned for the float4 broadcasts, which the packed layout does not guarantee (in_dim is odd for the
// one-hot domains), so the image is re-laid out:  [W1 (h*in) | b1 (h) | pad | W2 (h*h) | b2 | W3 | b3].
struct SmemNet {
  int in, h, out, oW1, ob1, oW2, ob2, oW3, ob3, size;
  __host__ __device__ SmemNet(int in_, int h_, int out_) : in(in_), h(h_), out(out_) {
    oW1 = 0;
    ob1 = h * in;
    oW2 = (ob1 + h + 3) & ~3;
    ob2 = oW2 + h * h;
    oW3 = ob2 + h;  // h % 4 == 0 so this stays aligned
    ob3 = oW3 + out * h;
    size = (ob3 + out + 3) & ~3;
  }
};

// Copy one arm's packed weights [P] from global into the SmemNet image (all threads of the CTA).
__device__ __forceinline__ void stage_net(float* dst, const SmemNet& L, const float* __restrict__ src) {
  const MlpLayout G(L.in, L.h, L.out);
  const int shift = L.oW2 - G.oW2;  // everything from W2 onward moves by the alignment pad
  for (int i = threadIdx.x; i < G.P; i += blockDim.x) dst[i < G.oW2 ? i : i + shift] = src[i];
}

// a1 = tanh(W1 x + b1 [+ add]) for x = onehot(s) ++ [lam]  or  [s/norm, lam]   (rmabppo_core.py:213-218).
// `add` (may be NULL) is an extra first-layer pre-activation term add[j * add_stride]: the MA-RMABPPO agent critic's
// W1_nature . bounded_nature_vector (ma_rmabppo_core.py:312-316), computed by the caller as one dense GEMM.
template <int H>
__device__ __forceinline__ void mlp_layer1(const float* __restrict__ w, const SmemNet& L, int one_hot, int s,
                                           float state_norm, float lam, float (&a1)[H],
                                           const float* __restrict__ add = nullptr, int64_t add_stride = 0) {
  const int in = L.in;
  if (one_hot) {
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const float* row = w + L.oW1 + j * in;
      float z = fmaf(row[in - 1], lam, row[s]) + w[L.ob1 + j];
      if (add) z += add[j * add_stride];
      a1[j] = tanh_fast(z);
    }
  } else {
    const float x0 = __fdiv_rn((float)s, state_norm);
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const float* row = w + L.oW1 + j * in;
      float z = fmaf(row[1], lam, fmaf(row[0], x0, w[L.ob1 + j]));
      if (add) z += add[j * add_stride];
      a1[j] = tanh_fast(z);
    }
  }
}

// out[o] = W3 tanh(W2 a1 + b2) + b3, streaming over the second hidden layer (a2 never materialised).
template <int H>
__device__ __forceinline__ void mlp_tail(const float* __restrict__ w, const SmemNet& L, const float (&a1)[H], int n_out,
                                         float* out) {
  float acc[kMaxA];
#pragma unroll
  for (int o = 0; o < kMaxA; ++o) acc[o] = (o < n_out) ? w[L.ob3 + o] : 0.f;
#pragma unroll 4
  for (int j = 0; j < H; ++j) {
    const float4* row = reinterpret_cast<const float4*>(w + L.oW2 + j * H);
    float z = w[L.ob2 + j];
#pragma unroll
    for (int k4 = 0; k4 < H / 4; ++k4) {
      const float4 q = row[k4];
      z = fmaf(q.x, a1[4 * k4 + 0], z);
      z = fmaf(q.y, a1[4 * k4 + 1], z);
      z = fmaf(q.z, a1[4 * k4 + 2], z);
      z = fmaf(q.w, a1[4 * k4 + 3], z);
    }
    const float a2 = tanh_fast(z);
#pragma unroll
    for (int o = 0; o < kMaxA; ++o)
      if (o < n_out) acc[o] = fmaf(w[L.oW3 + o * H + j], a2, acc[o]);
  }
#pragma unroll
  for (int o = 0; o < kMaxA; ++o)
    if (o < n_out) out[o] = acc[o];
}

// Categorical(logits) normalisation (rmabppo_core.py:79-81): logp = z - max - log(sum exp), p = softmax.
__device__ __forceinline__ void softmax_small(const float* logit, int A, float (&p)[kMaxA], float (&lp)[kMaxA]) {
  float m = logit[0];
#pragma unroll
  for (int k = 1; k < kMaxA; ++k)
    if (k < A) m = fmaxf(m, logit[k]);
  float e[kMaxA], sum = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxA; ++k) {
    e[k] = (k < A) ? expf(logit[k] - m) : 0.f;
    sum += e[k];
  }
  const float lse = logf(sum);
#pragma unroll
  for (int k = 0; k < kMaxA; ++k) {
    p[k] = (k < A) ? __fdiv_rn(e[k], sum) : 0.f;
    lp[k] = (k < A) ? (logit[k] - m) - lse : 0.f;
  }
}

__device__ __forceinline__ int categorical_rt(const float (&p)[kMaxA], int A, float u) {
  // entries >= A are 0 and never become `last`, so the fixed-size draw equals the A-way draw
  return categorical<kMaxA>(p, u);
}

// Rows of the statistics block of one (arm): stats[n][row][e]
template <int S, int A>
struct StatRows {
  static constexpr int SA = S * A;
  // policy update reads rows [0, 3SA+S); value update reads rows [3SA, 3SA+2S)
  static constexpr int kAp = 0, kAm = SA, kLp = 2 * SA, kCntS = 3 * SA, kRetMean = 3 * SA + S, kV = 3 * SA + 2 * S,
                       kCnt = 3 * SA + 3 * S, K = 4 * SA + 3 * S;
};


}  // namespace rmab
