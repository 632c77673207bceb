// Per-arm policy / value TABLES for hidden width 64 on tcgen05: p(a | s, lambda_e), log p and v(s, lambda_e) for every
// (state, env) of an arm -- the pre-pass of the fused epoch kernel (epoch.cu), which with one-hot inputs only ever needs
// these S * E rows per arm (rmabppo_core.py:198-244 evaluated once per (state, env) instead of once per step).
//
// At h = 64 that pre-pass is a real GEMM (S*E = 3072 rows x 64 x 64 per arm and net at BASELINE configs[2]) and the
// thread-per-row FFMA form of mlp.cuh spills its 64-wide activation vector; here it runs like the forward half of
// ppo_table_umma64.cu: one CTA per arm (16 warps), rows in tiles of 128 = TMEM lanes, warp (q, cq) owns lane quarter q and
// features 16 cq .. 16 cq + 15;  E1: a1 = tanh(W1 x + b1), hi / lo into TMEM;  G1: z2 = a1 . W2^T as three tf32 products
// (A from TMEM, W2 hi / lo K-major in shared memory);  E2: a2 = tanh(z2 + b2), the output layer as per-warp partial dots
// reduced through shared memory, softmax / value, coalesced stores into the statistics block.
// Two CTAs per SM (256 TMEM columns, 64 registers, ~42 KB of shared memory each) fill each other's barriers.
// Same tanh (ex2 + rcp) and the same 3xTF32 split as the update kernels, so rollout and update evaluate the nets alike.
#include "common.cuh"
#include "mlp.cuh"
#include "umma.cuh"

namespace rmab {

namespace {

using namespace umma;

constexpr uint32_t F_W2H = 0, F_W2L = 16384, F_END = 32768;      // K-major 64 x 64 hi / lo tiles
constexpr uint32_t FC_A1H = 0, FC_A1L = 64, FC_D = 128;           // TMEM columns (256 allocated)

struct FwdArgs {
  RmabNetDesc d;
  const float* W;      // packed nets [N, P]
  const float* lamb;   // [E]
  float* stats;        // [N, K, E]
  int K;               // statistics rows per arm
  int row_p, row_lp, row_v;   // first rows of p(a|s), log p(a|s) (policy) / of v(s) (value)
};

}  // namespace

template <int S, bool IS_PI>
__global__ void __launch_bounds__(NTHR, 2) table_fwd_umma64_kernel(FwdArgs g) {
  constexpr int A = 2, OUT = IS_PI ? A : 1, IN = S + 1;
  extern __shared__ unsigned char smraw_f[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  unsigned char* base = smraw_f + ((1024u - (smem_u32(smraw_f) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(base);
  const int E = g.d.n_envs;
  const MlpLayout L(IN, H, OUT);
  float* W2H = reinterpret_cast<float*>(base + F_W2H);
  float* W2L = reinterpret_cast<float*>(base + F_W2L);
  float* wb = reinterpret_cast<float*>(base + F_END);   // [S][H]  (W1[:, s] + b1) * kTanhScale
  float* wl = wb + S * H;                                // [H]     W1[:, lambda] * kTanhScale
  float* b2s = wl + H;                                   // [H]     b2 * kTanhScale
  float* w3 = b2s + H;                                   // [OUT][H]
  float* b3 = w3 + OUT * H;                              // [OUT] (padded to 4)
  float* plog = b3 + 4;                                  // [4 cq][128 rows][OUT] partial output dots
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const int q = warp & 3, cq = warp >> 2, r = 32 * q + lane, f0 = FW * cq;
  const float* Wn = g.W + (int64_t)n * L.P;
  for (int i = tid; i < HH; i += NTHR) {
    const int j = i >> 6, c = i & 63;
    uint32_t hi, lo;
    split_u(Wn[L.oW2 + i], hi, lo);
    W2H[kmaj_off(j, c)] = __uint_as_float(hi);
    W2L[kmaj_off(j, c)] = __uint_as_float(lo);
  }
  for (int i = tid; i < S * H; i += NTHR) {
    const int s = i >> 6, k = i & 63;
    wb[i] = kTanhScale * (Wn[L.oW1 + k * IN + s] + Wn[L.ob1 + k]);
  }
  for (int i = tid; i < H; i += NTHR) {
    wl[i] = kTanhScale * Wn[L.oW1 + i * IN + IN - 1];
    b2s[i] = kTanhScale * Wn[L.ob2 + i];
  }
  for (int i = tid; i < OUT * H; i += NTHR) w3[i] = Wn[L.oW3 + i];
  if (tid < OUT) b3[tid] = Wn[L.ob3 + tid];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  fence_async_smem();   // the W2 tiles were written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t barp = smem_u32(&bar);
  const int M = S * E;
  const int n_tiles = (M + TM - 1) / TM;
  float* st = g.stats + (int64_t)n * g.K * E;

  for (int t = 0; t < n_tiles; ++t) {
    const int m = t * TM + r;
    const bool valid = m < M;
    const int mm = valid ? m : M - 1;
    const int srow = (mm >= E) + (mm >= 2 * E);   // S <= 3
    const int e = mm - srow * E;
    const float lam = g.lamb[e];
    // ---- E1: a1 = tanh(W1 x + b1) for this thread's 16 features; hi / lo -> TMEM (A operand of G1) ----
#pragma unroll
    for (int k4 = 0; k4 < FW / 4; ++k4) {
      const float4 b = *reinterpret_cast<const float4*>(wb + srow * H + f0 + 4 * k4);
      const float4 l = *reinterpret_cast<const float4*>(wl + f0 + 4 * k4);
      const float x0 = tanh_scaled_u(fmaf(l.x, lam, b.x)), x1 = tanh_scaled_u(fmaf(l.y, lam, b.y));
      const float x2 = tanh_scaled_u(fmaf(l.z, lam, b.z)), x3 = tanh_scaled_u(fmaf(l.w, lam, b.w));
      uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
      split_u(x0, h0, l0); split_u(x1, h1, l1); split_u(x2, h2, l2); split_u(x3, h3, l3);
      tmem_st4(tl + FC_A1H + f0 + 4 * k4, h0, h1, h2, h3);
      tmem_st4(tl + FC_A1L + f0 + 4 * k4, l0, l1, l2, l3);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {   // G1: z2 = a1 . W2^T as lo.hi + hi.lo + hi.hi, eight k-steps each
        constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
        const uint32_t wh = desc_lo_k(sbase + F_W2H), wlo = desc_lo_k(sbase + F_W2L);
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t acol = tmem + (term == 0 ? FC_A1L : FC_A1H);
          const uint32_t bt = term == 1 ? wlo : wh;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) umma_ts(tmem + FC_D, acol + 8 * ks, bt + 16 * ks, DESC_HI_K, idesc, (term | ks) != 0);
        }
        umma_commit(barp);
      }
      __syncwarp();
    }
    mbar_wait(barp, t & 1);
    tc_fence_after();
    // ---- E2: a2 = tanh(z2 + b2); this warp's partial of the output layer over its 16 features ----
    {
      uint32_t v[FW];
      tmem_ld16(tl + FC_D + f0, v);
      tmem_wait_ld();
      float part[OUT];
#pragma unroll
      for (int a = 0; a < OUT; ++a) part[a] = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < FW / 4; ++k4) {
        const float4 b = *reinterpret_cast<const float4*>(b2s + f0 + 4 * k4);
        const float a2[4] = {tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4]), kTanhScale, b.x)),
                             tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 1]), kTanhScale, b.y)),
                             tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 2]), kTanhScale, b.z)),
                             tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 3]), kTanhScale, b.w))};
#pragma unroll
        for (int a = 0; a < OUT; ++a) {
          const float4 w = *reinterpret_cast<const float4*>(w3 + a * H + f0 + 4 * k4);
          part[a] = fmaf(w.x, a2[0], part[a]); part[a] = fmaf(w.y, a2[1], part[a]);
          part[a] = fmaf(w.z, a2[2], part[a]); part[a] = fmaf(w.w, a2[3], part[a]);
        }
      }
#pragma unroll
      for (int a = 0; a < OUT; ++a) plog[(cq * TM + r) * OUT + a] = part[a];
    }
    tc_fence_before();
    __syncthreads();   // partial dots visible; the TMEM operands / accumulator are free for the next tile
    if (cq == 0 && valid) {
      float logit[OUT];
#pragma unroll
      for (int a = 0; a < OUT; ++a)   // feature blocks in a fixed order
        logit[a] = b3[a] + (((plog[(0 * TM + r) * OUT + a] + plog[(1 * TM + r) * OUT + a]) + plog[(2 * TM + r) * OUT + a]) +
                            plog[(3 * TM + r) * OUT + a]);
      if (IS_PI) {
        // Categorical(logits) as softmax_small (mlp.cuh): logp = z - max - log(sum exp), p = exp / sum
        const float mx = fmaxf(logit[0], logit[OUT - 1]);
        const float e0 = expf(logit[0] - mx), e1 = expf(logit[OUT - 1] - mx);
        const float sum = e0 + e1, lse = logf(sum);
        st[(int64_t)(g.row_p + srow * A + 0) * E + e] = __fdiv_rn(e0, sum);
        st[(int64_t)(g.row_p + srow * A + 1) * E + e] = __fdiv_rn(e1, sum);
        st[(int64_t)(g.row_lp + srow * A + 0) * E + e] = (logit[0] - mx) - lse;
        st[(int64_t)(g.row_lp + srow * A + 1) * E + e] = (logit[OUT - 1] - mx) - lse;
      } else {
        st[(int64_t)(g.row_v + srow) * E + e] = logit[0];
      }
    }
    // the next tile's E1 only writes TMEM columns G1 has finished reading (the mbarrier above) and plog is rewritten after
    // the next tile's first __syncthreads, which every reader of this tile's plog has passed
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}

static size_t table_fwd_smem(int S, int OUT) {
  return 1024 + F_END + sizeof(float) * (size_t)(S * H + 2 * H + OUT * H + 4 + 4 * TM * OUT);
}

// Entry used by rmab_epoch_table (epoch.cu) for H = 64: fills the p / log p rows (policy net) and the v rows (value net)
// of the statistics block of every arm.
int table_fwd_umma64_run(const RmabNetDesc* net, const float* Wpi, const float* Wv, const float* lamb, float* stats, int K,
                         int row_p, int row_lp, int row_v, cudaStream_t st) {
  FwdArgs g = {};
  g.d = *net; g.lamb = lamb; g.stats = stats; g.K = K; g.row_p = row_p; g.row_lp = row_lp; g.row_v = row_v;
  const int S = net->n_states;
#define FWD_LAUNCH(S_, PI_, W_)                                                                                          \
  {                                                                                                                      \
    const size_t smem = table_fwd_smem(S_, PI_ ? 2 : 1);                                                                 \
    auto kern = table_fwd_umma64_kernel<S_, PI_>;                                                                        \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)               \
      return fail(RMAB_E_LAUNCH, "rmab_epoch_table: cannot reserve %zu B of shared memory for the table pre-pass", smem); \
    g.W = W_;                                                                                                            \
    kern<<<net->n_arms, NTHR, smem, st>>>(g);                                                                            \
    int rc = check_launch("rmab_epoch_table(table pre-pass, tcgen05 h64)");                                              \
    if (rc) return rc;                                                                                                   \
  }
  if (S == 2) {
    FWD_LAUNCH(2, true, Wpi)
    FWD_LAUNCH(2, false, Wv)
  } else {
    FWD_LAUNCH(3, true, Wpi)
    FWD_LAUNCH(3, false, Wv)
  }
#undef FWD_LAUNCH
  return RMAB_OK;
}

}  // namespace rmab
