// MLPActorCriticRMAB.step (rmabppo_core.py:198-244) for hidden width 64 on tcgen05: per arm, the E env rows of the policy
// net and then of the value net go through the forward tile of table_fwd_umma64.cu (E1: layer 1 + tanh, hi / lo into TMEM;
// G1: a1 . W2^T as three tf32 products; E2: tanh, output layer as per-warp partial dots), followed by the categorical draw.
// Inputs as rmab_ac_step: one-hot state or the scalar s / state_norm, lambda_e, and for the MA-RMABPPO agent critics the
// caller's first-layer addend extra[n][j][e] (ma_rmabppo_core.py:312-316).  One CTA (16 warps) per arm, two CTAs per SM.
// The thread-per-env FFMA kernel (ac.cu) keeps the small shapes (E < 64) and is the A/B arm (RMAB_AC_FFMA=1).
#include "common.cuh"
#include "mlp.cuh"
#include "umma.cuh"

namespace rmab {

namespace {

using namespace umma;

constexpr uint32_t A_W2H = 0, A_W2L = 16384, A_END = 32768;
constexpr uint32_t AC_A1H = 0, AC_A1L = 64, AC_D = 128;
constexpr int kAcMaxS = kMaxIn - 1;   // one-hot states the layer-1 table holds

struct AcArgs {
  RmabNetDesc d;
  const float* Wpi; const float* Wv;
  const uint8_t* s; const float* lamb; const float* extra;
  Rng g;
  const float* u_inject;
  uint8_t* a_out; float* v_out; float* logp_out; float* p1_out; float* probs_out;
};

}  // namespace

__global__ void __launch_bounds__(NTHR, 2) ac_step_umma64_kernel(AcArgs g) {
  extern __shared__ unsigned char smraw_a[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  unsigned char* base = smraw_a + ((1024u - (smem_u32(smraw_a) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(base);
  const int E = g.d.n_envs, S = g.d.n_states, A = g.d.n_actions;
  const bool one_hot = g.d.one_hot != 0;
  const int IN = (one_hot ? S : 1) + 1;
  float* W2H = reinterpret_cast<float*>(base + A_W2H);
  float* W2L = reinterpret_cast<float*>(base + A_W2L);
  float* wb = reinterpret_cast<float*>(base + A_END);   // one-hot: [S][H] (W1[:, s] + b1) k;  scalar input: [H] b1 k
  float* wa = wb + kAcMaxS * H;                          // [H]  W1[:, 0] k (scalar input)
  float* wl = wa + H;                                    // [H]  W1[:, lambda] k
  float* b2s = wl + H;                                   // [H]  b2 k
  float* w3 = b2s + H;                                   // [kMaxA][H]
  float* b3 = w3 + kMaxA * H;                            // [kMaxA]
  float* plog = b3 + kMaxA;                              // [4 cq][128 rows][kMaxA]
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
  const int q = warp & 3, cq = warp >> 2, r = 32 * q + lane, f0 = FW * cq;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" ::"r"(smem_u32(&tmem_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
  const uint32_t barp = smem_u32(&bar);
  const int n_tiles = (E + TM - 1) / TM;
  uint32_t gt = 0;

  for (int net = 0; net < 2; ++net) {
    const int OUT = net == 0 ? A : 1;
    const MlpLayout L(IN, H, OUT);
    const float* Wn = (net == 0 ? g.Wpi : g.Wv) + (int64_t)n * L.P;
    __syncthreads();   // the previous net's tiles (plog readers, MMAs: all waited for) are done with shared memory
    for (int i = tid; i < HH; i += NTHR) {
      const int j = i >> 6, c = i & 63;
      uint32_t hi, lo;
      split_u(Wn[L.oW2 + i], hi, lo);
      W2H[kmaj_off(j, c)] = __uint_as_float(hi);
      W2L[kmaj_off(j, c)] = __uint_as_float(lo);
    }
    if (one_hot) {
      for (int i = tid; i < S * H; i += NTHR) {
        const int s = i >> 6, k = i & 63;
        wb[i] = kTanhScale * (Wn[L.oW1 + k * IN + s] + Wn[L.ob1 + k]);
      }
    } else {
      for (int i = tid; i < H; i += NTHR) {
        wb[i] = kTanhScale * Wn[L.ob1 + i];
        wa[i] = kTanhScale * Wn[L.oW1 + i * IN];
      }
    }
    for (int i = tid; i < H; i += NTHR) {
      wl[i] = kTanhScale * Wn[L.oW1 + i * IN + IN - 1];
      b2s[i] = kTanhScale * Wn[L.ob2 + i];
    }
    for (int i = tid; i < kMaxA * H; i += NTHR) w3[i] = i < OUT * H ? Wn[L.oW3 + i] : 0.f;
    if (tid < kMaxA) b3[tid] = tid < OUT ? Wn[L.ob3 + tid] : 0.f;
    fence_async_smem();
    __syncthreads();

    for (int t = 0; t < n_tiles; ++t, ++gt) {
      const int e = t * TM + r;
      const bool valid = e < E;
      const int ee = valid ? e : E - 1;
      const int64_t idx = (int64_t)n * E + ee;
      int si = g.s[idx];
      si = si < S ? si : S - 1;
      const float lam = g.lamb[ee];
      const float x0 = one_hot ? 0.f : __fdiv_rn((float)si, g.d.state_norm);
      const float* wrow = one_hot ? wb + si * H : wb;
      const float* ex = (net == 1 && g.extra) ? g.extra + ((int64_t)n * H + f0) * E + ee : nullptr;
      // ---- E1 ----
#pragma unroll
      for (int k4 = 0; k4 < FW / 4; ++k4) {
        const float4 b = *reinterpret_cast<const float4*>(wrow + f0 + 4 * k4);
        const float4 l = *reinterpret_cast<const float4*>(wl + f0 + 4 * k4);
        float z[4] = {fmaf(l.x, lam, b.x), fmaf(l.y, lam, b.y), fmaf(l.z, lam, b.z), fmaf(l.w, lam, b.w)};
        if (!one_hot) {
          const float4 c = *reinterpret_cast<const float4*>(wa + f0 + 4 * k4);
          z[0] = fmaf(c.x, x0, z[0]); z[1] = fmaf(c.y, x0, z[1]); z[2] = fmaf(c.z, x0, z[2]); z[3] = fmaf(c.w, x0, z[3]);
        }
        if (ex) {
#pragma unroll
          for (int j = 0; j < 4; ++j) z[j] = fmaf(kTanhScale, ex[(int64_t)(4 * k4 + j) * E], z[j]);
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split_u(tanh_scaled_u(z[j]), hi[j], lo[j]);
        tmem_st4(tl + AC_A1H + f0 + 4 * k4, hi[0], hi[1], hi[2], hi[3]);
        tmem_st4(tl + AC_A1L + f0 + 4 * k4, lo[0], lo[1], lo[2], lo[3]);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
          const uint32_t wh = desc_lo_k(sbase + A_W2H), wlo = desc_lo_k(sbase + A_W2L);
#pragma unroll
          for (int term = 0; term < 3; ++term) {
            const uint32_t acol = tmem + (term == 0 ? AC_A1L : AC_A1H);
            const uint32_t bt = term == 1 ? wlo : wh;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) umma_ts(tmem + AC_D, acol + 8 * ks, bt + 16 * ks, DESC_HI_K, idesc, (term | ks) != 0);
          }
          umma_commit(barp);
        }
        __syncwarp();
      }
      mbar_wait(barp, gt & 1);
      tc_fence_after();
      // ---- E2 ----
      {
        uint32_t v[FW];
        tmem_ld16(tl + AC_D + f0, v);
        tmem_wait_ld();
        float part[kMaxA] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const float4 b = *reinterpret_cast<const float4*>(b2s + f0 + 4 * k4);
          const float a2[4] = {tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4]), kTanhScale, b.x)),
                               tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 1]), kTanhScale, b.y)),
                               tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 2]), kTanhScale, b.z)),
                               tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 3]), kTanhScale, b.w))};
#pragma unroll
          for (int a = 0; a < kMaxA; ++a) {
            if (a < OUT) {
              const float4 w = *reinterpret_cast<const float4*>(w3 + a * H + f0 + 4 * k4);
              part[a] = fmaf(w.x, a2[0], part[a]); part[a] = fmaf(w.y, a2[1], part[a]);
              part[a] = fmaf(w.z, a2[2], part[a]); part[a] = fmaf(w.w, a2[3], part[a]);
            }
          }
        }
        *reinterpret_cast<float4*>(plog + (cq * TM + r) * kMaxA) = make_float4(part[0], part[1], part[2], part[3]);
      }
      tc_fence_before();
      __syncthreads();
      if (cq == 0 && valid) {
        float logit[kMaxA];
#pragma unroll
        for (int a = 0; a < kMaxA; ++a)
          logit[a] = b3[a] + (((plog[(0 * TM + r) * kMaxA + a] + plog[(1 * TM + r) * kMaxA + a]) + plog[(2 * TM + r) * kMaxA + a]) +
                              plog[(3 * TM + r) * kMaxA + a]);
        if (net == 0) {
          float p[kMaxA], lp[kMaxA];
          softmax_small(logit, A, p, lp);
          const float u = draw_uniform(g.g, g.u_inject, idx, ee, n);
          const int act = categorical_rt(p, A, u);
          g.a_out[idx] = (uint8_t)act;
          float lpa = lp[0];
#pragma unroll
          for (int k = 1; k < kMaxA; ++k) lpa = (act == k) ? lp[k] : lpa;
          g.logp_out[idx] = lpa;
          if (g.p1_out) g.p1_out[idx] = p[1];
          if (g.probs_out)
            for (int k = 0; k < A; ++k) g.probs_out[((int64_t)n * A + k) * E + ee] = p[k];
        } else {
          g.v_out[idx] = logit[0];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" ::"r"(tmem) : "memory");
}

// Entry used by rmab_ac_step (ac.cu) for hidden 64 and E >= 64.
int ac_step_umma64_run(const RmabNetDesc* net, const float* Wpi, const float* Wv, const uint8_t* s, const float* lamb,
                       const float* extra, const Rng& rng, const float* u_inject, uint8_t* a, float* v, float* logp, float* p1,
                       float* probs, cudaStream_t st) {
  AcArgs g = {};
  g.d = *net; g.Wpi = Wpi; g.Wv = Wv; g.s = s; g.lamb = lamb; g.extra = extra; g.g = rng; g.u_inject = u_inject;
  g.a_out = a; g.v_out = v; g.logp_out = logp; g.p1_out = p1; g.probs_out = probs;
  const size_t smem = 1024 + A_END + sizeof(float) * (size_t)(kAcMaxS * H + 3 * H + kMaxA * H + kMaxA + 4 * TM * kMaxA);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(ac_step_umma64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return fail(RMAB_E_LAUNCH, "rmab_ac_step: cannot reserve %zu B of shared memory", smem);
    attr = true;
  }
  ac_step_umma64_kernel<<<net->n_arms, NTHR, smem, st>>>(g);
  return check_launch("rmab_ac_step(tcgen05 h64)");
}

}  // namespace rmab
