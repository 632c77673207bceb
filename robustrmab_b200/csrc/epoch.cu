// Fused epoch front half for the one-hot table domains (counterexample, ARMMAN) under a per-arm nature table:
// rollout (actor sample + env transition) -> GAE / rewards-to-go -> per-arm advantage normalisation -> the
// per-(env, state, action) sufficient statistics of the PPO losses.
//
// Reference behaviour: robust_rmab/algos/rmabppo/agent_oracle.py rollout loop :506-570 (ac.step, env.step,
// buf.store, bootstrap ac.step, finish_path), RMABPPO_Buffer.finish_path :90-139, get :144-161,
// compute_loss_pi :285-322, compute_loss_v :325-339; env laws bandit_env_robust.py:860-893 / :575-634.
//
// Why statistics are enough: with one-hot inputs a net only ever sees (state, lambda_env), so for a fixed
// (arm, env) the new log-probabilities, entropies and values take S distinct values, and
//   sum_t min(r_t A_t, clip(r_t) A_t) = sum_{s,a} [ min(r_sa, 1+c) * Aplus_sa + max(r_sa, 1-c) * Aminus_sa ]
//   sum_t (v(s_t) - ret_t)^2          = sum_s   [ n_s (v_s - mean_ret_s)^2 ] + const
// with Aplus/Aminus the sums of the positive / negative normalised advantages of the samples that visited
// (s, a).  The update kernel (ppo_table.cu) then runs every Adam iteration on E*S rows instead of E*T samples.
//
// One CTA walks arms blockIdx.x, blockIdx.x + gridDim.x, ...; one thread per env.  HBM traffic per arm-step:
// the trajectory write (s' 1 B + a 1 B) and its two L2-resident re-reads by the same thread; the statistics
// are (4*S*A + 3*S) floats per (arm, env), i.e. < 1 B per arm-step at T = 100.
#include <stdlib.h>
#include "common.cuh"
#include "mlp.cuh"

namespace rmab {

template <int S>
__device__ __forceinline__ float pick_f(const float (&t)[S], int s) {
  float v = t[0];
#pragma unroll
  for (int k = 1; k < S; ++k) v = (s == k) ? t[k] : v;
  return v;
}

// Integer threshold of the two-category draw: (p1 > 0 && c0 <= m 2^-24)  <=>  m >= thr_draw(c0, p1) for every 24-bit m.
// c0 2^24 and its ceiling are exact in fp32; c0 > 1, p1 <= 0 and NaNs give 2^24 (never), a negative c0 gives 0 (always).
__device__ __forceinline__ uint32_t thr_draw(float c0, float p1) {
  if (!(p1 > 0.f) || !(c0 <= 1.f)) return 0x1000000u;
  return (uint32_t)ceilf(fmaxf(c0, 0.f) * 16777216.f);
}

template <int DOMAIN>
__device__ __forceinline__ int table_next(int s, int a, float p, float u) {
  if (DOMAIN == RMAB_DOMAIN_CE) {
    // T[s=0] = [.5,.5]; T[1,0] = [1,0]; T[1,1] = [1-p, p]   (bandit_env_robust.py:831-836, :876-877)
    float pr[2];
    if (s == 0) { pr[0] = 0.5f; pr[1] = 0.5f; }
    else if (a == 0) { pr[0] = 1.f; pr[1] = 0.f; }
    else { pr[0] = __fsub_rn(1.f, p); pr[1] = p; }
    return categorical<2>(pr, u);
  } else {
    // s=0: [p,1-p,0]; s=1,a=0: [0,1-p,p]; s=1,a=1: [p,1-p,0]; s=2: [0,1-p,p]   (:595-616)
    const bool to0 = (s == 0) || (s == 1 && a == 1);
    float pr[3];
    pr[0] = to0 ? p : 0.f;
    pr[1] = __fsub_rn(1.f, p);
    pr[2] = to0 ? 0.f : p;
    return categorical<3>(pr, u);
  }
}

struct EpochArgs {
  RmabNetDesc d;
  int T;
  const float* Wpi; const float* Wv; const float* lamb; const float* nature; const float* ranges;
  const float* R; const float* C;
  double gamma, gl;  // discount, gamma * lam_OTHER
  uint32_t k0, k1, t_act0, t_env0, env0, arm0;
  uint8_t* s_traj; uint8_t* a;
  float* stats; float* arm_stats; double* env_part;
  float* adv; float* ret; float* last_val;
  // env-sharded ranks (the reference's MPI layout: every rank all arms, a slice of the envs): the advantage moments are
  // per arm over ALL envs, so a first launch exports the raw sums of this rank's envs and a second one is given the global
  // (mean, std) -- mpi_statistics_scalar of get() (agent_oracle.py:152-155, mpi_tools.py:76-97)
  double* mom_out;        // [N,2] sum adv, sum adv^2 over this launch's envs (may be NULL)
  const float* mom_in;    // [N,2] (mean, std) to normalise with instead of the launch's own (may be NULL)
  int tables_ready;   // the p / log p / v rows of `stats` were filled by the tcgen05 pre-pass (table_fwd_umma64.cu)
};

// ONCHIP: the trajectory of every (arm, env) chain stays in shared memory as bit planes -- one 32-bit word per 32 time
// steps for the action bit and for each state bit (S = 2: one plane, S = 3: two), words of env e at [word][plane][e] so
// that a warp's accesses are consecutive -- instead of the [N,T+1,E] / [N,T,E] byte arrays in HBM that only this kernel
// re-reads (two backward passes).  The byte arrays are still written when the caller passes `a` (MA-RMABPPO, tests,
// keep_buffers); with a == NULL nothing but the statistics leaves the SM.
// TRAJ: 0 = trajectory through HBM (s_traj [N,T+1,E], a [N,T,E]); 1 = on chip as bit planes, per-sample outputs (a, adv, ret)
// written when the caller passes them; 2 = on chip and no per-sample output requested (the trainers' epoch): the per-step
// pointer tests are compiled out.
template <int H, int DOMAIN, int TRAJ>
__global__ void __launch_bounds__(256, (H <= 16 ? 4 : 2)) epoch_table_kernel(EpochArgs g) {
  constexpr bool ONCHIP = TRAJ != 0, LEAN = TRAJ == 2;
  constexpr int S = DOMAIN == RMAB_DOMAIN_CE ? 2 : 3;
  constexpr int A = 2;
  constexpr int PL = S == 2 ? 2 : 3;   // bit planes per 32 steps: action, state bit 0 (, state bit 1)
  using SR = StatRows<S, A>;
  extern __shared__ __align__(16) float smem[];
  __shared__ double dred[3][8];
  __shared__ float s_ms[2];
  const int E = g.d.n_envs, T = g.T, N = g.d.n_arms;
  const SmemNet Lp(S + 1, H, A), Lv(S + 1, H, 1);
  const MlpLayout Gp(S + 1, H, A), Gv(S + 1, H, 1);
  float* sp = smem;
  float* sv = smem + Lp.size;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem + Lp.size + Lv.size);   // [n_words][PL][E]
  const int64_t wstride = (int64_t)PL * E;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const double gamma = g.gamma, gl = g.gl;
  const float C0 = __ldg(g.C), C1 = __ldg(g.C + 1);

  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    __syncthreads();  // previous arm's readers of the staged nets are done
    if (!g.tables_ready) {
      stage_net(sp, Lp, g.Wpi + (int64_t)n * Gp.P);
      stage_net(sv, Lv, g.Wv + (int64_t)n * Gv.P);
    }
    __syncthreads();
    // clamped nature parameter per (state, action) of this arm (:867-874 / :583-592)
    float pn0[S], pn1[S], Rn[S];
#pragma unroll
    for (int si = 0; si < S; ++si) {
      Rn[si] = __ldg(g.R + (int64_t)n * S + si);
      if (DOMAIN == RMAB_DOMAIN_CE) {
        const float p = fminf(fmaxf(__ldg(g.nature + n), __ldg(g.ranges + 2 * (int64_t)n)), __ldg(g.ranges + 2 * (int64_t)n + 1));
        pn0[si] = p;
        pn1[si] = p;
      } else {
        const int64_t c0 = ((int64_t)n * S + si) * A;
        pn0[si] = fminf(fmaxf(__ldg(g.nature + c0), __ldg(g.ranges + 2 * c0)), __ldg(g.ranges + 2 * c0 + 1));
        pn1[si] = fminf(fmaxf(__ldg(g.nature + c0 + 1), __ldg(g.ranges + 2 * c0 + 2)), __ldg(g.ranges + 2 * c0 + 3));
      }
    }
    // CE transition out of (s = 1, a = 1): next = 1 iff p > 0 and fsub_rn(1, p) <= u (table_next, categorical<2>)
    const uint32_t thr_e11 = thr_draw(__fsub_rn(1.f, pn1[S - 1]), pn1[S - 1]);
    const int64_t traj0 = (int64_t)n * (g.a ? (int64_t)(T + 1) * E : (int64_t)E);   // a == NULL: s_traj is the [N,E] start states
    const int64_t step0 = (int64_t)n * T * E;
    const uint32_t gn = g.arm0 + (uint32_t)n;

    // ---- pass 0: tables + forward chain; pass 1: backward scan for the advantage moments ----
    double asum = 0.0, asq = 0.0;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const float lam = g.lamb[e];
      float p0t[S], p1t[S], vt[S];
      if (g.tables_ready) {
        // hidden 64: the tcgen05 pre-pass parked p(a | s) in the (still unused) advantage rows and wrote log p and v
        const float* st = g.stats + ((int64_t)n * SR::K) * E + e;
#pragma unroll
        for (int si = 0; si < S; ++si) {
          p0t[si] = st[(int64_t)(SR::kAp + si * A + 0) * E];
          p1t[si] = st[(int64_t)(SR::kAp + si * A + 1) * E];
          vt[si] = st[(int64_t)(SR::kV + si) * E];
        }
      } else
#pragma unroll
      for (int si = 0; si < S; ++si) {
        float a1[H], logit[kMaxA], vout[kMaxA], p[kMaxA], lp[kMaxA];
        mlp_layer1<H>(sp, Lp, 1, si, 1.f, lam, a1);
        mlp_tail<H>(sp, Lp, a1, A, logit);
        mlp_layer1<H>(sv, Lv, 1, si, 1.f, lam, a1);
        mlp_tail<H>(sv, Lv, a1, 1, vout);
        softmax_small(logit, A, p, lp);
        p0t[si] = p[0]; p1t[si] = p[1]; vt[si] = vout[0];
        float* st = g.stats + ((int64_t)n * SR::K) * E + e;
        st[(int64_t)(SR::kLp + si * A + 0) * E] = lp[0];
        st[(int64_t)(SR::kLp + si * A + 1) * E] = lp[1];
        st[(int64_t)(SR::kV + si) * E] = vout[0];
      }
      uint8_t* tr = g.s_traj + traj0 + e;
      uint8_t* ac = !LEAN && g.a ? g.a + step0 + e : nullptr;
      int si = tr[0];
      si = si < S ? si : S - 1;
      const uint32_t ge = g.env0 + (uint32_t)e;
      uint4 blk = make_uint4(0u, 0u, 0u, 0u);
      // bit t of the state planes = s_t (t = 0..T), bit t of the action plane = a_t (t = 0..T-1)
      uint32_t aw = 0u, sw0 = (uint32_t)(si & 1), sw1 = (uint32_t)(si >> 1);
      uint32_t* bw = bits + e;
      // Draws in the integer domain: u = m 2^-24 with m = word >> 8, so `c <= u` is `m >= ceil(c 2^24)` exactly (c 2^24 and
      // the ceiling are exact in fp32).  The action rule categorical<2>({p0, p1}, u) = (p1 > 0 && p0 <= u) becomes one
      // threshold per state, computed once per (arm, env); thr_draw() maps "never" (p1 <= 0, NaN) to 2^24.
      uint32_t thr_a[S];
#pragma unroll
      for (int k = 0; k < S; ++k) thr_a[k] = thr_draw(p0t[k], p1t[k]);
      // one step of the chain given its two Philox words (action draw, transition draw)
      auto chain_step = [&](int t, uint32_t w_act, uint32_t w_env) {
        uint32_t ta = thr_a[0];
#pragma unroll
        for (int k = 1; k < S; ++k) ta = (si == k) ? thr_a[k] : ta;
        const int act = (w_act >> 8) >= ta ? 1 : 0;
        if (!ONCHIP || (!LEAN && ac)) ac[(int64_t)t * E] = (uint8_t)act;
        if (DOMAIN == RMAB_DOMAIN_CE) {
          // table_next<CE>: s = 0 -> (0.5 <= u); s = 1, a = 0 -> 0; s = 1, a = 1 -> (p > 0 && 1 - p <= u)
          const uint32_t te = si == 0 ? 0x800000u : (act ? thr_e11 : 0x1000000u);
          si = (w_env >> 8) >= te ? 1 : 0;
        } else {
          const float p = act ? pick_f<S>(pn1, si) : pick_f<S>(pn0, si);
          si = table_next<DOMAIN>(si, act, p, u24(w_env));
        }
        if (!ONCHIP || (!LEAN && ac)) tr[(int64_t)(t + 1) * E] = (uint8_t)si;
        if (ONCHIP) {
          aw |= (uint32_t)act << (t & 31);
          const int b = (t + 1) & 31;
          if (b == 0) {   // the word holding s_{t-31..t} and a_{t-31..t} is complete
            uint32_t* wp = bw + (int64_t)(t >> 5) * wstride;
            wp[0] = aw; wp[E] = sw0;
            if (PL == 3) wp[2 * E] = sw1;
            aw = sw0 = sw1 = 0u;
          }
          sw0 |= (uint32_t)(si & 1) << b;
          sw1 |= (uint32_t)(si >> 1) << b;
        }
      };
      // One Philox block serves the two steps 2k, 2k + 1 of the GLOBAL step counter (x, y = transition / action draw of the
      // even step, z, w of the odd one).  Even t_act0 (every epoch when T is even): steps walked in pairs, no parity tests.
      int t = 0;
      if (g.t_act0 & 1u) {   // odd start: the first step uses the upper half of its block
        blk = philox_step_block(g.k0, g.k1, g.t_act0, ge, gn);
        chain_step(0, blk.w, blk.z);
        t = 1;
      }
      for (; t + 1 < T; t += 2) {
        blk = philox_step_block(g.k0, g.k1, g.t_act0 + (uint32_t)t, ge, gn);
        chain_step(t, blk.y, blk.x);
        chain_step(t + 1, blk.w, blk.z);
      }
      if (t < T) {           // a last even step without its partner
        blk = philox_step_block(g.k0, g.k1, g.t_act0 + (uint32_t)t, ge, gn);
        chain_step(t, blk.y, blk.x);
      }
      if (ONCHIP) {       // the last (partial) word: s_T sits at bit T & 31
        uint32_t* wp = bw + (int64_t)(T >> 5) * wstride;
        wp[0] = aw; wp[E] = sw0;
        if (PL == 3) wp[2 * E] = sw1;
      }
      // backward scan 1: only the advantage moments are needed here, so the recurrence runs in fp32 (relative error
      // ~1e-7 per step, damped by gamma*lam); pass 2 below redoes it in fp64 for the values that are kept
      const float lvf = pick_f<S>(vt, si);
      if (g.last_val) g.last_val[(int64_t)n * E + e] = lvf;
      const float gam_f = (float)gamma, gl_f = (float)gl;
      float next_v = lvf, adv_acc = 0.f;
      int sn = si;
      uint32_t ca = aw, cs0 = sw0, cs1 = sw1;   // the last word is still in registers
      for (int t = T - 1; t >= 0; --t) {
        int a_t, s_t;
        if (ONCHIP) {
          if ((t & 31) == 31) {
            const uint32_t* wp = bw + (int64_t)(t >> 5) * wstride;
            ca = wp[0]; cs0 = wp[E];
            if (PL == 3) cs1 = wp[2 * E];
          }
          a_t = (int)((ca >> (t & 31)) & 1u);
          s_t = (int)((cs0 >> (t & 31)) & 1u) | (PL == 3 ? (int)(((cs1 >> (t & 31)) & 1u) << 1) : 0);
        } else {
          a_t = ac[(int64_t)t * E];
          s_t = tr[(int64_t)t * E];
        }
        const float v = pick_f<S>(vt, s_t);
        const float rt = fmaf(-lam, a_t ? C1 : C0, pick_f<S>(Rn, sn));
        const float delta = fmaf(gam_f, next_v, rt) - v;
        adv_acc = fmaf(gl_f, adv_acc, delta);
        asum += (double)adv_acc;
        asq = fma((double)adv_acc, (double)adv_acc, asq);
        next_v = v;
        sn = s_t;
      }
    }
    // ---- per-arm mean / population std over all T*E samples (get() :152-155, mpi_tools.py:76-97) ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      asum += __shfl_xor_sync(0xffffffffu, asum, o);
      asq += __shfl_xor_sync(0xffffffffu, asq, o);
    }
    if (lane == 0) { dred[0][warp] = asum; dred[1][warp] = asq; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s1 = 0.0, s2 = 0.0;
      for (int w = 0; w < nwarp; ++w) { s1 += dred[0][w]; s2 += dred[1][w]; }
      const double cnt = (double)T * (double)E;
      const double mean = s1 / cnt;
      double var = s2 / cnt - mean * mean;
      var = var > 0.0 ? var : 0.0;
      s_ms[0] = (float)mean;
      s_ms[1] = (float)sqrt(var);
      if (g.mom_out) { g.mom_out[2 * (int64_t)n] = s1; g.mom_out[2 * (int64_t)n + 1] = s2; }
      if (g.mom_in) { s_ms[0] = g.mom_in[2 * (int64_t)n]; s_ms[1] = g.mom_in[2 * (int64_t)n + 1]; }
      if (g.arm_stats) { g.arm_stats[4 * (int64_t)n] = s_ms[0]; g.arm_stats[4 * (int64_t)n + 1] = s_ms[1]; }
    }
    __syncthreads();
    const float mean = s_ms[0], stdv = s_ms[1];
    const float inv_std = __fdiv_rn(1.f, stdv);  // std = 0 -> inf, and (adv - mean) * inf = NaN like the reference's 0 / 0

    // ---- pass 2: backward scan again, normalised advantages into the per-(s,a) statistics ----
    double var_sum = 0.0;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      float* st = g.stats + ((int64_t)n * SR::K) * E + e;
      float vt[S];
#pragma unroll
      for (int si = 0; si < S; ++si) vt[si] = st[(int64_t)(SR::kV + si) * E];
      const uint8_t* tr = g.s_traj + traj0 + e;
      const uint8_t* ac = g.a ? g.a + step0 + e : nullptr;
      const double lm = (double)g.lamb[e];
      const uint32_t* bw = bits + e;
      uint32_t ca = 0u, cs0 = 0u, cs1 = 0u;
      int sn;
      if (ONCHIP) {
        const uint32_t* wp = bw + (int64_t)(T >> 5) * wstride;
        ca = wp[0]; cs0 = wp[E];
        if (PL == 3) cs1 = wp[2 * E];
        sn = (int)((cs0 >> (T & 31)) & 1u) | (PL == 3 ? (int)(((cs1 >> (T & 31)) & 1u) << 1) : 0);
      } else {
        sn = tr[(int64_t)T * E];
      }
      const double lv = (double)pick_f<S>(vt, sn);
      double next_v = lv, adv_acc = 0.0, ret_acc = lv, dcost = 0.0, rsum = 0.0;
      float Ap[S * A], Am[S * A];
      double Sr[S], Sr2[S];
      int cnt[S * A];
#pragma unroll
      for (int c = 0; c < S * A; ++c) { Ap[c] = 0.f; Am[c] = 0.f; cnt[c] = 0; }
#pragma unroll
      for (int s = 0; s < S; ++s) { Sr[s] = 0.0; Sr2[s] = 0.0; }
      for (int t = T - 1; t >= 0; --t) {
        int a_t, s_t;
        if (ONCHIP) {
          if ((t & 31) == 31) {
            const uint32_t* wp = bw + (int64_t)(t >> 5) * wstride;
            ca = wp[0]; cs0 = wp[E];
            if (PL == 3) cs1 = wp[2 * E];
          }
          a_t = (int)((ca >> (t & 31)) & 1u);
          s_t = (int)((cs0 >> (t & 31)) & 1u) | (PL == 3 ? (int)(((cs1 >> (t & 31)) & 1u) << 1) : 0);
        } else {
          a_t = ac[(int64_t)t * E];
          s_t = tr[(int64_t)t * E];
        }
        const double r = (double)pick_f<S>(Rn, sn), c = (double)(a_t ? C1 : C0), v = (double)pick_f<S>(vt, s_t);
        const double rt = __dsub_rn(r, __dmul_rn(lm, c));
        const double qv = __dadd_rn(rt, __dmul_rn(gamma, next_v));
        const double delta = __dsub_rn(qv, v);
        adv_acc = __dadd_rn(delta, __dmul_rn(gl, adv_acc));
        ret_acc = __dadd_rn(rt, __dmul_rn(gamma, ret_acc));
        dcost = __dadd_rn(c, __dmul_rn(gamma, dcost));
        rsum += r;
        const float x = __fmul_rn(__fsub_rn((float)adv_acc, mean), inv_std);  // (adv - mean) / std in fp32 (:155)
        const float rf = (float)ret_acc;
        if (!LEAN && g.adv) g.adv[step0 + (int64_t)t * E + e] = x;
        if (!LEAN && g.ret) g.ret[step0 + (int64_t)t * E + e] = rf;
        const int cell = s_t * A + a_t;
        const float xp = fmaxf(x, 0.f), xm = fminf(x, 0.f);
        // predicated accumulates (one @P FADD / DADD / DFMA each) rather than select-then-add pairs
        const double rd = (double)rf;
#pragma unroll
        for (int cc = 0; cc < S * A; ++cc) {
          if (cell == cc) {
            Ap[cc] += xp;
            Am[cc] += xm;
            cnt[cc] += 1;
          }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
          if (s_t == s) {
            Sr[s] += rd;
            Sr2[s] = fma(rd, rd, Sr2[s]);
          }
        }
        next_v = v;
        sn = s_t;
      }
#pragma unroll
      for (int c = 0; c < S * A; ++c) {
        st[(int64_t)(SR::kAp + c) * E] = Ap[c];
        st[(int64_t)(SR::kAm + c) * E] = Am[c];
        st[(int64_t)(SR::kCnt + c) * E] = (float)cnt[c];
      }
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int ns = cnt[s * A] + cnt[s * A + 1];
        const double m = ns > 0 ? Sr[s] / (double)ns : 0.0;
        const double mf = (double)(float)m;  // the update kernel sees the fp32-rounded mean
        // sum (ret - mf)^2 = Sr2 - 2 mf Sr + n mf^2
        const double centred = ns > 0 ? fma((double)ns * mf, mf, fma(-2.0 * mf, Sr[s], Sr2[s])) : 0.0;
        var_sum += centred > 0.0 ? centred : 0.0;
        st[(int64_t)(SR::kRetMean + s) * E] = (float)m;
        st[(int64_t)(SR::kCntS + s) * E] = (float)ns;
      }
      if (g.env_part) {
        // accumulated over the arms this CTA walks; written once at the end (below)
        double* ep = g.env_part + ((int64_t)blockIdx.x * 2) * E + e;
        const bool first = n == (int)blockIdx.x;
        ep[0] = (first ? 0.0 : ep[0]) + rsum;
        ep[E] = (first ? 0.0 : ep[E]) + dcost;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) var_sum += __shfl_xor_sync(0xffffffffu, var_sum, o);
    if (lane == 0) dred[2][warp] = var_sum;
    __syncthreads();
    if (threadIdx.x == 0 && g.arm_stats) {
      double s = 0.0;
      for (int w = 0; w < nwarp; ++w) s += dred[2][w];
      g.arm_stats[4 * (int64_t)n + 2] = (float)s;  // sum_t (ret_t - mean_ret[s_t, e])^2 over the arm's batch
      g.arm_stats[4 * (int64_t)n + 3] = 0.f;
    }
  }
}

// out[i] = sum_b in[b, i] in fp64 (i over the 2*E sums).  One warp per output: lane l adds partials l, l+32, ... in
// order, then a fixed butterfly -- deterministic for a given grid size.
__global__ void part_sum_kernel(int nb, int64_t ke, const double* __restrict__ in, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= ke) return;
  double acc = 0.0;
  for (int b = lane; b < nb; b += 32) acc += in[(int64_t)b * ke + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[i] = (float)acc;
}

int table_fwd_umma64_run(const RmabNetDesc* net, const float* Wpi, const float* Wv, const float* lamb, float* stats, int K,
                         int row_p, int row_lp, int row_v, cudaStream_t st);   // table_fwd_umma64.cu

template <typename K>
static int set_smem_attr(K kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess ? 0 : -1;
}

static int epoch_grid(int N) {
  const int cap = kNumSMs * 4;
  return N < cap ? N : cap;
}

}  // namespace rmab

using namespace rmab;

extern "C" int rmab_stat_rows(int domain) {
  if (domain == RMAB_DOMAIN_CE) return StatRows<2, 2>::K;
  if (domain == RMAB_DOMAIN_ARMMAN) return StatRows<3, 2>::K;
  return -1;
}

extern "C" size_t rmab_epoch_ws_bytes(int N, int E) { return (size_t)epoch_grid(N) * 2 * (size_t)E * sizeof(double); }

static int epoch_table_impl(int domain, const RmabNetDesc* net, int T, const float* Wpi, const float* Wv,
                            const float* lamb, const float* nature, const float* ranges, const float* R,
                            const float* C, double gamma, double lam, uint64_t seed, uint32_t t_act0,
                            uint32_t t_env0, uint32_t env0, uint32_t arm0, uint8_t* s_traj, uint8_t* a,
                            float* stats, float* arm_stats, float* env_sums, void* ws, size_t ws_bytes, float* adv,
                            float* ret, float* last_val, double* mom_out, const float* mom_in, rmab_stream_t stream) {
  RMAB_REQUIRE(net, RMAB_E_BADARG, "rmab_epoch_table: null net descriptor");
  RMAB_REQUIRE(domain == RMAB_DOMAIN_CE || domain == RMAB_DOMAIN_ARMMAN, RMAB_E_BADARG,
               "rmab_epoch_table: domain %d is not a table domain", domain);
  const int S = domain == RMAB_DOMAIN_CE ? 2 : 3;
  const int h = net->hidden;
  RMAB_REQUIRE(h == 8 || h == 16 || h == 32 || h == 64, RMAB_E_SHAPE,
               "rmab_epoch_table: hidden width %d unsupported (8, 16, 32 or 64; two hidden layers)", h);
  RMAB_REQUIRE(net->one_hot && net->n_states == S && net->n_actions == 2 && net->n_extra == 0, RMAB_E_SHAPE,
               "rmab_epoch_table: needs one-hot S=%d, A=2 nets", S);
  RMAB_REQUIRE(T > 0 && net->n_envs > 0 && net->n_arms >= 0, RMAB_E_SHAPE, "rmab_epoch_table: bad T/E/N");
  RMAB_REQUIRE(t_act0 == t_env0, RMAB_E_BADARG, "rmab_epoch_table: the action and transition draws of a step share one "
               "Philox block, so t_act0 must equal t_env0");
  RMAB_REQUIRE(Wpi && Wv && lamb && nature && ranges && R && C && s_traj && stats, RMAB_E_BADARG,
               "rmab_epoch_table: null pointer");
  RMAB_REQUIRE(!env_sums || (ws && ws_bytes >= rmab_epoch_ws_bytes(net->n_arms, net->n_envs)), RMAB_E_BADARG,
               "rmab_epoch_table: env_sums needs a workspace of rmab_epoch_ws_bytes()");
  if (net->n_arms == 0) return RMAB_OK;
  EpochArgs g;
  g.d = *net; g.T = T; g.Wpi = Wpi; g.Wv = Wv; g.lamb = lamb; g.nature = nature; g.ranges = ranges; g.R = R; g.C = C;
  g.gamma = gamma; g.gl = gamma * lam;
  g.k0 = (uint32_t)(seed & 0xFFFFFFFFull); g.k1 = (uint32_t)(seed >> 32);
  g.t_act0 = t_act0; g.t_env0 = t_env0; g.env0 = env0; g.arm0 = arm0;
  g.s_traj = s_traj; g.a = a; g.stats = stats; g.arm_stats = arm_stats; g.env_part = env_sums ? (double*)ws : nullptr;
  g.adv = adv; g.ret = ret; g.last_val = last_val;
  g.mom_out = mom_out; g.mom_in = mom_in;
  size_t smem = sizeof(float) * (SmemNet(S + 1, h, 2).size + SmemNet(S + 1, h, 1).size);
  // on-chip trajectory: (T / 32 + 1) words x (1 action + 1 or 2 state) planes x E envs; two CTAs per SM must still fit
  const size_t bit_bytes = (size_t)(T / 32 + 1) * (S == 2 ? 2 : 3) * (size_t)net->n_envs * sizeof(uint32_t);
  static const bool force_hbm = getenv("RMAB_EPOCH_HBM_TRAJ") != nullptr;   // A/B switch: round-1 behaviour
  const bool onchip = !force_hbm && smem + bit_bytes <= 100 * 1024;
  RMAB_REQUIRE(a || onchip, RMAB_E_BADARG, "rmab_epoch_table: a == NULL needs the on-chip trajectory (T = %d, E = %d is too "
               "large for shared memory)", T, net->n_envs);
  if (onchip) smem += bit_bytes;
  cudaStream_t st0 = (cudaStream_t)stream;
  static const bool ffma_tables = getenv("RMAB_EPOCH_FFMA_TABLES") != nullptr;   // A/B switch: thread-per-row pre-pass at h = 64
  g.tables_ready = 0;
  if (h == 64 && !ffma_tables) {
    int rc = S == 2 ? table_fwd_umma64_run(net, Wpi, Wv, lamb, stats, StatRows<2, 2>::K, StatRows<2, 2>::kAp,
                                           StatRows<2, 2>::kLp, StatRows<2, 2>::kV, st0)
                    : table_fwd_umma64_run(net, Wpi, Wv, lamb, stats, StatRows<3, 2>::K, StatRows<3, 2>::kAp,
                                           StatRows<3, 2>::kLp, StatRows<3, 2>::kV, st0);
    if (rc) return rc;
    g.tables_ready = 1;
  }
  const int grid = epoch_grid(net->n_arms);
  int threads = ((net->n_envs + 31) / 32) * 32;
  threads = threads > 256 ? 256 : threads;
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH2(H, D, O)                                                                                       \
  {                                                                                                            \
    RMAB_REQUIRE(set_smem_attr(epoch_table_kernel<H, D, O>, smem) == 0, RMAB_E_LAUNCH, "rmab_epoch_table: smem %zu B", smem); \
    epoch_table_kernel<H, D, O><<<grid, threads, smem, st>>>(g);                                               \
  }
#define LAUNCH(H, D) \
  if (onchip && !a && !adv && !ret) LAUNCH2(H, D, 2) else if (onchip) LAUNCH2(H, D, 1) else LAUNCH2(H, D, 0)
#define LAUNCH_H(D)               \
  switch (h) {                    \
    case 8: LAUNCH(8, D) break;   \
    case 16: LAUNCH(16, D) break; \
    case 32: LAUNCH(32, D) break; \
    default: LAUNCH(64, D) break; \
  }
  if (domain == RMAB_DOMAIN_CE) { LAUNCH_H(RMAB_DOMAIN_CE) } else { LAUNCH_H(RMAB_DOMAIN_ARMMAN) }
#undef LAUNCH_H
#undef LAUNCH
#undef LAUNCH2
  int rc = check_launch("rmab_epoch_table");
  if (rc || !env_sums) return rc;
  const int64_t ke = 2 * (int64_t)net->n_envs;
  part_sum_kernel<<<(int)((ke + 7) / 8), 256, 0, st>>>(grid, ke, (const double*)ws, env_sums);
  return check_launch("rmab_epoch_table(sum)");
}

extern "C" int rmab_epoch_table(int domain, const RmabNetDesc* net, int T, const float* Wpi, const float* Wv,
                                const float* lamb, const float* nature, const float* ranges, const float* R,
                                const float* C, double gamma, double lam, uint64_t seed, uint32_t t_act0,
                                uint32_t t_env0, uint32_t env0, uint32_t arm0, uint8_t* s_traj, uint8_t* a,
                                float* stats, float* arm_stats, float* env_sums, void* ws, size_t ws_bytes, float* adv,
                                float* ret, float* last_val, rmab_stream_t stream) {
  return epoch_table_impl(domain, net, T, Wpi, Wv, lamb, nature, ranges, R, C, gamma, lam, seed, t_act0, t_env0, env0, arm0,
                          s_traj, a, stats, arm_stats, env_sums, ws, ws_bytes, adv, ret, last_val, nullptr, nullptr, stream);
}

extern "C" int rmab_epoch_table_envshard(int domain, const RmabNetDesc* net, int T, const float* Wpi, const float* Wv,
                                         const float* lamb, const float* nature, const float* ranges, const float* R,
                                         const float* C, double gamma, double lam, uint64_t seed, uint32_t t_act0,
                                         uint32_t t_env0, uint32_t env0, uint32_t arm0, uint8_t* s_traj, uint8_t* a,
                                         float* stats, float* arm_stats, float* env_sums, void* ws, size_t ws_bytes,
                                         float* last_val, double* adv_moments_out, const float* adv_moments_in,
                                         rmab_stream_t stream) {
  RMAB_REQUIRE((adv_moments_out != nullptr) != (adv_moments_in != nullptr), RMAB_E_BADARG,
               "rmab_epoch_table_envshard: give exactly one of adv_moments_out (first launch) / adv_moments_in (second)");
  return epoch_table_impl(domain, net, T, Wpi, Wv, lamb, nature, ranges, R, C, gamma, lam, seed, t_act0, t_env0, env0, arm0,
                          s_traj, a, stats, arm_stats, env_sums, ws, ws_bytes, nullptr, nullptr, last_val, adv_moments_out,
                          adv_moments_in, stream);
}
