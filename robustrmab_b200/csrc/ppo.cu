// PPO-clip update for N independent per-arm nets: all train_pi_iters + train_v_iters full-batch Adam
// steps of one update() call, one CTA per arm, weights / gradients / Adam moments resident in shared
// memory across iterations.
//
// Reference behaviour: robust_rmab/algos/rmabppo/agent_oracle.py compute_loss_pi :285-322,
// compute_loss_v :325-339, update :399-436, Adam optimisers :383-388 (torch defaults).
//
// Mapping ("lanes over features"): a group of LPS = min(H,32) lanes owns one sample, lane f owns hidden
// feature(s) f (+32 when H = 64) of both hidden layers.  The weight-gradient accumulators of the lane's
// features live in registers for the whole pass over the batch, so the only cross-thread reductions are
// the A-logit sums (warp shuffles) and one fixed-order merge of the 8 warps per iteration; activations
// are exchanged through per-warp shared-memory rows and read back as float4 broadcasts.
// FFMA-bound: 6 * P flops per sample per iteration (fwd + two backward products).
#include "common.cuh"
#include "mlp.cuh"

namespace rmab {

template <int H>
struct PpoCfg {
  static constexpr int LPS = H < 32 ? H : 32;  // lanes per sample
  static constexpr int FPL = H / LPS;          // features per lane
  static constexpr int SPW = 32 / LPS;         // samples per warp pass
};

// TABLE = false: one row per sample (s_buf, a_buf, adv, logp_old, ret are [N,T,E] arrays) -- the general path
//                (any input encoding, any A <= 4; SIS / MA-RMABPPO).
// TABLE = true : one row per (state, env) with the statistics of rmab_epoch_table; kept as the simple
//                reference formulation of the table update -- the shipped table path is ppo_table.cu.
template <int H, bool IS_PI, bool TABLE>
__global__ void __launch_bounds__(256, (H >= 64 ? 1 : 2))
ppo_update_kernel(RmabNetDesc d, RmabHyper hp, int T, int s_rows, float* __restrict__ Wg, float* __restrict__ adam,
                  const uint8_t* __restrict__ s_buf, const uint8_t* __restrict__ a_buf, const float* __restrict__ adv,
                  const float* __restrict__ logp_old, const float* __restrict__ ret, const float* __restrict__ lamb,
                  const float* __restrict__ stats, const float* __restrict__ arm_stats, int stage_stats,
                  float* __restrict__ loss_out, const float* __restrict__ z1_extra, float* __restrict__ dz1_out,
                  int64_t z1_arm_stride, int64_t z1_sample_stride) {
  using Cfg = PpoCfg<H>;
  constexpr int LPS = Cfg::LPS, FPL = Cfg::FPL, SPW = Cfg::SPW;
  constexpr int NW = 8;
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];

  const int E = d.n_envs, S = d.n_states;
  const int A = IS_PI ? d.n_actions : 1;
  const int in = (d.one_hot ? S : 1) + 1;
  const MlpLayout L(in, H, A);
  const int P = L.P, Pp = L.padded();
  float* w = smem;              // packed weights [P]
  float* w2t = w + Pp;          // W2 transposed [k][j]
  float* G = w2t + H * H;       // gradient [P]
  float* am = G + Pp;           // Adam m
  float* av = am + Pp;          // Adam v
  float* scr = av + Pp;         // per-warp scratch: a1 rows then d2 rows
  const int n = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LPS, f0 = lane % LPS;
  float* sa1 = scr + warp * (2 * SPW * H) + sub * H;
  float* sd2 = sa1 + SPW * H;

  float* Wn = Wg + (int64_t)n * P;
  float* An = adam + (int64_t)n * 2 * P;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    w[i] = Wn[i];
    am[i] = An[i];
    av[i] = An[P + i];
    G[i] = 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * H; i += blockDim.x) w2t[(i % H) * H + (i / H)] = w[L.oW2 + i];
  __syncthreads();

  // statistics rows this pass needs: policy [0, 3SA+S) = Ap Am Lp CntS, value [3SA, 3SA+2S) = CntS RetMean
  const int SA = S * d.n_actions;
  const int row0 = IS_PI ? 0 : 3 * SA, nrows = IS_PI ? 3 * SA + S : 2 * S;
  const int K = 4 * SA + 3 * S;
  const float* st = nullptr;  // st[(row - row0) * E + e]
  if (TABLE) {
    const float* gst = stats + ((int64_t)n * K + row0) * E;
    if (stage_stats) {
      float* sst = scr + NW * 2 * SPW * H;
      for (int i = threadIdx.x; i < nrows * E; i += blockDim.x) sst[i] = gst[i];
      __syncthreads();
      st = sst;
    } else {
      st = gst;
    }
  }
  const int M = TABLE ? S * E : T * E;
  const float invM = 1.f / (float)((int64_t)T * E);
  const int64_t sbase = (int64_t)n * s_rows * E;
  const int64_t base = (int64_t)n * T * E;
  const int iters = IS_PI ? hp.pi_iters : hp.v_iters;
  const double lr = IS_PI ? hp.pi_lr : hp.vf_lr;
  const int step0 = IS_PI ? hp.adam_step0_pi : hp.adam_step0_v;
  const float clip_lo = (float)(1.0 - hp.clip_ratio), clip_hi = (float)(1.0 + hp.clip_ratio);
  const float beta_ent = (float)hp.entropy_coeff;
  const float lerp_w = (float)(1.0 - hp.beta1), beta2f = (float)hp.beta2, omb2 = (float)(1.0 - hp.beta2);
  const float epsf = (float)hp.adam_eps;

  for (int it = 0; it < iters; ++it) {
    float gW1[FPL][kMaxIn], gb1[FPL], gW2[FPL][H], gb2[FPL], gW3[kMaxA][FPL], gb3[kMaxA];
#pragma unroll
    for (int i = 0; i < FPL; ++i) {
      gb1[i] = 0.f;
      gb2[i] = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxIn; ++c) gW1[i][c] = 0.f;
#pragma unroll
      for (int k = 0; k < H; ++k) gW2[i][k] = 0.f;
#pragma unroll
      for (int a = 0; a < kMaxA; ++a) gW3[a][i] = 0.f;
    }
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) gb3[a] = 0.f;
    float loss_acc = 0.f;

    const int n_groups = (M + SPW - 1) / SPW;
    for (int grp = warp; grp < n_groups; grp += NW) {
      const int m = grp * SPW + sub;
      const bool valid = m < M;
      const int mm = valid ? m : M - 1;
      const int t = mm / E, e = mm - t * E;  // TABLE: t is the state of the row
      int si;
      if (TABLE) si = t;
      else {
        si = s_buf[sbase + (int64_t)t * E + e];
        si = si < S ? si : S - 1;
      }
      const float lam = lamb[e];
      const float x0 = d.one_hot ? 0.f : __fdiv_rn((float)si, d.state_norm);

      // ---- forward ----
      float a1[FPL], a2[FPL];
#pragma unroll
      for (int i = 0; i < FPL; ++i) {
        const int f = f0 + LPS * i;
        const float* row = w + L.oW1 + f * in;
        float z;
        if (d.one_hot) z = fmaf(row[in - 1], lam, row[si]) + w[L.ob1 + f];
        else z = fmaf(row[1], lam, fmaf(row[0], x0, w[L.ob1 + f]));
        // MA agent critic: first-layer term of the nature inputs (one dense GEMM by the caller), arm-major
        // z1_extra[n][sample][f] or sample-major z1_extra[sample][n][f]
        if (z1_extra) z += z1_extra[n * z1_arm_stride + mm * z1_sample_stride + f];
        a1[i] = tanhf(z);
        sa1[f] = a1[i];
      }
      __syncwarp();
      float z2[FPL];
#pragma unroll
      for (int i = 0; i < FPL; ++i) z2[i] = w[L.ob2 + f0 + LPS * i];
#pragma unroll
      for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 x = reinterpret_cast<const float4*>(sa1)[k4];
#pragma unroll
        for (int i = 0; i < FPL; ++i) {
          const int f = f0 + LPS * i;
          z2[i] = fmaf(w2t[(4 * k4 + 0) * H + f], x.x, z2[i]);
          z2[i] = fmaf(w2t[(4 * k4 + 1) * H + f], x.y, z2[i]);
          z2[i] = fmaf(w2t[(4 * k4 + 2) * H + f], x.z, z2[i]);
          z2[i] = fmaf(w2t[(4 * k4 + 3) * H + f], x.w, z2[i]);
        }
      }
      float logit[kMaxA];
#pragma unroll
      for (int a = 0; a < kMaxA; ++a) logit[a] = 0.f;
#pragma unroll
      for (int i = 0; i < FPL; ++i) {
        a2[i] = tanhf(z2[i]);
#pragma unroll
        for (int a = 0; a < kMaxA; ++a)
          if (a < A) logit[a] = fmaf(w[L.oW3 + a * H + f0 + LPS * i], a2[i], logit[a]);
      }
#pragma unroll
      for (int a = 0; a < kMaxA; ++a)
        if (a < A) {
#pragma unroll
          for (int o = LPS / 2; o > 0; o >>= 1) logit[a] += __shfl_xor_sync(0xffffffffu, logit[a], o);
          logit[a] += w[L.ob3 + a];
        }

      // ---- loss and d(loss)/d(logit) ----
      float g[kMaxA];
#pragma unroll
      for (int a = 0; a < kMaxA; ++a) g[a] = 0.f;
      const int64_t o = base + (int64_t)t * E + e;
      if (TABLE) {
        const float wrow = st[(int64_t)(3 * SA - row0 + si) * E + e];  // CntS: samples of this env in state si
        if (IS_PI) {
          float p[kMaxA], lp[kMaxA], ga[kMaxA];
          softmax_small(logit, A, p, lp);
          float ent = 0.f, sumg = 0.f, lrow = 0.f;
#pragma unroll
          for (int a = 0; a < kMaxA; ++a)
            if (a < A) ent = fmaf(-p[a], lp[a], ent);
#pragma unroll
          for (int a = 0; a < kMaxA; ++a) {
            ga[a] = 0.f;
            if (a < A) {
              const int c = si * A + a;
              const float Ap = st[(int64_t)c * E + e], Am = st[(int64_t)(SA + c) * E + e];
              const float ratio = expf(lp[a] - st[(int64_t)(2 * SA + c) * E + e]);
              // positive advantages: min(r, 1+c) * A ; negative: max(r, 1-c) * A   (:301-305)
              lrow -= fminf(ratio, clip_hi) * Ap + fmaxf(ratio, clip_lo) * Am;
              ga[a] = -((ratio <= clip_hi ? ratio * Ap : 0.f) + (ratio >= clip_lo ? ratio * Am : 0.f));
              sumg += ga[a];
            }
          }
          if (valid) {
            loss_acc += lrow - beta_ent * wrow * ent;
#pragma unroll
            for (int a = 0; a < kMaxA; ++a)
              if (a < A) g[a] = invM * (ga[a] - p[a] * sumg + beta_ent * wrow * p[a] * (lp[a] + ent));
          }
        } else {
          const float diff = logit[0] - st[(int64_t)(3 * SA + S - row0 + si) * E + e];
          if (valid) {
            loss_acc = fmaf(wrow * diff, diff, loss_acc);
            g[0] = invM * 2.f * wrow * diff;
          }
        }
      } else if (IS_PI) {
        float p[kMaxA], lp[kMaxA];
        softmax_small(logit, A, p, lp);
        const int act = a_buf[o];
        float ent = 0.f, lpa = lp[0];
#pragma unroll
        for (int a = 0; a < kMaxA; ++a) {
          if (a < A) ent = fmaf(-p[a], lp[a], ent);
          if (a > 0 && act == a) lpa = lp[a];
        }
        const float advv = adv[o];
        const float ratio = expf(lpa - logp_old[o]);
        const float surr1 = ratio * advv;
        const float surr2 = fminf(fmaxf(ratio, clip_lo), clip_hi) * advv;
        const bool pass = (ratio >= clip_lo && ratio <= clip_hi) || (surr1 < surr2);
        const float coef = pass ? -advv * ratio : 0.f;
        if (valid) {
          loss_acc += -fminf(surr1, surr2) - beta_ent * ent;
#pragma unroll
          for (int a = 0; a < kMaxA; ++a)
            if (a < A) g[a] = invM * (coef * ((act == a ? 1.f : 0.f) - p[a]) + beta_ent * p[a] * (lp[a] + ent));
        }
      } else {
        const float diff = logit[0] - ret[o];
        if (valid) {
          loss_acc = fmaf(diff, diff, loss_acc);
          g[0] = invM * 2.f * diff;
        }
      }

      // ---- backward ----
      float d2[FPL];
#pragma unroll
      for (int i = 0; i < FPL; ++i) {
        const int f = f0 + LPS * i;
        float up = 0.f;
#pragma unroll
        for (int a = 0; a < kMaxA; ++a)
          if (a < A) {
            gW3[a][i] = fmaf(g[a], a2[i], gW3[a][i]);
            up = fmaf(g[a], w[L.oW3 + a * H + f], up);
          }
        d2[i] = up * (1.f - a2[i] * a2[i]);
        gb2[i] += d2[i];
        sd2[f] = d2[i];
      }
#pragma unroll
      for (int a = 0; a < kMaxA; ++a) gb3[a] += g[a];
#pragma unroll
      for (int k4 = 0; k4 < H / 4; ++k4) {
        const float4 x = reinterpret_cast<const float4*>(sa1)[k4];
#pragma unroll
        for (int i = 0; i < FPL; ++i) {
          gW2[i][4 * k4 + 0] = fmaf(d2[i], x.x, gW2[i][4 * k4 + 0]);
          gW2[i][4 * k4 + 1] = fmaf(d2[i], x.y, gW2[i][4 * k4 + 1]);
          gW2[i][4 * k4 + 2] = fmaf(d2[i], x.z, gW2[i][4 * k4 + 2]);
          gW2[i][4 * k4 + 3] = fmaf(d2[i], x.w, gW2[i][4 * k4 + 3]);
        }
      }
      __syncwarp();
      float d1[FPL];
#pragma unroll
      for (int i = 0; i < FPL; ++i) d1[i] = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < H / 4; ++j4) {
        const float4 x = reinterpret_cast<const float4*>(sd2)[j4];
#pragma unroll
        for (int i = 0; i < FPL; ++i) {
          const int f = f0 + LPS * i;
          d1[i] = fmaf(w[L.oW2 + (4 * j4 + 0) * H + f], x.x, d1[i]);
          d1[i] = fmaf(w[L.oW2 + (4 * j4 + 1) * H + f], x.y, d1[i]);
          d1[i] = fmaf(w[L.oW2 + (4 * j4 + 2) * H + f], x.z, d1[i]);
          d1[i] = fmaf(w[L.oW2 + (4 * j4 + 3) * H + f], x.w, d1[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < FPL; ++i) {
        d1[i] *= (1.f - a1[i] * a1[i]);
        if (dz1_out && valid) dz1_out[n * z1_arm_stride + mm * z1_sample_stride + f0 + LPS * i] = d1[i];  // dL/dz1 for dW1_nature
        gb1[i] += d1[i];
        if (d.one_hot) {
#pragma unroll
          for (int c = 0; c < kMaxIn - 1; ++c) gW1[i][c] += (si == c) ? d1[i] : 0.f;
          gW1[i][kMaxIn - 1] = fmaf(d1[i], lam, gW1[i][kMaxIn - 1]);  // lambda column kept in the last slot
        } else {
          gW1[i][0] = fmaf(d1[i], x0, gW1[i][0]);
          gW1[i][kMaxIn - 1] = fmaf(d1[i], lam, gW1[i][kMaxIn - 1]);
        }
      }
      __syncwarp();
    }

    // ---- merge: sample sub-groups inside the warp (fixed butterfly), then the 8 warps in order ----
#define RMAB_SUBRED(x)                                                                       \
  {                                                                                          \
    _Pragma("unroll") for (int o_ = 16; o_ >= LPS; o_ >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o_); \
  }
#pragma unroll
    for (int i = 0; i < FPL; ++i) {
      RMAB_SUBRED(gb1[i]);
      RMAB_SUBRED(gb2[i]);
#pragma unroll
      for (int c = 0; c < kMaxIn; ++c) RMAB_SUBRED(gW1[i][c]);
#pragma unroll
      for (int k = 0; k < H; ++k) RMAB_SUBRED(gW2[i][k]);
#pragma unroll
      for (int a = 0; a < kMaxA; ++a) RMAB_SUBRED(gW3[a][i]);
    }
#pragma unroll
    for (int a = 0; a < kMaxA; ++a) RMAB_SUBRED(gb3[a]);
#undef RMAB_SUBRED
    for (int r = 0; r < NW; ++r) {
      if (warp == r && lane < LPS) {
#pragma unroll
        for (int i = 0; i < FPL; ++i) {
          const int f = f0 + LPS * i;
          G[L.ob1 + f] += gb1[i];
          G[L.ob2 + f] += gb2[i];
          if (d.one_hot) {
#pragma unroll
            for (int c = 0; c < kMaxIn - 1; ++c)
              if (c < in - 1) G[L.oW1 + f * in + c] += gW1[i][c];
          } else {
            G[L.oW1 + f * in] += gW1[i][0];
          }
          G[L.oW1 + f * in + in - 1] += gW1[i][kMaxIn - 1];
#pragma unroll
          for (int k = 0; k < H; ++k) G[L.oW2 + f * H + k] += gW2[i][k];
#pragma unroll
          for (int a = 0; a < kMaxA; ++a)
            if (a < A) G[L.oW3 + a * H + f] += gW3[a][i];
        }
        if (lane == 0)
#pragma unroll
          for (int a = 0; a < kMaxA; ++a)
            if (a < A) G[L.ob3 + a] += gb3[a];
      }
      __syncthreads();
    }
    const float loss_tot = block_sum(f0 == 0 ? loss_acc : 0.f, red);
    if (loss_out && threadIdx.x == 0) {
      float extra = 0.f;
      if (TABLE && !IS_PI && arm_stats) extra = arm_stats[4 * (int64_t)n + 2];  // sum (ret - mean_ret[s,e])^2
      loss_out[(int64_t)n * iters + it] = (loss_tot + extra) * invM;
    }

    // ---- Adam (torch.optim.Adam defaults; _single_tensor_adam form) ----
    const int step = step0 + it + 1;
    // python-double scalars of torch's _single_tensor_adam, cast to fp32 where torch casts them
    const double bc1 = 1.0 - pow(hp.beta1, (double)step);
    const double bc2 = 1.0 - pow(hp.beta2, (double)step);
    const float step_size = (float)(lr / bc1);
    const float bc2s = (float)sqrt(bc2);
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
      const float gi = G[i];
      const float mi = am[i] + lerp_w * (gi - am[i]);                     // exp_avg.lerp_(grad, 1-beta1)
      const float vi = av[i] * beta2f + (omb2 * gi) * gi;                 // mul_(beta2).addcmul_(g, g, 1-beta2)
      am[i] = mi;
      av[i] = vi;
      const float denom = sqrtf(vi) / bc2s + epsf;
      const float wi = w[i] - step_size * (mi / denom);
      w[i] = wi;
      G[i] = 0.f;
      if (i >= L.oW2 && i < L.ob2) {
        const int r = i - L.oW2;
        w2t[(r % H) * H + (r / H)] = wi;
      }
    }
    __syncthreads();
  }

  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    Wn[i] = w[i];
    An[i] = am[i];
    An[P + i] = av[i];
  }
}

template <int H, bool IS_PI, bool TABLE>
static int launch_ppo(const RmabNetDesc& d, const RmabHyper& hp, int T, int s_rows, float* W, float* adam,
                      const uint8_t* s, const uint8_t* a, const float* adv, const float* logp_old, const float* ret,
                      const float* lamb, const float* stats, const float* arm_stats, float* loss, cudaStream_t st,
                      const float* z1_extra = nullptr, float* dz1_out = nullptr, int z1_sample_major = 0) {
  using Cfg = PpoCfg<H>;
  const int in = (d.one_hot ? d.n_states : 1) + 1;
  const MlpLayout L(in, H, IS_PI ? d.n_actions : 1);
  size_t smem = sizeof(float) * (4 * (size_t)L.padded() + H * H + 8 * 2 * Cfg::SPW * H);
  int stage = 0;
  if (TABLE) {
    const int SA = d.n_states * d.n_actions;
    const size_t rows = IS_PI ? 3 * SA + d.n_states : 2 * d.n_states;
    const size_t extra = rows * (size_t)d.n_envs * sizeof(float);
    if (smem + extra <= 200 * 1024) { smem += extra; stage = 1; }
  }
  auto kern = ppo_update_kernel<H, IS_PI, TABLE>;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(RMAB_E_LAUNCH, "rmab_ppo_update: cannot reserve %zu B of shared memory", smem);
  kern<<<d.n_arms, 256, smem, st>>>(d, hp, T, s_rows, W, adam, s, a, adv, logp_old, ret, lamb, stats, arm_stats, stage,
                                    loss, z1_extra, dz1_out,
                                    z1_sample_major ? (int64_t)H : (int64_t)T * d.n_envs * H,
                                    z1_sample_major ? (int64_t)d.n_arms * H : (int64_t)H);
  return check_launch(IS_PI ? "rmab_ppo_update(pi)" : "rmab_ppo_update(v)");
}

static int check_ppo_args(const RmabNetDesc* net, const RmabHyper* hp, int T, const char* who) {
  RMAB_REQUIRE(net && hp, RMAB_E_BADARG, "%s: null descriptor", who);
  const int h = net->hidden;
  RMAB_REQUIRE(h == 8 || h == 16 || h == 32 || h == 64, RMAB_E_SHAPE,
               "%s: hidden width %d unsupported (8, 16, 32 or 64; two hidden layers)", who, h);
  RMAB_REQUIRE(net->n_actions >= 2 && net->n_actions <= kMaxA, RMAB_E_SHAPE, "%s: n_actions %d not in [2,%d]", who,
               net->n_actions, kMaxA);
  RMAB_REQUIRE(net->n_extra == 0, RMAB_E_SHAPE, "%s: extra critic inputs not supported here", who);
  RMAB_REQUIRE(!net->one_hot || net->n_states + 1 <= kMaxIn, RMAB_E_SHAPE, "%s: S=%d too large for one-hot", who,
               net->n_states);
  RMAB_REQUIRE(T > 0 && net->n_envs > 0 && (int64_t)T * net->n_envs < (1ll << 30), RMAB_E_SHAPE, "%s: bad T/E", who);
  RMAB_REQUIRE(hp->pi_iters >= 0 && hp->v_iters >= 0, RMAB_E_BADARG, "%s: negative iteration count", who);
  return RMAB_OK;
}

}  // namespace rmab

using namespace rmab;

namespace rmab {
// tcgen05 / TMEM kernels for hidden 64 with the scalar state input (ppo_table_umma64.cu, per-sample row modes)
bool ppo_sample_umma64_fits(const RmabNetDesc* net);
int ppo_sample_umma64_run(const RmabNetDesc* net, const RmabHyper* hp, int T, int s_rows, float* Wpi, float* Wv,
                          float* adam_pi, float* adam_v, const uint8_t* s, const uint8_t* a, const float* adv,
                          const float* logp_old, const float* ret, const float* lamb, float* loss_pi, float* loss_v,
                          cudaStream_t st);
int ppo_sample_umma64_critic_extra(const RmabNetDesc* net, const RmabHyper* hp, int T, int s_rows, float* Wv, float* adam_v,
                                   const uint8_t* s, const float* ret, const float* lamb, const float* z1_extra,
                                   float* dz1_out, float* loss_v, cudaStream_t st);
// RMAB_H64_MMA_SYNC=1 keeps hidden 64 off the tcgen05 kernels (A/B measurements)
static bool h64_legacy() {
  static const bool v = getenv("RMAB_H64_MMA_SYNC") != nullptr;
  return v;
}

}  // namespace rmab

extern "C" int rmab_ppo_update(const RmabNetDesc* net, const RmabHyper* hp, int T, int s_rows, float* Wpi, float* Wv,
                               float* adam_pi, float* adam_v, const uint8_t* s, const uint8_t* a, const float* adv,
                               const float* logp_old, const float* ret, const float* lamb, float* loss_pi,
                               float* loss_v, rmab_stream_t stream) {
  int rc = check_ppo_args(net, hp, T, "rmab_ppo_update");
  if (rc) return rc;
  RMAB_REQUIRE(s_rows >= T, RMAB_E_SHAPE, "rmab_ppo_update: s_rows < T");
  RMAB_REQUIRE(Wpi && Wv && adam_pi && adam_v && s && a && adv && logp_old && ret && lamb, RMAB_E_BADARG,
               "rmab_ppo_update: null pointer");
  if (net->n_arms == 0) return RMAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (net->hidden == 64 && net->n_extra == 0 && !h64_legacy() && ppo_sample_umma64_fits(net))
    return ppo_sample_umma64_run(net, hp, T, s_rows, Wpi, Wv, adam_pi, adam_v, s, a, adv, logp_old, ret, lamb, loss_pi,
                                 loss_v, st);
#define RUN(H)                                                                                                    \
  {                                                                                                               \
    if (hp->pi_iters > 0)                                                                                         \
      rc = launch_ppo<H, true, false>(*net, *hp, T, s_rows, Wpi, adam_pi, s, a, adv, logp_old, ret, lamb, nullptr, \
                                      nullptr, loss_pi, st);                                                      \
    if (rc == RMAB_OK && hp->v_iters > 0)                                                                         \
      rc = launch_ppo<H, false, false>(*net, *hp, T, s_rows, Wv, adam_v, s, a, adv, logp_old, ret, lamb, nullptr,  \
                                       nullptr, loss_v, st);                                                      \
  }
  switch (net->hidden) {
    case 8: RUN(8) break;
    case 16: RUN(16) break;
    case 32: RUN(32) break;
    default: RUN(64) break;
  }
#undef RUN
  return rc;
}

extern "C" int rmab_critic_extra_iter(const RmabNetDesc* net, const RmabHyper* hp, int T, int s_rows, float* Wv,
                                      float* adam_v, const uint8_t* s, const float* ret, const float* lamb,
                                      const float* z1_extra, float* dz1_out, int z1_sample_major, float* loss_v,
                                      rmab_stream_t stream) {
  RMAB_REQUIRE(net && hp, RMAB_E_BADARG, "rmab_critic_extra_iter: null descriptor");
  RmabNetDesc d = *net;
  d.n_extra = 0;  // the packed part of the critic has the plain (state, lambda) inputs
  RmabHyper h1 = *hp;
  h1.pi_iters = 0;
  h1.v_iters = 1;
  int rc = check_ppo_args(&d, &h1, T, "rmab_critic_extra_iter");
  if (rc) return rc;
  RMAB_REQUIRE(s_rows >= T && Wv && adam_v && s && ret && lamb && z1_extra && dz1_out, RMAB_E_BADARG,
               "rmab_critic_extra_iter: null pointer or s_rows < T");
  if (d.n_arms == 0) return RMAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (d.hidden == 64 && z1_sample_major && !h64_legacy() && ppo_sample_umma64_fits(&d))
    return ppo_sample_umma64_critic_extra(&d, &h1, T, s_rows, Wv, adam_v, s, ret, lamb, z1_extra, dz1_out, loss_v, st);
#define RUN(H) rc = launch_ppo<H, false, false>(d, h1, T, s_rows, Wv, adam_v, s, nullptr, nullptr, nullptr, ret, lamb, \
                                                nullptr, nullptr, loss_v, st, z1_extra, dz1_out, z1_sample_major);
  switch (d.hidden) {
    case 8: RUN(8) break;
    case 16: RUN(16) break;
    case 32: RUN(32) break;
    default: RUN(64) break;
  }
#undef RUN
  return rc;
}
