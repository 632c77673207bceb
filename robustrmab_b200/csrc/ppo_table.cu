// Statistics-based PPO update for the one-hot table domains: update() (agent_oracle.py:399-436) on the rows
// written by rmab_epoch_table -- one row per (state, env) instead of one per sample (identity in epoch.cu).
// All train_pi_iters (or train_v_iters) full-batch Adam steps of one arm run inside one CTA with the weights,
// Adam moments and the arm's statistics resident in shared memory.
//
// Reference behaviour: compute_loss_pi :285-322, compute_loss_v :325-339, Adam :383-388 (torch defaults).
//
// Structure per iteration, per chunk of RC rows (activations live in shared memory as [feature][row] matrices):
//   P1  a1 = tanh(W1 x + b1)                      thread per row (x = onehot(s) ++ lambda, so W1 x is a lookup)
//   P2  a2 = tanh(W2 a1 + b2)                     register-tiled GEMM, TJ features x 4 rows per thread
//   P3  logits, loss, dL/dlogit, d2 = W3^T g (1-a2^2)   thread per row
//   P5  dW2 += d2 a1^T                            register-tiled GEMM over the rows, 4x4 tile per thread
//   P4  d1 = (W2^T d2)(1-a1^2)                    register-tiled GEMM (overwrites a2); the bias / W1 / W3
//                                                 gradients are row sums of tiles this phase already holds
// then one fixed-order reduction of the partial sums, Adam, and the refresh of the derived weight tables.
// FFMA-bound: 6 * P flops per row per iteration; every shared-memory wavefront feeds >= 8 FMA instructions.
#include <stdlib.h>
#include "common.cuh"
#include "mlp.cuh"

namespace rmab {

template <int H>
struct TblCfg {
  static constexpr int RC = H >= 64 ? 128 : 256;  // rows per chunk
  static constexpr int LDA = RC + 4;              // padded row length of the activation matrices
  static constexpr int TJ = H * RC / 1024;        // features per thread in P2 / P4 (x 4 rows)
  static constexpr int NJB = H / TJ;              // feature blocks
  static constexpr int WT = H >= 16 ? 4 : 2;      // dW2 tile edge
  static constexpr int NT = (H / WT) * (H / WT);  // dW2 tiles
  static constexpr int NSL = 256 / NT;            // row slices per tile
  static constexpr int RPS = RC / NSL;            // rows per slice
  static constexpr int NB2 = H / NJB;             // db2 features accumulated per thread (j = jb + NJB * i)
};

// tanh_fast: mlp.cuh

struct TblArgs {
  RmabNetDesc d;
  RmabHyper hp;
  int T, stage_stats;
  float* W; float* adam;
  const float* lamb; const float* stats; const float* arm_stats;
  float* loss_out;
};

template <int H, int S, bool IS_PI>
__global__ void __launch_bounds__(256, (H >= 64 ? 1 : 2)) ppo_table_kernel(TblArgs g) {
  using C = TblCfg<H>;
  constexpr int RC = C::RC, LDA = C::LDA, TJ = C::TJ, NJB = C::NJB, WT = C::WT, NT = C::NT, NSL = C::NSL, RPS = C::RPS,
                NB2 = C::NB2;
  constexpr int A = 2, OUT = IS_PI ? A : 1, IN = S + 1, SA = S * A;
  constexpr int ROW0 = IS_PI ? 0 : 3 * SA, NROWS = IS_PI ? 3 * SA + S : 2 * S, KST = 4 * SA + 3 * S;
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int E = g.d.n_envs, T = g.T;
  const MlpLayout L(IN, H, OUT);
  const int P = L.P, Pp = L.padded();
  // ---- shared memory carve-up (every region starts on a multiple of 4 floats) ----
  float* w = smem;                    // packed weights [P]: W1 b1 W2 b2 W3 b3
  float* G = w + Pp;                  // gradient [P]
  float* am = G + Pp;                 // Adam m
  float* av = am + Pp;                // Adam v
  float* w2t = av + Pp;               // W2 transposed [k][j]
  float* wb = w2t + H * H;            // W1[:, s] + b1 per state [S][H]
  float* wl = wb + S * H;             // W1[:, lambda] [H]
  float* B1 = wl + H;                 // a1  [H][LDA]
  float* B2 = B1 + H * LDA;           // a2, later d1  [H][LDA]
  float* B3 = B2 + H * LDA;           // d2  [H][LDA]
  float* g3 = B3 + H * LDA;           // dL/dlogit [OUT][RC]
  float* rlam = g3 + OUT * RC;        // lambda of each row of the chunk [RC]
  float* roh = rlam + RC;             // one-hot state of each row of the chunk [S][RC]
  float* slam = roh + S * RC;         // lambda per env [E]
  float* sst = slam + ((E + 3) & ~3); // staged statistics rows [NROWS][E]
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  float* Wn = g.W + (int64_t)n * P;
  float* An = g.adam + (int64_t)n * 2 * P;
  for (int i = tid; i < P; i += 256) {
    w[i] = Wn[i];
    am[i] = An[i];
    av[i] = An[P + i];
  }
  for (int i = tid; i < E; i += 256) slam[i] = g.lamb[i];
  const float* st;  // st[(row - ROW0) * E + e]
  {
    const float* gst = g.stats + ((int64_t)n * KST + ROW0) * E;
    if (g.stage_stats) {
      for (int i = tid; i < NROWS * E; i += 256) sst[i] = gst[i];
      st = sst;
    } else {
      st = gst;
    }
  }
  __syncthreads();

  const int M = S * E;
  const int n_chunks = (M + RC - 1) / RC;
  const float invM = 1.f / (float)((int64_t)T * E);
  const int iters = IS_PI ? g.hp.pi_iters : g.hp.v_iters;
  const double lr = IS_PI ? g.hp.pi_lr : g.hp.vf_lr;
  const int step0 = IS_PI ? g.hp.adam_step0_pi : g.hp.adam_step0_v;
  const float clip_lo = (float)(1.0 - g.hp.clip_ratio), clip_hi = (float)(1.0 + g.hp.clip_ratio);
  const float beta_ent = (float)g.hp.entropy_coeff;
  const float lerp_w = (float)(1.0 - g.hp.beta1), beta2f = (float)g.hp.beta2, omb2 = (float)(1.0 - g.hp.beta2);
  const float epsf = (float)g.hp.adam_eps;

  // thread roles
  const int jb = tid % NJB, rb = tid / NJB;          // P2 / P4: features [TJ*jb, +TJ), rows [4*rb, +4)
  const int tile = tid % NT, slice = tid / NT;       // P5a: dW2 tile and row slice
  const int tj = tile / (H / WT), tk = tile % (H / WT);

  double b1p = pow(g.hp.beta1, (double)step0), b2p = pow(g.hp.beta2, (double)step0);
  for (int it = 0; it < iters; ++it) {
    // ---- derived weight tables for this iteration ----
    for (int i = tid; i < H * H; i += 256) w2t[(i % H) * H + (i / H)] = w[L.oW2 + i];
    for (int i = tid; i < S * H; i += 256) {
      const int s = i / H, k = i - s * H;
      wb[i] = w[L.oW1 + k * IN + s] + w[L.ob1 + k];
    }
    for (int i = tid; i < H; i += 256) wl[i] = w[L.oW1 + i * IN + IN - 1];
    __syncthreads();

    float accW2[WT][WT];
#pragma unroll
    for (int i = 0; i < WT; ++i)
#pragma unroll
      for (int j = 0; j < WT; ++j) accW2[i][j] = 0.f;
    // row-sum gradients accumulated by the P4 tile owner (features TJ*jb + j, rows of block rb)
    float s_w3[OUT][TJ], s_b1[TJ], s_wl[TJ], s_w1[S][TJ], s_b2[NB2], s_b3[OUT];
#pragma unroll
    for (int j = 0; j < TJ; ++j) {
      s_b1[j] = 0.f; s_wl[j] = 0.f;
#pragma unroll
      for (int a = 0; a < OUT; ++a) s_w3[a][j] = 0.f;
#pragma unroll
      for (int c = 0; c < S; ++c) s_w1[c][j] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB2; ++i) s_b2[i] = 0.f;
#pragma unroll
    for (int a = 0; a < OUT; ++a) s_b3[a] = 0.f;
    float loss_acc = 0.f;

    for (int ch = 0; ch < n_chunks; ++ch) {
      const int m0 = ch * RC;
      // ---- P1: a1[k][row] = tanh(wb[s][k] + wl[k] * lambda) ----
      if (tid < RC) {
        const int m = m0 + tid < M ? m0 + tid : M - 1;
        const int s = m / E, e = m - s * E;
        const float lam = slam[e];
        const float* wbs = wb + s * H;
        const bool live = m0 + tid < M;
        rlam[tid] = live ? lam : 0.f;
#pragma unroll
        for (int c = 0; c < S; ++c) roh[c * RC + tid] = (live && s == c) ? 1.f : 0.f;
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          const float4 b = *reinterpret_cast<const float4*>(wbs + 4 * k4);
          const float4 l = *reinterpret_cast<const float4*>(wl + 4 * k4);
          B1[(4 * k4 + 0) * LDA + tid] = tanh_fast(fmaf(l.x, lam, b.x));
          B1[(4 * k4 + 1) * LDA + tid] = tanh_fast(fmaf(l.y, lam, b.y));
          B1[(4 * k4 + 2) * LDA + tid] = tanh_fast(fmaf(l.z, lam, b.z));
          B1[(4 * k4 + 3) * LDA + tid] = tanh_fast(fmaf(l.w, lam, b.w));
        }
      }
      __syncthreads();
      // ---- P2: a2 = tanh(W2 a1 + b2), tile TJ x 4 ----
      {
        float acc[TJ][4];
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
          const float b = w[L.ob2 + TJ * jb + j];
          acc[j][0] = b; acc[j][1] = b; acc[j][2] = b; acc[j][3] = b;
        }
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
          const float4 x = *reinterpret_cast<const float4*>(B1 + k * LDA + 4 * rb);
          float wv[TJ];
          if (TJ >= 4) {
#pragma unroll
            for (int q = 0; q < TJ / 4; ++q) {
              const float4 t = *reinterpret_cast<const float4*>(w2t + k * H + TJ * jb + 4 * q);
              wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < TJ; ++j) wv[j] = w2t[k * H + TJ * jb + j];
          }
#pragma unroll
          for (int j = 0; j < TJ; ++j) {
            acc[j][0] = fmaf(wv[j], x.x, acc[j][0]);
            acc[j][1] = fmaf(wv[j], x.y, acc[j][1]);
            acc[j][2] = fmaf(wv[j], x.z, acc[j][2]);
            acc[j][3] = fmaf(wv[j], x.w, acc[j][3]);
          }
        }
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
          float4 o;
          o.x = tanh_fast(acc[j][0]); o.y = tanh_fast(acc[j][1]); o.z = tanh_fast(acc[j][2]); o.w = tanh_fast(acc[j][3]);
          *reinterpret_cast<float4*>(B2 + (TJ * jb + j) * LDA + 4 * rb) = o;
        }
      }
      __syncthreads();
      // ---- P3: logits, loss, dL/dlogit, d2 (thread per row) ----
      if (tid < RC) {
        const int m = m0 + tid;
        const bool valid = m < M;
        const int mm = valid ? m : M - 1;
        const int s = mm / E, e = mm - s * E;
        float logit[OUT], gl[OUT];
#pragma unroll
        for (int a = 0; a < OUT; ++a) { logit[a] = w[L.ob3 + a]; gl[a] = 0.f; }
        float a2v[H];
#pragma unroll
        for (int j = 0; j < H; ++j) {
          a2v[j] = B2[j * LDA + tid];
#pragma unroll
          for (int a = 0; a < OUT; ++a) logit[a] = fmaf(w[L.oW3 + a * H + j], a2v[j], logit[a]);
        }
        const float wrow = st[(int64_t)(3 * SA - ROW0 + s) * E + e];  // CntS: samples of env e in state s
        if (IS_PI) {
          // Categorical(logits) (rmabppo_core.py:79-81): logp = z - max - log(sum exp)
          const float mx = fmaxf(logit[0], logit[OUT - 1]);
          const float e0 = __expf(logit[0] - mx), e1 = __expf(logit[OUT - 1] - mx);
          const float sum = e0 + e1, lse = __logf(sum), inv = __fdividef(1.f, sum);
          const float p[2] = {e0 * inv, e1 * inv};
          const float lp[2] = {(logit[0] - mx) - lse, (logit[OUT - 1] - mx) - lse};
          const float ent = -(p[0] * lp[0] + p[1] * lp[1]);
          float ga[2], sumg = 0.f, lrow = 0.f;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int c = s * A + a;
            const float Ap = st[(int64_t)c * E + e], Am = st[(int64_t)(SA + c) * E + e];
            const float ratio = __expf(lp[a] - st[(int64_t)(2 * SA + c) * E + e]);
            // positive advantages: min(r, 1+c) * A ; negative: max(r, 1-c) * A   (:301-305)
            lrow -= fminf(ratio, clip_hi) * Ap + fmaxf(ratio, clip_lo) * Am;
            ga[a] = -((ratio <= clip_hi ? ratio * Ap : 0.f) + (ratio >= clip_lo ? ratio * Am : 0.f));
            sumg += ga[a];
          }
          if (valid) {
            loss_acc += lrow - beta_ent * wrow * ent;
#pragma unroll
            for (int a = 0; a < 2; ++a)
              gl[a < OUT ? a : 0] = invM * (ga[a] - p[a] * sumg + beta_ent * wrow * p[a] * (lp[a] + ent));
          }
        } else {
          const float diff = logit[0] - st[(int64_t)(3 * SA + S - ROW0 + s) * E + e];
          if (valid) {
            loss_acc = fmaf(wrow * diff, diff, loss_acc);
            gl[0] = invM * 2.f * wrow * diff;
          }
        }
#pragma unroll
        for (int a = 0; a < OUT; ++a) g3[a * RC + tid] = gl[a];
#pragma unroll
        for (int j = 0; j < H; ++j) {
          float up = 0.f;
#pragma unroll
          for (int a = 0; a < OUT; ++a) up = fmaf(gl[a], w[L.oW3 + a * H + j], up);
          B3[j * LDA + tid] = up * (1.f - a2v[j] * a2v[j]);
        }
      }
      __syncthreads();
      // ---- P5: dW2 tile += d2 a1^T over this thread's row slice ----
      {
        const int r0 = slice * RPS;
#pragma unroll 2
        for (int r = 0; r < RPS; r += 4) {
          float4 dv[WT], xv[WT];
#pragma unroll
          for (int i = 0; i < WT; ++i) {
            dv[i] = *reinterpret_cast<const float4*>(B3 + (WT * tj + i) * LDA + r0 + r);
            xv[i] = *reinterpret_cast<const float4*>(B1 + (WT * tk + i) * LDA + r0 + r);
          }
#pragma unroll
          for (int i = 0; i < WT; ++i)
#pragma unroll
            for (int j = 0; j < WT; ++j) {
              accW2[i][j] = fmaf(dv[i].x, xv[j].x, accW2[i][j]);
              accW2[i][j] = fmaf(dv[i].y, xv[j].y, accW2[i][j]);
              accW2[i][j] = fmaf(dv[i].z, xv[j].z, accW2[i][j]);
              accW2[i][j] = fmaf(dv[i].w, xv[j].w, accW2[i][j]);
            }
        }
      }
      __syncthreads();
      // ---- P4: d1 = (W2^T d2) (1 - a1^2), tile TJ x 4, written over a2; row-sum gradients on the way ----
      {
        float acc[TJ][4];
#pragma unroll
        for (int j = 0; j < TJ; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
#pragma unroll 8
        for (int j2 = 0; j2 < H; ++j2) {
          const float4 x = *reinterpret_cast<const float4*>(B3 + j2 * LDA + 4 * rb);
          float wv[TJ];
          if (TJ >= 4) {
#pragma unroll
            for (int q = 0; q < TJ / 4; ++q) {
              const float4 t = *reinterpret_cast<const float4*>(w + L.oW2 + j2 * H + TJ * jb + 4 * q);
              wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < TJ; ++j) wv[j] = w[L.oW2 + j2 * H + TJ * jb + j];
          }
#pragma unroll
          for (int j = 0; j < TJ; ++j) {
            acc[j][0] = fmaf(wv[j], x.x, acc[j][0]);
            acc[j][1] = fmaf(wv[j], x.y, acc[j][1]);
            acc[j][2] = fmaf(wv[j], x.z, acc[j][2]);
            acc[j][3] = fmaf(wv[j], x.w, acc[j][3]);
          }
        }
        // db2[j] = sum_rows d2[j][row] for this thread's share of the features (j = jb + NJB * i)
#pragma unroll
        for (int i = 0; i < NB2; ++i) {
          const float4 x = *reinterpret_cast<const float4*>(B3 + (jb + NJB * i) * LDA + 4 * rb);
          s_b2[i] += (x.x + x.y) + (x.z + x.w);
        }
        float4 gv[OUT];
#pragma unroll
        for (int a = 0; a < OUT; ++a) {
          gv[a] = *reinterpret_cast<const float4*>(g3 + a * RC + 4 * rb);
          if (jb == 0) s_b3[a] += (gv[a].x + gv[a].y) + (gv[a].z + gv[a].w);
        }
        const float4 lm = *reinterpret_cast<const float4*>(rlam + 4 * rb);
        float4 oh[S];
#pragma unroll
        for (int c = 0; c < S; ++c) oh[c] = *reinterpret_cast<const float4*>(roh + c * RC + 4 * rb);
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
          float* cell = B2 + (TJ * jb + j) * LDA + 4 * rb;
          const float4 a2 = *reinterpret_cast<const float4*>(cell);   // dW3[a][j] += g_a * a2 before a2 is overwritten
#pragma unroll
          for (int a = 0; a < OUT; ++a)
            s_w3[a][j] += fmaf(gv[a].x, a2.x, gv[a].y * a2.y) + fmaf(gv[a].z, a2.z, gv[a].w * a2.w);
          const float4 a1 = *reinterpret_cast<const float4*>(B1 + (TJ * jb + j) * LDA + 4 * rb);
          float4 o;
          o.x = acc[j][0] * (1.f - a1.x * a1.x);
          o.y = acc[j][1] * (1.f - a1.y * a1.y);
          o.z = acc[j][2] * (1.f - a1.z * a1.z);
          o.w = acc[j][3] * (1.f - a1.w * a1.w);
          s_b1[j] += (o.x + o.y) + (o.z + o.w);
          s_wl[j] += fmaf(o.x, lm.x, o.y * lm.y) + fmaf(o.z, lm.z, o.w * lm.w);
#pragma unroll
          for (int c = 0; c < S; ++c)
            s_w1[c][j] += fmaf(o.x, oh[c].x, o.y * oh[c].y) + fmaf(o.z, oh[c].z, o.w * oh[c].w);
        }
      }
      __syncthreads();  // the next chunk's P1 overwrites B1 / the row metadata
    }

    // ---- reduce the partial sums in a fixed order into G ----
    // dW2: every thread parks its tile partial in the (now free) activation region, then NSL slices are summed
    float* park = B1;  // NSL * H * H floats <= 3 * H * LDA
#pragma unroll
    for (int i = 0; i < WT; ++i)
#pragma unroll
      for (int j = 0; j < WT; ++j) park[slice * (H * H) + (WT * tj + i) * H + WT * tk + j] = accW2[i][j];
    // row sums: threads with the same jb (lanes congruent mod NJB) hold partials of the same features; butterfly over
    // the row-block bits of the lane, then the 8 warps through shared memory in warp order
    constexpr int NV = TJ * (OUT + 2 + S) + NB2 + OUT;
    float* rs = park + NSL * H * H;  // [8 warps][NJB][NV]
    {
      float vals[NV];
      int q = 0;
#pragma unroll
      for (int j = 0; j < TJ; ++j) {
        vals[q++] = s_b1[j];
        vals[q++] = s_wl[j];
#pragma unroll
        for (int c = 0; c < S; ++c) vals[q++] = s_w1[c][j];
#pragma unroll
        for (int a = 0; a < OUT; ++a) vals[q++] = s_w3[a][j];
      }
#pragma unroll
      for (int i = 0; i < NB2; ++i) vals[q++] = s_b2[i];
#pragma unroll
      for (int a = 0; a < OUT; ++a) vals[q++] = s_b3[a];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float x = vals[v];
#pragma unroll
        for (int o = 16; o >= NJB; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane < NJB) rs[(warp * NJB + lane) * NV + v] = x;
      }
    }
    const float loss_tot = block_sum(loss_acc, red);  // contains __syncthreads
    __syncthreads();
    for (int i = tid; i < H * H; i += 256) {
      float v = park[i];
#pragma unroll
      for (int sl = 1; sl < NSL; ++sl) v += park[sl * (H * H) + i];
      G[L.oW2 + i] = v;
    }
    for (int i = tid; i < NJB * NV; i += 256) {
      const int b = i / NV, v = i - b * NV;
      float x = rs[i];
#pragma unroll
      for (int wq = 1; wq < 8; ++wq) x += rs[wq * NJB * NV + i];
      // decode slot v of feature block b (same order as vals[] above)
      if (v < TJ * (OUT + 2 + S)) {
        const int j = v / (OUT + 2 + S), r = v - j * (OUT + 2 + S), k = TJ * b + j;
        if (r == 0) G[L.ob1 + k] = x;
        else if (r == 1) G[L.oW1 + k * IN + IN - 1] = x;
        else if (r < 2 + S) G[L.oW1 + k * IN + (r - 2)] = x;
        else G[L.oW3 + (r - 2 - S) * H + k] = x;
      } else if (v < TJ * (OUT + 2 + S) + NB2) {
        G[L.ob2 + b + NJB * (v - TJ * (OUT + 2 + S))] = x;
      } else if (b == 0) {
        G[L.ob3 + (v - TJ * (OUT + 2 + S) - NB2)] = x;
      }
    }
    if (g.loss_out && tid == 0) {
      float extra = 0.f;
      if (!IS_PI && g.arm_stats) extra = g.arm_stats[4 * (int64_t)n + 2];  // sum (ret - mean_ret[s,e])^2
      g.loss_out[(int64_t)n * iters + it] = (loss_tot + extra) * invM;
    }
    __syncthreads();

    // ---- Adam (torch.optim.Adam defaults; _single_tensor_adam form) ----
    b1p *= g.hp.beta1;  // beta^step, step = step0 + it + 1
    b2p *= g.hp.beta2;
    const double bc1 = 1.0 - b1p, bc2 = 1.0 - b2p;
    const float step_size = (float)(lr / bc1);
    const float bc2s = (float)sqrt(bc2);
    for (int i = tid; i < P; i += 256) {
      const float gi = G[i];
      const float mi = am[i] + lerp_w * (gi - am[i]);                     // exp_avg.lerp_(grad, 1-beta1)
      const float vi = av[i] * beta2f + (omb2 * gi) * gi;                 // mul_(beta2).addcmul_(g, g, 1-beta2)
      am[i] = mi;
      av[i] = vi;
      const float denom = sqrtf(vi) / bc2s + epsf;
      w[i] = w[i] - step_size * (mi / denom);
    }
    __syncthreads();
  }

  for (int i = tid; i < P; i += 256) {
    Wn[i] = w[i];
    An[i] = am[i];
    An[P + i] = av[i];
  }
}

template <int H, int S, bool IS_PI>
static int launch_table(const TblArgs& args, cudaStream_t st) {
  using C = TblCfg<H>;
  TblArgs g = args;
  constexpr int OUT = IS_PI ? 2 : 1, SA = S * 2;
  const MlpLayout L(S + 1, H, OUT);
  const int E = g.d.n_envs;
  size_t fl = 4 * (size_t)L.padded() + H * H + S * H + H + 3 * (size_t)H * C::LDA + (OUT + 1 + S) * C::RC + (((size_t)E + 3) & ~(size_t)3);
  const size_t rows = IS_PI ? 3 * SA + S : 2 * S;
  g.stage_stats = 0;
  if ((fl + rows * E) * sizeof(float) <= 100 * 1024 || (H >= 64 && (fl + rows * E) * sizeof(float) <= 220 * 1024)) {
    fl += rows * E;
    g.stage_stats = 1;
  }
  const size_t smem = fl * sizeof(float);
  if (smem > 227 * 1024) return fail(RMAB_E_SHAPE, "rmab_ppo_update_table: E=%d needs %zu B of shared memory", E, smem);
  auto kern = ppo_table_kernel<H, S, IS_PI>;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(RMAB_E_LAUNCH, "rmab_ppo_update_table: cannot reserve %zu B of shared memory", smem);
  kern<<<g.d.n_arms, 256, smem, st>>>(g);
  return check_launch(IS_PI ? "rmab_ppo_update_table(pi)" : "rmab_ppo_update_table(v)");
}

template <int H, int S>
static int run_table(const TblArgs& base, float* Wpi, float* Wv, float* adam_pi, float* adam_v, float* loss_pi,
                     float* loss_v, cudaStream_t st) {
  int rc = RMAB_OK;
  TblArgs gp = base, gv = base;
  gp.W = Wpi; gp.adam = adam_pi; gp.loss_out = loss_pi;
  gv.W = Wv; gv.adam = adam_v; gv.loss_out = loss_v;
  if (base.hp.pi_iters > 0) rc = launch_table<H, S, true>(gp, st);
  if (rc == RMAB_OK && base.hp.v_iters > 0) rc = launch_table<H, S, false>(gv, st);
  return rc;
}

// tensor-core variant for H = 16 / 32 (ppo_table_mma.cu)
int ppo_table_mma_run(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv, float* adam_pi,
                      float* adam_v, const float* stats, const float* arm_stats, const float* lamb, float* loss_pi,
                      float* loss_v, cudaStream_t st);
// tensor-core variant for H = 64 (ppo_table_mma64.cu)
int ppo_table_mma64_run(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv, float* adam_pi,
                        float* adam_v, const float* stats, const float* arm_stats, const float* lamb, float* loss_pi,
                        float* loss_v, cudaStream_t st);
// tcgen05 / TMEM variant for H = 16 (ppo_table_umma16.cu)
int ppo_table_umma16_run(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv, float* adam_pi,
                         float* adam_v, const float* stats, const float* arm_stats, const float* lamb, float* loss_pi,
                         float* loss_v, cudaStream_t st);
// tcgen05 / TMEM variant for H = 64 (ppo_table_umma64.cu)
bool ppo_table_umma64_fits(const RmabNetDesc* net);
int ppo_table_umma64_run(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv, float* adam_pi,
                         float* adam_v, const float* stats, const float* arm_stats, const float* lamb, float* loss_pi,
                         float* loss_v, cudaStream_t st);

}  // namespace rmab

using namespace rmab;

extern "C" int rmab_ppo_update_table(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv,
                                     float* adam_pi, float* adam_v, const float* stats, const float* arm_stats,
                                     const float* lamb, float* loss_pi, float* loss_v, rmab_stream_t stream) {
  RMAB_REQUIRE(net && hp, RMAB_E_BADARG, "rmab_ppo_update_table: null descriptor");
  const int h = net->hidden, S = net->n_states;
  RMAB_REQUIRE(h == 8 || h == 16 || h == 32 || h == 64, RMAB_E_SHAPE,
               "rmab_ppo_update_table: hidden width %d unsupported (8, 16, 32 or 64; two hidden layers)", h);
  RMAB_REQUIRE(net->one_hot && (S == 2 || S == 3) && net->n_actions == 2 && net->n_extra == 0, RMAB_E_SHAPE,
               "rmab_ppo_update_table: needs the one-hot table domains (S = 2 or 3, A = 2)");
  RMAB_REQUIRE(T > 0 && net->n_envs > 0 && net->n_arms >= 0, RMAB_E_SHAPE, "rmab_ppo_update_table: bad T/E/N");
  RMAB_REQUIRE(hp->pi_iters >= 0 && hp->v_iters >= 0, RMAB_E_BADARG, "rmab_ppo_update_table: negative iteration count");
  RMAB_REQUIRE(Wpi && Wv && adam_pi && adam_v && stats && lamb, RMAB_E_BADARG, "rmab_ppo_update_table: null pointer");
  RMAB_REQUIRE(arm_stats || !loss_v, RMAB_E_BADARG, "rmab_ppo_update_table: loss_v needs arm_stats");
  if (net->n_arms == 0) return RMAB_OK;
  TblArgs g = {};
  g.d = *net; g.hp = *hp; g.T = T; g.lamb = lamb; g.stats = stats; g.arm_stats = arm_stats;
  cudaStream_t st = (cudaStream_t)stream;
  // H = 16 / 32: 3xTF32 mma.sync kernels; H = 64: tcgen05 kernel (RMAB_H64_MMA_SYNC=1 selects the mma.sync one, also
  // used when E is too large for its shared-memory plan); H = 8 (and RMAB_TABLE_FFMA=1, for A/B measurements): FFMA kernel
  static const bool force_ffma = getenv("RMAB_TABLE_FFMA") != nullptr;
  static const bool h64_mma_sync = getenv("RMAB_H64_MMA_SYNC") != nullptr;
  static const bool h16_mma_sync = getenv("RMAB_H16_MMA_SYNC") != nullptr;
  if (h == 16 && !force_ffma && !h16_mma_sync)
    return ppo_table_umma16_run(net, hp, T, Wpi, Wv, adam_pi, adam_v, stats, arm_stats, lamb, loss_pi, loss_v, st);
  if ((h == 16 || h == 32) && !force_ffma)
    return ppo_table_mma_run(net, hp, T, Wpi, Wv, adam_pi, adam_v, stats, arm_stats, lamb, loss_pi, loss_v, st);
  if (h == 64 && !force_ffma && !h64_mma_sync && ppo_table_umma64_fits(net))
    return ppo_table_umma64_run(net, hp, T, Wpi, Wv, adam_pi, adam_v, stats, arm_stats, lamb, loss_pi, loss_v, st);
  if (h == 64 && !force_ffma)
    return ppo_table_mma64_run(net, hp, T, Wpi, Wv, adam_pi, adam_v, stats, arm_stats, lamb, loss_pi, loss_v, st);
#define RUN(H)                                                                                   \
  return S == 2 ? run_table<H, 2>(g, Wpi, Wv, adam_pi, adam_v, loss_pi, loss_v, st)              \
                : run_table<H, 3>(g, Wpi, Wv, adam_pi, adam_v, loss_pi, loss_v, st);
  switch (h) {
    case 8: RUN(8)
    case 16: RUN(16)
    case 32: RUN(32)
    default: RUN(64)
  }
#undef RUN
}
