// Statistics-based PPO update, tensor-core variant for hidden widths 16 and 32: the three H x H products of every
// Adam iteration (a2 = W2 a1, d1 = W2^T d2, dW2 = d2 a1^T) run on mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, c += lo*hi + hi*lo + hi*hi), which keeps fp32-level accuracy (the 1e-5 parity bar) at ~1/10 of
// the FFMA issue slots.  Everything else is as in ppo_table.cu (same statistics rows, same loss, same Adam).
//
// Warp-local pipeline: every warp owns blocks of 32 consecutive rows (row = state * E + env) and keeps their
// activations in its own shared-memory tiles [feature][row], so the chunk loop needs __syncwarp only:
//   P1  a1 = tanh(W1 x + b1)                  lane per row
//   F   a2 = tanh(W2 a1 + b2)                 A = W2 fragments (pre-split, shared), B = a1 tile, C -> a2 tile
//   P3  logits, loss, dL/dlogit, d2           lane per row
//   B   d1 = (W2^T d2)(1 - a1^2)              A = W2^T fragments, B = d2 tile; db1 / dW1 are row sums of C
//   W   dW2 += d2 a1^T                        A = d2 tile (rows as k), B = a1 tile; db2 / dW3 ride along
// Fragment layouts are those of PTX mma.m16n8k8 (.tf32): g = lane / 4, t = lane % 4,
//   A: a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  B: b0 (t, g) b1 (t+4, g);  C: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
//
// Reference behaviour: agent_oracle.py compute_loss_pi :285-322, compute_loss_v :325-339, update :399-436.
#include "common.cuh"
#include "mlp.cuh"

namespace rmab {

namespace {

constexpr int kLDT = 40;  // tile row length: 32 rows + pad; 40 = 8 (mod 32) makes the B-fragment loads conflict-free

// tanh(x) = 1 - 2 / (2^(2 log2(e) x) + 1): FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA.  Absolute error < 2e-7 (what
// propagates through the next layer); saturates correctly (ex2 -> inf gives 1, ex2 -> 0 gives -1).
__device__ __forceinline__ float tanh_fast_m(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

// x = hi + lo with hi = the top 19 bits of x (a valid TF32 operand) and lo = x - hi exactly; the tensor core reads
// only the TF32 bits of lo, which leaves a relative error of ~2^-21 -- the same order as the dropped lo*lo term.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// c += A * B in 3xTF32: the two cross terms first, the leading term last
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t bhi0,
                                     uint32_t bhi1, uint32_t blo0, uint32_t blo1) {
  mma_tf32(c, alo, bhi0, bhi1);
  mma_tf32(c, ahi, blo0, blo1);
  mma_tf32(c, ahi, bhi0, bhi1);
}

struct MmaArgs {
  RmabNetDesc d;
  RmabHyper hp;
  int T, stage_stats;
  float* W; float* adam;
  const float* lamb; const float* stats; const float* arm_stats;
  float* loss_out;
};

}  // namespace

template <int H, int S, bool IS_PI>
__global__ void __launch_bounds__(256, (H >= 32 ? 1 : 2)) ppo_table_mma_kernel(MmaArgs g) {
  constexpr int MT = H / 16, KS = H / 8, NT8 = H / 8, LDT = kLDT;
  constexpr int A = 2, OUT = IS_PI ? A : 1, IN = S + 1, SA = S * A;
  constexpr int ROW0 = IS_PI ? 0 : 3 * SA, NROWS = IS_PI ? 3 * SA + S : 2 * S, KST = 4 * SA + 3 * S;
  constexpr int FRAG = KS * MT * 32 * 8;          // floats of one pre-split fragment set
  constexpr int WTILE = 3 * H * LDT + (OUT + 1 + S) * 32;  // per-warp floats: a1, a2, d2 tiles, g3, row lambda, row one-hot
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int E = g.d.n_envs, T = g.T;
  const MlpLayout L(IN, H, OUT);
  const int P = L.P, Pp = L.padded();
  float* w = smem;
  float* G = w + Pp;
  float* am = G + Pp;
  float* av = am + Pp;
  float* fragF = av + Pp;            // W2 fragments for the forward product   [KS][MT][32][8] (hi a0..a3, lo a0..a3)
  float* fragB = fragF + FRAG;       // W2^T fragments for the backward product
  float* wb = fragB + FRAG;          // W1[:, s] + b1 per state [S][H]
  float* wl = wb + S * H;            // W1[:, lambda] [H]
  float* tiles = wl + H;             // 8 warps x WTILE
  float* slam = tiles + 8 * WTILE;   // lambda per env [E]
  float* sst = slam + ((E + 3) & ~3);
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  float* A1w = tiles + warp * WTILE;
  float* A2w = A1w + H * LDT;
  float* D2w = A2w + H * LDT;
  float* g3w = D2w + H * LDT;        // [OUT][32]
  float* rlw = g3w + OUT * 32;       // [32]
  float* row = rlw + 32;             // [S][32]

  float* Wn = g.W + (int64_t)n * P;
  float* An = g.adam + (int64_t)n * 2 * P;
  for (int i = tid; i < P; i += 256) {
    w[i] = Wn[i];
    am[i] = An[i];
    av[i] = An[P + i];
  }
  for (int i = tid; i < E; i += 256) slam[i] = g.lamb[i];
  const float* st;
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
  const int n_blocks = (M + 31) / 32;
  const float invM = 1.f / (float)((int64_t)T * E);
  const int iters = IS_PI ? g.hp.pi_iters : g.hp.v_iters;
  const double lr = IS_PI ? g.hp.pi_lr : g.hp.vf_lr;
  const int step0 = IS_PI ? g.hp.adam_step0_pi : g.hp.adam_step0_v;
  const float clip_lo = (float)(1.0 - g.hp.clip_ratio), clip_hi = (float)(1.0 + g.hp.clip_ratio);
  const float beta_ent = (float)g.hp.entropy_coeff;
  const float lerp_w = (float)(1.0 - g.hp.beta1), beta2f = (float)g.hp.beta2, omb2 = (float)(1.0 - g.hp.beta2);
  const float epsf = (float)g.hp.adam_eps;
  double b1p = pow(g.hp.beta1, (double)step0), b2p = pow(g.hp.beta2, (double)step0);

  for (int it = 0; it < iters; ++it) {
    // ---- derived tables: pre-split W2 / W2^T fragments, first-layer lookups ----
    for (int i = tid; i < KS * MT * 32 * 4; i += 256) {
      const int q = i & 3, ln = (i >> 2) & 31, fm = i >> 7;  // fm = ks * MT + mt
      const int mt = fm % MT, ks = fm / MT;
      const int gg = ln >> 2, tt = ln & 3;
      const int r = 16 * mt + gg + ((q & 1) ? 8 : 0), c = 8 * ks + tt + ((q & 2) ? 4 : 0);
      uint32_t hi, lo;
      split_tf32(w[L.oW2 + r * H + c], hi, lo);          // forward: A[m = j][k] = W2[j][k]
      fragF[(fm * 32 + ln) * 8 + q] = __uint_as_float(hi);
      fragF[(fm * 32 + ln) * 8 + 4 + q] = __uint_as_float(lo);
      split_tf32(w[L.oW2 + c * H + r], hi, lo);          // backward: A[m = k][j] = W2[j][k]
      fragB[(fm * 32 + ln) * 8 + q] = __uint_as_float(hi);
      fragB[(fm * 32 + ln) * 8 + 4 + q] = __uint_as_float(lo);
    }
    for (int i = tid; i < S * H; i += 256) {
      const int s = i / H, k = i - s * H;
      wb[i] = w[L.oW1 + k * IN + s] + w[L.ob1 + k];
    }
    for (int i = tid; i < H; i += 256) wl[i] = w[L.oW1 + i * IN + IN - 1];
    __syncthreads();

    float accW2[MT][NT8][4];
    float s_b2[MT][2], s_b1[MT][2], s_wl[MT][2], s_w1[S][MT][2], s_w3[OUT][MT][2], s_b3[OUT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) accW2[mt][nt][q] = 0.f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        s_b2[mt][hh] = 0.f; s_b1[mt][hh] = 0.f; s_wl[mt][hh] = 0.f;
#pragma unroll
        for (int c = 0; c < S; ++c) s_w1[c][mt][hh] = 0.f;
#pragma unroll
        for (int a = 0; a < OUT; ++a) s_w3[a][mt][hh] = 0.f;
      }
    }
#pragma unroll
    for (int a = 0; a < OUT; ++a) s_b3[a] = 0.f;
    float loss_acc = 0.f;

    for (int blk = warp; blk < n_blocks; blk += 8) {
      // ---- P1 (lane per row) ----
      const int m = blk * 32 + lane;
      const bool valid = m < M;
      const int mm = valid ? m : M - 1;
      const int srow = (mm >= E) + (mm >= 2 * E), erow = mm - srow * E;  // S <= 3: no integer division
      {
        const float lam = slam[erow];
        rlw[lane] = valid ? lam : 0.f;
#pragma unroll
        for (int c = 0; c < S; ++c) row[c * 32 + lane] = (valid && srow == c) ? 1.f : 0.f;
        const float* wbs = wb + srow * H;
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          const float4 b = *reinterpret_cast<const float4*>(wbs + 4 * k4);
          const float4 l = *reinterpret_cast<const float4*>(wl + 4 * k4);
          A1w[(4 * k4 + 0) * LDT + lane] = tanh_fast_m(fmaf(l.x, lam, b.x));
          A1w[(4 * k4 + 1) * LDT + lane] = tanh_fast_m(fmaf(l.y, lam, b.y));
          A1w[(4 * k4 + 2) * LDT + lane] = tanh_fast_m(fmaf(l.z, lam, b.z));
          A1w[(4 * k4 + 3) * LDT + lane] = tanh_fast_m(fmaf(l.w, lam, b.w));
        }
      }
      __syncwarp();
      // ---- F: a2 = tanh(W2 a1 + b2) ----
      {
        float c[MT][4][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const float blo = w[L.ob2 + 16 * mt + gq], bhi = w[L.ob2 + 16 * mt + gq + 8];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) { c[mt][nt][0] = blo; c[mt][nt][1] = blo; c[mt][nt][2] = bhi; c[mt][nt][3] = bhi; }
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            split_tf32(A1w[(8 * ks + tq) * LDT + 8 * nt + gq], bh[nt][0], bl[nt][0]);
            split_tf32(A1w[(8 * ks + tq + 4) * LDT + 8 * nt + gq], bh[nt][1], bl[nt][1]);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const float4 fh = *reinterpret_cast<const float4*>(fragF + ((ks * MT + mt) * 32 + lane) * 8);
            const float4 fl = *reinterpret_cast<const float4*>(fragF + ((ks * MT + mt) * 32 + lane) * 8 + 4);
            const uint32_t ahi[4] = {__float_as_uint(fh.x), __float_as_uint(fh.y), __float_as_uint(fh.z), __float_as_uint(fh.w)};
            const uint32_t alo[4] = {__float_as_uint(fl.x), __float_as_uint(fl.y), __float_as_uint(fl.z), __float_as_uint(fl.w)};
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma3(c[mt][nt], ahi, alo, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            float2 lo2, hi2;
            lo2.x = tanh_fast_m(c[mt][nt][0]); lo2.y = tanh_fast_m(c[mt][nt][1]);
            hi2.x = tanh_fast_m(c[mt][nt][2]); hi2.y = tanh_fast_m(c[mt][nt][3]);
            *reinterpret_cast<float2*>(A2w + (16 * mt + gq) * LDT + 8 * nt + 2 * tq) = lo2;
            *reinterpret_cast<float2*>(A2w + (16 * mt + gq + 8) * LDT + 8 * nt + 2 * tq) = hi2;
          }
      }
      __syncwarp();
      // ---- P3 (lane per row): logits, loss, dL/dlogit, d2 ----
      {
        float logit[OUT], gl[OUT];
#pragma unroll
        for (int a = 0; a < OUT; ++a) { logit[a] = w[L.ob3 + a]; gl[a] = 0.f; }
        float a2v[H];
#pragma unroll
        for (int j = 0; j < H; ++j) a2v[j] = A2w[j * LDT + lane];
#pragma unroll
        for (int a = 0; a < OUT; ++a)
#pragma unroll
          for (int j4 = 0; j4 < H / 4; ++j4) {
            const float4 w3 = *reinterpret_cast<const float4*>(w + L.oW3 + a * H + 4 * j4);
            logit[a] = fmaf(w3.x, a2v[4 * j4], logit[a]);
            logit[a] = fmaf(w3.y, a2v[4 * j4 + 1], logit[a]);
            logit[a] = fmaf(w3.z, a2v[4 * j4 + 2], logit[a]);
            logit[a] = fmaf(w3.w, a2v[4 * j4 + 3], logit[a]);
          }
        const float* sp = st + erow;
        const float wrow = sp[(3 * SA - ROW0 + srow) * E];  // CntS
        if (IS_PI) {
          const float mx = fmaxf(logit[0], logit[OUT - 1]);
          const float e0 = __expf(logit[0] - mx), e1 = __expf(logit[OUT - 1] - mx);
          const float sum = e0 + e1, lse = __logf(sum), inv = __fdividef(1.f, sum);
          const float p[2] = {e0 * inv, e1 * inv};
          const float lp[2] = {(logit[0] - mx) - lse, (logit[OUT - 1] - mx) - lse};
          const float ent = -(p[0] * lp[0] + p[1] * lp[1]);
          float ga[2], sumg = 0.f, lrow = 0.f;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int c = srow * A + a;
            const float Ap = sp[c * E], Am = sp[(SA + c) * E];
            const float ratio = __expf(lp[a] - sp[(2 * SA + c) * E]);
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
          const float diff = logit[0] - sp[(3 * SA + S - ROW0 + srow) * E];
          if (valid) {
            loss_acc = fmaf(wrow * diff, diff, loss_acc);
            gl[0] = invM * 2.f * wrow * diff;
          }
        }
#pragma unroll
        for (int a = 0; a < OUT; ++a) {
          g3w[a * 32 + lane] = gl[a];
          s_b3[a] += gl[a];
        }
#pragma unroll
        for (int j4 = 0; j4 < H / 4; ++j4) {
          float up[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int a = 0; a < OUT; ++a) {
            const float4 w3 = *reinterpret_cast<const float4*>(w + L.oW3 + a * H + 4 * j4);
            up[0] = fmaf(gl[a], w3.x, up[0]);
            up[1] = fmaf(gl[a], w3.y, up[1]);
            up[2] = fmaf(gl[a], w3.z, up[2]);
            up[3] = fmaf(gl[a], w3.w, up[3]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a2 = a2v[4 * j4 + q];
            D2w[(4 * j4 + q) * LDT + lane] = up[q] * fmaf(-a2, a2, 1.f);
          }
        }
      }
      __syncwarp();
      // ---- B: d1 = (W2^T d2)(1 - a1^2); db1, dW1 row sums straight from the C fragments ----
      {
        float c[MT][4][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) { c[mt][nt][0] = 0.f; c[mt][nt][1] = 0.f; c[mt][nt][2] = 0.f; c[mt][nt][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            split_tf32(D2w[(8 * ks + tq) * LDT + 8 * nt + gq], bh[nt][0], bl[nt][0]);
            split_tf32(D2w[(8 * ks + tq + 4) * LDT + 8 * nt + gq], bh[nt][1], bl[nt][1]);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const float4 fh = *reinterpret_cast<const float4*>(fragB + ((ks * MT + mt) * 32 + lane) * 8);
            const float4 fl = *reinterpret_cast<const float4*>(fragB + ((ks * MT + mt) * 32 + lane) * 8 + 4);
            const uint32_t ahi[4] = {__float_as_uint(fh.x), __float_as_uint(fh.y), __float_as_uint(fh.z), __float_as_uint(fh.w)};
            const uint32_t alo[4] = {__float_as_uint(fl.x), __float_as_uint(fl.y), __float_as_uint(fl.z), __float_as_uint(fl.w)};
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma3(c[mt][nt], ahi, alo, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
          }
        }
        // rows of a block share one state whenever E is a multiple of 32 (warp-uniform test): then dW1[:, s] gets
        // the same row sum as db1 and the one-hot products are skipped
        const int m_first = blk * 32, m_last = min(blk * 32 + 31, M - 1);
        const int s_first = (m_first >= E) + (m_first >= 2 * E), s_last = (m_last >= E) + (m_last >= 2 * E);
        const bool uniform = s_first == s_last;
        float t_b1[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { t_b1[mt][0] = 0.f; t_b1[mt][1] = 0.f; }
        if (uniform) {   // warp-uniform branch: all rows of the block are in state s_first
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float2 lm = *reinterpret_cast<const float2*>(rlw + 8 * nt + 2 * tq);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const float2 a1 = *reinterpret_cast<const float2*>(A1w + (16 * mt + gq + 8 * hh) * LDT + 8 * nt + 2 * tq);
                const float dx = c[mt][nt][2 * hh] * fmaf(-a1.x, a1.x, 1.f), dy = c[mt][nt][2 * hh + 1] * fmaf(-a1.y, a1.y, 1.f);
                t_b1[mt][hh] += dx + dy;
                s_wl[mt][hh] += fmaf(dx, lm.x, dy * lm.y);
              }
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              s_b1[mt][hh] += t_b1[mt][hh];
#pragma unroll
              for (int cc = 0; cc < S; ++cc) s_w1[cc][mt][hh] += (s_first == cc) ? t_b1[mt][hh] : 0.f;
            }
        } else {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float2 lm = *reinterpret_cast<const float2*>(rlw + 8 * nt + 2 * tq);
            float2 oh[S];
#pragma unroll
            for (int cc = 0; cc < S; ++cc) oh[cc] = *reinterpret_cast<const float2*>(row + cc * 32 + 8 * nt + 2 * tq);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const float2 a1 = *reinterpret_cast<const float2*>(A1w + (16 * mt + gq + 8 * hh) * LDT + 8 * nt + 2 * tq);
                const float dx = c[mt][nt][2 * hh] * fmaf(-a1.x, a1.x, 1.f), dy = c[mt][nt][2 * hh + 1] * fmaf(-a1.y, a1.y, 1.f);
                s_b1[mt][hh] += dx + dy;
                s_wl[mt][hh] += fmaf(dx, lm.x, dy * lm.y);
#pragma unroll
                for (int cc = 0; cc < S; ++cc) s_w1[cc][mt][hh] += fmaf(dx, oh[cc].x, dy * oh[cc].y);
              }
          }
        }
      }
      // ---- W: dW2 += d2 a1^T over the block's 32 rows (4 k-steps); db2 and dW3 ride along ----
#pragma unroll
      for (int kr = 0; kr < 4; ++kr) {
        uint32_t bh[NT8][2], bl[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
          split_tf32(A1w[(8 * nt + gq) * LDT + 8 * kr + tq], bh[nt][0], bl[nt][0]);
          split_tf32(A1w[(8 * nt + gq) * LDT + 8 * kr + tq + 4], bh[nt][1], bl[nt][1]);
        }
        float gv[OUT][2];
#pragma unroll
        for (int a = 0; a < OUT; ++a) { gv[a][0] = g3w[a * 32 + 8 * kr + tq]; gv[a][1] = g3w[a * 32 + 8 * kr + tq + 4]; }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const float d00 = D2w[(16 * mt + gq) * LDT + 8 * kr + tq], d10 = D2w[(16 * mt + gq + 8) * LDT + 8 * kr + tq];
          const float d01 = D2w[(16 * mt + gq) * LDT + 8 * kr + tq + 4], d11 = D2w[(16 * mt + gq + 8) * LDT + 8 * kr + tq + 4];
          s_b2[mt][0] += d00 + d01;
          s_b2[mt][1] += d10 + d11;
          const float x00 = A2w[(16 * mt + gq) * LDT + 8 * kr + tq], x10 = A2w[(16 * mt + gq + 8) * LDT + 8 * kr + tq];
          const float x01 = A2w[(16 * mt + gq) * LDT + 8 * kr + tq + 4], x11 = A2w[(16 * mt + gq + 8) * LDT + 8 * kr + tq + 4];
#pragma unroll
          for (int a = 0; a < OUT; ++a) {
            s_w3[a][mt][0] += fmaf(gv[a][0], x00, gv[a][1] * x01);
            s_w3[a][mt][1] += fmaf(gv[a][0], x10, gv[a][1] * x11);
          }
          uint32_t ahi[4], alo[4];
          split_tf32(d00, ahi[0], alo[0]);
          split_tf32(d10, ahi[1], alo[1]);
          split_tf32(d01, ahi[2], alo[2]);
          split_tf32(d11, ahi[3], alo[3]);
#pragma unroll
          for (int nt = 0; nt < NT8; ++nt) mma3(accW2[mt][nt], ahi, alo, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
        }
      }
      __syncwarp();  // the next block's P1 overwrites the tiles
    }

    // ---- reduce in a fixed order into G ----
    __syncthreads();  // every warp is done with its tiles; reuse them as scratch
    float* pW2 = tiles;                       // [8 warps][H*H]
    float* pRS = tiles + 8 * H * H;           // [8 warps][H][2 + S + OUT + 1]  (b1, wl, w1[S], w3[OUT], b2)
    constexpr int NRS = 2 + S + OUT + 1;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        float* o = pW2 + warp * (H * H);
        o[(16 * mt + gq) * H + 8 * nt + 2 * tq] = accW2[mt][nt][0];
        o[(16 * mt + gq) * H + 8 * nt + 2 * tq + 1] = accW2[mt][nt][1];
        o[(16 * mt + gq + 8) * H + 8 * nt + 2 * tq] = accW2[mt][nt][2];
        o[(16 * mt + gq + 8) * H + 8 * nt + 2 * tq + 1] = accW2[mt][nt][3];
      }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float vals[NRS];
        vals[0] = s_b1[mt][hh];
        vals[1] = s_wl[mt][hh];
#pragma unroll
        for (int c = 0; c < S; ++c) vals[2 + c] = s_w1[c][mt][hh];
#pragma unroll
        for (int a = 0; a < OUT; ++a) vals[2 + S + a] = s_w3[a][mt][hh];
        vals[2 + S + OUT] = s_b2[mt][hh];
#pragma unroll
        for (int v = 0; v < NRS; ++v) {
          float x = vals[v];
          x += __shfl_xor_sync(0xffffffffu, x, 1);
          x += __shfl_xor_sync(0xffffffffu, x, 2);
          if (tq == 0) pRS[(warp * H + 16 * mt + gq + 8 * hh) * NRS + v] = x;
        }
      }
    float b3tot[OUT];
#pragma unroll
    for (int a = 0; a < OUT; ++a) b3tot[a] = warp_sum(s_b3[a]);
    const float loss_tot = g.loss_out ? block_sum(loss_acc, red) : 0.f;  // block-wide; skipped when nobody reads it
    if (lane == 0)
#pragma unroll
      for (int a = 0; a < OUT; ++a) red[8 + warp * OUT + a] = b3tot[a];
    __syncthreads();
    for (int i = tid; i < H * H; i += 256) {
      float v = pW2[i];
#pragma unroll
      for (int wq = 1; wq < 8; ++wq) v += pW2[wq * (H * H) + i];
      G[L.oW2 + i] = v;
    }
    for (int i = tid; i < H * NRS; i += 256) {
      const int k = i / NRS, v = i - k * NRS;
      float x = pRS[i];
#pragma unroll
      for (int wq = 1; wq < 8; ++wq) x += pRS[wq * H * NRS + i];
      if (v == 0) G[L.ob1 + k] = x;
      else if (v == 1) G[L.oW1 + k * IN + IN - 1] = x;
      else if (v < 2 + S) G[L.oW1 + k * IN + (v - 2)] = x;
      else if (v < 2 + S + OUT) G[L.oW3 + (v - 2 - S) * H + k] = x;
      else G[L.ob2 + k] = x;
    }
    if (tid < OUT) {
      float x = 0.f;
      for (int wq = 0; wq < 8; ++wq) x += red[8 + wq * OUT + tid];
      G[L.ob3 + tid] = x;
    }
    if (g.loss_out && tid == 0) {
      float extra = 0.f;
      if (!IS_PI && g.arm_stats) extra = g.arm_stats[4 * (int64_t)n + 2];
      g.loss_out[(int64_t)n * iters + it] = (loss_tot + extra) * invM;
    }
    __syncthreads();

    // ---- Adam (torch.optim.Adam defaults; _single_tensor_adam form) ----
    b1p *= g.hp.beta1;
    b2p *= g.hp.beta2;
    if (tid == 0) {  // the double-precision bias corrections once per iteration, not once per thread
      red[30] = (float)(lr / (1.0 - b1p));
      red[31] = (float)sqrt(1.0 - b2p);
    }
    __syncthreads();
    const float step_size = red[30], bc2s = red[31];
    for (int i = tid; i < P; i += 256) {
      const float gi = G[i];
      const float mi = am[i] + lerp_w * (gi - am[i]);
      const float vi = av[i] * beta2f + (omb2 * gi) * gi;
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
static int launch_mma(const MmaArgs& args, cudaStream_t st) {
  MmaArgs g = args;
  constexpr int OUT = IS_PI ? 2 : 1, SA = S * 2, MT = H / 16, KS = H / 8;
  const MlpLayout L(S + 1, H, OUT);
  const int E = g.d.n_envs;
  size_t fl = 4 * (size_t)L.padded() + 2 * (size_t)(KS * MT * 32 * 8) + S * H + H +
              8 * (size_t)(3 * H * kLDT + (OUT + 1 + S) * 32) + (((size_t)E + 3) & ~(size_t)3);
  const size_t rows = IS_PI ? 3 * SA + S : 2 * S;
  g.stage_stats = 0;
  const size_t cap = (H >= 32 ? 220 : 110) * 1024;
  if ((fl + rows * E) * sizeof(float) <= cap) {
    fl += rows * E;
    g.stage_stats = 1;
  }
  const size_t smem = fl * sizeof(float);
  if (smem > 227 * 1024) return fail(RMAB_E_SHAPE, "rmab_ppo_update_table: E=%d needs %zu B of shared memory", E, smem);
  auto kern = ppo_table_mma_kernel<H, S, IS_PI>;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(RMAB_E_LAUNCH, "rmab_ppo_update_table: cannot reserve %zu B of shared memory", smem);
  kern<<<g.d.n_arms, 256, smem, st>>>(g);
  return check_launch(IS_PI ? "rmab_ppo_update_table(pi, mma)" : "rmab_ppo_update_table(v, mma)");
}

// Entry used by rmab_ppo_update_table (ppo_table.cu) for H = 16 / 32.
int ppo_table_mma_run(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv, float* adam_pi,
                      float* adam_v, const float* stats, const float* arm_stats, const float* lamb, float* loss_pi,
                      float* loss_v, cudaStream_t st) {
  MmaArgs g = {};
  g.d = *net; g.hp = *hp; g.T = T; g.lamb = lamb; g.stats = stats; g.arm_stats = arm_stats;
  MmaArgs gp = g, gv = g;
  gp.W = Wpi; gp.adam = adam_pi; gp.loss_out = loss_pi;
  gv.W = Wv; gv.adam = adam_v; gv.loss_out = loss_v;
  int rc = RMAB_OK;
#define RUN(H, S)                                                                   \
  {                                                                                 \
    if (hp->pi_iters > 0) rc = launch_mma<H, S, true>(gp, st);                      \
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_mma<H, S, false>(gv, st);     \
    return rc;                                                                      \
  }
  if (net->hidden == 16 && net->n_states == 2) RUN(16, 2)
  if (net->hidden == 16 && net->n_states == 3) RUN(16, 3)
  if (net->hidden == 32 && net->n_states == 2) RUN(32, 2)
  if (net->hidden == 32 && net->n_states == 3) RUN(32, 3)
#undef RUN
  return fail(RMAB_E_SHAPE, "ppo_table_mma_run: unsupported shape");
}

}  // namespace rmab
