// Statistics-based PPO update, tensor-core variant for hidden width 64 (the ARMMAN N = 100k x 1024 envs shape).
// Same algorithm and fragment conventions as ppo_table_mma.cu (3xTF32 mma.sync m16n8k8, warp-local tiles); what
// changes is the shared-memory plan, because a 64 x 64 layer does not fit the 8-warp / 32-row tiles in 227 KB:
//   * blocks of 16 rows (two n-tiles), tile row length 24 (24t + g hits 32 distinct banks), 8 warps per CTA;
//   * two lanes per row in the lane-per-row phases: lane half hf owns the features {4i + 2hf, 4i + 2hf + 1};
//   * W2 lives ONLY in its two pre-split fragment sets (hi + lo is the exact fp32 value): the Adam step of W2 reads
//     the summed gradient straight from the reduction scratch and writes the new hi / lo back into both sets;
//   * Adam moments stay in global memory (read-modify-write once per iteration, L2 resident); only the non-W2
//     weights and their gradient are staged in shared memory; the statistics rows are read from global (L2).
// Per 16-row block a warp issues 3 x (8 x 4 x 2 + 8 x 4 x 2 + 2 x 4 x 8) x 3 = 576 mma, ~85 % of its instructions'
// tensor time; FFMA equivalent would be 3 x 64 x 64 x 16 / 32 = 6144 FFMA instructions.
//
// Reference behaviour: agent_oracle.py compute_loss_pi :285-322, compute_loss_v :325-339, update :399-436.
#include "common.cuh"
#include "mlp.cuh"

namespace rmab {

namespace {

constexpr int H = 64, MT = 4, KS = 8, NT8 = 8, RB = 16, NTL = 2, LDT = 24, NW = 8, NTHR = NW * 32, HH = H * H;

// Position of W2[r][c] inside a pre-split fragment set [KS][MT][32 lanes][8] (hi a0..a3, lo a0..a3) whose A operand is
// A[m = r][k = c] (fragment layout of mma.m16n8k8: a_q at row g + 8 (q & 1), column t + 4 (q >> 1)).
__host__ __device__ __forceinline__ int frag_pos(int r, int c) {
  const int mt = r >> 4, rr = r & 15, ks = c >> 3, cc = c & 7;
  return ((ks * MT + mt) * 32 + (rr & 7) * 4 + (cc & 3)) * 8 + (rr >> 3) + 2 * (cc >> 2);
}

__device__ __forceinline__ float tanh_fast_6(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

__device__ __forceinline__ void split6(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma6(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma6x3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t bhi0,
                                       uint32_t bhi1, uint32_t blo0, uint32_t blo1) {
  mma6(c, alo, bhi0, bhi1);
  mma6(c, ahi, blo0, blo1);
  mma6(c, ahi, bhi0, bhi1);
}

struct Mma64Args {
  RmabNetDesc d;
  RmabHyper hp;
  int T;
  float* W; float* adam;
  const float* lamb; const float* stats; const float* arm_stats;
  float* loss_out;
};

}  // namespace

template <int S, bool IS_PI>
__global__ void __launch_bounds__(NTHR, 1) ppo_table_mma64_kernel(Mma64Args g) {
  constexpr int A = 2, OUT = IS_PI ? A : 1, IN = S + 1, SA = S * A;
  constexpr int ROW0 = IS_PI ? 0 : 3 * SA, KST = 4 * SA + 3 * S;
  constexpr int FRAG = KS * MT * 32 * 8;
  constexpr int WTILE = 3 * H * LDT + (OUT + 1 + S) * RB;
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[32];
  const int E = g.d.n_envs, T = g.T;
  const MlpLayout L(IN, H, OUT);
  const int P = L.P, Pp = L.padded();
  // w / G hold the packed layout WITHOUT the W2 block: index i < oW2 as is, i >= ob2 shifted down by H*H
  const int Ps = Pp - HH;
  float* w = smem;
  float* G = w + Ps;
  float* fragF = G + Ps;
  float* fragB = fragF + FRAG;
  float* wb = fragB + FRAG;          // [S][H]
  float* wl = wb + S * H;            // [H]
  float* tiles = wl + H;             // NW x WTILE
  float* slam = tiles + NW * WTILE;  // [E]
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int rr = lane & 15, hf = lane >> 4;   // lane-per-row phases: row of the block, feature half
  float* A1w = tiles + warp * WTILE;
  float* A2w = A1w + H * LDT;
  float* D2w = A2w + H * LDT;
  float* g3w = D2w + H * LDT;        // [OUT][RB]
  float* rlw = g3w + OUT * RB;       // [RB]
  float* row = rlw + RB;             // [S][RB]

  float* Wn = g.W + (int64_t)n * P;
  float* An = g.adam + (int64_t)n * 2 * P;
  for (int i = tid; i < P; i += NTHR) {
    const float x = Wn[i];
    if (i < L.oW2) w[i] = x;
    else if (i >= L.ob2) w[i - HH] = x;
    else {
      const int r = (i - L.oW2) / H, c = (i - L.oW2) - r * H;
      uint32_t hi, lo;
      split6(x, hi, lo);
      fragF[frag_pos(r, c)] = __uint_as_float(hi);
      fragF[frag_pos(r, c) + 4] = __uint_as_float(lo);
      fragB[frag_pos(c, r)] = __uint_as_float(hi);    // backward operand A[m = c][k = r] = W2[r][c]
      fragB[frag_pos(c, r) + 4] = __uint_as_float(lo);
    }
  }
  for (int i = tid; i < E; i += NTHR) slam[i] = g.lamb[i];
  const float* st = g.stats + ((int64_t)n * KST + ROW0) * E;
  __syncthreads();

  const int M = S * E;
  const int n_blocks = (M + RB - 1) / RB;
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
    for (int i = tid; i < S * H; i += NTHR) {
      const int s = i / H, k = i - s * H;
      wb[i] = w[L.oW1 + k * IN + s] + w[L.ob1 + k];
    }
    for (int i = tid; i < H; i += NTHR) wl[i] = w[L.oW1 + i * IN + IN - 1];
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

    for (int blk = warp; blk < n_blocks; blk += NW) {
      // ---- P1: two lanes per row, lane half hf owns features {4i + 2hf, 4i + 2hf + 1} ----
      const int m = blk * RB + rr;
      const bool valid = m < M;
      const int mm = valid ? m : M - 1;
      const int srow = (mm >= E) + (mm >= 2 * E), erow = mm - srow * E;  // S <= 3: no integer division
      {
        const float lam = slam[erow];
        if (hf == 0) {
          rlw[rr] = valid ? lam : 0.f;
#pragma unroll
          for (int c = 0; c < S; ++c) row[c * RB + rr] = (valid && srow == c) ? 1.f : 0.f;
        }
        const float* wbs = wb + srow * H;
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          const float4 b = *reinterpret_cast<const float4*>(wbs + 4 * k4);
          const float4 l = *reinterpret_cast<const float4*>(wl + 4 * k4);
          const float b0 = hf ? b.z : b.x, b1 = hf ? b.w : b.y, l0 = hf ? l.z : l.x, l1 = hf ? l.w : l.y;
          A1w[(4 * k4 + 2 * hf) * LDT + rr] = tanh_fast_6(fmaf(l0, lam, b0));
          A1w[(4 * k4 + 2 * hf + 1) * LDT + rr] = tanh_fast_6(fmaf(l1, lam, b1));
        }
      }
      __syncwarp();
      // ---- F: a2 = tanh(W2 a1 + b2) ----
      {
        float c[MT][NTL][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const float blo = w[L.ob2 - HH + 16 * mt + gq], bhi = w[L.ob2 - HH + 16 * mt + gq + 8];
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) { c[mt][nt][0] = blo; c[mt][nt][1] = blo; c[mt][nt][2] = bhi; c[mt][nt][3] = bhi; }
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t bh[NTL][2], bl[NTL][2];
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            split6(A1w[(8 * ks + tq) * LDT + 8 * nt + gq], bh[nt][0], bl[nt][0]);
            split6(A1w[(8 * ks + tq + 4) * LDT + 8 * nt + gq], bh[nt][1], bl[nt][1]);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const float4 fh = *reinterpret_cast<const float4*>(fragF + ((ks * MT + mt) * 32 + lane) * 8);
            const float4 fl = *reinterpret_cast<const float4*>(fragF + ((ks * MT + mt) * 32 + lane) * 8 + 4);
            const uint32_t ahi[4] = {__float_as_uint(fh.x), __float_as_uint(fh.y), __float_as_uint(fh.z), __float_as_uint(fh.w)};
            const uint32_t alo[4] = {__float_as_uint(fl.x), __float_as_uint(fl.y), __float_as_uint(fl.z), __float_as_uint(fl.w)};
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) mma6x3(c[mt][nt], ahi, alo, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            float2 lo2, hi2;
            lo2.x = tanh_fast_6(c[mt][nt][0]); lo2.y = tanh_fast_6(c[mt][nt][1]);
            hi2.x = tanh_fast_6(c[mt][nt][2]); hi2.y = tanh_fast_6(c[mt][nt][3]);
            *reinterpret_cast<float2*>(A2w + (16 * mt + gq) * LDT + 8 * nt + 2 * tq) = lo2;
            *reinterpret_cast<float2*>(A2w + (16 * mt + gq + 8) * LDT + 8 * nt + 2 * tq) = hi2;
          }
      }
      __syncwarp();
      // ---- P3: two lanes per row; each sums its 32 features, the halves meet through one shuffle ----
      {
        float logit[OUT], gl[OUT];
#pragma unroll
        for (int a = 0; a < OUT; ++a) { logit[a] = 0.f; gl[a] = 0.f; }
        float a2v[H / 2];
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          a2v[2 * k4] = A2w[(4 * k4 + 2 * hf) * LDT + rr];
          a2v[2 * k4 + 1] = A2w[(4 * k4 + 2 * hf + 1) * LDT + rr];
        }
#pragma unroll
        for (int a = 0; a < OUT; ++a) {
#pragma unroll
          for (int k4 = 0; k4 < H / 4; ++k4) {
            const float4 w3 = *reinterpret_cast<const float4*>(w + L.oW3 - HH + a * H + 4 * k4);
            logit[a] = fmaf(hf ? w3.z : w3.x, a2v[2 * k4], logit[a]);
            logit[a] = fmaf(hf ? w3.w : w3.y, a2v[2 * k4 + 1], logit[a]);
          }
          logit[a] += __shfl_xor_sync(0xffffffffu, logit[a], 16);
          logit[a] += w[L.ob3 - HH + a];
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
            if (hf == 0) loss_acc += lrow - beta_ent * wrow * ent;
#pragma unroll
            for (int a = 0; a < 2; ++a)
              gl[a < OUT ? a : 0] = invM * (ga[a] - p[a] * sumg + beta_ent * wrow * p[a] * (lp[a] + ent));
          }
        } else {
          const float diff = logit[0] - sp[(3 * SA + S - ROW0 + srow) * E];
          if (valid) {
            if (hf == 0) loss_acc = fmaf(wrow * diff, diff, loss_acc);
            gl[0] = invM * 2.f * wrow * diff;
          }
        }
        if (hf == 0) {
#pragma unroll
          for (int a = 0; a < OUT; ++a) {
            g3w[a * RB + rr] = gl[a];
            s_b3[a] += gl[a];
          }
        }
#pragma unroll
        for (int k4 = 0; k4 < H / 4; ++k4) {
          float up0 = 0.f, up1 = 0.f;
#pragma unroll
          for (int a = 0; a < OUT; ++a) {
            const float4 w3 = *reinterpret_cast<const float4*>(w + L.oW3 - HH + a * H + 4 * k4);
            up0 = fmaf(gl[a], hf ? w3.z : w3.x, up0);
            up1 = fmaf(gl[a], hf ? w3.w : w3.y, up1);
          }
          const float x0 = a2v[2 * k4], x1 = a2v[2 * k4 + 1];
          D2w[(4 * k4 + 2 * hf) * LDT + rr] = up0 * fmaf(-x0, x0, 1.f);
          D2w[(4 * k4 + 2 * hf + 1) * LDT + rr] = up1 * fmaf(-x1, x1, 1.f);
        }
      }
      __syncwarp();
      // ---- B: d1 = (W2^T d2)(1 - a1^2); db1, dW1 row sums from the C fragments ----
      {
        float c[MT][NTL][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) { c[mt][nt][0] = 0.f; c[mt][nt][1] = 0.f; c[mt][nt][2] = 0.f; c[mt][nt][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t bh[NTL][2], bl[NTL][2];
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
            split6(D2w[(8 * ks + tq) * LDT + 8 * nt + gq], bh[nt][0], bl[nt][0]);
            split6(D2w[(8 * ks + tq + 4) * LDT + 8 * nt + gq], bh[nt][1], bl[nt][1]);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const float4 fh = *reinterpret_cast<const float4*>(fragB + ((ks * MT + mt) * 32 + lane) * 8);
            const float4 fl = *reinterpret_cast<const float4*>(fragB + ((ks * MT + mt) * 32 + lane) * 8 + 4);
            const uint32_t ahi[4] = {__float_as_uint(fh.x), __float_as_uint(fh.y), __float_as_uint(fh.z), __float_as_uint(fh.w)};
            const uint32_t alo[4] = {__float_as_uint(fl.x), __float_as_uint(fl.y), __float_as_uint(fl.z), __float_as_uint(fl.w)};
#pragma unroll
            for (int nt = 0; nt < NTL; ++nt) mma6x3(c[mt][nt], ahi, alo, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
          }
        }
        const int m_first = blk * RB, m_last = min(blk * RB + RB - 1, M - 1);
        const int s_first = (m_first >= E) + (m_first >= 2 * E), s_last = (m_last >= E) + (m_last >= 2 * E);
        const bool uniform = s_first == s_last;
        float t_b1[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) { t_b1[mt][0] = 0.f; t_b1[mt][1] = 0.f; }
        if (uniform) {   // warp-uniform branch: all rows of the block are in state s_first
#pragma unroll
          for (int nt = 0; nt < NTL; ++nt) {
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
          for (int nt = 0; nt < NTL; ++nt) {
            const float2 lm = *reinterpret_cast<const float2*>(rlw + 8 * nt + 2 * tq);
            float2 oh[S];
#pragma unroll
            for (int cc = 0; cc < S; ++cc) oh[cc] = *reinterpret_cast<const float2*>(row + cc * RB + 8 * nt + 2 * tq);
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
      // ---- W: dW2 += d2 a1^T over the block's 16 rows (2 k-steps); db2 and dW3 ride along ----
#pragma unroll
      for (int kr = 0; kr < RB / 8; ++kr) {
        uint32_t bh[NT8][2], bl[NT8][2];
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
          split6(A1w[(8 * nt + gq) * LDT + 8 * kr + tq], bh[nt][0], bl[nt][0]);
          split6(A1w[(8 * nt + gq) * LDT + 8 * kr + tq + 4], bh[nt][1], bl[nt][1]);
        }
        float gv[OUT][2];
#pragma unroll
        for (int a = 0; a < OUT; ++a) { gv[a][0] = g3w[a * RB + 8 * kr + tq]; gv[a][1] = g3w[a * RB + 8 * kr + tq + 4]; }
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
          split6(d00, ahi[0], alo[0]);
          split6(d10, ahi[1], alo[1]);
          split6(d01, ahi[2], alo[2]);
          split6(d11, ahi[3], alo[3]);
#pragma unroll
          for (int nt = 0; nt < NT8; ++nt) mma6x3(accW2[mt][nt], ahi, alo, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
        }
      }
      __syncwarp();
    }

    // ---- reduce in a fixed order into G ----
    __syncthreads();
    float* pW2 = tiles;                       // [NW][H*H]
    float* pRS = tiles + NW * H * H;          // [NW][H][NRS]
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
    for (int i = tid; i < H * NRS; i += NTHR) {
      const int k = i / NRS, v = i - k * NRS;
      float x = pRS[i];
#pragma unroll
      for (int wq = 1; wq < NW; ++wq) x += pRS[wq * H * NRS + i];
      if (v == 0) G[L.ob1 + k] = x;
      else if (v == 1) G[L.oW1 + k * IN + IN - 1] = x;
      else if (v < 2 + S) G[L.oW1 + k * IN + (v - 2)] = x;
      else if (v < 2 + S + OUT) G[L.oW3 - HH + (v - 2 - S) * H + k] = x;
      else G[L.ob2 - HH + k] = x;
    }
    if (tid < OUT) {
      float x = 0.f;
      for (int wq = 0; wq < NW; ++wq) x += red[8 + wq * OUT + tid];
      G[L.ob3 - HH + tid] = x;
    }
    if (g.loss_out && tid == 0) {
      float extra = 0.f;
      if (!IS_PI && g.arm_stats) extra = g.arm_stats[4 * (int64_t)n + 2];
      g.loss_out[(int64_t)n * iters + it] = (loss_tot + extra) * invM;
    }
    __syncthreads();

    // ---- Adam; moments live in global memory (one read-modify-write per iteration) ----
    b1p *= g.hp.beta1;
    b2p *= g.hp.beta2;
    if (tid == 0) {  // the double-precision bias corrections once per iteration, not once per thread
      red[30] = (float)(lr / (1.0 - b1p));
      red[31] = (float)sqrt(1.0 - b2p);
    }
    __syncthreads();
    const float step_size = red[30], bc2s = red[31];
    for (int i = tid; i < P; i += NTHR) {
      const bool in_w2 = i >= L.oW2 && i < L.ob2;
      float gi, x;
      int r = 0, c = 0;
      if (in_w2) {
        const int j = i - L.oW2;
        r = j / H; c = j - r * H;
        gi = pW2[j];
#pragma unroll
        for (int wq = 1; wq < NW; ++wq) gi += pW2[wq * HH + j];     // fixed warp order
        x = fragF[frag_pos(r, c)] + fragF[frag_pos(r, c) + 4];      // hi + lo is the exact fp32 weight
      } else {
        const int k = i < L.oW2 ? i : i - HH;
        gi = G[k];
        x = w[k];
      }
      const float m0 = An[i], v0 = An[P + i];
      const float mi = m0 + lerp_w * (gi - m0);
      const float vi = v0 * beta2f + (omb2 * gi) * gi;
      An[i] = mi;
      An[P + i] = vi;
      const float denom = sqrtf(vi) / bc2s + epsf;
      const float xn = x - step_size * (mi / denom);
      if (in_w2) {
        uint32_t hi, lo;
        split6(xn, hi, lo);
        fragF[frag_pos(r, c)] = __uint_as_float(hi);
        fragF[frag_pos(r, c) + 4] = __uint_as_float(lo);
        fragB[frag_pos(c, r)] = __uint_as_float(hi);
        fragB[frag_pos(c, r) + 4] = __uint_as_float(lo);
      } else {
        w[i < L.oW2 ? i : i - HH] = xn;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < P; i += NTHR) {
    if (i < L.oW2) Wn[i] = w[i];
    else if (i >= L.ob2) Wn[i] = w[i - HH];
    else {
      const int r = (i - L.oW2) / H, c = (i - L.oW2) - r * H;
      Wn[i] = fragF[frag_pos(r, c)] + fragF[frag_pos(r, c) + 4];
    }
  }
}

template <int S, bool IS_PI>
static int launch_mma64(const Mma64Args& g, cudaStream_t st) {
  constexpr int OUT = IS_PI ? 2 : 1;
  const MlpLayout L(S + 1, H, OUT);
  const int E = g.d.n_envs;
  const size_t fl = 2 * (size_t)(L.padded() - HH) + 2 * (size_t)(KS * MT * 32 * 8) + S * H + H +
                    NW * (size_t)(3 * H * LDT + (OUT + 1 + S) * RB) + (((size_t)E + 3) & ~(size_t)3);
  const size_t smem = fl * sizeof(float);
  if (smem > 227 * 1024) return fail(RMAB_E_SHAPE, "rmab_ppo_update_table: E=%d needs %zu B of shared memory", E, smem);
  auto kern = ppo_table_mma64_kernel<S, IS_PI>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(RMAB_E_LAUNCH, "rmab_ppo_update_table: cannot reserve %zu B of shared memory", smem);
  kern<<<g.d.n_arms, NTHR, smem, st>>>(g);
  return check_launch(IS_PI ? "rmab_ppo_update_table(pi, mma64)" : "rmab_ppo_update_table(v, mma64)");
}

// Entry used by rmab_ppo_update_table (ppo_table.cu) for H = 64.
int ppo_table_mma64_run(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv, float* adam_pi,
                        float* adam_v, const float* stats, const float* arm_stats, const float* lamb, float* loss_pi,
                        float* loss_v, cudaStream_t st) {
  Mma64Args g = {};
  g.d = *net; g.hp = *hp; g.T = T; g.lamb = lamb; g.stats = stats; g.arm_stats = arm_stats;
  Mma64Args gp = g, gv = g;
  gp.W = Wpi; gp.adam = adam_pi; gp.loss_out = loss_pi;
  gv.W = Wv; gv.adam = adam_v; gv.loss_out = loss_v;
  int rc = RMAB_OK;
  if (net->n_states == 2) {
    if (hp->pi_iters > 0) rc = launch_mma64<2, true>(gp, st);
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_mma64<2, false>(gv, st);
  } else {
    if (hp->pi_iters > 0) rc = launch_mma64<3, true>(gp, st);
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_mma64<3, false>(gv, st);
  }
  return rc;
}

}  // namespace rmab
