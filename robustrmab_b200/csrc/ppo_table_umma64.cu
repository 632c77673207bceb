// Statistics-based PPO update for hidden width 64 on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM) -- the ARMMAN N = 100k x 1024 envs shape, where S*E = 3072 rows x 64 x 64 per arm per Adam step is a real GEMM.
// Same algorithm as ppo_table_mma64.cu (loss statistics, 3xTF32 hi/lo split, fp32 accumulation); one CTA per arm.
//
// A CTA walks the arm's rows in tiles of 128 (TMEM lane = tile row).  16 warps: warp (q, cq) owns rows 32q..32q+31 (its
// TMEM lane quarter) and the 16 hidden features 16cq..16cq+15.  Per tile, three GEMMs, each issued by one elected lane
// (G2 by warp 0 after the tile's barrier, G1 of the next tile by warp 1 once G2 has completed, G3 in two halves by warps 2, 3):
//   G1  z2   [128x64] = a1 . W2^T    A = a1 hi/lo in TMEM (written with tcgen05.st), B = W2   K-major in smem
//   G2  d1'  [128x64] = d2 . W2      A = d2 hi/lo in TMEM,                           B = W2^T K-major in smem
//   G3  dW2  [ 64x64] += d2^T . a1   A = d2, B = a1 tiles in smem, both MN-major (128-byte rows, 32-byte swizzle atoms:
//                                    the one MN-major layout tf32 operands have); accumulates over ALL tiles in TMEM
// each as three tf32 products (lo.hi + hi.lo + hi.hi).  tools/umma_probe.cu pins every descriptor encoding used here
// against a CPU product on the device.  Everything else (layer 1, tanh, softmax / clip terms, the row sums that make
// db1 / dW1 / db2 / dW3) runs on the CUDA cores between the GEMMs; a1/d2 tiles are single-buffered (192 of 227 KB).
// Schedule: ONE block barrier per tile.  E2(t) computes d2 (FMA pipe) interleaved with layer 1 of tile t+1 (MUFU pipe), then
//   | barrier, issue G2(t), G1(t+1), G3(t) |  E3(t) (row sums of d1, needs G2)  ->  E2(t+1) (needs G1)  ->  ...
// so the tensor pipe works under E3 and E2; G3 is never waited for inside the pass.
// Row modes (MODE): statistics rows (rmab_ppo_update_table), or one row per sample with the scalar state input for
// rmab_ppo_update / rmab_critic_extra_iter (SIS, MA-RMABPPO agent nets) -- see kRowsTable / kRowsSample / kRowsSampleExtra.
//
// Reference behaviour: agent_oracle.py compute_loss_pi :285-322, compute_loss_v :325-339, update :399-436.
#include "common.cuh"
#include "mlp.cuh"
#include "umma.cuh"

namespace rmab {

namespace {

using namespace umma;

struct UmmaArgs {
  RmabNetDesc d;
  RmabHyper hp;
  int T;
  float* W; float* adam;
  const float* lamb; const float* stats; const float* arm_stats;
  float* loss_out;
  // per-sample rows (MODE >= 1): the buffers of rmab_ppo_update / rmab_critic_extra_iter
  int s_rows;
  const uint8_t* s_buf; const uint8_t* a_buf;
  const float* adv; const float* logp_old; const float* ret;
  const float* z1_extra; float* dz1_out;   // MODE 2, sample-major [T*E, N, h]
  int stage_lamb;                          // lambda [E] staged in shared memory (else read through L1 / L2)
  // float images of the hyper-parameters, rounded once on the host (the tile loop otherwise re-derives them from the doubles)
  float clip_lo, clip_hi, beta_ent, lerp_w, beta2f, omb2, epsf, invM;
};

// Row sources.  MODE 0: one row per (state, env) carrying the loss statistics of rmab_epoch_table (one-hot inputs, S <= 3).
// MODE 1: one row per sample (t, env) of the [N,T,E] buffers, scalar state input s / state_norm (the SIS domain).
// MODE 2: MODE 1 for the MA agent critics: the caller's dense-GEMM term is added to the first layer, dLoss/dz1 is written.
constexpr int kRowsTable = 0, kRowsSample = 1, kRowsSampleExtra = 2;
constexpr int kAdamTab64 = 128;

}  // namespace

// RAGGED = tiles may hold rows past the end (their a1 is forced to 0, they are excluded from the sums) and, for table rows,
// straddle two states.  The other instantiation (table rows: E % 128 == 0; sample rows: T * E % 128 == 0 -- every BASELINE
// shape) drops the row predication and takes the state of a tile from its first row.
template <int S, int A, bool IS_PI, int MODE, bool RAGGED>
__global__ void __launch_bounds__(NTHR, 1) ppo_table_umma64_kernel(const __grid_constant__ UmmaArgs g) {
  constexpr bool TABLE = MODE == kRowsTable, EXTRA = MODE == kRowsSampleExtra;
  static_assert(TABLE ? (A == 2 && (S == 2 || S == 3)) : (S == 1 && A >= 2 && A <= 3), "row mode / shape");
  constexpr int OUT = IS_PI ? A : 1, IN = S + 1, SA = S * A;   // sample rows: IN = 2 = (s / state_norm, lambda)
  constexpr int ROW0 = IS_PI ? 0 : 3 * SA, KST = 4 * SA + 3 * S;
  // per-feature row sums: db2, dW1[:, lambda], then table: dW1[:, s] x S (db1 is their sum) | sample: dW1[:, 0], db1; dW3 x OUT
  constexpr int NW1 = TABLE ? S : 2;
  // Two-action table policy: softmax, entropy and their gradients depend on the logit DIFFERENCE only and the two logit
  // gradients of a row sum to zero, so the output layer runs on W3[0] - W3[1] with ONE gradient per row (OUTL = 1).
  constexpr bool DIFF = TABLE && IS_PI;
  constexpr int OUTL = DIFF ? 1 : OUT;
  constexpr int NQ = 2 + NW1 + OUTL;
  extern __shared__ unsigned char smraw[];
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_s;
  __shared__ float red[64];
  __shared__ float2 adam_tab[kAdamTab64];   // (lr / (1 - beta1^t), sqrt(1 - beta2^t)) of every Adam step of this launch
  // 1024-aligned window address, pinned in one register (left alone the compiler re-derives it from SR_CgaCtaId per phase)
  uint32_t sbase = (smem_u32(smraw) + 1023u) & ~1023u;
  asm volatile("" : "+r"(sbase));
  unsigned char* const base = reinterpret_cast<unsigned char*>(__cvta_shared_to_generic(sbase));
  const int E = g.d.n_envs, T = g.T;
  const MlpLayout L(IN, H, OUT);
  const int P = L.P, Ps = L.padded() - HH;   // w / G hold the packed layout WITHOUT the W2 block
  float* w = reinterpret_cast<float*>(base + T_END);
  float* G = w + Ps;
  float* wb = G + Ps;                 // [S][H]  (W1[:, s] + b1) * kTanhScale        (sample rows: b1 * kTanhScale)
  float* wl = wb + S * H;             // [H]     W1[:, lambda] * kTanhScale
  float* b2s = wl + H;                // [H]     b2 * kTanhScale
  float* wa = b2s + H;                // [H]     W1[:, 0] * kTanhScale (sample rows)
  float* w3d = wa + H;                // [H] table policy: W3[0] - W3[1]
  float* plog = w3d + H;              // [4 cq][128 rows][OUTL] partial logits
  float* redw = plog + 4 * TM * OUTL;  // [NW][NQ][16]
  float* slam_s = redw + NW * NQ * FW;  // [E] when it fits
  const float* slam = g.stage_lamb ? slam_s : g.lamb;   // one generic pointer (typed LDS / LDG selects measured slower: 44.2 vs 43.9 ms)
  float* W2H = reinterpret_cast<float*>(base + T_W2H);
  float* W2L = reinterpret_cast<float*>(base + T_W2L);
  float* WTH = reinterpret_cast<float*>(base + T_WTH);
  float* WTL = reinterpret_cast<float*>(base + T_WTL);
  float* scratch = reinterpret_cast<float*>(base + T_A1H);   // dW2 [64][64] between the last GEMM of an iteration and Adam

  int tid = threadIdx.x;
#ifndef U64_NO_PIN_TID
  asm volatile("" : "+r"(tid));   // pinned (no SR_TID.X re-reads inside the tile loop)
#endif
  const int n = blockIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform (see the h16 kernel)
  const int q = warp & 3, cq = warp >> 2, r = 32 * q + lane, f0 = FW * cq;
  float* Wn = g.W + (int64_t)n * P;
  float* An = g.adam + (int64_t)n * 2 * P;
  for (int i = tid; i < P; i += NTHR) {
    const float x = Wn[i];
    if (i < L.oW2) w[i] = x;
    else if (i >= L.ob2) w[i - HH] = x;
    else {
      const int j = (i - L.oW2) >> 6, c = (i - L.oW2) & 63;
      uint32_t hi, lo;
      split_u(x, hi, lo);
      W2H[kmaj_off(j, c)] = __uint_as_float(hi); W2L[kmaj_off(j, c)] = __uint_as_float(lo);
      WTH[kmaj_off(c, j)] = __uint_as_float(hi); WTL[kmaj_off(c, j)] = __uint_as_float(lo);
    }
  }
  if (g.stage_lamb)
    for (int i = tid; i < E; i += NTHR) slam_s[i] = g.lamb[i];
  if (tid == 0) {
    for (int b = 0; b < 3; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&bars[b])), "r"(b == 2 ? 2 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tmem_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);   // this warp's lane quarter
  const uint32_t bar1 = smem_u32(&bars[0]), bar2 = smem_u32(&bars[1]), bar3 = smem_u32(&bars[2]);
  // this thread's two 32-byte chunks of its row in the swizzled MN-major tiles
  const uint32_t rowb = (uint32_t)(cq >> 1) * BLK + (uint32_t)(r >> 2) * 512 + (uint32_t)(r & 3) * 128;
  const uint32_t off_c0 = rowb + (uint32_t)(((2 * (cq & 1)) ^ (r & 3)) * 32), off_c1 = rowb + (uint32_t)(((2 * (cq & 1) + 1) ^ (r & 3)) * 32);

  const float* st = TABLE ? g.stats + ((int64_t)n * KST + ROW0) * E : nullptr;
  asm volatile("" : "+l"(st));   // opaque base: 32-bit offsets per load instead of a 64-bit re-derivation each
  const int M = TABLE ? S * E : T * E;
  // floor(2^32 / E): m / E with one correction step.  E = 1 (the reference's own shape) would overflow the 32-bit magic to
  // 0 and leave env indices >= 1 -- lambda was then read past its single entry; 2^32 - 1 gives m - 1 there, corrected to m.
  const uint32_t e_magic = E > 1 ? (uint32_t)(0x100000000ull / (uint32_t)E) : 0xFFFFFFFFu;
  const int64_t smp_base = (int64_t)n * T * E, st_base = (int64_t)n * g.s_rows * E;
  const int64_t z_stride = (int64_t)g.d.n_arms * H;                    // sample-major [T*E, N, h]
  const int s_max = g.d.n_states - 1;
  const float state_norm = g.d.state_norm;
  const int n_tiles = (M + TM - 1) / TM;
  const float invM = g.invM;
  const int iters = IS_PI ? g.hp.pi_iters : g.hp.v_iters;
  const double lr = IS_PI ? g.hp.pi_lr : g.hp.vf_lr;
  const int step0 = IS_PI ? g.hp.adam_step0_pi : g.hp.adam_step0_v;
  const float clip_lo = g.clip_lo, clip_hi = g.clip_hi, beta_ent = g.beta_ent;
  const float lerp_w = g.lerp_w, beta2f = g.beta2f, omb2 = g.omb2, epsf = g.epsf;
  double b1p = pow(g.hp.beta1, (double)step0), b2p = pow(g.hp.beta2, (double)step0);
  for (int k = tid; k < min(iters, kAdamTab64); k += NTHR)   // the double-precision bias corrections once, in parallel
    adam_tab[k] = make_float2((float)(lr / (1.0 - pow(g.hp.beta1, (double)(step0 + k + 1)))),
                              (float)sqrt(1.0 - pow(g.hp.beta2, (double)(step0 + k + 1))));
  uint32_t gt = 0;   // tiles issued so far: every barrier completes one phase per tile

  for (int it = 0; it < iters; ++it) {
    for (int i = tid; i < S * H; i += NTHR) {
      const int s = i >> 6, k = i & 63;
      wb[i] = kTanhScale * (TABLE ? w[L.oW1 + k * IN + s] + w[L.ob1 + k] : w[L.ob1 + k]);
    }
    for (int i = tid; i < H; i += NTHR) {
      wl[i] = kTanhScale * w[L.oW1 + i * IN + IN - 1];
      b2s[i] = kTanhScale * w[L.ob2 - HH + i];
      wa[i] = kTanhScale * w[L.oW1 + i * IN];
      if (DIFF) w3d[i] = w[L.oW3 - HH + i] - w[L.oW3 - HH + H + i];
    }
    __syncthreads();

    float s_b2[FW], s_w3[OUTL][FW], s_w1r[NW1], s_b3[OUTL], s_wlr = 0.f;
#pragma unroll
    for (int k = 0; k < FW; ++k) {
      s_b2[k] = 0.f;
#pragma unroll
      for (int a = 0; a < OUTL; ++a) s_w3[a][k] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < NW1; ++c) s_w1r[c] = 0.f;
#pragma unroll
    for (int a = 0; a < OUTL; ++a) s_b3[a] = 0.f;
    float loss_acc = 0.f;
    struct Row { int srow; int mm; float x0, lam, keep; };   // table: state of the row | sample: index and scalar input

    // E1 chunk: a1 = tanh(W1 x + b1) for features f0 + 4 k4 .. + 3 of row rw; hi/lo into TMEM (operand of G1)
    auto layer1_chunk = [&](int k4, const Row& rw, bool last1) {
      const float4 b = *reinterpret_cast<const float4*>(wb + rw.srow * H + f0 + 4 * k4);
      const float4 l = *reinterpret_cast<const float4*>(wl + f0 + 4 * k4);
      float z0 = fmaf(l.x, rw.lam, b.x), z1 = fmaf(l.y, rw.lam, b.y), z2 = fmaf(l.z, rw.lam, b.z), z3 = fmaf(l.w, rw.lam, b.w);
      if (!TABLE) {
        const float4 c = *reinterpret_cast<const float4*>(wa + f0 + 4 * k4);
        z0 = fmaf(c.x, rw.x0, z0); z1 = fmaf(c.y, rw.x0, z1); z2 = fmaf(c.z, rw.x0, z2); z3 = fmaf(c.w, rw.x0, z3);
      }
      if (EXTRA) {   // the caller's dense GEMM over the nature inputs (ma_rmabppo_core.py:312-316)
        const float4 x = *reinterpret_cast<const float4*>(g.z1_extra + rw.mm * z_stride + (int64_t)n * H + f0 + 4 * k4);
        z0 = fmaf(kTanhScale, x.x, z0); z1 = fmaf(kTanhScale, x.y, z1); z2 = fmaf(kTanhScale, x.z, z2); z3 = fmaf(kTanhScale, x.w, z3);
      }
      float x0 = tanh_scaled_u(z0), x1 = tanh_scaled_u(z1), x2 = tanh_scaled_u(z2), x3 = tanh_scaled_u(z3);
      if (RAGGED && last1) { x0 *= rw.keep; x1 *= rw.keep; x2 *= rw.keep; x3 *= rw.keep; }   // rows past the end: a1 = 0 (so d2 = d1 = 0)
      uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
      split_u(x0, h0, l0); split_u(x1, h1, l1); split_u(x2, h2, l2); split_u(x3, h3, l3);
      tmem_st4(tl + C_A1H + f0 + 4 * k4, h0, h1, h2, h3);
      tmem_st4(tl + C_A1L + f0 + 4 * k4, l0, l1, l2, l3);
    };
    // row of this thread in tile tt
    auto row_of = [&](int tt, Row& rw) {
      const int m = tt * TM + r;
      const int mm = !RAGGED || m < M ? m : M - 1;
      rw.mm = mm;
      rw.keep = !RAGGED || m < M ? 1.f : 0.f;
      if (TABLE) {
        const int mrow = RAGGED ? mm : tt * TM;   // aligned: the tile's first row decides the state
        rw.srow = (mrow >= E) + (mrow >= 2 * E);   // S <= 3: no integer division
        rw.lam = slam[mm - rw.srow * E];
        rw.x0 = 0.f;
      } else {
        uint32_t ts = __umulhi((uint32_t)mm, e_magic);
        int e = mm - (int)ts * E;
        if (e >= E) e -= E;
        const int si = min((int)g.s_buf[st_base + mm], s_max);   // sample m = t * E + e of the first T rows of s
        rw.srow = 0;
        rw.lam = slam[e];
        rw.x0 = __fdiv_rn((float)si, state_norm);
      }
    };
    // G1: z2 = a1 . W2^T, three tf32 products (lo.hi, hi.lo, hi.hi)
    auto issue_g1 = [&]() {
      constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
      const uint32_t wh = desc_lo_k(sbase + T_W2H), wlo = desc_lo_k(sbase + T_W2L);
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t acol = tmem + (term == 0 ? C_A1L : C_A1H);
        const uint32_t bt = term == 1 ? wlo : wh;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_ts(tmem + C_D1, acol + 8 * ks, bt + 16 * ks, DESC_HI_K, idesc, (term | ks) != 0);
      }
      umma_commit(bar1);
    };

    // G3 half (k-steps 8 half .. 8 half + 7 of this tile's 16): dW2 += d2^T . a1; DW2 is zeroed at the start of the pass.
    // Two warps issue one half each (issue spread), both accumulating into the same TMEM columns -- so the halves are
    // ORDERED: warp 3 issues only after warp 2 has (named barrier 5 + the tcgen05 thread-sync fences), which fixes the
    // order of the fp32 sum.  Round 1 let them race: the sum, and with it every later trajectory, varied from run to run
    // in the last bits (tools/determinism_probe.py).
    auto issue_g3_half = [&](int half) {
      constexpr uint32_t idesc3 = make_idesc(64, 64, 1, 1);
      const uint32_t dh = desc_lo_mn(sbase + T_D2H), dl = desc_lo_mn(sbase + T_D2L);
      const uint32_t ah = desc_lo_mn(sbase + T_A1H), al = desc_lo_mn(sbase + T_A1L);
      if (half == 1) {
        asm volatile("bar.sync 5, 64;\n" ::: "memory");
        tc_fence_after();
      }
      if (elect_one()) {
#pragma unroll
        for (int term = 0; term < 3; ++term) {
          const uint32_t at = term == 0 ? dl : dh, bt = term == 1 ? al : ah;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            umma_ss(tmem + C_DW2, at + 64 * (8 * half + ks), bt + 64 * (8 * half + ks), DESC_HI_MN, idesc3, 1u);
        }
        umma_commit(bar3);
      }
      __syncwarp();
      if (half == 0) {
        tc_fence_before();
        asm volatile("bar.arrive 5, 64;\n" ::: "memory");
      }
    };
    // E3 of tile tt (its G2 has been issued): d1 = d1' (1 - a1^2); column sums for db1 / dW1.
    auto backward_rowsums = [&](int tt) {
      Row rw;
      row_of(tt, rw);
      mbar_wait(bar2, (gt - 1) & 1);
      tc_fence_after();
      if (warp == 1 && tt + 1 < n_tiles) {   // G1 of the next tile goes behind the finished G2, ahead of G3
        if (elect_one()) issue_g1();
        __syncwarp();
      }
      const bool last_tile = tt + 1 == n_tiles;
      if (last_tile && (warp == 2 || warp == 3))     // no G1 follows the last tile: its G3 runs under E3 and the reductions
        issue_g3_half(warp == 2 ? 0 : 1);
      uint32_t v[FW], om[FW];
      tmem_ld16(tl + C_D1P + f0, v);
      tmem_ld16(tl + C_OM1 + f0, om);   // 1 - a1^2, parked in TMEM by E2 of the same tile
      tmem_wait_ld();
      float d1[FW], d1l[FW];
#pragma unroll
      for (int k = 0; k < FW; ++k) {
        d1[k] = __uint_as_float(v[k]) * __uint_as_float(om[k]);
        d1l[k] = d1[k] * rw.lam;
      }
      s_wlr += colsum16(d1l, lane);
      if (TABLE) {
        // db1 / dW1[:, s]: the tile's column sums go straight to one register per state (lane l holds feature f0 + l / 2)
        const int m_lo = min(tt * TM + 32 * q, M - 1), m_hi = min(tt * TM + 32 * q + 31, M - 1);
        const int s_lo = (m_lo >= E) + (m_lo >= 2 * E), s_hi = (m_hi >= E) + (m_hi >= 2 * E);
        if (!RAGGED || s_lo == s_hi) {   // warp-uniform: all 32 rows are in one state
          const float x = colsum16(d1, lane);
#pragma unroll
          for (int c = 0; c < NW1; ++c) s_w1r[c] += (c == (RAGGED ? s_lo : rw.srow)) ? x : 0.f;
        } else {              // the warp's rows straddle a state boundary: masked sums for this tile
#pragma unroll
          for (int c = 0; c < NW1; ++c) {
            float msk[FW];
#pragma unroll
            for (int k = 0; k < FW; ++k) msk[k] = (rw.srow == c) ? d1[k] : 0.f;
            s_w1r[c] += colsum16(msk, lane);
          }
        }
      } else {
        if (EXTRA && rw.keep != 0.f) {   // dLoss / d(first-layer pre-activation) for the caller's dW1_nature GEMM
          float4* o = reinterpret_cast<float4*>(g.dz1_out + rw.mm * z_stride + (int64_t)n * H + f0);
#pragma unroll
          for (int k4 = 0; k4 < FW / 4; ++k4) o[k4] = make_float4(d1[4 * k4], d1[4 * k4 + 1], d1[4 * k4 + 2], d1[4 * k4 + 3]);
        }
        s_w1r[1] += colsum16(d1, lane);                       // db1
#pragma unroll
        for (int k = 0; k < FW; ++k) d1l[k] = d1[k] * rw.x0;
        s_w1r[0] += colsum16(d1l, lane);                      // dW1[:, 0]
      }
      if (!last_tile && (warp == 2 || warp == 3))     // G3(tt) behind G1(tt + 1): never waited for before the next tile's d2 is ready
        issue_g3_half(warp == 2 ? 0 : 1);
    };
    {
      Row rw;
      row_of(0, rw);
#pragma unroll
      for (int k4 = 0; k4 < FW / 4; ++k4) layer1_chunk(k4, rw, n_tiles == 1);
      const uint32_t z[FW] = {};
      tmem_st16(tl + C_DW2 + f0, z);   // G3 only ever accumulates
    }
    tmem_wait_st();
    fence_async_smem();   // the W2 tiles (initial load / Adam) were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) issue_g1();
      __syncwarp();
    }

    for (int t = 0; t < n_tiles; ++t, ++gt) {
      const int m = t * TM + r;
      const bool valid = RAGGED ? m < M : true;
      const int mm = valid ? m : M - 1;
      // loss inputs of this row (L2 resident), in flight while G1 runs.  Table rows: statistics Ap Am Lp per action, CntS
      // (policy) / RetMean, CntS (value).  Sample rows: action, advantage, old log-prob (policy) / return (value).
      float wrow = 1.f;
      float stA[IS_PI ? 6 : 1];
      int act = 0;
      if (TABLE) {
        const int mrow = RAGGED ? mm : t * TM;
        const int srow = (mrow >= E) + (mrow >= 2 * E), erow = mm - srow * E;
        const uint32_t o_s = (uint32_t)erow + (uint32_t)srow * (uint32_t)E;   // row `srow` of a per-state block
        wrow = __ldg(st + (o_s + (uint32_t)((3 * SA - ROW0) * E)));   // CntS
        if (IS_PI) {
          const uint32_t o_sa = o_s + (uint32_t)srow * (uint32_t)((A - 1) * E);   // row `srow * A` of a per-(state, action) block
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            stA[a] = __ldg(st + (o_sa + (uint32_t)(a * E)));
            stA[2 + a] = __ldg(st + (o_sa + (uint32_t)((SA + a) * E)));
            stA[4 + a] = __ldg(st + (o_sa + (uint32_t)((2 * SA + a) * E)));
          }
        } else {
          stA[0] = __ldg(st + (o_s + (uint32_t)((3 * SA + S - ROW0) * E)));
        }
      } else if (IS_PI) {
        act = g.a_buf[smp_base + mm];
        stA[0] = g.adv[smp_base + mm];
        stA[1] = g.logp_old[smp_base + mm];
      } else {
        stA[0] = g.ret[smp_base + mm];
      }

      // ---- E3 of the previous tile, under this tile's G1 ----
      if (t > 0) backward_rowsums(t - 1);

      // ---- E2: a2 = tanh(z2 + b2); logits (partials meet through smem); loss terms; d2 ----
      mbar_wait(bar1, gt & 1);
      tc_fence_after();
      float a2[FW];
      {
        uint32_t v[FW];
        tmem_ld16(tl + C_D1 + f0, v);
        tmem_wait_ld();
        const float4* b24 = reinterpret_cast<const float4*>(b2s + f0);   // b2 * kTanhScale
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const float4 b = b24[k4];
          a2[4 * k4] = tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4]), kTanhScale, b.x));
          a2[4 * k4 + 1] = tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 1]), kTanhScale, b.y));
          a2[4 * k4 + 2] = tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 2]), kTanhScale, b.z));
          a2[4 * k4 + 3] = tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 3]), kTanhScale, b.w));
        }
      }
#pragma unroll
      for (int a = 0; a < OUTL; ++a) {
        const float4* w34 = reinterpret_cast<const float4*>(DIFF ? w3d + f0 : w + L.oW3 - HH + a * H + f0);
        float pl = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const float4 w3 = w34[k4];
          pl = fmaf(w3.x, a2[4 * k4], pl); pl = fmaf(w3.y, a2[4 * k4 + 1], pl);
          pl = fmaf(w3.z, a2[4 * k4 + 2], pl); pl = fmaf(w3.w, a2[4 * k4 + 3], pl);
        }
        plog[(cq * TM + r) * OUTL + a] = pl;
      }
      asm volatile("bar.sync %0, 128;\n" ::"r"(1 + q) : "memory");   // the four warps that share these 32 rows
      float gl[OUTL];
      {
        float logit[OUTL];
#pragma unroll
        for (int a = 0; a < OUTL; ++a) {
          logit[a] = ((plog[(0 * TM + r) * OUTL + a] + plog[(1 * TM + r) * OUTL + a]) + plog[(2 * TM + r) * OUTL + a]) +
                     plog[(3 * TM + r) * OUTL + a] + (DIFF ? w[L.ob3 - HH] - w[L.ob3 - HH + 1] : w[L.ob3 - HH + a]);
          gl[a] = 0.f;
        }
        if (DIFF) {   // logit[0] = logit of action 0 - logit of action 1
          const float pl = logit[0];
          const float e = __expf(-fabsf(pl));   // exp(smaller logit - larger logit); the larger one contributes exp(0) = 1
          const float sum = 1.f + e, lse = __logf(sum), inv = __fdividef(1.f, sum);
          const bool z = pl >= 0.f;
          const float p[2] = {z ? inv : e * inv, z ? e * inv : inv};
          const float lp[2] = {(z ? 0.f : pl) - lse, (z ? -pl : 0.f) - lse};
          const float ent = -(p[0] * lp[0] + p[1] * lp[1]);
          float ga[2], sumg = 0.f, lrow = 0.f;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const float Ap = stA[a], Am = stA[IS_PI ? 2 + a : 0];
            const float ratio = __expf(lp[a] - stA[IS_PI ? 4 + a : 0]);
            lrow -= fminf(ratio, clip_hi) * Ap + fmaxf(ratio, clip_lo) * Am;
            ga[a] = -((ratio <= clip_hi ? ratio * Ap : 0.f) + (ratio >= clip_lo ? ratio * Am : 0.f));
            sumg += ga[a];
          }
          if (valid) {
            if (cq == 0) loss_acc += lrow - beta_ent * wrow * ent;
            gl[0] = invM * (ga[0] - p[0] * sumg + beta_ent * wrow * p[0] * (lp[0] + ent));
          }
        } else
        if (IS_PI) {
          float mx = logit[0];
#pragma unroll
          for (int a = 1; a < OUT; ++a) mx = fmaxf(mx, logit[a]);
          float ex[OUT], sum = 0.f;
#pragma unroll
          for (int a = 0; a < OUT; ++a) { ex[a] = __expf(logit[a] - mx); sum += ex[a]; }
          const float lse = __logf(sum), inv = __fdividef(1.f, sum);
          float p[OUT], lp[OUT], ent = 0.f;
#pragma unroll
          for (int a = 0; a < OUT; ++a) {
            p[a] = ex[a] * inv;
            lp[a] = (logit[a] - mx) - lse;
            ent = fmaf(-p[a], lp[a], ent);
          }
          if (TABLE) {
            float ga[OUT], sumg = 0.f, lrow = 0.f;
#pragma unroll
            for (int a = 0; a < OUT; ++a) {
              const float Ap = stA[a], Am = stA[TABLE ? 2 + a : 0];
              const float ratio = __expf(lp[a] - stA[TABLE ? 4 + a : 0]);
              lrow -= fminf(ratio, clip_hi) * Ap + fmaxf(ratio, clip_lo) * Am;
              ga[a] = -((ratio <= clip_hi ? ratio * Ap : 0.f) + (ratio >= clip_lo ? ratio * Am : 0.f));
              sumg += ga[a];
            }
            if (valid) {
              if (cq == 0) loss_acc += lrow - beta_ent * wrow * ent;
#pragma unroll
              for (int a = 0; a < OUT; ++a) gl[a] = invM * (ga[a] - p[a] * sumg + beta_ent * wrow * p[a] * (lp[a] + ent));
            }
          } else {   // one sample: compute_loss_pi (agent_oracle.py:285-322) on (act, adv, logp_old)
            float lpa = lp[0];
#pragma unroll
            for (int a = 1; a < OUT; ++a) lpa = act == a ? lp[a] : lpa;
            const float advv = stA[0];
            const float ratio = __expf(lpa - stA[IS_PI && !TABLE ? 1 : 0]);
            const float surr1 = ratio * advv;
            const float surr2 = fminf(fmaxf(ratio, clip_lo), clip_hi) * advv;
            const bool pass = (ratio >= clip_lo && ratio <= clip_hi) || (surr1 < surr2);
            const float coef = pass ? -advv * ratio : 0.f;
            if (valid) {
              if (cq == 0) loss_acc += -fminf(surr1, surr2) - beta_ent * ent;
#pragma unroll
              for (int a = 0; a < OUT; ++a)
                gl[a] = invM * (coef * ((act == a ? 1.f : 0.f) - p[a]) + beta_ent * p[a] * (lp[a] + ent));
            }
          }
        } else {
          const float diff = logit[0] - stA[0];
          if (valid) {
            if (cq == 0) loss_acc = fmaf(wrow * diff, diff, loss_acc);
            gl[0] = invM * 2.f * wrow * diff;
          }
        }
        if (cq == 0)
#pragma unroll
          for (int a = 0; a < OUTL; ++a) s_b3[a] += gl[a];
      }
      // d2 (FMA pipe) interleaved with the NEXT tile's layer 1 (MUFU pipe); z2's a1 operand in TMEM is free once G1 is done
      {
        const bool more = t + 1 < n_tiles;
        Row rw1 = {};
        if (more) row_of(t + 1, rw1);
        if (gt > 0) mbar_wait(bar3, (gt - 1) & 1);   // the previous tile's G3 has finished reading the a1 / d2 tiles
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const uint32_t off = (k4 < 2 ? off_c0 : off_c1) + 16 * (k4 & 1);
          {   // this tile's a1 (still in TMEM) -> the smem tile G3 contracts over; 1 - a1^2 parked in TMEM for E3
            uint32_t ah[4], al[4];
            tmem_ld4(tl + C_A1H + f0 + 4 * k4, ah);
            tmem_ld4(tl + C_A1L + f0 + 4 * k4, al);
            tmem_wait_ld();
            *reinterpret_cast<uint4*>(base + T_A1H + off) = make_uint4(ah[0], ah[1], ah[2], ah[3]);
            *reinterpret_cast<uint4*>(base + T_A1L + off) = make_uint4(al[0], al[1], al[2], al[3]);
            uint32_t om[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a1 = __uint_as_float(ah[j]) + __uint_as_float(al[j]);   // hi + lo is the exact fp32 activation
              om[j] = __float_as_uint(fmaf(-a1, a1, 1.f));
            }
            tmem_st4(tl + C_OM1 + f0 + 4 * k4, om[0], om[1], om[2], om[3]);
          }
          float up[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int a = 0; a < OUTL; ++a) {
            const float4 w3 = *reinterpret_cast<const float4*>((DIFF ? w3d : w + L.oW3 - HH + a * H) + f0 + 4 * k4);
            up[0] = fmaf(gl[a], w3.x, up[0]); up[1] = fmaf(gl[a], w3.y, up[1]);
            up[2] = fmaf(gl[a], w3.z, up[2]); up[3] = fmaf(gl[a], w3.w, up[3]);
          }
          uint32_t h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 4 * k4 + j;
            const float x = a2[k], d2 = up[j] * fmaf(-x, x, 1.f);
            s_b2[k] += d2;
#pragma unroll
            for (int a = 0; a < OUTL; ++a) s_w3[a][k] = fmaf(gl[a], x, s_w3[a][k]);
            split_u(d2, h[j], l[j]);
          }
          tmem_st4(tl + C_D2H + f0 + 4 * k4, h[0], h[1], h[2], h[3]);
          tmem_st4(tl + C_D2L + f0 + 4 * k4, l[0], l[1], l[2], l[3]);
          *reinterpret_cast<uint4*>(base + T_D2H + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(base + T_D2L + off) = make_uint4(l[0], l[1], l[2], l[3]);
          if (more) layer1_chunk(k4, rw1, t + 2 == n_tiles);
        }
      }
      tmem_wait_st();
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        if (elect_one()) {
          {   // G2(t): d1' = d2 . W2
            constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
            const uint32_t wh = desc_lo_k(sbase + T_WTH), wlo = desc_lo_k(sbase + T_WTL);
#pragma unroll
            for (int term = 0; term < 3; ++term) {
              const uint32_t acol = tmem + (term == 0 ? C_D2L : C_D2H);
              const uint32_t bt = term == 1 ? wlo : wh;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) umma_ts(tmem + C_D1P, acol + 8 * ks, bt + 16 * ks, DESC_HI_K, idesc, (term | ks) != 0);
            }
            umma_commit(bar2);
          }
        }
        __syncwarp();
      }
    }
    backward_rowsums(n_tiles - 1);   // E3 of the last tile


    // ---- end of the pass over the rows: per-warp row sums -> smem, dW2 from TMEM -> smem ----
    {
      float vals[NQ];
      vals[0] = colsum16(s_b2, lane);
      vals[1] = s_wlr;
#pragma unroll
      for (int c = 0; c < NW1; ++c) vals[2 + c] = s_w1r[c];
#pragma unroll
      for (int a = 0; a < OUTL; ++a) vals[2 + NW1 + a] = colsum16(s_w3[a], lane);
      if ((lane & 1) == 0)
#pragma unroll
        for (int v = 0; v < NQ; ++v) redw[(warp * NQ + v) * FW + (lane >> 1)] = vals[v];
    }
    float b3tot[OUTL];
#pragma unroll
    for (int a = 0; a < OUTL; ++a) b3tot[a] = warp_sum(s_b3[a]);
    const float loss_tot = g.loss_out ? block_sum(loss_acc, red) : 0.f;
    if (lane == 0 && cq == 0)
#pragma unroll
      for (int a = 0; a < OUTL; ++a) red[40 + q * OUTL + a] = b3tot[a];
    mbar_wait(bar3, (gt - 1) & 1);   // every G3 of this pass has landed in TMEM
    tc_fence_after();
    __syncthreads();                 // all E3 reads of the a1 tile are done: its space becomes the dW2 scratch
    {
      uint32_t v[FW];
      tmem_ld16(tl + C_DW2 + f0, v);
      tmem_wait_ld();
      if (lane < 16) {   // M = 64 accumulator: row j = 16q + lane sits on lane 32q + lane; scratch in K-major tile order
        const int j = 16 * q + lane;
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4)
          *reinterpret_cast<float4*>(scratch + kmaj_off(j, f0 + 4 * k4)) =
              make_float4(__uint_as_float(v[4 * k4]), __uint_as_float(v[4 * k4 + 1]), __uint_as_float(v[4 * k4 + 2]), __uint_as_float(v[4 * k4 + 3]));
      }
    }
    tc_fence_before();
    __syncthreads();
    for (int i = tid; i < H * NQ; i += NTHR) {
      const int k = i / NQ, v = i - k * NQ, kc = k >> 4, kk = k & 15;
      float x = redw[((kc * 4 + 0) * NQ + v) * FW + kk];
#pragma unroll
      for (int qq = 1; qq < 4; ++qq) x += redw[((kc * 4 + qq) * NQ + v) * FW + kk];   // fixed order over the row quarters
      if (v == 0) G[L.ob2 - HH + k] = x;
      else if (v == 1) G[L.oW1 + k * IN + IN - 1] = x;
      else if (v < 2 + NW1) {
        if (TABLE || v == 2) G[L.oW1 + k * IN + (v - 2)] = x;   // table: dW1[:, state]; sample: dW1[:, 0]
        else G[L.ob1 + k] = x;                                   // sample: db1
      } else {
        G[L.oW3 - HH + (v - 2 - NW1) * H + k] = x;
        if (DIFF) G[L.oW3 - HH + H + k] = -x;   // dW3[1] = -dW3[0]
      }
    }
    if (tid < OUTL) {
      float x = 0.f;
      for (int qq = 0; qq < 4; ++qq) x += red[40 + qq * OUTL + tid];
      G[L.ob3 - HH + tid] = x;
      if (DIFF) G[L.ob3 - HH + 1] = -x;
    }
    if (g.loss_out && tid == 0) {
      float extra = 0.f;
      if (TABLE && !IS_PI && g.arm_stats) extra = g.arm_stats[4 * (int64_t)n + 2];
      g.loss_out[(int64_t)n * iters + it] = (loss_tot + extra) * invM;
    }
    __syncthreads();

    // ---- Adam; moments live in global memory (one read-modify-write per iteration) ----
    float step_size, bc2s;
    if (it < kAdamTab64) {
      step_size = adam_tab[it].x; bc2s = adam_tab[it].y;
    } else {   // beyond the table: one thread computes, everybody waits
      b1p = pow(g.hp.beta1, (double)(step0 + it + 1));
      b2p = pow(g.hp.beta2, (double)(step0 + it + 1));
      if (tid == 0) {
        red[62] = (float)(lr / (1.0 - b1p));
        red[63] = (float)sqrt(1.0 - b2p);
      }
      __syncthreads();
      step_size = red[62]; bc2s = red[63];
    }
    auto adam = [&](float x, float gi, float m0, float v0, int i) {
      const float mi = m0 + lerp_w * (gi - m0);
      const float vi = v0 * beta2f + (omb2 * gi) * gi;
      An[i] = mi;
      An[P + i] = vi;
      const float denom = sqrtf(vi) / bc2s + epsf;
      return x - step_size * (mi / denom);
    };
    // W2 block, walked in the order of the K-major tile (conflict-free shared-memory access; the gradient scratch has the
    // same layout): tile position p <-> W2[j][c]
    {
      float m0[HH / NTHR], v0[HH / NTHR];
#pragma unroll
      for (int k = 0; k < HH / NTHR; ++k) {   // all moment loads in flight before the first use
        const int p = tid + k * NTHR;
        const int j = ((p >> 9) << 3) | ((p >> 2) & 7), c = (((p >> 5) & 15) << 2) | (p & 3);
        m0[k] = An[L.oW2 + j * H + c];
        v0[k] = An[P + L.oW2 + j * H + c];
      }
#pragma unroll
      for (int k = 0; k < HH / NTHR; ++k) {
        const int p = tid + k * NTHR;
        const int j = ((p >> 9) << 3) | ((p >> 2) & 7), c = (((p >> 5) & 15) << 2) | (p & 3);
        const float xn = adam(W2H[p] + W2L[p], scratch[p], m0[k], v0[k], L.oW2 + j * H + c);   // hi + lo = exact fp32 weight
        uint32_t hi, lo;
        split_u(xn, hi, lo);
        W2H[p] = __uint_as_float(hi); W2L[p] = __uint_as_float(lo);
        WTH[kmaj_off(c, j)] = __uint_as_float(hi); WTL[kmaj_off(c, j)] = __uint_as_float(lo);
      }
    }
    for (int k = tid; k < P - HH; k += NTHR) {   // everything else, packed order without the W2 block
      const int i = k < L.oW2 ? k : k + HH;
      float gi = G[k];
      if (TABLE && k >= L.ob1 && k < L.oW2) {      // db1 = sum over states of dW1[:, s]
        gi = 0.f;
#pragma unroll
        for (int c = 0; c < S; ++c) gi += G[L.oW1 + (k - L.ob1) * IN + c];
      }
      w[k] = adam(w[k], gi, An[i], An[P + i], i);
    }
    fence_async_smem();   // the new W2 tiles are read by the tensor core in the next pass
    __syncthreads();
  }
  for (int i = tid; i < P; i += NTHR) {
    if (i < L.oW2) Wn[i] = w[i];
    else if (i >= L.ob2) Wn[i] = w[i - HH];
    else {
      const int j = (i - L.oW2) >> 6, c = (i - L.oW2) & 63;
      Wn[i] = W2H[kmaj_off(j, c)] + W2L[kmaj_off(j, c)];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tmem) : "memory");
}

static size_t umma64_smem_bytes(int S, int OUT, bool table, int E) {
  const MlpLayout L(S + 1, H, OUT);
  const int NQ = 2 + (table ? S : 2) + OUT;
  const size_t fl = 2 * (size_t)(L.padded() - HH) + S * H + 4 * H + 4 * TM * OUT + NW * NQ * FW + (((size_t)E + 3) & ~(size_t)3);
  return 1024 + T_END + fl * sizeof(float);
}

constexpr size_t kUmmaSmemLimit = 227 * 1024 - 512;   // 512 B of static __shared__

template <int S, int A, bool IS_PI, int MODE, bool RAGGED>
static int launch_umma64_r(const UmmaArgs& g0, cudaStream_t st, const char* what) {
  constexpr int OUT = IS_PI ? A : 1;
  UmmaArgs g = g0;
  g.clip_lo = (float)(1.0 - g.hp.clip_ratio); g.clip_hi = (float)(1.0 + g.hp.clip_ratio);
  g.beta_ent = (float)g.hp.entropy_coeff;
  g.lerp_w = (float)(1.0 - g.hp.beta1); g.beta2f = (float)g.hp.beta2; g.omb2 = (float)(1.0 - g.hp.beta2);
  g.epsf = (float)g.hp.adam_eps;
  g.invM = 1.f / (float)((int64_t)g.T * g.d.n_envs);
  size_t smem = umma64_smem_bytes(S, OUT, MODE == kRowsTable, g.d.n_envs);
  g.stage_lamb = smem <= kUmmaSmemLimit;
  if (!g.stage_lamb) smem = umma64_smem_bytes(S, OUT, MODE == kRowsTable, 0);   // large E: lambda stays in global memory
  auto kern = ppo_table_umma64_kernel<S, A, IS_PI, MODE, RAGGED>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(RMAB_E_LAUNCH, "%s: cannot reserve %zu B of shared memory", what, smem);
  kern<<<g.d.n_arms, NTHR, smem, st>>>(g);
  return check_launch(what);
}

template <int S, int A, bool IS_PI, int MODE>
static int launch_umma64(const UmmaArgs& g, cudaStream_t st, const char* what) {
  const int64_t rows = MODE == kRowsTable ? (int64_t)g.d.n_envs : (int64_t)g.T * g.d.n_envs;   // table rows: per state
  return rows % TM ? launch_umma64_r<S, A, IS_PI, MODE, true>(g, st, what) : launch_umma64_r<S, A, IS_PI, MODE, false>(g, st, what);
}

// Does the tcgen05 kernel's shared-memory plan fit this shape?  (Always: lambda is staged only when there is room.)
bool ppo_table_umma64_fits(const RmabNetDesc* net) {
  return umma64_smem_bytes(net->n_states, 2, true, 0) <= kUmmaSmemLimit;
}

// Entry used by rmab_ppo_update_table (ppo_table.cu) for H = 64.
int ppo_table_umma64_run(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv, float* adam_pi,
                         float* adam_v, const float* stats, const float* arm_stats, const float* lamb, float* loss_pi,
                         float* loss_v, cudaStream_t st) {
  UmmaArgs g = {};
  g.d = *net; g.hp = *hp; g.T = T; g.lamb = lamb; g.stats = stats; g.arm_stats = arm_stats;
  UmmaArgs gp = g, gv = g;
  gp.W = Wpi; gp.adam = adam_pi; gp.loss_out = loss_pi;
  gv.W = Wv; gv.adam = adam_v; gv.loss_out = loss_v;
  int rc = RMAB_OK;
  if (net->n_states == 2) {
    if (hp->pi_iters > 0) rc = launch_umma64<2, 2, true, kRowsTable>(gp, st, "rmab_ppo_update_table(pi, tcgen05)");
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_umma64<2, 2, false, kRowsTable>(gv, st, "rmab_ppo_update_table(v, tcgen05)");
  } else {
    if (hp->pi_iters > 0) rc = launch_umma64<3, 2, true, kRowsTable>(gp, st, "rmab_ppo_update_table(pi, tcgen05)");
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_umma64<3, 2, false, kRowsTable>(gv, st, "rmab_ppo_update_table(v, tcgen05)");
  }
  return rc;
}

// ---- per-sample rows, scalar state input (rmab_ppo_update / rmab_critic_extra_iter with hidden 64, one_hot = 0) ----
bool ppo_sample_umma64_fits(const RmabNetDesc* net) {
  return !net->one_hot && net->hidden == H && net->n_actions >= 2 && net->n_actions <= 3 &&
         umma64_smem_bytes(1, 3, false, 0) <= kUmmaSmemLimit;
}

// One policy pass (hp->pi_iters Adam steps) and / or one value pass over the [N,T,E] sample buffers.
int ppo_sample_umma64_run(const RmabNetDesc* net, const RmabHyper* hp, int T, int s_rows, float* Wpi, float* Wv,
                          float* adam_pi, float* adam_v, const uint8_t* s, const uint8_t* a, const float* adv,
                          const float* logp_old, const float* ret, const float* lamb, float* loss_pi, float* loss_v,
                          cudaStream_t st) {
  UmmaArgs g = {};
  g.d = *net; g.hp = *hp; g.T = T; g.lamb = lamb; g.s_rows = s_rows; g.s_buf = s; g.a_buf = a; g.adv = adv;
  g.logp_old = logp_old; g.ret = ret;
  UmmaArgs gp = g, gv = g;
  gp.W = Wpi; gp.adam = adam_pi; gp.loss_out = loss_pi;
  gv.W = Wv; gv.adam = adam_v; gv.loss_out = loss_v;
  int rc = RMAB_OK;
  if (hp->pi_iters > 0)
    rc = net->n_actions == 2 ? launch_umma64<1, 2, true, kRowsSample>(gp, st, "rmab_ppo_update(pi, tcgen05)")
                             : launch_umma64<1, 3, true, kRowsSample>(gp, st, "rmab_ppo_update(pi, tcgen05)");
  if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_umma64<1, 2, false, kRowsSample>(gv, st, "rmab_ppo_update(v, tcgen05)");
  return rc;
}

// One Adam step of the MA agent critics (first-layer term of the nature inputs from / to the caller's dense GEMMs).
int ppo_sample_umma64_critic_extra(const RmabNetDesc* net, const RmabHyper* hp, int T, int s_rows, float* Wv, float* adam_v,
                                   const uint8_t* s, const float* ret, const float* lamb, const float* z1_extra,
                                   float* dz1_out, float* loss_v, cudaStream_t st) {
  UmmaArgs g = {};
  g.d = *net; g.hp = *hp; g.T = T; g.lamb = lamb; g.s_rows = s_rows; g.s_buf = s; g.ret = ret;
  g.z1_extra = z1_extra; g.dz1_out = dz1_out; g.W = Wv; g.adam = adam_v; g.loss_out = loss_v;
  return launch_umma64<1, 2, false, kRowsSampleExtra>(g, st, "rmab_critic_extra_iter(tcgen05)");
}

}  // namespace rmab
