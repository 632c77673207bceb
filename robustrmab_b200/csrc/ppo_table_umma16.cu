// Statistics-based PPO update for hidden width 16 on tcgen05 / TMEM -- the headline shape (BASELINE configs[1]: CE,
// 4098 arms x 256 envs, hidden [16,16]).  Same algorithm and building blocks as ppo_table_umma64.cu (umma.cuh); what
// changes is the plan, because a 16 x 16 layer leaves the tensor core idle and the CUDA-core work is everything:
//   * one CTA = 4 warps = one 128-row tile at a time (thread = tile row = TMEM lane, all 16 features), FOUR CTAs per SM
//     (48 KB of shared memory, 128 TMEM columns, 128 registers each) -- the other CTAs' work hides a CTA's barriers and its
//     waits for the tensor pipe, so the schedule inside a CTA is simply E1 | G1 | E2 | G2, G3 | E3 per tile;
//   * the a1 and d2 tiles G3 contracts over hold hi AND lo side by side in one 128-byte swizzled row (features 0..15 = hi,
//     16..31 = lo), so ONE pass of 16 MMAs with M = 64, N = 32 produces all hi/lo cross products of d2^T a1 at once
//     (rows 0..15 x cols 0..15 = hi.hi, rows 0..15 x cols 16..31 = hi.lo, rows 16..31 x cols 0..15 = lo.hi); the d2 tile is
//     followed by the a1 tile in memory, which is what the M = 64 operand's second 32-feature block reads (ignored rows);
//   * G1 / G2 are six MMAs each (M = 128, N = 16): A = hi / lo in TMEM, B = W2 / W2^T K-major 16 x 16 tiles in smem.
// tools/umma_probe.cu modes 6 and 7 pin these two patterns against a CPU product on the device.
// The kernel is bound by instruction issue (4 warps per scheduler, ~ 900 instructions per warp and tile), so what is left is
// what the instructions are spent on: all of shared memory at compile-time offsets of ONE pinned window
// register, float hyper-parameters from the host, the two-action policy head on the logit difference with one logit gradient
// per row, and two instantiations -- `RAGGED` for env counts that are not a multiple of the tile height, the aligned one with
// one state per tile and no row predication for the BASELINE shapes.
//
// Reference behaviour: agent_oracle.py compute_loss_pi :285-322, compute_loss_v :325-339, update :399-436.
#include "common.cuh"
#include "mlp.cuh"
#include "umma.cuh"

namespace rmab {

namespace {

namespace u = umma;

constexpr int H16 = 16, HH16 = 256, TM16 = 128, NT16 = 128;
#ifndef U16_SWAP
#define U16_SWAP 1      // half-swapped 16-byte tile stores (0: natural order, 2-way bank conflicts, 32 selects fewer per store_row)
#endif
constexpr int CTAS16 = 4;   // measured: 3 CTAs with the E3 row sums in registers (170 regs) is 5 % slower, occupancy wins
// shared memory (bytes from a 1024-byte aligned base): swizzled [128 x (hi16 | lo16)] tiles, K-major 16 x 16 tiles
constexpr uint32_t S_D2 = 0, S_A1 = 16384, S_W2H = 32768, S_W2L = 33792, S_WTH = 34816, S_WTL = 35840, S_END = 36864;
// the former static shared objects live in the dynamic block as well, so that every shared address is a compile-time
// constant (+ a per-thread offset): mbarriers, the TMEM base word, the reduction scratch, the Adam table, then the floats
constexpr uint32_t S_BARS = S_END, S_TMEM = S_END + 24, S_RED = S_END + 32, S_ADAM = S_RED + 256;
// TMEM columns (128 allocated per CTA)
constexpr int kAdamTab = 128;
constexpr uint32_t S_FLOATS = S_ADAM + 8 * kAdamTab;
constexpr uint32_t Q_D = 0, Q_OM1 = 16, Q_A1H = 32, Q_A1L = 48, Q_D2H = 64, Q_D2L = 80, Q_DW2 = 96;
constexpr int SCRA = 36, SCRB = 20;   // row pitches (floats) of the dW2 readback scratch: 16-byte stores at 144 B / 80 B pitch
constexpr uint32_t DESC16_HI_K = (512u >> 4) | (1u << 14);                 // SBO = 512 (next 8 rows of a 16-column tile)
constexpr uint32_t DESC16_HI_MN = (512u >> 4) | (1u << 14) | (1u << 29);   // SBO = 512 (next 4 rows), layout type 1

__host__ __device__ __forceinline__ int kmaj16(int r, int f) { return (r >> 3) * 128 + (f >> 2) * 32 + (r & 7) * 4 + (f & 3); }
__device__ __forceinline__ uint32_t lo_k16(uint32_t saddr) { return (saddr >> 4) | ((128u >> 4) << 16); }
__device__ __forceinline__ uint32_t lo_mn16(uint32_t saddr) { return (saddr >> 4) | ((16384u >> 4) << 16); }

struct U16Args {
  RmabNetDesc d;
  RmabHyper hp;
  int T;
  float* W; float* adam;
  const float* lamb; const float* stats; const float* arm_stats;
  float* loss_out;
  int stage_lamb;
  // float images of the hyper-parameters, rounded once on the host: inside the tile loop the compiler would otherwise
  // re-derive them from the doubles (DADD + F2F per tile at the 128-register cap)
  float clip_lo, clip_hi, beta_ent, lerp_w, beta2f, omb2, epsf, invM;
};

}  // namespace

// RAGGED = the env count is not a multiple of the tile height: tiles may straddle two states and the last one may hold rows
// past the end (their a1 is forced to 0 and they are excluded from the sums).  The other instantiation (E % 128 == 0, the
// BASELINE shapes) knows every tile lies in ONE state -- the state index is uniform arithmetic on the tile number -- and drops
// the row predication (16 predicated FMULs in E1 alone) and the straddling branch of E3.
template <int S, bool IS_PI, bool RAGGED>
__global__ void __launch_bounds__(NT16, CTAS16) ppo_table_umma16_kernel(const __grid_constant__ U16Args g) {
  constexpr int H = H16, HH = HH16, A = 2, OUT = IS_PI ? A : 1, IN = S + 1, SA = S * A, FW = 16;
  constexpr int ROW0 = IS_PI ? 0 : 3 * SA, KST = 4 * SA + 3 * S;
  constexpr int NQ = 3 + S;         // per-feature row sums: db2, dW1[:, lambda], dW1[:, s] x S, dW3 (policy: of action 0)
  // packed net layout (MlpLayout) as compile-time constants
  constexpr int oW1 = 0, ob1 = H * IN, oW2 = ob1 + H, ob2 = oW2 + HH, oW3 = ob2 + H, ob3 = oW3 + OUT * H, P = ob3 + OUT;
  constexpr int Ps = ((P + 3) & ~3) - HH;   // w holds the packed layout WITHOUT the W2 block
  // The whole block is dynamic and declared 1024-byte aligned (no static shared objects in this kernel), so the tiles sit
  // at the constants S_* of the CTA's shared window and no address is derived from a run-time alignment fix-up.
  extern __shared__ __align__(1024) unsigned char smraw16[];
  // The window address is pinned in one register (an opaque asm): left alone, the compiler re-derives it from
  // SR_CgaCtaId in every phase of the tile loop (S2R + MOV + LEA, eight times per tile) to save that register.
  uint32_t sbase = u::smem_u32(smraw16);
  asm volatile("" : "+r"(sbase));
  if (sbase & 1023u) __trap();   // the swizzled tiles and the UMMA descriptors rely on the declared alignment
  unsigned char* const base = reinterpret_cast<unsigned char*>(__cvta_shared_to_generic(sbase));
  float* const red = reinterpret_cast<float*>(base + S_RED);                // [64]
  float2* const adam_tab = reinterpret_cast<float2*>(base + S_ADAM);        // (lr / (1 - beta1^t), sqrt(1 - beta2^t)) per step
  const int E = g.d.n_envs;
  float* const w = reinterpret_cast<float*>(base + S_FLOATS);
  float* const wb = w + Ps;                 // [S][H]  (W1[:, s] + b1) * kTanhScale
  float* const wl = wb + S * H;             // [H]     W1[:, lambda] * kTanhScale
  float* const b2s = wl + H;                // [H]     b2 * kTanhScale
  float* const w3d = b2s + H;               // [H + 4] policy: W3[0] - W3[1] then b3[0] - b3[1]; value: W3[0], b3[0]
  float* const redw = w3d + H + 4;          // [4 warps][NQ][16]
  float* const scrA = redw + 4 * NQ * FW;   // [16][SCRA]  dW2 readback: hi rows (cols 0..15 = a1 hi, 16..31 = a1 lo)
  float* const scrB = scrA + H * SCRA;      // [16][SCRB]  lo rows x hi cols
  float* const mom = scrB + H * SCRB;       // [2][P]  Adam m, v of this arm: resident in shared memory for the whole launch
  float* const slam_s = mom + 2 * ((P + 3) & ~3);   // [E] when it fits
  float* const W2H = reinterpret_cast<float*>(base + S_W2H);
  float* const W2L = reinterpret_cast<float*>(base + S_W2L);
  float* const WTH = reinterpret_cast<float*>(base + S_WTH);
  float* const WTL = reinterpret_cast<float*>(base + S_WTL);

  // (tid is NOT pinned like the window base: measured 4.03 vs 3.99 ms -- the register is worth more than the S2R re-reads)
  // the warp index through a shuffle broadcast: the compiler then knows it is warp-uniform, keeps the TMEM addresses derived
  // from it in uniform registers (no R2UR in front of every tcgen05.ld / st) and turns `warp == 0` into uniform branches
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), r = tid;
  float* Wn = g.W + (int64_t)n * P;
  float* An = g.adam + (int64_t)n * 2 * P;
  for (int i = tid; i < P; i += NT16) {
    const float x = Wn[i];
    mom[i] = An[i];
    mom[P + i] = An[P + i];
    if (i < oW2) w[i] = x;
    else if (i >= ob2) w[i - HH] = x;
    else {
      const int j = (i - oW2) >> 4, c = (i - oW2) & 15;
      uint32_t hi, lo;
      u::split_u(x, hi, lo);
      W2H[kmaj16(j, c)] = __uint_as_float(hi); W2L[kmaj16(j, c)] = __uint_as_float(lo);
      WTH[kmaj16(c, j)] = __uint_as_float(hi); WTL[kmaj16(c, j)] = __uint_as_float(lo);
    }
  }
  if (g.stage_lamb)
    for (int i = tid; i < E; i += NT16) slam_s[i] = g.lamb[i];
  if (tid == 0) {
    for (int b = 0; b < 3; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(sbase + S_BARS + 8 * b));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(sbase + S_TMEM) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  u::tc_fence_before();
  __syncthreads();
  u::tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<const uint32_t*>(base + S_TMEM);
  const uint32_t tl = tmem + ((uint32_t)(32 * warp) << 16);   // this warp's lane quarter
  const uint32_t bar1 = sbase + S_BARS, bar2 = sbase + S_BARS + 8, bar3 = sbase + S_BARS + 16;
  const float* st = g.stats + ((int64_t)n * KST + ROW0) * E;
  asm volatile("" : "+l"(st));   // opaque: one 64-bit base + 32-bit offsets per load instead of a 64-bit re-derivation each
  const int M = S * E;
  const int n_tiles = (M + TM16 - 1) / TM16;
  const float invM = g.invM;
  const int iters = IS_PI ? g.hp.pi_iters : g.hp.v_iters;
  const double lr = IS_PI ? g.hp.pi_lr : g.hp.vf_lr;
  const int step0 = IS_PI ? g.hp.adam_step0_pi : g.hp.adam_step0_v;
  const float clip_lo = g.clip_lo, clip_hi = g.clip_hi, beta_ent = g.beta_ent;
  const float lerp_w = g.lerp_w, beta2f = g.beta2f, omb2 = g.omb2, epsf = g.epsf;
  double b1p = pow(g.hp.beta1, (double)step0), b2p = pow(g.hp.beta2, (double)step0);
  for (int k = tid; k < min(iters, kAdamTab); k += NT16)   // the double-precision bias corrections once, in parallel
    adam_tab[k] = make_float2((float)(lr / (1.0 - pow(g.hp.beta1, (double)(step0 + k + 1)))),
                              (float)sqrt(1.0 - pow(g.hp.beta2, (double)(step0 + k + 1))));
  uint32_t gt = 0;   // tiles issued so far: every barrier completes one phase per tile

  // A row's 32-byte chunks sit at only four bank positions (128-byte pitch), so the 8 rows of a warp that share a position
  // would serialise 8-fold on a 16-byte store; rows with bit 2 of the row index set write the upper half of every chunk first
  // (4-fold = the minimum for 512 bytes).  The data swap is selects, the address swap one per-thread constant.
  // This thread's row of the swizzled tiles is four 32-byte chunks (hi 0..7, hi 8..15, lo 0..7, lo 8..15), chunk c at
  // position c ^ (r & 3).  One address register serves all eight stores: position and half are XORs of its bits 4..6.
  const bool up = U16_SWAP ? (r >> 2) & 1 : false;
  const uint32_t arow = sbase + (uint32_t)(r >> 2) * 512 + (uint32_t)(r & 3) * 160 + (up ? 16u : 0u);   // (r&3) * (128 + 32)
  auto st_pair = [&](uint32_t addr, const uint32_t* v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr), "r"(up ? v[4] : v[0]), "r"(up ? v[5] : v[1]),
                 "r"(up ? v[6] : v[2]), "r"(up ? v[7] : v[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(addr ^ 16u), "r"(up ? v[0] : v[4]), "r"(up ? v[1] : v[5]),
                 "r"(up ? v[2] : v[6]), "r"(up ? v[3] : v[7]) : "memory");
  };
  auto store_row = [&](uint32_t tile, const uint32_t (&hi)[FW], const uint32_t (&lo)[FW]) {
    st_pair(arow + tile, &hi[0]);
    st_pair((arow ^ 32u) + tile, &hi[8]);
    st_pair((arow ^ 64u) + tile, &lo[0]);
    st_pair((arow ^ 96u) + tile, &lo[8]);
  };

  for (int it = 0; it < iters; ++it) {
    for (int i = tid; i < S * H; i += NT16) {
      const int s = i >> 4, k = i & 15;
      wb[i] = u::kTanhScale * (w[oW1 + k * IN + s] + w[ob1 + k]);
    }
    if (tid < H) {
      wl[tid] = u::kTanhScale * w[oW1 + tid * IN + IN - 1];
      b2s[tid] = u::kTanhScale * w[ob2 - HH + tid];
      // Two actions: softmax, entropy and their gradients depend on the logit DIFFERENCE only, and the two logit
      // gradients of a row sum to zero (gl[1] = -gl[0]); so the output layer runs on W3[0] - W3[1] with one gradient.
      w3d[tid] = IS_PI ? w[oW3 - HH + tid] - w[oW3 - HH + H + tid] : w[oW3 - HH + tid];
      if (tid == 0) w3d[H] = IS_PI ? w[ob3 - HH] - w[ob3 - HH + 1] : w[ob3 - HH];
    }
    {
      const uint32_t z[FW] = {};
      u::tmem_st16(tl + Q_DW2, z);   // G3 only ever accumulates
      u::tmem_st16(tl + Q_DW2 + 16, z);
    }
    __syncthreads();

    float s_b2[FW], s_w3[FW], s_w1r[S], s_b3 = 0.f, s_wlr = 0.f;
#pragma unroll
    for (int k = 0; k < FW; ++k) s_b2[k] = s_w3[k] = 0.f;
#pragma unroll
    for (int c = 0; c < S; ++c) s_w1r[c] = 0.f;
    float loss_acc = 0.f;

    for (int t = 0; t < n_tiles; ++t, ++gt) {
      const int m = t * TM16 + r;
      const bool valid = RAGGED ? m < M : true;
      const int mm = valid ? m : M - 1;
      const int mrow = RAGGED ? mm : t * TM16;                             // aligned: the tile's first row decides the state
      const int srow = (mrow >= E) + (mrow >= 2 * E), erow = mm - srow * E;   // S <= 3: no integer division
      const float lam = g.stage_lamb ? slam_s[erow] : g.lamb[erow];
      // statistics of this row (L2 resident), in flight during E1 and G1; one uniform base + 32-bit element offsets
      const uint32_t o_s = (uint32_t)erow + (uint32_t)srow * (uint32_t)E;   // row `srow` of a per-state block
      const float wrow = __ldg(st + (o_s + (uint32_t)((3 * SA - ROW0) * E)));   // CntS
      float stA[IS_PI ? 6 : 1];
      if (IS_PI) {
        const uint32_t o_sa = o_s + (uint32_t)srow * (uint32_t)((A - 1) * E);    // row `srow * A` of a per-(state, action) block
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          stA[a] = __ldg(st + (o_sa + (uint32_t)(a * E)));
          stA[2 + a] = __ldg(st + (o_sa + (uint32_t)((SA + a) * E)));
          stA[4 + a] = __ldg(st + (o_sa + (uint32_t)((2 * SA + a) * E)));
        }
      } else {
        stA[0] = __ldg(st + (o_s + (uint32_t)((3 * SA + S - ROW0) * E)));
      }

      // ---- E1: a1 = tanh(W1 x + b1); hi/lo -> TMEM (A of G1) and the smem tile (B of G3); 1 - a1^2 parked in TMEM ----
      {
        uint32_t hi[FW], lo[FW], om[FW];
        const float keep = valid ? 1.f : 0.f;   // rows past the end contribute a1 = 0 (and so d2 = d1 = 0)
        const bool last = t == n_tiles - 1;     // only the last tile can hold such rows
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const float4 b = *reinterpret_cast<const float4*>(wb + srow * H + 4 * k4);
          const float4 l = *reinterpret_cast<const float4*>(wl + 4 * k4);
          float x[4] = {u::tanh_scaled_u(fmaf(l.x, lam, b.x)), u::tanh_scaled_u(fmaf(l.y, lam, b.y)),
                        u::tanh_scaled_u(fmaf(l.z, lam, b.z)), u::tanh_scaled_u(fmaf(l.w, lam, b.w))};
          if (RAGGED && last) { x[0] *= keep; x[1] *= keep; x[2] *= keep; x[3] *= keep; }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            u::split_u(x[j], hi[4 * k4 + j], lo[4 * k4 + j]);
            om[4 * k4 + j] = __float_as_uint(fmaf(-x[j], x[j], 1.f));
          }
        }
        u::tmem_st16(tl + Q_A1H, hi);
        u::tmem_st16(tl + Q_A1L, lo);
        u::tmem_st16(tl + Q_OM1, om);
        if (gt > 0) u::mbar_wait(bar3, (gt - 1) & 1);   // the previous tile's G3 has finished reading the a1 / d2 tiles
        store_row(S_A1, hi, lo);
        u::tmem_wait_st();
      }
      u::fence_async_smem();
      u::tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        u::tc_fence_after();
        if (u::elect_one()) {   // G1: z2 = a1 . W2^T as lo.hi + hi.lo + hi.hi, two k-steps each
          constexpr uint32_t idesc = u::make_idesc(128, 16, 0, 0);
          const uint32_t wh = lo_k16(sbase + S_W2H), wlo = lo_k16(sbase + S_W2L);
#pragma unroll
          for (int term = 0; term < 3; ++term)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              u::umma_ts(tmem + Q_D, tmem + (term == 0 ? Q_A1L : Q_A1H) + 8 * ks, (term == 1 ? wlo : wh) + 16 * ks, DESC16_HI_K, idesc,
                         (term | ks) != 0);
          u::umma_commit(bar1);
        }
        __syncwarp();
      }

      // ---- E2: a2 = tanh(z2 + b2); logits, loss terms, d2 ----
      u::mbar_wait(bar1, gt & 1);
      u::tc_fence_after();
      float a2[FW];
      {
        uint32_t v[FW];
        u::tmem_ld16(tl + Q_D, v);
        u::tmem_wait_ld();
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const float4 b = *reinterpret_cast<const float4*>(b2s + 4 * k4);
          a2[4 * k4] = u::tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4]), u::kTanhScale, b.x));
          a2[4 * k4 + 1] = u::tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 1]), u::kTanhScale, b.y));
          a2[4 * k4 + 2] = u::tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 2]), u::kTanhScale, b.z));
          a2[4 * k4 + 3] = u::tanh_scaled_u(fmaf(__uint_as_float(v[4 * k4 + 3]), u::kTanhScale, b.w));
        }
      }
      float gl = 0.f;   // dLoss / d(logit 0) (policy; logit 1 gets -gl) or dLoss / dv (value)
      {
        float pl = w3d[H];
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const float4 w3 = *reinterpret_cast<const float4*>(w3d + 4 * k4);
          pl = fmaf(w3.x, a2[4 * k4], pl); pl = fmaf(w3.y, a2[4 * k4 + 1], pl);
          pl = fmaf(w3.z, a2[4 * k4 + 2], pl); pl = fmaf(w3.w, a2[4 * k4 + 3], pl);
        }
        if (IS_PI) {   // pl = logit 0 - logit 1
          const float e = __expf(-fabsf(pl));   // exp(smaller logit - larger logit); the larger one contributes exp(0) = 1
          const float sum = 1.f + e, lse = __logf(sum), inv = __fdividef(1.f, sum);
          const bool z = pl >= 0.f;             // logit 0 is the larger one
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
            loss_acc += lrow - beta_ent * wrow * ent;
            gl = invM * (ga[0] - p[0] * sumg + beta_ent * wrow * p[0] * (lp[0] + ent));
          }
        } else {
          const float diff = pl - stA[0];
          if (valid) {
            loss_acc = fmaf(wrow * diff, diff, loss_acc);
            gl = invM * 2.f * wrow * diff;
          }
        }
        s_b3 += gl;
      }
      {
        uint32_t hi[FW], lo[FW];
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const float4 w3 = *reinterpret_cast<const float4*>(w3d + 4 * k4);
          const float up[4] = {gl * w3.x, gl * w3.y, gl * w3.z, gl * w3.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = 4 * k4 + j;
            const float x = a2[k], d2 = up[j] * fmaf(-x, x, 1.f);
            s_b2[k] += d2;
            s_w3[k] = fmaf(gl, x, s_w3[k]);
            u::split_u(d2, hi[k], lo[k]);
          }
        }
        u::tmem_st16(tl + Q_D2H, hi);
        u::tmem_st16(tl + Q_D2L, lo);
        store_row(S_D2, hi, lo);
        u::tmem_wait_st();
      }
      u::fence_async_smem();
      u::tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        u::tc_fence_after();
        if (u::elect_one()) {
          {   // G2: d1' = d2 . W2 (the accumulator of G1: E2 has consumed z2)
            constexpr uint32_t idesc = u::make_idesc(128, 16, 0, 0);
            const uint32_t wh = lo_k16(sbase + S_WTH), wlo = lo_k16(sbase + S_WTL);
#pragma unroll
            for (int term = 0; term < 3; ++term)
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                u::umma_ts(tmem + Q_D, tmem + (term == 0 ? Q_D2L : Q_D2H) + 8 * ks, (term == 1 ? wlo : wh) + 16 * ks, DESC16_HI_K, idesc,
                           (term | ks) != 0);
            u::umma_commit(bar2);
          }
          {   // G3: [d2 hi | d2 lo | (a1 tile)]^T . [a1 hi | a1 lo] over the tile's 128 rows, accumulated across the pass
            constexpr uint32_t idesc3 = u::make_idesc(64, 32, 1, 1);
            const uint32_t da = lo_mn16(sbase + S_D2), db = lo_mn16(sbase + S_A1);
#pragma unroll
            for (int ks = 0; ks < 16; ++ks) u::umma_ss(tmem + Q_DW2, da + 64 * ks, db + 64 * ks, DESC16_HI_MN, idesc3, 1u);
            u::umma_commit(bar3);
          }
        }
        __syncwarp();
      }

      // ---- E3: d1 = d1' (1 - a1^2); column sums for db1 / dW1 ----
      u::mbar_wait(bar2, gt & 1);
      u::tc_fence_after();
      {
        uint32_t v[FW], om[FW];
        u::tmem_ld16(tl + Q_D, v);
        u::tmem_ld16(tl + Q_OM1, om);
        u::tmem_wait_ld();
        float d1[FW];
#pragma unroll
        for (int k = 0; k < FW; ++k) d1[k] = __uint_as_float(v[k]) * __uint_as_float(om[k]);
        // db1 / dW1[:, s] / dW1[:, lambda]: the tile's column sums go straight to one register each (lane l holds feature l / 2)
        const int m_lo = min(t * TM16 + 32 * warp, M - 1), m_hi = min(t * TM16 + 32 * warp + 31, M - 1);
        const int s_lo = (m_lo >= E) + (m_lo >= 2 * E), s_hi = (m_hi >= E) + (m_hi >= 2 * E);
        float d1l[FW];
#pragma unroll
        for (int k = 0; k < FW; ++k) d1l[k] = d1[k] * lam;
        s_wlr += u::colsum16(d1l, lane);
        if (!RAGGED || s_lo == s_hi) {   // warp-uniform: all 32 rows are in one state
          const float x = u::colsum16(d1, lane);
#pragma unroll
          for (int c = 0; c < S; ++c) s_w1r[c] += (c == (RAGGED ? s_lo : srow)) ? x : 0.f;
        } else {              // the warp's rows straddle a state boundary: masked sums for this tile
#pragma unroll
          for (int c = 0; c < S; ++c) {
            float msk[FW];
#pragma unroll
            for (int k = 0; k < FW; ++k) msk[k] = (srow == c) ? d1[k] : 0.f;
            s_w1r[c] += u::colsum16(msk, lane);
          }
        }
      }
    }

    // ---- end of the pass over the rows: per-warp row sums -> smem, dW2 blocks from TMEM -> smem ----
    {
      float vals[NQ];
      vals[0] = u::colsum16(s_b2, lane);
      vals[1] = s_wlr;
#pragma unroll
      for (int c = 0; c < S; ++c) vals[2 + c] = s_w1r[c];
      vals[2 + S] = u::colsum16(s_w3, lane);
      if ((lane & 1) == 0)
#pragma unroll
        for (int v = 0; v < NQ; ++v) redw[(warp * NQ + v) * FW + (lane >> 1)] = vals[v];
    }
    const float b3tot = warp_sum(s_b3);
    const float loss_tot = g.loss_out ? block_sum(loss_acc, red) : 0.f;
    if (lane == 0) red[40 + warp] = b3tot;
    u::mbar_wait(bar3, (gt - 1) & 1);   // every G3 of this pass has landed in TMEM
    u::tc_fence_after();
    if (warp < 2) {   // M = 64 accumulator: row m sits on lane (m % 16) + 32 (m / 16): warp 0 = d2 hi rows, warp 1 = d2 lo rows
      uint32_t v0[FW], v1[FW];
      u::tmem_ld16(tl + Q_DW2, v0);        // columns 0..15  = a1 hi
      u::tmem_ld16(tl + Q_DW2 + 16, v1);   // columns 16..31 = a1 lo
      u::tmem_wait_ld();
      if (lane < 16) {   // 16-byte stores at a pitch of 144 B (80 B): consecutive lanes land 4 banks apart, no conflict
        if (warp == 0) {
#pragma unroll
          for (int k = 0; k < FW; k += 4) {
            *reinterpret_cast<uint4*>(scrA + lane * SCRA + k) = make_uint4(v0[k], v0[k + 1], v0[k + 2], v0[k + 3]);
            *reinterpret_cast<uint4*>(scrA + lane * SCRA + 16 + k) = make_uint4(v1[k], v1[k + 1], v1[k + 2], v1[k + 3]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < FW; k += 4)
            *reinterpret_cast<uint4*>(scrB + lane * SCRB + k) = make_uint4(v0[k], v0[k + 1], v0[k + 2], v0[k + 3]);
        }
      }
    }
    u::tc_fence_before();
    __syncthreads();
    if (g.loss_out && tid == 0) {
      float extra = 0.f;
      if (!IS_PI && g.arm_stats) extra = g.arm_stats[4 * (int64_t)n + 2];
      g.loss_out[(int64_t)n * iters + it] = (loss_tot + extra) * invM;
    }

    // ---- Adam; the moments stay in shared memory for the whole launch (round 1: one L2 read-modify-write per step) ----
    float step_size, bc2s;
    if (it < kAdamTab) {
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
    auto adam = [&](float x, float gi, int i) {
      const float m0 = mom[i], v0 = mom[P + i];
      const float mi = m0 + lerp_w * (gi - m0);
      const float vi = v0 * beta2f + (omb2 * gi) * gi;
      mom[i] = mi;
      mom[P + i] = vi;
      const float denom = sqrtf(vi) / bc2s + epsf;
      return x - step_size * (mi / denom);
    };
#pragma unroll
    for (int k = 0; k < HH / NT16; ++k) {   // W2 block in the order of the K-major tile
      const int p = tid + k * NT16;
      const int j = ((p >> 7) << 3) | ((p >> 2) & 7), c = (((p >> 5) & 3) << 2) | (p & 3);
      const float gi = (scrA[j * SCRA + c] + scrA[j * SCRA + 16 + c]) + scrB[j * SCRB + c];   // hi.hi + hi.lo + lo.hi
      const float xn = adam(W2H[p] + W2L[p], gi, oW2 + j * H + c);                    // hi + lo = exact fp32 weight
      uint32_t hi, lo;
      u::split_u(xn, hi, lo);
      W2H[p] = __uint_as_float(hi); W2L[p] = __uint_as_float(lo);
      WTH[kmaj16(c, j)] = __uint_as_float(hi); WTL[kmaj16(c, j)] = __uint_as_float(lo);
    }
    // everything else, packed order without the W2 block: [W1 | b1 | b2 | W3 | b3]; its gradient is the sum over the four
    // row quarters (fixed order) of the per-warp row sums
    auto sum4 = [&](int f, int v) {
      return ((redw[(0 * NQ + v) * FW + f] + redw[(1 * NQ + v) * FW + f]) + redw[(2 * NQ + v) * FW + f]) + redw[(3 * NQ + v) * FW + f];
    };
    for (int k = tid; k < P - HH; k += NT16) {
      const int i = k < oW2 ? k : k + HH;
      float gi;
      if (k < ob1) {                       // W1[f][c]: c < S state columns, c == S the lambda column
        const int f = k / IN, c = k - f * IN;
        gi = sum4(f, c == IN - 1 ? 1 : 2 + c);
      } else if (k < oW2) {                // b1 = sum over states of dW1[:, s]
        gi = 0.f;
#pragma unroll
        for (int c = 0; c < S; ++c) gi += sum4(k - ob1, 2 + c);
      } else if (k < oW3 - HH) {           // b2
        gi = sum4(k - (ob2 - HH), 0);
      } else if (k < ob3 - HH) {           // W3[a][f]
        const int a = (k - (oW3 - HH)) >> 4, f = (k - (oW3 - HH)) & 15;
        gi = sum4(f, 2 + S);
        if (a == 1) gi = -gi;
      } else {                               // b3[a]
        const int a = k - (ob3 - HH);
        gi = ((red[40] + red[41]) + red[42]) + red[43];
        if (a == 1) gi = -gi;
      }
      w[k] = adam(w[k], gi, i);
    }
    u::fence_async_smem();   // the new W2 tiles are read by the tensor core in the next pass
    __syncthreads();
  }
  for (int i = tid; i < P; i += NT16) {
    An[i] = mom[i];
    An[P + i] = mom[P + i];
    if (i < oW2) Wn[i] = w[i];
    else if (i >= ob2) Wn[i] = w[i - HH];
    else {
      const int j = (i - oW2) >> 4, c = (i - oW2) & 15;
      Wn[i] = W2H[kmaj16(j, c)] + W2L[kmaj16(j, c)];
    }
  }
  u::tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(tmem) : "memory");
}

static size_t umma16_smem_bytes(int S, int OUT, int E) {
  const MlpLayout L(S + 1, H16, OUT);
  const int NQ = 3 + S;
  const size_t fl = (size_t)(L.padded() - HH16) + S * H16 + 3 * H16 + 4 + 4 * NQ * 16 + H16 * SCRA + H16 * SCRB +
                    2 * (size_t)((L.P + 3) & ~3) + (((size_t)E + 3) & ~(size_t)3);
  return S_FLOATS + fl * sizeof(float);
}

template <int S, bool IS_PI, bool RAGGED>
static int launch_umma16_r(const U16Args& g0, cudaStream_t st) {
  constexpr int OUT = IS_PI ? 2 : 1;
  U16Args g = g0;
  size_t smem = umma16_smem_bytes(S, OUT, g.d.n_envs);
  g.stage_lamb = smem <= (size_t)(226 / CTAS16) * 1024;   // CTAS16 CTAs per SM
  if (!g.stage_lamb) smem = umma16_smem_bytes(S, OUT, 0);
  auto kern = ppo_table_umma16_kernel<S, IS_PI, RAGGED>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(RMAB_E_LAUNCH, "rmab_ppo_update_table: cannot reserve %zu B of shared memory", smem);
  kern<<<g.d.n_arms, NT16, smem, st>>>(g);
  return check_launch(IS_PI ? "rmab_ppo_update_table(pi, tcgen05 h16)" : "rmab_ppo_update_table(v, tcgen05 h16)");
}

template <int S, bool IS_PI>
static int launch_umma16(const U16Args& g, cudaStream_t st) {
  return g.d.n_envs % TM16 ? launch_umma16_r<S, IS_PI, true>(g, st) : launch_umma16_r<S, IS_PI, false>(g, st);
}

// Entry used by rmab_ppo_update_table (ppo_table.cu) for H = 16.
int ppo_table_umma16_run(const RmabNetDesc* net, const RmabHyper* hp, int T, float* Wpi, float* Wv, float* adam_pi,
                         float* adam_v, const float* stats, const float* arm_stats, const float* lamb, float* loss_pi,
                         float* loss_v, cudaStream_t st) {
  U16Args g = {};
  g.d = *net; g.hp = *hp; g.T = T; g.lamb = lamb; g.stats = stats; g.arm_stats = arm_stats;
  g.clip_lo = (float)(1.0 - hp->clip_ratio); g.clip_hi = (float)(1.0 + hp->clip_ratio);
  g.beta_ent = (float)hp->entropy_coeff;
  g.lerp_w = (float)(1.0 - hp->beta1); g.beta2f = (float)hp->beta2; g.omb2 = (float)(1.0 - hp->beta2);
  g.epsf = (float)hp->adam_eps;
  g.invM = 1.f / (float)((int64_t)T * net->n_envs);
  U16Args gp = g, gv = g;
  gp.W = Wpi; gp.adam = adam_pi; gp.loss_out = loss_pi;
  gv.W = Wv; gv.adam = adam_v; gv.loss_out = loss_v;
  int rc = RMAB_OK;
  if (net->n_states == 2) {
    if (hp->pi_iters > 0) rc = launch_umma16<2, true>(gp, st);
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_umma16<2, false>(gv, st);
  } else {
    if (hp->pi_iters > 0) rc = launch_umma16<3, true>(gp, st);
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_umma16<3, false>(gv, st);
  }
  return rc;
}

}  // namespace rmab
