// Statistics-based PPO update for hidden width 64 on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM) -- the ARMMAN N = 100k x 1024 envs shape, where S*E = 3072 rows x 64 x 64 per arm per Adam step is a real GEMM.
// Same algorithm as ppo_table_mma64.cu (loss statistics, 3xTF32 hi/lo split, fp32 accumulation); one CTA per arm.
//
// A CTA walks the arm's rows in tiles of 128 (TMEM lane = tile row).  16 warps: warp (q, cq) owns rows 32q..32q+31 (its
// TMEM lane quarter) and the 16 hidden features 16cq..16cq+15.  Per tile, three GEMMs issued by one thread:
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
//
// Reference behaviour: agent_oracle.py compute_loss_pi :285-322, compute_loss_v :325-339, update :399-436.
#include "common.cuh"
#include "mlp.cuh"

namespace rmab {

namespace {

constexpr int H = 64, HH = H * H, TM = 128, NW = 16, NTHR = NW * 32, FW = 16;
constexpr float kTanhScale = 2.885390081777927f;   // 2 / ln 2: tanh(x) = 1 - 2 / (2^(x * kTanhScale) + 1)
// shared-memory tiles (byte offsets from a 1024-byte aligned base)
constexpr uint32_t T_A1H = 0, T_A1L = 32768, T_D2H = 65536, T_D2L = 98304;             // 128 x 64, swizzled MN-major
constexpr uint32_t T_W2H = 131072, T_W2L = 147456, T_WTH = 163840, T_WTL = 180224;     // 64 x 64, K-major core matrices
constexpr uint32_t T_END = 196608;
constexpr uint32_t BLK = 16384;   // bytes per 32-feature block of a 128-row MN-major tile
// TMEM columns (512 allocated: the CTA owns its SM)
constexpr uint32_t C_D1 = 0, C_D1P = 64, C_DW2 = 128, C_A1H = 192, C_A1L = 256, C_D2H = 320, C_D2L = 384, C_OM1 = 448;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major 8-row x 16-byte core matrices: element (r, f) of a [rows][64] tile, in floats
__host__ __device__ __forceinline__ int kmaj_off(int r, int f) { return (r >> 3) * 512 + (f >> 2) * 32 + (r & 7) * 4 + (f & 3); }

// Shared-memory matrix descriptors, as (lo, hi) words: start address >> 4 | LBO >> 4 << 16 ; SBO >> 4 | version 1 << 14 | layout << 29.
// K-major core matrices: LBO = 128 (next 4 k), SBO = 2048 (next 8 rows), no swizzle; one k-step (8 k) = +256 B.
// MN-major swizzled tiles: LBO = BLK (next 32 features), SBO = 512 (next 4 rows), layout type 1; one k-step (8 rows) = +1024 B.
constexpr uint32_t DESC_HI_K = (2048u >> 4) | (1u << 14);
constexpr uint32_t DESC_HI_MN = (512u >> 4) | (1u << 14) | (1u << 29);
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t saddr) { return (saddr >> 4) | ((128u >> 4) << 16); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t saddr) { return (saddr >> 4) | ((BLK >> 4) << 16); }

// kind::tf32, fp32 accumulate: c_format F32 (bit 4), a/b format TF32 (bits 7, 10), majors (bits 15, 16), N>>3, M>>4
constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 ad, {%1, %3};\n\tmov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], ad, bd, %4, p;\n\t}\n" ::"r"(d_tmem),
      "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc)
      : "memory");
}

// One lane of a converged warp (the compiler then knows a single thread issues, and keeps the operands uniform).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}

// Bounded wait: a wrong phase must fail the launch, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int spin = 0; spin < (1 << 26); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return;
  }
  __trap();
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n" ::"r"(v[0]),
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0,%1,%2,%3};\n" ::"r"(a), "r"(b), "r"(c), "r"(d), "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];\n" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}

__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ float tanh_fast_u(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

// tanh(x) given y = x * 2 / ln 2 (callers fold the scale into the layer's weights / the accumulator read)
__device__ __forceinline__ float tanh_scaled_u(float y) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

__device__ __forceinline__ void split_u(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// Column sums over the 32 lanes of 16 per-lane values: lane l returns the sum of column l >> 1 (both lanes of a pair).
__device__ __forceinline__ float colsum16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = h4 ? v[i + 8] : v[i], send = h4 ? v[i] : v[i + 8];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h3 ? a[i + 4] : a[i], send = h3 ? a[i] : a[i + 4];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h2 ? b[i + 2] : b[i], send = h2 ? b[i] : b[i + 2];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h1 ? c[1] : c[0], send = h1 ? c[0] : c[1];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

struct UmmaArgs {
  RmabNetDesc d;
  RmabHyper hp;
  int T;
  float* W; float* adam;
  const float* lamb; const float* stats; const float* arm_stats;
  float* loss_out;
};

}  // namespace

template <int S, bool IS_PI>
__global__ void __launch_bounds__(NTHR, 1) ppo_table_umma64_kernel(UmmaArgs g) {
  constexpr int A = 2, OUT = IS_PI ? A : 1, IN = S + 1, SA = S * A;
  constexpr int ROW0 = IS_PI ? 0 : 3 * SA, KST = 4 * SA + 3 * S;
  constexpr int NQ = 2 + S + OUT;   // per-feature row sums: db2, dW1[:, lambda], dW1[:, s] x S, dW3 x OUT
  extern __shared__ unsigned char smraw[];
  __shared__ __align__(8) uint64_t bars[3];
  __shared__ uint32_t tmem_s;
  __shared__ float red[64];
  unsigned char* base = smraw + ((1024u - (smem_u32(smraw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(base);
  const int E = g.d.n_envs, T = g.T;
  const MlpLayout L(IN, H, OUT);
  const int P = L.P, Ps = L.padded() - HH;   // w / G hold the packed layout WITHOUT the W2 block
  float* w = reinterpret_cast<float*>(base + T_END);
  float* G = w + Ps;
  float* wb = G + Ps;                 // [S][H]  (W1[:, s] + b1) * kTanhScale
  float* wl = wb + S * H;             // [H]     W1[:, lambda] * kTanhScale
  float* b2s = wl + H;                // [H]     b2 * kTanhScale
  float* plog = b2s + H;              // [4 cq][128 rows][OUT] partial logits
  float* redw = plog + 4 * TM * OUT;  // [NW][NQ][16]
  float* slam = redw + NW * NQ * FW;  // [E]
  float* W2H = reinterpret_cast<float*>(base + T_W2H);
  float* W2L = reinterpret_cast<float*>(base + T_W2L);
  float* WTH = reinterpret_cast<float*>(base + T_WTH);
  float* WTL = reinterpret_cast<float*>(base + T_WTL);
  float* scratch = reinterpret_cast<float*>(base + T_A1H);   // dW2 [64][64] between the last GEMM of an iteration and Adam

  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
  for (int i = tid; i < E; i += NTHR) slam[i] = g.lamb[i];
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

  const float* st = g.stats + ((int64_t)n * KST + ROW0) * E;
  const int M = S * E;
  const int n_tiles = (M + TM - 1) / TM;
  const float invM = 1.f / (float)((int64_t)T * E);
  const int iters = IS_PI ? g.hp.pi_iters : g.hp.v_iters;
  const double lr = IS_PI ? g.hp.pi_lr : g.hp.vf_lr;
  const int step0 = IS_PI ? g.hp.adam_step0_pi : g.hp.adam_step0_v;
  const float clip_lo = (float)(1.0 - g.hp.clip_ratio), clip_hi = (float)(1.0 + g.hp.clip_ratio);
  const float beta_ent = (float)g.hp.entropy_coeff;
  const float lerp_w = (float)(1.0 - g.hp.beta1), beta2f = (float)g.hp.beta2, omb2 = (float)(1.0 - g.hp.beta2);
  const float epsf = (float)g.hp.adam_eps;
  double b1p = pow(g.hp.beta1, (double)step0), b2p = pow(g.hp.beta2, (double)step0);
  uint32_t gt = 0;   // tiles issued so far: every barrier completes one phase per tile

  for (int it = 0; it < iters; ++it) {
    for (int i = tid; i < S * H; i += NTHR) {
      const int s = i >> 6, k = i & 63;
      wb[i] = kTanhScale * (w[L.oW1 + k * IN + s] + w[L.ob1 + k]);
    }
    for (int i = tid; i < H; i += NTHR) {
      wl[i] = kTanhScale * w[L.oW1 + i * IN + IN - 1];
      b2s[i] = kTanhScale * w[L.ob2 - HH + i];
    }
    __syncthreads();

    float s_b2[FW], s_w3[OUT][FW], s_w1r[S], s_b3[OUT], s_wlr = 0.f;
#pragma unroll
    for (int k = 0; k < FW; ++k) {
      s_b2[k] = 0.f;
#pragma unroll
      for (int a = 0; a < OUT; ++a) s_w3[a][k] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < S; ++c) s_w1r[c] = 0.f;
#pragma unroll
    for (int a = 0; a < OUT; ++a) s_b3[a] = 0.f;
    float loss_acc = 0.f;

    // E1 chunk: a1 = tanh(W1 x + b1) for features f0 + 4 k4 .. + 3 of row (srow1, lam1); hi/lo into TMEM (operand of G1)
    auto layer1_chunk = [&](int k4, int srow1, float lam1, float keep1, bool last1) {
      const float4 b = *reinterpret_cast<const float4*>(wb + srow1 * H + f0 + 4 * k4);
      const float4 l = *reinterpret_cast<const float4*>(wl + f0 + 4 * k4);
      float x0 = tanh_scaled_u(fmaf(l.x, lam1, b.x)), x1 = tanh_scaled_u(fmaf(l.y, lam1, b.y));
      float x2 = tanh_scaled_u(fmaf(l.z, lam1, b.z)), x3 = tanh_scaled_u(fmaf(l.w, lam1, b.w));
      if (last1) { x0 *= keep1; x1 *= keep1; x2 *= keep1; x3 *= keep1; }   // rows past the end: a1 = 0 (so d2 = d1 = 0)
      uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
      split_u(x0, h0, l0); split_u(x1, h1, l1); split_u(x2, h2, l2); split_u(x3, h3, l3);
      tmem_st4(tl + C_A1H + f0 + 4 * k4, h0, h1, h2, h3);
      tmem_st4(tl + C_A1L + f0 + 4 * k4, l0, l1, l2, l3);
    };
    // row of this thread in tile tt: state, lambda, validity
    auto row_of = [&](int tt, int& srow1, float& lam1, float& keep1) {
      const int m = tt * TM + r;
      const int mm = m < M ? m : M - 1;
      srow1 = (mm >= E) + (mm >= 2 * E);   // S <= 3: no integer division
      lam1 = slam[mm - srow1 * E];
      keep1 = m < M ? 1.f : 0.f;
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

    // G3 half (k-steps ks0 .. ks0+7 of this tile's 16): dW2 += d2^T . a1; DW2 is zeroed at the start of the pass
    auto issue_g3_half = [&](int ks0) {
      constexpr uint32_t idesc3 = make_idesc(64, 64, 1, 1);
      const uint32_t dh = desc_lo_mn(sbase + T_D2H), dl = desc_lo_mn(sbase + T_D2L);
      const uint32_t ah = desc_lo_mn(sbase + T_A1H), al = desc_lo_mn(sbase + T_A1L);
#pragma unroll
      for (int term = 0; term < 3; ++term) {
        const uint32_t at = term == 0 ? dl : dh, bt = term == 1 ? al : ah;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_ss(tmem + C_DW2, at + 64 * (ks0 + ks), bt + 64 * (ks0 + ks), DESC_HI_MN, idesc3, 1u);
      }
      umma_commit(bar3);
    };
    // E3 of tile tt (its G2 has been issued): d1 = d1' (1 - a1^2); column sums for db1 / dW1.
    auto backward_rowsums = [&](int tt) {
      int srow1; float lam1, keep1;
      row_of(tt, srow1, lam1, keep1);
      mbar_wait(bar2, (gt - 1) & 1);
      tc_fence_after();
      if (warp == 1 && tt + 1 < n_tiles) {   // G1 of the next tile goes behind the finished G2, ahead of G3
        if (elect_one()) issue_g1();
        __syncwarp();
      }
      uint32_t v[FW], om[FW];
      tmem_ld16(tl + C_D1P + f0, v);
      tmem_ld16(tl + C_OM1 + f0, om);   // 1 - a1^2, parked in TMEM by E2 of the same tile
      tmem_wait_ld();
      float d1[FW], d1l[FW];
#pragma unroll
      for (int k = 0; k < FW; ++k) {
        d1[k] = __uint_as_float(v[k]) * __uint_as_float(om[k]);
        d1l[k] = d1[k] * lam1;
      }
      s_wlr += colsum16(d1l, lane);
      // db1 / dW1[:, s]: the tile's column sums go straight to one register per state (lane l holds feature f0 + l / 2)
      const int m_lo = min(tt * TM + 32 * q, M - 1), m_hi = min(tt * TM + 32 * q + 31, M - 1);
      const int s_lo = (m_lo >= E) + (m_lo >= 2 * E), s_hi = (m_hi >= E) + (m_hi >= 2 * E);
      if (s_lo == s_hi) {   // warp-uniform: all 32 rows are in one state
        const float x = colsum16(d1, lane);
#pragma unroll
        for (int c = 0; c < S; ++c) s_w1r[c] += (c == s_lo) ? x : 0.f;
      } else {              // the warp's rows straddle a state boundary: masked sums for this tile
#pragma unroll
        for (int c = 0; c < S; ++c) {
          float msk[FW];
#pragma unroll
          for (int k = 0; k < FW; ++k) msk[k] = (srow1 == c) ? d1[k] : 0.f;
          s_w1r[c] += colsum16(msk, lane);
        }
      }
      if (warp == 2 || warp == 3) {   // G3(tt) behind G1(tt + 1): never waited for before the next tile's d2 is ready
        if (elect_one()) issue_g3_half(warp == 2 ? 0 : 8);
        __syncwarp();
      }
    };
    {
      int srow1; float lam1, keep1;
      row_of(0, srow1, lam1, keep1);
#pragma unroll
      for (int k4 = 0; k4 < FW / 4; ++k4) layer1_chunk(k4, srow1, lam1, keep1, n_tiles == 1);
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
      const bool valid = m < M;
      const int mm = valid ? m : M - 1;
      const int srow = (mm >= E) + (mm >= 2 * E), erow = mm - srow * E;
      // statistics of this row (L2 resident), in flight while G1 runs
      const float* sp = st + erow;
      const float wrow = sp[(3 * SA - ROW0 + srow) * E];   // CntS
      float stA[IS_PI ? 6 : 1];
      if (IS_PI) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int c = srow * A + a;
          stA[a] = sp[c * E]; stA[2 + a] = sp[(SA + c) * E]; stA[4 + a] = sp[(2 * SA + c) * E];
        }
      } else {
        stA[0] = sp[(3 * SA + S - ROW0 + srow) * E];
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
      for (int a = 0; a < OUT; ++a) {
        const float4* w34 = reinterpret_cast<const float4*>(w + L.oW3 - HH + a * H + f0);
        float pl = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < FW / 4; ++k4) {
          const float4 w3 = w34[k4];
          pl = fmaf(w3.x, a2[4 * k4], pl); pl = fmaf(w3.y, a2[4 * k4 + 1], pl);
          pl = fmaf(w3.z, a2[4 * k4 + 2], pl); pl = fmaf(w3.w, a2[4 * k4 + 3], pl);
        }
        plog[(cq * TM + r) * OUT + a] = pl;
      }
      asm volatile("bar.sync %0, 128;\n" ::"r"(1 + q) : "memory");   // the four warps that share these 32 rows
      float gl[OUT];
      {
        float logit[OUT];
#pragma unroll
        for (int a = 0; a < OUT; ++a) {
          logit[a] = ((plog[(0 * TM + r) * OUT + a] + plog[(1 * TM + r) * OUT + a]) + plog[(2 * TM + r) * OUT + a]) +
                     plog[(3 * TM + r) * OUT + a] + w[L.ob3 - HH + a];
          gl[a] = 0.f;
        }
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
            const float Ap = stA[a], Am = stA[IS_PI ? 2 + a : 0];
            const float ratio = __expf(lp[a] - stA[IS_PI ? 4 + a : 0]);
            lrow -= fminf(ratio, clip_hi) * Ap + fmaxf(ratio, clip_lo) * Am;
            ga[a] = -((ratio <= clip_hi ? ratio * Ap : 0.f) + (ratio >= clip_lo ? ratio * Am : 0.f));
            sumg += ga[a];
          }
          if (valid) {
            if (cq == 0) loss_acc += lrow - beta_ent * wrow * ent;
#pragma unroll
            for (int a = 0; a < 2; ++a) gl[a < OUT ? a : 0] = invM * (ga[a] - p[a] * sumg + beta_ent * wrow * p[a] * (lp[a] + ent));
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
          for (int a = 0; a < OUT; ++a) s_b3[a] += gl[a];
      }
      // d2 (FMA pipe) interleaved with the NEXT tile's layer 1 (MUFU pipe); z2's a1 operand in TMEM is free once G1 is done
      {
        const bool more = t + 1 < n_tiles;
        int srow1 = 0; float lam1 = 0.f, keep1 = 0.f;
        if (more) row_of(t + 1, srow1, lam1, keep1);
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
          for (int a = 0; a < OUT; ++a) {
            const float4 w3 = *reinterpret_cast<const float4*>(w + L.oW3 - HH + a * H + f0 + 4 * k4);
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
            for (int a = 0; a < OUT; ++a) s_w3[a][k] = fmaf(gl[a], x, s_w3[a][k]);
            split_u(d2, h[j], l[j]);
          }
          tmem_st4(tl + C_D2H + f0 + 4 * k4, h[0], h[1], h[2], h[3]);
          tmem_st4(tl + C_D2L + f0 + 4 * k4, l[0], l[1], l[2], l[3]);
          *reinterpret_cast<uint4*>(base + T_D2H + off) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(base + T_D2L + off) = make_uint4(l[0], l[1], l[2], l[3]);
          if (more) layer1_chunk(k4, srow1, lam1, keep1, t + 2 == n_tiles);
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
      for (int c = 0; c < S; ++c) vals[2 + c] = s_w1r[c];
#pragma unroll
      for (int a = 0; a < OUT; ++a) vals[2 + S + a] = colsum16(s_w3[a], lane);
      if ((lane & 1) == 0)
#pragma unroll
        for (int v = 0; v < NQ; ++v) redw[(warp * NQ + v) * FW + (lane >> 1)] = vals[v];
    }
    float b3tot[OUT];
#pragma unroll
    for (int a = 0; a < OUT; ++a) b3tot[a] = warp_sum(s_b3[a]);
    const float loss_tot = g.loss_out ? block_sum(loss_acc, red) : 0.f;
    if (lane == 0 && cq == 0)
#pragma unroll
      for (int a = 0; a < OUT; ++a) red[40 + q * OUT + a] = b3tot[a];
    mbar_wait(bar3, (gt - 1) & 1);   // every G3 of this pass has landed in TMEM
    tc_fence_after();
    __syncthreads();                 // all E3 reads of the a1 tile are done: its space becomes the dW2 scratch
    {
      uint32_t v[FW];
      tmem_ld16(tl + C_DW2 + f0, v);
      tmem_wait_ld();
      if (lane < 16) {   // M = 64 accumulator: row j = 16q + lane sits on lane 32q + lane
        float4* o = reinterpret_cast<float4*>(scratch + (16 * q + lane) * H + f0);
        o[0] = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
        o[1] = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
        o[2] = make_float4(__uint_as_float(v[8]), __uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
        o[3] = make_float4(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]), __uint_as_float(v[15]));
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
      else if (v < 2 + S) G[L.oW1 + k * IN + (v - 2)] = x;
      else G[L.oW3 - HH + (v - 2 - S) * H + k] = x;
    }
    if (tid < OUT) {
      float x = 0.f;
      for (int qq = 0; qq < 4; ++qq) x += red[40 + qq * OUT + tid];
      G[L.ob3 - HH + tid] = x;
    }
    if (g.loss_out && tid == 0) {
      float extra = 0.f;
      if (!IS_PI && g.arm_stats) extra = g.arm_stats[4 * (int64_t)n + 2];
      g.loss_out[(int64_t)n * iters + it] = (loss_tot + extra) * invM;
    }
    __syncthreads();
    for (int k = tid; k < H; k += NTHR) {   // db1 = sum over states of dW1[:, s]
      float x = 0.f;
#pragma unroll
      for (int c = 0; c < S; ++c) x += G[L.oW1 + k * IN + c];
      G[L.ob1 + k] = x;
    }

    // ---- Adam; moments live in global memory (one read-modify-write per iteration) ----
    b1p *= g.hp.beta1;
    b2p *= g.hp.beta2;
    if (tid == 0) {
      red[62] = (float)(lr / (1.0 - b1p));
      red[63] = (float)sqrt(1.0 - b2p);
    }
    __syncthreads();
    const float step_size = red[62], bc2s = red[63];
    for (int i = tid; i < P; i += NTHR) {
      const bool in_w2 = i >= L.oW2 && i < L.ob2;
      float gi, x;
      int j = 0, c = 0;
      if (in_w2) {
        j = (i - L.oW2) >> 6; c = (i - L.oW2) & 63;
        gi = scratch[i - L.oW2];
        x = W2H[kmaj_off(j, c)] + W2L[kmaj_off(j, c)];   // hi + lo is the exact fp32 weight
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
        split_u(xn, hi, lo);
        W2H[kmaj_off(j, c)] = __uint_as_float(hi); W2L[kmaj_off(j, c)] = __uint_as_float(lo);
        WTH[kmaj_off(c, j)] = __uint_as_float(hi); WTL[kmaj_off(c, j)] = __uint_as_float(lo);
      } else {
        w[i < L.oW2 ? i : i - HH] = xn;
      }
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

static size_t umma64_smem_bytes(int S, int OUT, int E) {
  const MlpLayout L(S + 1, H, OUT);
  const int NQ = 2 + S + OUT;
  const size_t fl = 2 * (size_t)(L.padded() - HH) + S * H + 2 * H + 4 * TM * OUT + NW * NQ * FW + (((size_t)E + 3) & ~(size_t)3);
  return 1024 + T_END + fl * sizeof(float);
}

template <int S, bool IS_PI>
static int launch_umma64(const UmmaArgs& g, cudaStream_t st) {
  constexpr int OUT = IS_PI ? 2 : 1;
  const size_t smem = umma64_smem_bytes(S, OUT, g.d.n_envs);
  auto kern = ppo_table_umma64_kernel<S, IS_PI>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(RMAB_E_LAUNCH, "rmab_ppo_update_table: cannot reserve %zu B of shared memory", smem);
  kern<<<g.d.n_arms, NTHR, smem, st>>>(g);
  return check_launch(IS_PI ? "rmab_ppo_update_table(pi, tcgen05)" : "rmab_ppo_update_table(v, tcgen05)");
}

// Does the tcgen05 kernel's shared-memory plan fit this shape?  (E bounds the staged lambda vector.)
bool ppo_table_umma64_fits(const RmabNetDesc* net) {
  return umma64_smem_bytes(net->n_states, 2, net->n_envs) <= 227 * 1024 - 512;   // 512 B of static __shared__
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
    if (hp->pi_iters > 0) rc = launch_umma64<2, true>(gp, st);
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_umma64<2, false>(gv, st);
  } else {
    if (hp->pi_iters > 0) rc = launch_umma64<3, true>(gp, st);
    if (rc == RMAB_OK && hp->v_iters > 0) rc = launch_umma64<3, false>(gv, st);
  }
  return rc;
}

}  // namespace rmab
