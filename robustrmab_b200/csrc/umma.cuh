// tcgen05 / TMEM building blocks shared by the hidden-64 update kernels (ppo_table_umma64.cu: statistics rows;
// ppo_sample_umma64.cu: per-sample rows).  Tile geometry, shared-memory / TMEM maps, descriptor encodings (every one of
// them pinned on the device by tools/umma_probe.cu), TMEM loads / stores, mbarrier wait, the 3xTF32 split.
#pragma once
#include "common.cuh"

namespace rmab {
namespace umma {

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

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};\n" ::"r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(taddr)
               : "memory");
}

// Column sums over the 32 lanes of 8 per-lane values: 3 halving levels + 2 plain ones; lane l ends with column l >> 2.
__device__ __forceinline__ float colsum8(const float (&v)[8], int lane) {
  float a[4], b[2];
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = (h4 ? v[i + 4] : v[i]) + __shfl_xor_sync(0xffffffffu, h4 ? v[i] : v[i + 4], 16);
#pragma unroll
  for (int i = 0; i < 2; ++i) b[i] = (h3 ? a[i + 2] : a[i]) + __shfl_xor_sync(0xffffffffu, h3 ? a[i] : a[i + 2], 8);
  float c = (h2 ? b[1] : b[0]) + __shfl_xor_sync(0xffffffffu, h2 ? b[0] : b[1], 4);
  c += __shfl_xor_sync(0xffffffffu, c, 2);
  c += __shfl_xor_sync(0xffffffffu, c, 1);
  return c;
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

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100a): one issue slot for two IEEE-rn fp32 operations.  The
// update kernels are bound by instruction issue, not by the FMA pipe, so the per-feature elementwise work runs on
// feature PAIRS held in aligned 64-bit registers; results are bit-identical to the scalar fmaf / * / + forms.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ f32x2 pack2u(uint32_t a, uint32_t b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void unpack2u(f32x2 v, uint32_t& a, uint32_t& b) { asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 dup2(float a) { return pack2(a, a); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// tanh of a pair given y = x * 2 / ln 2:  t = 1 - 2 / (2^y + 1)  and  -t = 2 / (2^y + 1) - 1  (both exact mirror images)
__device__ __forceinline__ void tanh2_scaled(f32x2 y, f32x2& t, f32x2& nt) {
  float y0, y1, e0, e1, r0, r1;
  unpack2(y, y0, y1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(y0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(y1));
  float d0, d1;
  unpack2(add2(pack2(e0, e1), dup2(1.f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  const f32x2 r = pack2(r0, r1);
  t = fma2(r, dup2(-2.f), dup2(1.f));
  nt = fma2(r, dup2(2.f), dup2(-1.f));
}
// 3xTF32 split of a pair: hi = the tf32-representable head, lo = x - hi (exact)
__device__ __forceinline__ void split2(f32x2 x, f32x2& hi, f32x2& lo) {
  uint32_t a, b;
  unpack2u(x, a, b);
  hi = pack2u(a & 0xffffe000u, b & 0xffffe000u);
  lo = fma2(hi, dup2(-1.f), x);
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

}  // namespace umma
}  // namespace rmab
