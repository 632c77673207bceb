// Dense GEMM of the MA-RMABPPO agent critics' nature block, fp32 accuracy on the fp16 tensor-core path.
//
// Reference behaviour: every agent critic is MLP([obs_i, lambda, a_nature] -> h -> h -> 1) (ma_rmabppo_core.py:203), so
// the first layer of all N critics over the D nature inputs is one dense product per batch,
//     z1[m, i*h + j] = sum_d nat[m, d] * W1n[i, j, d]              (forward, nature_oracle.py:388-416)
//     dW1n[i, j, d]  = sum_m dz1[m, i*h + j] * nat[m, d]           (its gradient, :505-513)
// with m = (t, env) over the T*E samples of an epoch: [5120, 8192] x [8192, 131072] at BASELINE configs[3].
//
// Arithmetic: x = s^-1 (hi + lo) with s a per-row power of two that puts the row's largest magnitude in [2^14, 2^15),
// hi = fp16(s x), lo = fp16(s x - hi): 22 significant bits, like the 3xTF32 split, but on kind::f16 MMAs (twice the
// tf32 rate).  a.b = s_a^-1 s_b^-1 (hi_a hi_b + lo_a hi_b + hi_a lo_b); products are exact in fp32, accumulation is
// fp32 in TMEM; the dropped lo.lo term is 2^-22 relative.
//
// Kernel: persistent, one CTA pair per SM pair (cta_group::2), output tile 256 x 256 per pair (each CTA: 128 rows of A,
// 128 rows of B, a 128 x 256 fp32 accumulator in its TMEM, double buffered = all 512 columns).  K is walked in blocks
// of 64 fp16 (128-byte swizzled rows); per block TMA brings A hi/lo and B hi/lo (64 KB per CTA) through a 3-stage
// mbarrier ring, the leader CTA's MMA warp issues 4 k-steps x 3 products of tcgen05.mma M=256 N=256 K=16.
// Warps: 0 = TMA producer, 1 = MMA issuer (leader CTA), 2..5 = epilogue (TMEM -> registers -> scale -> global).
#include <cuda.h>
#include <math.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "umma.cuh"

namespace rmab {
namespace g3 {

using umma::elect_one;
using umma::smem_u32;

constexpr int BM = 128;            // rows of A per CTA (UMMA M = 256 across the pair)
constexpr int BN = 256;            // columns of the output tile (UMMA N); each CTA stages 128 of the B rows
constexpr int BNH = BN / 2;
constexpr int BK = 64;             // fp16 elements per k-block = one 128-byte swizzled row
constexpr int STAGES = 3;
constexpr uint32_t TILE_BYTES = BM * BK * 2;            // 16 KB: one 128 x 64 fp16 operand tile
constexpr uint32_t STAGE_BYTES = 4 * TILE_BYTES;        // A hi, A lo, B hi, B lo
constexpr uint32_t SMEM_TILES = STAGES * STAGE_BYTES;   // 192 KB
constexpr uint32_t EPI_BYTES = 4 * 32 * 33 * 4;           // per-epilogue-warp transpose buffers
constexpr uint32_t SMEM_BYTES = SMEM_TILES + 1024 /*alignment slack*/ + 256 /*barriers*/ + EPI_BYTES;
constexpr int NTHREADS = 192;
constexpr int ACC_COLS = BN;       // fp32 accumulator columns per stage

// kind::f16: D fp32 (bit 4), A/B fp16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
// K-major, 128-byte swizzle: SBO = 1024 B (8 rows x 128 B), LBO unused, descriptor version 1, layout type 2
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Bounded wait with cluster-scope acquire: a wrong phase must fail the launch, not hang the GPU.
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity) {
#pragma unroll 1
  for (int spin = 0; spin < (1 << 28); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_f16_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 ad, {%1, %3};\n\tmov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], ad, bd, %4, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(IDESC), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_mc2(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

struct Shape {
  int M, N, K;
  int tiles_m, tiles_n;
  int vec_ok;     // C and rsB are 16-byte aligned and ldc % 4 == 0: float4 epilogue stores
  int n_fast;     // 1: consecutive tiles walk N (the A block stays), 0: they walk M (the B block stays)
  long long ldc;
  // fused Adam epilogue (adam != 0): the product is a gradient; C is the parameter tensor, updated in place with its
  // moments m, v (same layout) exactly as rmab_adam_step does -- the gradient never reaches HBM, and the 6 x |C| bytes of
  // optimiser traffic run under the next tile's main loop
  int adam;
  float* m; float* v;
  float lerp_w, beta2, omb2, step_size, bc2s, eps;
};

__device__ __forceinline__ float adam_elem(const Shape& sh, float g, float p, float& m, float& v) {
  return adam_update(p, g, m, v, sh.lerp_w, sh.beta2, sh.omb2, sh.step_size, sh.bc2s, sh.eps);
}

__device__ __forceinline__ void tile_coords(const Shape& sh, int t, int& mb, int& nb) {
  if (sh.n_fast) { mb = t / sh.tiles_n; nb = t - mb * sh.tiles_n; }
  else { nb = t / sh.tiles_m; mb = t - nb * sh.tiles_m; }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gemm3h_kernel(const __grid_constant__ CUtensorMap mAh, const __grid_constant__ CUtensorMap mAl,
              const __grid_constant__ CUtensorMap mBh, const __grid_constant__ CUtensorMap mBl, Shape sh,
              const float* __restrict__ rsA, const float* __restrict__ rsB, float* __restrict__ C) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + SMEM_TILES;
  // barrier slots (8 bytes each): full[3] empty[3] tmem_full[2] tmem_empty[2]; then the TMEM base address word
  const uint32_t bar_full = bars, bar_empty = bars + 24, bar_tfull = bars + 48, bar_tempty = bars + 64, tmem_slot = bars + 80;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int n_tiles = sh.tiles_m * sh.tiles_n;
  const int KB = (sh.K + BK - 1) / BK;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 8);   // 4 epilogue warps x 2 CTAs arrive on the leader's barrier
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mAh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mAl)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mBh)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&mBl)) : "memory");
  }
  umma::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  umma::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer: this CTA's 128 rows of A and 128 rows of B, hi and lo, signalled on the LEADER's barrier =====
    if (elect_one()) {
      const uint32_t full0 = mapa_rank(bar_full, 0);
      int s = 0;
      uint32_t ph = 0;
      for (int t = pair; t < n_tiles; t += n_pairs) {
        int mb, nb;
        tile_coords(sh, t, mb, nb);
        const int row_a = mb * (2 * BM) + (int)rank * BM, row_b = nb * BN + (int)rank * BNH;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_cl(bar_empty + 8 * s, ph ^ 1u);
          if (leader) mbar_expect_tx(bar_full + 8 * s, 2 * STAGE_BYTES);
          const uint32_t dst = base + s * STAGE_BYTES, fb = full0 + 8 * s;
          tma_load_2sm(dst, &mAh, fb, kb * BK, row_a);
          tma_load_2sm(dst + TILE_BYTES, &mAl, fb, kb * BK, row_a);
          tma_load_2sm(dst + 2 * TILE_BYTES, &mBh, fb, kb * BK, row_b);
          tma_load_2sm(dst + 3 * TILE_BYTES, &mBl, fb, kb * BK, row_b);
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA): D[acc stage] += Ah Bh + Al Bh + Ah Bl over the k-blocks of a tile =====
    if (leader) {
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
        const int acc = it & 1;
        mbar_wait_cl(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1u);
        umma::tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * ACC_COLS);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_cl(bar_full + 8 * s, ph);
          umma::tc_fence_after();
          if (elect_one()) {
            const uint32_t a_h = (base + s * STAGE_BYTES) >> 4, a_l = a_h + (TILE_BYTES >> 4);
            const uint32_t b_h = a_h + (2 * TILE_BYTES >> 4), b_l = a_h + (3 * TILE_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t o = (uint32_t)(k * 32) >> 4;      // 16 fp16 = 32 bytes along the swizzled row
              mma_f16_2sm(d, a_h + o, b_h + o, (kb | k) != 0 ? 1u : 0u);
              mma_f16_2sm(d, a_l + o, b_h + o, 1u);
              mma_f16_2sm(d, a_h + o, b_l + o, 1u);
            }
            commit_mc2(bar_empty + 8 * s);                      // frees the stage in both CTAs once these MMAs retire
            if (kb == KB - 1) commit_mc2(bar_tfull + 8 * acc);  // accumulator complete: both CTAs' epilogues may read
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    // ===== epilogue: TMEM lane = output row; 16 columns per tcgen05.ld; scale and store fp32 =====
    const int q = warp & 3;                                     // TMEM lane quarter this warp may access
    float* es = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + SMEM_TILES + 256) + (warp - 2) * 32 * 33;
    const uint32_t tempty0 = mapa_rank(bar_tempty, 0);
    int it = 0;
    for (int t = pair; t < n_tiles; t += n_pairs, ++it) {
      int mb, nb;
      tile_coords(sh, t, mb, nb);
      const int acc = it & 1;
      mbar_wait_cl(bar_tfull + 8 * acc, (it >> 1) & 1);
      umma::tc_fence_after();
      const int m_base = mb * (2 * BM) + (int)rank * BM + q * 32;
      const int n0 = nb * BN;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS);
      // TMEM gives a thread 32 consecutive columns of ONE row; global rows are ldc floats apart, so the chunk goes through a
      // per-warp shared-memory transpose and leaves as 4 rows x 128 contiguous bytes per warp instruction
      const int col4 = (lane & 7) * 4, rsub = lane >> 3;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v0[16], v1[16];
        umma::tmem_ld16(taddr + c * 32, v0);
        umma::tmem_ld16(taddr + c * 32 + 16, v1);
        umma::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          es[lane * 33 + j] = __uint_as_float(v0[j]);
          es[lane * 33 + 16 + j] = __uint_as_float(v1[j]);
        }
        __syncwarp();
        const int nn = n0 + c * 32 + col4;
        const bool vec = sh.vec_ok && nn + 3 < sh.N;
        float sb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) sb[i] = (rsB && nn + i < sh.N) ? rsB[nn + i] : 1.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int row = k * 4 + rsub, mm = m_base + row;
          if (mm >= sh.M) continue;
          const float sa = rsA ? rsA[mm] : 1.f;
          float g[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) g[i] = es[row * 33 + col4 + i] * (sa * sb[i]);
          const long long off = (long long)mm * sh.ldc + nn;
          if (vec) {
            if (sh.adam) {
              float4 pm = *reinterpret_cast<const float4*>(sh.m + off), pv = *reinterpret_cast<const float4*>(sh.v + off);
              const float4 pp = *reinterpret_cast<const float4*>(C + off);
              g[0] = adam_elem(sh, g[0], pp.x, pm.x, pv.x);
              g[1] = adam_elem(sh, g[1], pp.y, pm.y, pv.y);
              g[2] = adam_elem(sh, g[2], pp.z, pm.z, pv.z);
              g[3] = adam_elem(sh, g[3], pp.w, pm.w, pv.w);
              *reinterpret_cast<float4*>(sh.m + off) = pm;
              *reinterpret_cast<float4*>(sh.v + off) = pv;
            }
            *reinterpret_cast<float4*>(C + off) = make_float4(g[0], g[1], g[2], g[3]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (nn + i < sh.N) {
                float x = g[i];
                if (sh.adam) x = adam_elem(sh, x, C[off + i], sh.m[off + i], sh.v[off + i]);
                C[off + i] = x;
              }
          }
        }
        __syncwarp();
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc);
    }
  }
  __syncwarp();
  // teardown: the peer's shared memory and barriers must outlive the leader's last MMA / the last remote arrive
  umma::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ splits
__device__ __forceinline__ float row_scale(float amax) {
  // power of two s with s * amax in [2^14, 2^15); 1 for an all-zero row
  if (!(amax > 0.f)) return 1.f;
  int ex;
  frexpf(amax, &ex);                    // amax = f * 2^ex, f in [0.5, 1)
  int sh = 15 - ex;
  sh = sh > 100 ? 100 : (sh < -100 ? -100 : sh);
  return ldexpf(1.f, sh);
}

__device__ __forceinline__ void split_h(float x, float s, __half& hi, __half& lo) {
  const float y = x * s;                // exact (power of two) unless it underflows
  hi = __float2half_rn(y);
  lo = __float2half_rn(y - __half2float(hi));
}

// X [R, K] (row stride ldx) -> Xh, Xl [R, ldk] fp16, rs[R] = 1 / scale.  One CTA per row.
__global__ void __launch_bounds__(256) split16_rows_kernel(int K, const float* __restrict__ X, long long ldx, int ldk,
                                                           __half* __restrict__ Xh, __half* __restrict__ Xl,
                                                           float* __restrict__ rs) {
  __shared__ float red[32];
  const long long r = blockIdx.x;
  const float* x = X + r * ldx;
  float amax = 0.f;
  const bool vec = ((ldx & 3) == 0) && ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (vec) {
    for (int k = threadIdx.x * 4; k < K; k += 1024) {
      const float4 v = *reinterpret_cast<const float4*>(x + k);
      amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  } else {
    for (int k = threadIdx.x; k < K; k += 256) amax = fmaxf(amax, fabsf(x[k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  amax = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) amax = fmaxf(amax, red[i]);
  const float s = row_scale(amax);
  if (threadIdx.x == 0) rs[r] = 1.f / s;
  __half* oh = Xh + r * ldk;
  __half* ol = Xl + r * ldk;
  if (vec && (ldk & 3) == 0) {
    for (int k = threadIdx.x * 4; k < K; k += 1024) {
      const float4 v = *reinterpret_cast<const float4*>(x + k);
      __half h[4], l[4];
      split_h(v.x, s, h[0], l[0]);
      split_h(v.y, s, h[1], l[1]);
      split_h(v.z, s, h[2], l[2]);
      split_h(v.w, s, h[3], l[3]);
      *reinterpret_cast<uint2*>(oh + k) = *reinterpret_cast<uint2*>(h);
      *reinterpret_cast<uint2*>(ol + k) = *reinterpret_cast<uint2*>(l);
    }
  } else {
    for (int k = threadIdx.x; k < K; k += 256) split_h(x[k], s, oh[k], ol[k]);
  }
  for (int k = K + threadIdx.x; k < ldk; k += 256) { oh[k] = __float2half(0.f); ol[k] = __float2half(0.f); }
}

// column maxima of X [Rk, Cn]: amax[c] (as uint bits of a non-negative float), zero-initialised by the caller
__global__ void __launch_bounds__(256) colmax_kernel(int Rk, int Cn, const float* __restrict__ X, long long ldx,
                                                     int rows_per_cta, unsigned int* __restrict__ amax) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= Cn) return;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(Rk, r0 + rows_per_cta);
  float m = 0.f;
  for (int r = r0; r < r1; ++r) m = fmaxf(m, fabsf(X[(long long)r * ldx + c]));
  atomicMax(amax + c, __float_as_uint(m));
}

// X [Rk, Cn] -> transposed split Xh, Xl [Cn, ldk] (k = the row index of X), rs[c] = 1 / scale of column c.
__global__ void __launch_bounds__(256) split16_cols_kernel(int Rk, int Cn, const float* __restrict__ X, long long ldx,
                                                           int ldk, const unsigned int* __restrict__ amax,
                                                           __half* __restrict__ Xh, __half* __restrict__ Xl,
                                                           float* __restrict__ rs) {
  __shared__ float tile[64][65];
  const int c0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
  for (int r = ty; r < 64; r += 4) {
    const int k = k0 + r, c = c0 + tx;
    tile[r][tx] = (k < Rk && c < Cn) ? X[(long long)k * ldx + c] : 0.f;
  }
  __syncthreads();
  // thread -> (column c = c0 + cc, 8 consecutive k): 64 columns x 8 k-groups = 512 items over 256 threads
  for (int item = threadIdx.x; item < 512; item += 256) {
    const int cc = item >> 3, kg = (item & 7) * 8;
    const int c = c0 + cc;
    if (c >= Cn) continue;
    const float s = row_scale(__uint_as_float(amax[c]));
    if (blockIdx.y == 0 && kg == 0) rs[c] = 1.f / s;
    __half h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_h(tile[kg + i][cc], s, h[i], l[i]);
    const int k = k0 + kg;
    if (k + 8 <= ldk) {
      *reinterpret_cast<uint4*>(Xh + (long long)c * ldk + k) = *reinterpret_cast<uint4*>(h);
      *reinterpret_cast<uint4*>(Xl + (long long)c * ldk + k) = *reinterpret_cast<uint4*>(l);
    } else {
      for (int i = 0; i < 8 && k + i < ldk; ++i) { Xh[(long long)c * ldk + k + i] = h[i]; Xl[(long long)c * ldk + k + i] = l[i]; }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, K] fp16, row stride ldk elements, box = 64 (k) x 128 (rows), 128-byte swizzle, zero fill outside
static bool make_map(CUtensorMap* m, const void* ptr, int rows, int K, int ldk) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ldk * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace g3
}  // namespace rmab

using namespace rmab;

extern "C" int rmab_split16_rows(int R, int K, const float* X, int64_t ldx, int ldk, void* Xh, void* Xl, float* rs,
                                 rmab_stream_t stream) {
  RMAB_REQUIRE(R >= 0 && K > 0 && X && Xh && Xl && rs && ldk >= K && (ldk % 8) == 0 && ldx >= K, RMAB_E_BADARG,
               "rmab_split16_rows: bad argument (ldk must be a multiple of 8 and >= K)");
  if (R == 0) return RMAB_OK;
  g3::split16_rows_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(K, X, (long long)ldx, ldk, static_cast<__half*>(Xh),
                                                               static_cast<__half*>(Xl), rs);
  return check_launch("rmab_split16_rows");
}

extern "C" int rmab_split16_cols(int Rk, int Cn, const float* X, int64_t ldx, int ldk, void* Xh, void* Xl, float* rs,
                                 uint32_t* amax_ws, rmab_stream_t stream) {
  RMAB_REQUIRE(Rk > 0 && Cn >= 0 && X && Xh && Xl && rs && amax_ws && ldk >= Rk && (ldk % 8) == 0 && ldx >= Cn, RMAB_E_BADARG,
               "rmab_split16_cols: bad argument (ldk must be a multiple of 8 and >= Rk)");
  if (Cn == 0) return RMAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(amax_ws, 0, (size_t)Cn * 4, st);
  const int gy = Rk >= 2048 ? 16 : (Rk >= 256 ? 4 : 1);
  const int rows_per = (Rk + gy - 1) / gy;
  g3::colmax_kernel<<<dim3((Cn + 255) / 256, gy), 256, 0, st>>>(Rk, Cn, X, (long long)ldx, rows_per, amax_ws);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g3::split16_cols_kernel<<<dim3((Cn + 63) / 64, (ldk + 63) / 64), 256, 0, st>>>(Rk, Cn, X, (long long)ldx, ldk, amax_ws,
                                                                                static_cast<__half*>(Xh),
                                                                                static_cast<__half*>(Xl), rs);
  return check_launch("rmab_split16_cols");
}

static int gemm3h_launch(int M, int N, int K, int ldk, const void* Ah, const void* Al, const float* rsA, const void* Bh,
                         const void* Bl, const float* rsB, float* C, int64_t ldc, int n_fast, const g3::Shape* adam,
                         rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && N > 0 && K > 0 && Ah && Al && Bh && Bl && C && ldk >= K && (ldk % 8) == 0 && ldc >= N, RMAB_E_BADARG,
               "rmab_gemm3h: bad argument (ldk must be a multiple of 8 and >= K)");
  RMAB_REQUIRE(((uintptr_t)Ah % 16) == 0 && ((uintptr_t)Al % 16) == 0 && ((uintptr_t)Bh % 16) == 0 && ((uintptr_t)Bl % 16) == 0,
               RMAB_E_BADARG, "rmab_gemm3h: operand pointers must be 16-byte aligned");
  CUtensorMap mAh, mAl, mBh, mBl;
  RMAB_REQUIRE(g3::make_map(&mAh, Ah, M, K, ldk) && g3::make_map(&mAl, Al, M, K, ldk) && g3::make_map(&mBh, Bh, N, K, ldk) &&
                   g3::make_map(&mBl, Bl, N, K, ldk),
               RMAB_E_LAUNCH, "rmab_gemm3h: cuTensorMapEncodeTiled failed");
  static bool attr_set = false;
  if (!attr_set) {
    RMAB_REQUIRE(cudaFuncSetAttribute(g3::gemm3h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g3::SMEM_BYTES) ==
                     cudaSuccess,
                 RMAB_E_LAUNCH, "rmab_gemm3h: cannot reserve %u B of shared memory", g3::SMEM_BYTES);
    attr_set = true;
  }
  g3::Shape sh = {};
  if (adam) sh = *adam;
  sh.M = M; sh.N = N; sh.K = K;
  sh.tiles_m = (M + 2 * g3::BM - 1) / (2 * g3::BM);
  sh.tiles_n = (N + g3::BN - 1) / g3::BN;
  sh.n_fast = n_fast ? 1 : 0;
  sh.ldc = (long long)ldc;
  sh.vec_ok = ((ldc & 3) == 0 && ((uintptr_t)C % 16) == 0 && (!rsB || ((uintptr_t)rsB % 16) == 0) &&
               (!adam || (((uintptr_t)sh.m % 16) == 0 && ((uintptr_t)sh.v % 16) == 0))) ? 1 : 0;
  const int tiles = sh.tiles_m * sh.tiles_n;
  const int pairs = tiles < kNumSMs / 2 ? tiles : kNumSMs / 2;
  g3::gemm3h_kernel<<<2 * pairs, g3::NTHREADS, g3::SMEM_BYTES, (cudaStream_t)stream>>>(mAh, mAl, mBh, mBl, sh, rsA, rsB, C);
  return check_launch("rmab_gemm3h");
}

extern "C" int rmab_gemm3h(int M, int N, int K, int ldk, const void* Ah, const void* Al, const float* rsA, const void* Bh,
                           const void* Bl, const float* rsB, float* C, int64_t ldc, int n_fast, rmab_stream_t stream) {
  return gemm3h_launch(M, N, K, ldk, Ah, Al, rsA, Bh, Bl, rsB, C, ldc, n_fast, nullptr, stream);
}

extern "C" int rmab_gemm3h_adam(int M, int N, int K, int ldk, const void* Ah, const void* Al, const float* rsA, const void* Bh,
                                const void* Bl, const float* rsB, float* P, float* Mo, float* Vo, int64_t ldc, double lr,
                                double beta1, double beta2, double eps, int step, int n_fast, rmab_stream_t stream) {
  RMAB_REQUIRE(P && Mo && Vo && step >= 1, RMAB_E_BADARG, "rmab_gemm3h_adam: bad argument");
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  g3::Shape a = {};
  a.adam = 1; a.m = Mo; a.v = Vo;
  a.lerp_w = (float)(1.0 - beta1); a.beta2 = (float)beta2; a.omb2 = (float)(1.0 - beta2);
  a.step_size = (float)(lr / bc1); a.bc2s = (float)sqrt(bc2); a.eps = (float)eps;
  return gemm3h_launch(M, N, K, ldk, Ah, Al, rsA, Bh, Bl, rsB, P, ldc, n_fast, &a, stream);
}
