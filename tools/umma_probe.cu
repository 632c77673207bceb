// Probe: pins the tcgen05 (UMMA) encodings the h=64 update kernel relies on, one variant per launch, against a CPU product.
//   mode 0: D[128x64] = A[128x64] . B[64x64]^T      A, B K-major, no swizzle            (forward GEMM pattern)
//   mode 1: D[128x64] = A[128x64] . W[64x64]        B = W read MN-major from the SAME un-swizzled tile  -- NOT a tf32 layout:
//   mode 2: D[ 64x64] = X[128x64]^T . Y[128x64]     A = X, B = Y both MN-major, un-swizzled            -- expected to fail; kept
//           as the record of why the kernel keeps a transposed copy of W2 and uses mode 5's layout for the a1 / d2 tiles
//   mode 3: 3xTF32 split accuracy of mode 0 on random fp32 data against an fp64 product
//   mode 4: as mode 0 with A read from TMEM (written with tcgen05.st, lane = row, column = k)
//   mode 5: as mode 2 with both tiles in the 128B/32B-atom swizzled MN-major layout (layout type 1), the only MN-major
//           layout tf32 operands have: addr(r, f) = (f/32)*BLK + (r/4)*512 + (r%4)*128 + (((f%32)/8) ^ (r%4))*32 + (f%8)*4
//   mode 6: hidden-16 weight-gradient pattern: 128 x 32 tiles (one 128-byte row per tile row, features 0..15 = hi, 16..31 = lo),
//           X tile followed by the Y tile; A = X with M = 64 (its second 32-feature block, one LBO further, is the Y tile),
//           B = Y with N = 32: D[m][n] = sum_r X[r][m] Y[r][n] (m < 32), sum_r Y[r][m-32] Y[r][n] (m >= 32) -- all four
//           hi/lo cross products of d2^T a1 in one pass of 16 MMAs
//   mode 7: hidden-16 forward pattern: A [128 x 16] from TMEM, B [16 x 16] K-major with 512-byte row groups, N = 16
// Tiles live in shared memory as 8-row x 16-byte core matrices: addr(r, f) = (r/8)*LDG + (f/4)*128 + (r%8)*16 + (f%4)*4.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe tools/umma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
  uint64_t d = (uint64_t)layout_type << 61;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;                // base offset 0, layout type 0 = no swizzle
}

__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

constexpr int BLK = 16384;  // bytes per 32-feature block of a 128-row tile in the swizzled MN-major layout
__device__ __forceinline__ int tile_off_mn(int r, int f) {
  return ((f >> 5) * BLK + (r >> 2) * 512 + (r & 3) * 128 + ((((f & 31) >> 3) ^ (r & 3)) * 32) + (f & 7) * 4) >> 2;
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (int spin = 0; spin < (1 << 24); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return true;
  }
  return false;
}

constexpr int LDG = 2048;  // bytes between 8-row groups (16 core matrices of 128 B)
__device__ __forceinline__ int tile_off(int r, int f) { return (r >> 3) * (LDG / 4) + (f >> 2) * 32 + (r & 7) * 4 + (f & 3); }

__device__ __forceinline__ void split(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}

// variant bit 0: swap LBO/SBO of the MN-major descriptors
__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ gA, const float* __restrict__ gB, float* __restrict__ gD, int mode,
                                                       int variant, int* status) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* tA = reinterpret_cast<float*>(smem_raw);           // 128 x 64 tile (hi)
  float* tAl = tA + 128 * 64;                               // lo
  float* tB = tAl + 128 * 64;                               // 128 x 64 (modes 0/1/3 use 64 rows)
  float* tBl = tB + 128 * 64;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rowsB = (mode == 2 || mode == 5) ? 128 : 64;
  if (mode == 6) {   // packed 32-wide swizzled tiles: X at tA, Y right behind it (16 KB each)
    for (int i = tid; i < 128 * 32; i += 128) {
      const int r = i >> 5, f = i & 31;
      const int o = ((r >> 2) * 512 + (r & 3) * 128 + (((f >> 3) ^ (r & 3)) * 32) + (f & 7) * 4) >> 2;
      tA[o] = gA[r * 64 + f];
      tA[4096 + o] = gB[r * 64 + f];
    }
  }
  if (mode == 7) {   // B [16 n][16 k] K-major, 8-row groups 512 B apart
    for (int i = tid; i < 16 * 16; i += 128) {
      const int r = i >> 4, f = i & 15;
      tB[(r >> 3) * 128 + (f >> 2) * 32 + (r & 7) * 4 + (f & 3)] = gB[r * 64 + f];
    }
  }
  for (int i = tid; i < 128 * 64 && mode < 6; i += 128) {
    const int r = i >> 6, f = i & 63;
    float hi, lo;
    if (mode == 3) split(gA[i], hi, lo); else { hi = gA[i]; lo = 0.f; }
    const int o = (mode == 5) ? tile_off_mn(r, f) : tile_off(r, f);
    tA[o] = hi; tAl[o] = lo;
  }
  for (int i = tid; i < rowsB * 64 && mode < 6; i += 128) {
    const int r = i >> 6, f = i & 63;
    float hi, lo;
    if (mode == 3) split(gB[i], hi, lo); else { hi = gB[i]; lo = 0.f; }
    const int o = (mode == 5) ? tile_off_mn(r, f) : tile_off(r, f);
    tB[o] = hi; tBl[o] = lo;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;\n" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (mode == 4 || mode == 7) {
    __syncthreads();  // tmem_base_s
    for (int c0 = 0; c0 < (mode == 7 ? 16 : 64); c0 += 16) {
      uint32_t v[16];
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(gA[tid * 64 + c0 + j]);
      const uint32_t taddr = tmem_base_s + ((uint32_t)(warp * 32) << 16) + 64 + c0;
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};\n" ::"r"(v[0]),
          "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
          "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr)
          : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy tile writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (tid == 0) {
    const uint32_t a0 = smem_u32(tA), al = smem_u32(tAl), b0 = smem_u32(tB), bl = smem_u32(tBl);
    const uint32_t K_LBO = 128, K_SBO = LDG;                                  // K-major: LBO along K (next 4 elements), SBO along M/N (next 8 rows)
    const uint32_t MN_LBO = (variant & 1) ? 128 : LDG, MN_SBO = (variant & 1) ? LDG : 128;  // MN-major: SBO along M/N (next 4 elements), LBO along K (next 8 rows)
    if (mode == 0 || mode == 3) {
      const uint32_t idesc = make_idesc(128, 64, 0, 0);
      const int terms = (mode == 3) ? 3 : 1;
      uint32_t acc = 0;
      for (int t = 0; t < terms; ++t) {
        const uint32_t as = (t == 1) ? al : a0, bs = (t == 2) ? bl : b0;   // lo.hi, hi.lo, hi.hi  (order: t=0 hi.hi, 1 lo.hi, 2 hi.lo)
        for (int ks = 0; ks < 8; ++ks) {
          umma_tf32(tmem, make_desc(as + ks * 256, K_LBO, K_SBO), make_desc(bs + ks * 256, K_LBO, K_SBO), idesc, acc);
          acc = 1;
        }
      }
    } else if (mode == 1) {
      const uint32_t idesc = make_idesc(128, 64, 0, 1);
      for (int ks = 0; ks < 8; ++ks)
        umma_tf32(tmem, make_desc(a0 + ks * 256, K_LBO, K_SBO), make_desc(b0 + ks * LDG, MN_LBO, MN_SBO), idesc, ks > 0);
    } else if (mode == 2) {
      const uint32_t idesc = make_idesc(64, 64, 1, 1);
      for (int ks = 0; ks < 16; ++ks)
        umma_tf32(tmem, make_desc(a0 + ks * LDG, MN_LBO, MN_SBO), make_desc(b0 + ks * LDG, MN_LBO, MN_SBO), idesc, ks > 0);
    } else if (mode == 4) {
      const uint32_t idesc = make_idesc(128, 64, 0, 0);
      for (int ks = 0; ks < 8; ++ks) umma_tf32_ts(tmem, tmem + 64 + ks * 8, make_desc(b0 + ks * 256, K_LBO, K_SBO), idesc, ks > 0);
    } else if (mode == 6) {
      const uint32_t idesc = make_idesc(64, 32, 1, 1);
      for (int ks = 0; ks < 16; ++ks)
        umma_tf32(tmem, make_desc(a0 + ks * 1024, 16384, 512, 1), make_desc(a0 + 16384 + ks * 1024, 16384, 512, 1), idesc, ks > 0);
    } else if (mode == 7) {
      const uint32_t idesc = make_idesc(128, 16, 0, 0);
      for (int ks = 0; ks < 2; ++ks) umma_tf32_ts(tmem, tmem + 64 + ks * 8, make_desc(b0 + ks * 256, 128, 512), idesc, ks > 0);
    } else {
      const uint32_t idesc = make_idesc(64, 64, 1, 1);
      const uint32_t lbo = (variant & 1) ? 512 : BLK, sbo = (variant & 1) ? BLK : 512;   // expected: LBO = next 32 features, SBO = next 4 rows
      for (int ks = 0; ks < 16; ++ks)
        umma_tf32(tmem, make_desc(a0 + ks * 1024, lbo, sbo, 1), make_desc(b0 + ks * 1024, lbo, sbo, 1), idesc, ks > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
  }
  const bool ok = mbar_wait(smem_u32(&bar), 0);
  if (!ok && tid == 0) *status = 1;
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (ok) {
    // 32x32b: thread = TMEM lane 32*warp + lane, 16 consecutive columns per load
    for (int c0 = 0; c0 < 64; c0 += 16) {
      uint32_t v[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
            "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
      for (int j = 0; j < 16; ++j) gD[tid * 64 + c0 + j] = __uint_as_float(v[j]);  // raw dump: lane-major [128][64]
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;\n" ::"r"(tmem) : "memory");
}

static float q(float x) { return roundf(x * 8.f) / 8.f; }  // exactly representable in tf32

int main() {
  const int smem = 4 * 128 * 64 * 4;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> A(128 * 64), B(128 * 64), D(128 * 64);
  float *dA, *dB, *dD; int* dS;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dS, 4));
  int fails = 0;
  for (int mode = 0; mode < 8; ++mode) {
    for (int variant = 0; variant < ((mode == 1 || mode == 2 || mode == 5) ? 2 : 1); ++variant) {
      srand(1234 + mode);
      for (auto& x : A) x = (mode == 3) ? (rand() / (float)RAND_MAX * 2.f - 1.f) : q(rand() / (float)RAND_MAX * 4.f - 2.f);
      for (auto& x : B) x = (mode == 3) ? (rand() / (float)RAND_MAX * 2.f - 1.f) : q(rand() / (float)RAND_MAX * 4.f - 2.f);
      CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(dD, 0xff, D.size() * 4)); CK(cudaMemset(dS, 0, 4));
      probe_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, variant, dS);
      CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      int st; CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0, maxref = 0;
      const int M = (mode == 2 || mode == 5 || mode == 6) ? 64 : 128;
      const int NN = mode == 6 ? 32 : (mode == 7 ? 16 : 64);
      for (int m = 0; m < M; ++m)
        for (int n = 0; n < NN; ++n) {
          double ref = 0;
          if (mode == 6) for (int r = 0; r < 128; ++r) ref += (double)(m < 32 ? A[r * 64 + m] : B[r * 64 + m - 32]) * B[r * 64 + n];
          if (mode == 7) for (int k = 0; k < 16; ++k) ref += (double)A[m * 64 + k] * B[n * 64 + k];
          if (mode == 0 || mode == 3 || mode == 4) for (int k = 0; k < 64; ++k) ref += (double)A[m * 64 + k] * B[n * 64 + k];
          else if (mode == 1) for (int k = 0; k < 64; ++k) ref += (double)A[m * 64 + k] * B[k * 64 + n];
          else if (mode == 2 || mode == 5) for (int r = 0; r < 128; ++r) ref += (double)A[r * 64 + m] * B[r * 64 + n];
          const int lane = (mode == 2 || mode == 5 || mode == 6) ? ((m & 15) + 32 * (m >> 4)) : m;   // M = 64: 16 rows per 32-lane quarter
          const double got = D[lane * 64 + n];
          maxerr = fmax(maxerr, fabs(got - ref)); maxref = fmax(maxref, fabs(ref));
        }
      const double tol = (mode == 3) ? 2e-6 * maxref : 1e-6;
      const bool pass = (st == 0) && (maxerr <= tol);
      const bool expected_fail = mode == 1 || mode == 2 || variant == 1;
      printf("mode %d variant %d: timeout=%d max|err|=%.3g (max|ref| %.3g) %s\n", mode, variant, st, maxerr, maxref,
             pass ? "PASS" : (expected_fail ? "fail (expected: not a valid tf32 operand layout)" : "FAIL"));
      if (!pass && !expected_fail) ++fails;
    }
  }
  printf("probe done, unexpected failures %d\n", fails);
  return fails ? 1 : 0;
}
