// The dense nature nets of MA-RMABPPO on the device: the Gaussian nature actor N -> h -> h -> D and the nature critic
// 2N -> h -> h -> 1, forward, backward and the per-sample loss terms, as tiled fp32 kernels.
//
// Reference behaviour: ma_rmabppo_core.py MLPGaussianActor :84-113 / MLPCritic :116-123, step :274-331 (nature action
// from the state vector, nature value from [obs, a_agent]); nature_oracle.py compute_loss_pi_nature :359-386 (PPO clip
// on the summed Normal log-prob), compute_loss_v_nature :421-428 (MSE), update :558-576 (Adam steps).
//
// Layout: parameters packed as [W1T (Nin x H) | b1 | W2 (H x H, torch order [out][in]) | b2 | W3 (OUT x H) | b3 | log_std];
// W1 is kept TRANSPOSED so that the rows of one arm -- and therefore one rank's arm shard -- are contiguous, as are the D/N
// output rows of W3 / b3 / log_std that belong to an arm.  Inputs are never materialised as floats: layer 1 reads the u8
// state (and action) trajectories [n][t][e] directly; sample m = t * E + e.  Under arm sharding layer 1 is a
// partial product over the rank's arms (caller all-reduces [M, H]) and layer 3 covers the rank's output rows.
// Every reduction is two-stage in a fixed order (partials per chunk, then a sum kernel): results are deterministic.
#include <math.h>
#include "common.cuh"

namespace rmab {
namespace nat {

constexpr int TS = 64;        // tile edge (samples, columns, arms)
constexpr int PITCH = 68;     // smem row pitch in floats: float4-aligned, staggered banks
constexpr int NT = 256;

__global__ void __launch_bounds__(NT) sum_parts_kernel(int64_t count, int nparts, const float* __restrict__ parts,
                                                       float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * NT;
  for (int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x; i < count; i += stride) {
    float s = parts[i];
    for (int p = 1; p < nparts; ++p) s += parts[(int64_t)p * count + i];
    out[i] = s;
  }
}

// ---- layer 1 forward: part[y][m][j] = sum_{n in chunk y} sc * x0[n, m] * W0[n][j] (+ x1[n, m] * W1[n][j]) ----
__global__ void __launch_bounds__(NT) l1_fwd_kernel(int M, int E, int H, int n_loc, int per_chunk, const uint8_t* __restrict__ x0,
                                                    int64_t stride0, float sc0, const float* __restrict__ W0,
                                                    const uint8_t* __restrict__ x1, int64_t stride1,
                                                    const float* __restrict__ W1, float* __restrict__ part) {
  __shared__ __align__(16) float xs[TS][PITCH];   // [n][m]
  __shared__ __align__(16) float ws[TS][PITCH];   // [n][j]
  const int m0 = blockIdx.x * TS, y = blockIdx.y;
  const int tm = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const int nb = y * per_chunk, ne = min(n_loc, nb + per_chunk);
  float acc[4][4] = {};
  for (int src = 0; src < 2; ++src) {
    const uint8_t* x = src == 0 ? x0 : x1;
    if (!x) continue;
    const float* W = src == 0 ? W0 : W1;
    const int64_t strd = src == 0 ? stride0 : stride1;
    const float sc = src == 0 ? sc0 : 1.f;
    for (int n0 = nb; n0 < ne; n0 += TS) {
      __syncthreads();
      for (int idx = threadIdx.x; idx < TS * TS; idx += NT) {
        const int n = idx >> 6, mm = idx & 63, m = m0 + mm;
        float v = 0.f;
        if (n0 + n < ne && m < M) {
          const int t = m / E, e = m - t * E;
          v = sc * (float)x[(int64_t)(n0 + n) * strd + (int64_t)t * E + e];
        }
        xs[n][mm] = v;
      }
      for (int idx = threadIdx.x; idx < TS * TS; idx += NT) {
        const int n = idx >> 6, j = idx & 63;
        ws[n][j] = (n0 + n < ne && j < H) ? W[(int64_t)(n0 + n) * H + j] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int n = 0; n < TS; ++n) {
        const float4 a = *reinterpret_cast<const float4*>(&xs[n][4 * tm]);
        const float4 b = *reinterpret_cast<const float4*>(&ws[n][4 * tj]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }
  float* out = part + (int64_t)y * M * H;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + 4 * tm + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * tj + j < H) out[(int64_t)m * H + 4 * tj + j] = acc[i][j];
  }
}

// ---- layers 1 (bias, tanh) and 2: a1 = tanh(z1 + b1), a2 = tanh(W2 a1 + b2) ----
__global__ void __launch_bounds__(NT) mid_fwd_kernel(int M, int H, const float* __restrict__ z1, const float* __restrict__ b1,
                                                     const float* __restrict__ W2, const float* __restrict__ b2,
                                                     float* __restrict__ a1, float* __restrict__ a2) {
  extern __shared__ float sm[];
  float* w2t = sm;               // [k][j]
  float* a1s = sm + H * H;       // [samples per block][H]
  const int spb = NT / H;
  for (int i = threadIdx.x; i < H * H; i += NT) {
    const int j = i / H, k = i - j * H;
    w2t[k * H + j] = W2[i];
  }
  const int ls = threadIdx.x / H, j = threadIdx.x - ls * H;
  const int m = blockIdx.x * spb + ls;
  const bool live = ls < spb && m < M;
  float x = 0.f;
  if (live) {
    x = tanhf(z1[(int64_t)m * H + j] + b1[j]);
    a1[(int64_t)m * H + j] = x;
    a1s[ls * H + j] = x;
  }
  __syncthreads();
  if (live) {
    float z = b2[j];
    for (int k = 0; k < H; ++k) z = fmaf(a1s[ls * H + k], w2t[k * H + j], z);
    a2[(int64_t)m * H + j] = tanhf(z);
  }
}

// ---- layers 2, 1 backward: d2 = da2 (1 - a2^2), d1 = (W2^T d2) (1 - a1^2) ----
__global__ void __launch_bounds__(NT) mid_bwd_kernel(int M, int H, const float* __restrict__ da2, const float* __restrict__ a1,
                                                     const float* __restrict__ a2, const float* __restrict__ W2,
                                                     float* __restrict__ d1, float* __restrict__ d2) {
  extern __shared__ float sm[];
  float* w2 = sm;                // [j][k] (torch order)
  float* d2s = sm + H * H;
  const int spb = NT / H;
  for (int i = threadIdx.x; i < H * H; i += NT) w2[i] = W2[i];
  const int ls = threadIdx.x / H, k = threadIdx.x - ls * H;
  const int m = blockIdx.x * spb + ls;
  const bool live = ls < spb && m < M;
  if (live) {
    const float x = a2[(int64_t)m * H + k];
    const float v = da2[(int64_t)m * H + k] * (1.f - x * x);
    d2[(int64_t)m * H + k] = v;
    d2s[ls * H + k] = v;
  }
  __syncthreads();
  if (live) {
    float s = 0.f;
    for (int j = 0; j < H; ++j) s = fmaf(d2s[ls * H + j], w2[j * H + k], s);
    const float x = a1[(int64_t)m * H + k];
    d1[(int64_t)m * H + k] = s * (1.f - x * x);
  }
}

// ---- generic small outer product over samples: P[ch][p][q] = sum_{m in chunk} X[m][p] Y[m][q]; sX[ch][p], sZ[ch][p] ----
__global__ void __launch_bounds__(NT) outer_kernel(int M, int H, int per_chunk, const float* __restrict__ X,
                                                   const float* __restrict__ Y, const float* __restrict__ Z,
                                                   float* __restrict__ P, float* __restrict__ sX, float* __restrict__ sZ) {
  __shared__ __align__(16) float xs[TS][PITCH];
  __shared__ __align__(16) float ys[TS][PITCH];
  const int ch = blockIdx.x, mb = ch * per_chunk, me = min(M, mb + per_chunk);
  const int tp = threadIdx.x >> 4, tq = threadIdx.x & 15;
  float acc[4][4] = {};
  float cx = 0.f, cz = 0.f;
  for (int m0 = mb; m0 < me; m0 += TS) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TS * TS; idx += NT) {
      const int mm = idx >> 6, c = idx & 63;
      const bool ok = m0 + mm < me && c < H;
      xs[mm][c] = ok ? X[(int64_t)(m0 + mm) * H + c] : 0.f;
      ys[mm][c] = ok ? Y[(int64_t)(m0 + mm) * H + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int mm = 0; mm < TS; ++mm) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[mm][4 * tp]);
      const float4 b = *reinterpret_cast<const float4*>(&ys[mm][4 * tq]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (threadIdx.x < H)
      for (int mm = 0; mm < TS && m0 + mm < me; ++mm) {
        cx += xs[mm][threadIdx.x];
        cz += Z[(int64_t)(m0 + mm) * H + threadIdx.x];
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = 4 * tp + i, q = 4 * tq + j;
      if (p < H && q < H) P[((int64_t)ch * H + p) * H + q] = acc[i][j];
    }
  if (threadIdx.x < H) { sX[ch * H + threadIdx.x] = cx; sZ[ch * H + threadIdx.x] = cz; }
}

// ---- layer 1 backward: part[ch][n][j] = sum_{m in chunk} sc x[n, m] d1[m][j] ----
__global__ void __launch_bounds__(NT) l1_bwd_kernel(int M, int E, int H, int n_loc, int per_chunk, const uint8_t* __restrict__ x,
                                                    int64_t strd, float sc, const float* __restrict__ d1,
                                                    float* __restrict__ part) {
  __shared__ __align__(16) float xs[TS][PITCH];   // [m][n]
  __shared__ __align__(16) float ds[TS][PITCH];   // [m][j]
  const int n0 = blockIdx.x * TS, ch = blockIdx.y;
  const int mb = ch * per_chunk, me = min(M, mb + per_chunk);
  const int tn = threadIdx.x >> 4, tj = threadIdx.x & 15;
  float acc[4][4] = {};
  for (int m0 = mb; m0 < me; m0 += TS) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TS * TS; idx += NT) {
      const int n = idx >> 6, mm = idx & 63, m = m0 + mm;      // consecutive threads walk m: coalesced over envs
      float v = 0.f;
      if (n0 + n < n_loc && m < me) {
        const int t = m / E, e = m - t * E;
        v = sc * (float)x[(int64_t)(n0 + n) * strd + (int64_t)t * E + e];
      }
      xs[mm][n] = v;
    }
    for (int idx = threadIdx.x; idx < TS * TS; idx += NT) {
      const int mm = idx >> 6, j = idx & 63;
      ds[mm][j] = (m0 + mm < me && j < H) ? d1[(int64_t)(m0 + mm) * H + j] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int mm = 0; mm < TS; ++mm) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[mm][4 * tn]);
      const float4 b = *reinterpret_cast<const float4*>(&ds[mm][4 * tj]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  float* out = part + (int64_t)ch * n_loc * H;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * tn + i;
    if (n >= n_loc) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (4 * tj + j < H) out[(int64_t)n * H + 4 * tj + j] = acc[i][j];
  }
}

// ---- Gaussian actor output layer ----------------------------------------------------------------
enum { MODE_MU = 0, MODE_LOGP = 1, MODE_BWD_W = 2, MODE_BWD_A = 3 };
constexpr float kHalfLog2Pi = 0.9189385332046727f;

struct OutArgs {
  int M, H, Dl;
  const float* a2;        // [M, H]
  const float* W3;        // [Dl, H] (this rank's rows)
  const float* b3;        // [Dl]
  const float* log_std;   // [Dl]
  const float* act;       // [M, ld_act] raw nature actions; column 0 = this rank's first output row
  int64_t ld_act;
  const float* g;         // [M] dLoss/dlogp
  float* out;             // MU: [M, ld_out]; LOGP: partial [Dl/64][M]; BWD_W: dW3 partial [mch][Dl][H]; BWD_A: da2 partial [dch][M][H]
  int64_t ld_out;
  float* out_b3;          // BWD_W: [mch][Dl]
  float* out_ls;          // BWD_W: [mch][Dl]
  int per_chunk;          // BWD_W: samples per m-chunk; BWD_A: columns per d-chunk
};

// stage [rows x H] of a row-major [*, H] matrix into smem [row][PITCH] (zero padded)
__device__ __forceinline__ void stage_rows(float (*dst)[PITCH], const float* __restrict__ src, int r0, int r_end, int H) {
  for (int idx = threadIdx.x; idx < TS * TS; idx += NT) {
    const int r = idx >> 6, c = idx & 63;
    dst[r][c] = (r0 + r < r_end && c < H) ? src[(int64_t)(r0 + r) * H + c] : 0.f;
  }
}
// the same, transposed: dst[k][row]
__device__ __forceinline__ void stage_rows_t(float (*dst)[PITCH], const float* __restrict__ src, int r0, int r_end, int H) {
  for (int idx = threadIdx.x; idx < TS * TS; idx += NT) {
    const int r = idx >> 6, c = idx & 63;
    dst[c][r] = (r0 + r < r_end && c < H) ? src[(int64_t)(r0 + r) * H + c] : 0.f;
  }
}

// mu tile: thread (tm, td) -> samples 4 tm .. + 3, columns 4 td .. + 3;  a2m [m][k], w3t [k][d]
__device__ __forceinline__ void mu_tile(const float (*a2m)[PITCH], const float (*w3t)[PITCH], int H, int tm, int td,
                                        float (&mu)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) mu[i][j] = 0.f;
  for (int k = 0; k < H; ++k) {
    const float4 b = *reinterpret_cast<const float4*>(&w3t[k][4 * td]);
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a = a2m[4 * tm + i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) mu[i][j] = fmaf(a, bv[j], mu[i][j]);
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(NT) out_kernel(OutArgs g) {
  extern __shared__ __align__(16) float smo[];
  float (*a2m)[PITCH] = reinterpret_cast<float (*)[PITCH]>(smo);                     // [m][k]
  float (*w3t)[PITCH] = reinterpret_cast<float (*)[PITCH]>(smo + TS * PITCH);        // [k][d]
  float (*dms)[PITCH] = reinterpret_cast<float (*)[PITCH]>(smo + 2 * TS * PITCH);    // [m][d]  dLoss/dmu
  float (*aux)[PITCH] = reinterpret_cast<float (*)[PITCH]>(smo + 3 * TS * PITCH);    // BWD_W: [m][d] dlog_std terms; BWD_A: w3m [d][k]
  const int tm = threadIdx.x >> 4, td = threadIdx.x & 15;
  const int M = g.M, H = g.H, Dl = g.Dl;

  if (MODE == MODE_MU || MODE == MODE_LOGP) {
    const int m0 = blockIdx.x * TS, d0 = blockIdx.y * TS;
    stage_rows(a2m, g.a2, m0, M, H);
    stage_rows_t(w3t, g.W3, d0, Dl, H);
    __syncthreads();
    float mu[4][4];
    mu_tile(a2m, w3t, H, tm, td, mu);
    float rowsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + 4 * td + j;
      if (d >= Dl) continue;
      const float b = g.b3[d];
      float ls = 0.f, inv2var = 0.f;
      if (MODE == MODE_LOGP) {
        ls = g.log_std[d];
        const float sd = expf(ls);
        inv2var = 1.f / (2.f * (sd * sd));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + 4 * tm + i;
        if (m >= M) continue;
        const float v = mu[i][j] + b;
        if (MODE == MODE_MU) g.out[(int64_t)m * g.ld_out + d] = v;
        else {
          const float diff = g.act[(int64_t)m * g.ld_act + d] - v;
          rowsum[i] += -(diff * diff) * inv2var - ls - kHalfLog2Pi;
        }
      }
    }
    if (MODE == MODE_LOGP) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = rowsum[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // the 16 column groups of this sample row
        const int m = m0 + 4 * tm + i;
        if (td == 0 && m < M) g.out[(int64_t)blockIdx.y * M + m] = v;
      }
    }
    return;
  }

  // dLoss/dmu of one 64 x 64 tile into dms (and the log_std terms into aux for BWD_W)
  auto dmu_tile = [&](int m0, int m_end, int d0) {
    float mu[4][4];
    mu_tile(a2m, w3t, H, tm, td, mu);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + 4 * td + j;
      float b = 0.f, ivar = 0.f;
      if (d < Dl) {
        b = g.b3[d];
        const float sd = expf(g.log_std[d]);
        ivar = 1.f / (sd * sd);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = m0 + 4 * tm + i;
        float dm = 0.f, dl = 0.f;
        if (d < Dl && m < m_end) {
          const float diff = g.act[(int64_t)m * g.ld_act + d] - (mu[i][j] + b);
          const float gm = g.g[m];
          dm = gm * diff * ivar;                     // d logp / d mu = (a - mu) / var
          dl = gm * (diff * diff * ivar - 1.f);      // d logp / d log_std = (a - mu)^2 / var - 1
        }
        dms[4 * tm + i][4 * td + j] = dm;
        if (MODE == MODE_BWD_W) aux[4 * tm + i][4 * td + j] = dl;
      }
    }
  };

  if (MODE == MODE_BWD_W) {
    // grid (d tiles, m chunks): dW3[d][k] += sum_m dmu[m][d] a2[m][k]; thread (tp, tq) -> d 4 tp.., k 4 tq..
    const int d0 = blockIdx.x * TS, ch = blockIdx.y;
    const int mb = ch * g.per_chunk, me = min(M, mb + g.per_chunk);
    stage_rows_t(w3t, g.W3, d0, Dl, H);
    float acc[4][4] = {};
    float cb = 0.f, cl = 0.f;
    for (int m0 = mb; m0 < me; m0 += TS) {
      __syncthreads();
      stage_rows(a2m, g.a2, m0, me, H);
      __syncthreads();
      dmu_tile(m0, me, d0);
      __syncthreads();
#pragma unroll 8
      for (int mm = 0; mm < TS; ++mm) {
        const float4 a = *reinterpret_cast<const float4*>(&dms[mm][4 * tm]);
        const float4 b = *reinterpret_cast<const float4*>(&a2m[mm][4 * td]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (threadIdx.x < TS)
        for (int mm = 0; mm < TS; ++mm) { cb += dms[mm][threadIdx.x]; cl += aux[mm][threadIdx.x]; }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = d0 + 4 * tm + i;
      if (d >= Dl) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * td + j < H) g.out[((int64_t)ch * Dl + d) * H + 4 * td + j] = acc[i][j];
    }
    if (threadIdx.x < TS && d0 + threadIdx.x < Dl) {
      g.out_b3[(int64_t)ch * Dl + d0 + threadIdx.x] = cb;
      g.out_ls[(int64_t)ch * Dl + d0 + threadIdx.x] = cl;
    }
    return;
  }

  if (MODE == MODE_BWD_A) {
    // grid (m tiles, d chunks): da2[m][k] += sum_d dmu[m][d] W3[d][k]; thread (tm, tq) -> m 4 tm.., k 4 tq..
    const int m0 = blockIdx.x * TS, ch = blockIdx.y;
    const int db = ch * g.per_chunk, de = min(Dl, db + g.per_chunk);
    stage_rows(a2m, g.a2, m0, M, H);
    float acc[4][4] = {};
    for (int d0 = db; d0 < de; d0 += TS) {
      __syncthreads();
      stage_rows_t(w3t, g.W3, d0, de, H);
      stage_rows(aux, g.W3, d0, de, H);          // w3m [d][k]
      __syncthreads();
      dmu_tile(m0, M, d0);
      __syncthreads();
#pragma unroll 4
      for (int dd = 0; dd < TS; ++dd) {
        const float4 b = *reinterpret_cast<const float4*>(&aux[dd][4 * td]);
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = dms[4 * tm + i][dd];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a, bv[j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + 4 * tm + i;
      if (m >= M) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * td + j < H) g.out[((int64_t)ch * M + m) * H + 4 * td + j] = acc[i][j];
    }
  }
}

// note on dmu_tile inside BWD_A: columns past `de` read as zero weights but must not contribute: dmu_tile masks d >= Dl only,
// so chunk boundaries are kept multiples of the tile (host side).

// ---- critic output layer: v = W3 a2 + b3; backward of mean((v - ret)^2) ----
__global__ void __launch_bounds__(NT) v_out_kernel(int M, int H, const float* __restrict__ a2, const float* __restrict__ W3,
                                                   const float* __restrict__ b3, const float* __restrict__ ret,
                                                   float* __restrict__ v, float* __restrict__ da2,
                                                   float* __restrict__ pW3, float* __restrict__ pb3) {
  __shared__ float w3[64];
  __shared__ float dvs[NT];
  if (threadIdx.x < H) w3[threadIdx.x] = W3[threadIdx.x];
  __syncthreads();
  const int m = blockIdx.x * NT + threadIdx.x;
  float dv = 0.f;
  if (m < M) {
    float z = b3[0];
    for (int k = 0; k < H; ++k) z = fmaf(a2[(int64_t)m * H + k], w3[k], z);
    if (v) v[m] = z;
    if (ret) {
      dv = 2.f * (z - ret[m]) / (float)M;
      for (int k = 0; k < H; ++k) da2[(int64_t)m * H + k] = dv * w3[k];
    }
  }
  if (!ret) return;
  dvs[threadIdx.x] = dv;
  __syncthreads();
  // per-block partials in a fixed order: thread k sums dv[m] a2[m][k] over the block's samples; thread H sums dv
  if (threadIdx.x <= H) {
    float s = 0.f;
    const int mb = blockIdx.x * NT;
    for (int i = 0; i < NT && mb + i < M; ++i)
      s += threadIdx.x < H ? dvs[i] * a2[(int64_t)(mb + i) * H + threadIdx.x] : dvs[i];
    if (threadIdx.x < H) pW3[blockIdx.x * H + threadIdx.x] = s;
    else pb3[blockIdx.x] = s;
  }
}

// PPO-clip coefficient of the nature policy loss (nature_oracle.py:359-386): g = dLoss/dlogp
__global__ void __launch_bounds__(NT) pi_coef_kernel(int M, const float* __restrict__ logp, const float* __restrict__ logp_old,
                                                     const float* __restrict__ adv, float clip_lo, float clip_hi,
                                                     float* __restrict__ g) {
  const int m = blockIdx.x * NT + threadIdx.x;
  if (m >= M) return;
  const float ratio = expf(logp[m] - logp_old[m]);
  const float A = adv[m];
  const float clipped = fminf(fmaxf(ratio, clip_lo), clip_hi) * A;
  const float un = ratio * A;
  // min(un, clipped): the clipped branch has zero slope outside the clip range; inside it equals un
  const bool pass = (un <= clipped) || (ratio >= clip_lo && ratio <= clip_hi);
  g[m] = pass ? -(A * ratio) / (float)M : 0.f;
}

static int chunks_for(int total, int tile, int want) {
  int tiles = (total + tile - 1) / tile;
  int ch = tiles < want ? tiles : want;
  return ch < 1 ? 1 : ch;
}

}  // namespace nat
}  // namespace rmab

using namespace rmab;
using namespace rmab::nat;

// chunk counts of the two-stage reductions: fine enough that a rank's shard of the work still fills the GPU (a CTA walks
// at most a few 64-wide tiles; these kernels are latency-bound per CTA, not throughput-bound)
static const int kL1Chunks = 4, kMChunks = 32, kDChunks = 32;

// arm chunks of the layer-1 forward: at least kL1Chunks, more when there are few sample tiles (the rollout's per-step
// forward has M = E samples: 4 tiles at E = 256), so that the grid still covers the GPU
static int l1_chunks(int M, int n_loc) {
  const int mt = (M + TS - 1) / TS;
  int want = (2 * kNumSMs + mt - 1) / mt;
  if (want < kL1Chunks) want = kL1Chunks;
  if (want > 32) want = 32;
  return chunks_for(n_loc, TS, want);
}

extern "C" size_t rmab_nat_ws_bytes(int M, int H, int n_loc, int Dl) {
  size_t a = (size_t)l1_chunks(M, n_loc > 0 ? n_loc : 1) * M * H;   // l1 forward partials
  size_t b = (size_t)kMChunks * ((size_t)(Dl > 0 ? Dl : 1) * (H + 2));   // out backward (W): dW3, db3, dls partials
  size_t c = (size_t)kDChunks * M * H;                               // out backward (A): da2 partials
  size_t d = (size_t)kMChunks * ((size_t)2 * n_loc * H);             // l1 backward partials (state and action blocks)
  size_t e = (size_t)kMChunks * ((size_t)H * H + 2 * H);             // outer partials
  size_t f = (size_t)((Dl + TS - 1) / TS + 1) * M;                   // logp partials
  size_t g = (size_t)((M + NT - 1) / NT) * (H + 1);                  // critic output partials
  size_t m = a;
  for (size_t x : {b, c, d, e, f, g}) m = x > m ? x : m;
  return m * sizeof(float) + 256;
}

static void sum_parts(int64_t count, int nparts, const float* parts, float* out, cudaStream_t st) {
  int64_t b = (count + NT - 1) / NT;
  int grid = (int)(b < (int64_t)kNumSMs * 8 ? b : (int64_t)kNumSMs * 8);
  sum_parts_kernel<<<grid < 1 ? 1 : grid, NT, 0, st>>>(count, nparts, parts, out);
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

#define NAT_H_OK(H) ((H) >= 4 && (H) <= 64 && ((H) % 4) == 0 && (NT % (H)) == 0)

extern "C" int rmab_nat_l1_fwd(int M, int E, int H, int n_loc, const uint8_t* x0, int64_t stride0, float sc0, const float* W0,
                               const uint8_t* x1, int64_t stride1, const float* W1, float* z1, void* ws, size_t ws_bytes,
                               rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && E > 0 && NAT_H_OK(H) && n_loc >= 0 && z1 && ws && (n_loc == 0 || (x0 && W0)), RMAB_E_BADARG,
               "rmab_nat_l1_fwd: bad argument (hidden must be 8, 16, 32 or 64)");
  RMAB_REQUIRE(ws_bytes >= rmab_nat_ws_bytes(M, H, n_loc, 0), RMAB_E_BADARG, "rmab_nat_l1_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_loc == 0) {
    cudaMemsetAsync(z1, 0, (size_t)M * H * 4, st);
    return RMAB_OK;
  }
  const int ch = l1_chunks(M, n_loc);
  const int per = ((n_loc + ch - 1) / ch + TS - 1) / TS * TS;
  const int nch = (n_loc + per - 1) / per;
  l1_fwd_kernel<<<dim3((M + TS - 1) / TS, nch), NT, 0, st>>>(M, E, H, n_loc, per, x0, stride0, sc0, W0, x1, stride1, W1,
                                                            static_cast<float*>(ws));
  sum_parts((int64_t)M * H, nch, static_cast<float*>(ws), z1, st);
  return check_launch("rmab_nat_l1_fwd");
}

extern "C" int rmab_nat_mid_fwd(int M, int H, const float* z1, const float* b1, const float* W2, const float* b2, float* a1,
                                float* a2, rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && NAT_H_OK(H) && z1 && b1 && W2 && b2 && a1 && a2, RMAB_E_BADARG, "rmab_nat_mid_fwd: bad argument");
  const int spb = NT / H;
  mid_fwd_kernel<<<(M + spb - 1) / spb, NT, (size_t)(H * H + spb * H) * 4, (cudaStream_t)stream>>>(M, H, z1, b1, W2, b2, a1, a2);
  return check_launch("rmab_nat_mid_fwd");
}

static int out_smem_attr() {
  static bool done = false;
  if (!done) {
    const int bytes = 4 * TS * PITCH * 4;
    if (cudaFuncSetAttribute(out_kernel<MODE_MU>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess ||
        cudaFuncSetAttribute(out_kernel<MODE_LOGP>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess ||
        cudaFuncSetAttribute(out_kernel<MODE_BWD_W>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess ||
        cudaFuncSetAttribute(out_kernel<MODE_BWD_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
      return fail(RMAB_E_LAUNCH, "rmab_nat_*: cannot reserve %d B of shared memory", bytes);
    done = true;
  }
  return RMAB_OK;
}
static const size_t kOutSmem = 4 * TS * PITCH * 4;

extern "C" int rmab_nat_mu(int M, int H, int Dl, const float* a2, const float* W3, const float* b3, float* mu, int64_t ld_mu,
                           rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && NAT_H_OK(H) && Dl > 0 && a2 && W3 && b3 && mu && ld_mu >= Dl, RMAB_E_BADARG, "rmab_nat_mu: bad argument");
  if (int rc = out_smem_attr()) return rc;
  OutArgs g = {};
  g.M = M; g.H = H; g.Dl = Dl; g.a2 = a2; g.W3 = W3; g.b3 = b3; g.out = mu; g.ld_out = ld_mu;
  out_kernel<MODE_MU><<<dim3((M + TS - 1) / TS, (Dl + TS - 1) / TS), NT, kOutSmem, (cudaStream_t)stream>>>(g);
  return check_launch("rmab_nat_mu");
}

extern "C" int rmab_nat_logp(int M, int H, int Dl, const float* a2, const float* W3, const float* b3, const float* log_std,
                             const float* act, int64_t ld_act, float* logp, void* ws, size_t ws_bytes, rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && NAT_H_OK(H) && Dl >= 0 && a2 && logp && ws, RMAB_E_BADARG, "rmab_nat_logp: bad argument");
  RMAB_REQUIRE(ws_bytes >= rmab_nat_ws_bytes(M, H, 0, Dl), RMAB_E_BADARG, "rmab_nat_logp: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (Dl == 0) {
    cudaMemsetAsync(logp, 0, (size_t)M * 4, st);
    return RMAB_OK;
  }
  if (int rc = out_smem_attr()) return rc;
  OutArgs g = {};
  g.M = M; g.H = H; g.Dl = Dl; g.a2 = a2; g.W3 = W3; g.b3 = b3; g.log_std = log_std; g.act = act; g.ld_act = ld_act;
  g.out = static_cast<float*>(ws);
  const int dt = (Dl + TS - 1) / TS;
  out_kernel<MODE_LOGP><<<dim3((M + TS - 1) / TS, dt), NT, kOutSmem, st>>>(g);
  sum_parts(M, dt, static_cast<float*>(ws), logp, st);
  return check_launch("rmab_nat_logp");
}

extern "C" int rmab_nat_out_bwd(int M, int H, int Dl, const float* a2, const float* W3, const float* b3, const float* log_std,
                                const float* act, int64_t ld_act, const float* g_coef, float* dW3, float* db3, float* dls,
                                float* da2, void* ws, size_t ws_bytes, rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && NAT_H_OK(H) && Dl >= 0 && a2 && g_coef && da2 && ws, RMAB_E_BADARG, "rmab_nat_out_bwd: bad argument");
  RMAB_REQUIRE(ws_bytes >= rmab_nat_ws_bytes(M, H, 0, Dl), RMAB_E_BADARG, "rmab_nat_out_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (Dl == 0) {
    cudaMemsetAsync(da2, 0, (size_t)M * H * 4, st);
    return RMAB_OK;
  }
  if (int rc = out_smem_attr()) return rc;
  float* w = static_cast<float*>(ws);
  OutArgs g = {};
  g.M = M; g.H = H; g.Dl = Dl; g.a2 = a2; g.W3 = W3; g.b3 = b3; g.log_std = log_std; g.act = act; g.ld_act = ld_act;
  g.g = g_coef;
  {   // weight-side sums over samples
    const int ch = chunks_for(M, TS, kMChunks);
    const int per = ((M + ch - 1) / ch + TS - 1) / TS * TS;
    const int nch = (M + per - 1) / per;
    g.per_chunk = per;
    g.out = w;
    g.out_b3 = w + (size_t)nch * Dl * H;
    g.out_ls = g.out_b3 + (size_t)nch * Dl;
    out_kernel<MODE_BWD_W><<<dim3((Dl + TS - 1) / TS, nch), NT, kOutSmem, st>>>(g);
    sum_parts((int64_t)Dl * H, nch, g.out, dW3, st);
    sum_parts(Dl, nch, g.out_b3, db3, st);
    sum_parts(Dl, nch, g.out_ls, dls, st);
  }
  {   // activation-side sums over this rank's output rows
    const int ch = chunks_for(Dl, TS, kDChunks);
    const int per = ((Dl + ch - 1) / ch + TS - 1) / TS * TS;
    const int nch = (Dl + per - 1) / per;
    g.per_chunk = per;
    g.out = w;
    out_kernel<MODE_BWD_A><<<dim3((M + TS - 1) / TS, nch), NT, kOutSmem, st>>>(g);
    sum_parts((int64_t)M * H, nch, w, da2, st);
  }
  return check_launch("rmab_nat_out_bwd");
}

extern "C" int rmab_nat_mid_bwd(int M, int H, const float* da2, const float* a1, const float* a2, const float* W2, float* d1,
                                float* d2, float* dW2, float* db2, float* db1, void* ws, size_t ws_bytes, rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && NAT_H_OK(H) && da2 && a1 && a2 && W2 && d1 && d2 && dW2 && db2 && db1 && ws, RMAB_E_BADARG,
               "rmab_nat_mid_bwd: bad argument");
  RMAB_REQUIRE(ws_bytes >= rmab_nat_ws_bytes(M, H, 0, 0), RMAB_E_BADARG, "rmab_nat_mid_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int spb = NT / H;
  mid_bwd_kernel<<<(M + spb - 1) / spb, NT, (size_t)(H * H + spb * H) * 4, st>>>(M, H, da2, a1, a2, W2, d1, d2);
  const int ch = chunks_for(M, TS, kMChunks);
  const int per = ((M + ch - 1) / ch + TS - 1) / TS * TS;
  const int nch = (M + per - 1) / per;
  float* w = static_cast<float*>(ws);
  float* pX = w + (size_t)nch * H * H;
  float* pZ = pX + (size_t)nch * H;
  outer_kernel<<<nch, NT, 0, st>>>(M, H, per, d2, a1, d1, w, pX, pZ);     // dW2[j][k] = sum_m d2[m][j] a1[m][k]
  sum_parts((int64_t)H * H, nch, w, dW2, st);
  sum_parts(H, nch, pX, db2, st);
  sum_parts(H, nch, pZ, db1, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_launch("rmab_nat_mid_bwd");
}

extern "C" int rmab_nat_l1_bwd(int M, int E, int H, int n_loc, const uint8_t* x0, int64_t stride0, float sc0,
                               const uint8_t* x1, int64_t stride1, const float* d1, float* dW0, float* dW1, void* ws,
                               size_t ws_bytes, rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && E > 0 && NAT_H_OK(H) && n_loc >= 0 && d1 && ws, RMAB_E_BADARG, "rmab_nat_l1_bwd: bad argument");
  RMAB_REQUIRE(ws_bytes >= rmab_nat_ws_bytes(M, H, n_loc, 0), RMAB_E_BADARG, "rmab_nat_l1_bwd: workspace too small");
  if (n_loc == 0) return RMAB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int ch = chunks_for(M, TS, kMChunks);
  const int per = ((M + ch - 1) / ch + TS - 1) / TS * TS;
  const int nch = (M + per - 1) / per;
  float* w = static_cast<float*>(ws);
  for (int src = 0; src < 2; ++src) {
    const uint8_t* x = src == 0 ? x0 : x1;
    float* dW = src == 0 ? dW0 : dW1;
    if (!x || !dW) continue;
    l1_bwd_kernel<<<dim3((n_loc + TS - 1) / TS, nch), NT, 0, st>>>(M, E, H, n_loc, per, x, src == 0 ? stride0 : stride1,
                                                                  src == 0 ? sc0 : 1.f, d1, w);
    sum_parts((int64_t)n_loc * H, nch, w, dW, st);
  }
  return check_launch("rmab_nat_l1_bwd");
}

extern "C" int rmab_nat_v_out(int M, int H, const float* a2, const float* W3, const float* b3, const float* ret, float* v,
                              float* da2, float* dW3, float* db3, void* ws, size_t ws_bytes, rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && NAT_H_OK(H) && a2 && W3 && b3 && (v || ret), RMAB_E_BADARG, "rmab_nat_v_out: bad argument");
  RMAB_REQUIRE(!ret || (da2 && dW3 && db3 && ws && ws_bytes >= rmab_nat_ws_bytes(M, H, 0, 0)), RMAB_E_BADARG,
               "rmab_nat_v_out: backward needs da2, dW3, db3 and the workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = (M + NT - 1) / NT;
  float* w = static_cast<float*>(ws);
  v_out_kernel<<<nb, NT, 0, st>>>(M, H, a2, W3, b3, ret, v, da2, w, w ? w + (size_t)nb * H : nullptr);
  if (ret) {
    sum_parts(H, nb, w, dW3, st);
    sum_parts(1, nb, w + (size_t)nb * H, db3, st);
  }
  return check_launch("rmab_nat_v_out");
}

extern "C" int rmab_nat_pi_coef(int M, const float* logp, const float* logp_old, const float* adv, float clip_ratio, float* g,
                                rmab_stream_t stream) {
  RMAB_REQUIRE(M > 0 && logp && logp_old && adv && g, RMAB_E_BADARG, "rmab_nat_pi_coef: bad argument");
  pi_coef_kernel<<<(M + NT - 1) / NT, NT, 0, (cudaStream_t)stream>>>(M, logp, logp_old, adv, 1.f - clip_ratio, 1.f + clip_ratio, g);
  return check_launch("rmab_nat_pi_coef");
}
