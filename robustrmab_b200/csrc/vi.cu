// Batched value iteration over arms x lambda-probes and Whittle-index bisection.
//
// Reference behaviour: vendored pymdptoolbox, mdptoolbox/mdp.py  MDP._bellmanOperator :217-246,
// ValueIteration.__init__ :1293-1318, _boundIter :1320-1364, run :1366-1410 (local 'full' stop rule
// max(V - Vprev) < eps(1-g)/g, or iter == max_iter); call site robust_rmab/simulator.py:504-521
// (R_i[a,s,s'] = R[i][s]; Q = R + g V.T[s,a]); lambda charge and indifference condition from
// robust_rmab/lp_methods.py:63,181.  All arithmetic in fp64 like the reference.
//   small S (CE S=2, ARMMAN S=3): one thread per (arm, probe), the whole MDP in registers -- ALU/latency
//     bound (the tables are 64-144 B per arm, far below anything HBM could limit).
//   generic S (SIS S=pop+1): one CTA per arm, T staged once in shared memory (fp64), threads over
//     (probe, state); shared-memory bound.
#include "common.cuh"

namespace rmab {

struct ViParams {
  double gamma, eps, thresh;
  int stop_mode;
};

constexpr int kViCap = 1000000;  // safety cap; the reference has none (VI contracts, so it always stops)

// max_iter of _boundIter, or -1 where the reference raises (span == 0, g*k == 1 -> division by zero;
// g*k <= 0 -> math domain error).
__device__ __forceinline__ int bound_iter(double k, double span, const ViParams& p) {
  const double gk = p.gamma * k;
  if (!(span > 0.0) || !(gk > 0.0) || gk == 1.0) return -1;
  const double den = log(gk);
  if (den == 0.0) return -1;
  const double v = log(p.thresh / span) / den;
  return (int)ceil(v);
}
// The same with log(gamma * k) supplied by the caller: it depends on the arm only, and the Whittle bisection solves the
// same arm at 20 values of lambda (one fp64 log per solve less; the value is the one bound_iter computes).
__device__ __forceinline__ int bound_iter_den(double k, double den, double span, const ViParams& p) {
  const double gk = p.gamma * k;
  if (!(span > 0.0) || !(gk > 0.0) || gk == 1.0) return -1;
  if (den == 0.0) return -1;
  const double v = log(p.thresh / span) / den;
  return (int)ceil(v);
}

// ---- small MDPs: everything in registers ---------------------------------------------------------
template <int S, int A>
struct SmallMdp {
  double P[A][S][S];  // P[a][s][s']
  double R[S];
  double C[A];
  double k;           // 1 - sum_ss min_{a,s} P[a][s][ss]
  double Ra[S][A];    // general reward R[s][a] (rmab_vi_general); used instead of R[s] - lam * C[a] when use_ra
  bool use_ra;
  __device__ __forceinline__ double rew(int s, int a, double lam) const {
    double v = 0.0;
#pragma unroll
    for (int ss = 0; ss < S; ++ss)
#pragma unroll
      for (int aa = 0; aa < A; ++aa)
        if (ss == s && aa == a) v = use_ra ? Ra[ss][aa] : (R[ss] - lam * C[aa]);
    return v;
  }
  __device__ void load(const double* __restrict__ T, const double* __restrict__ Rg, const double* __restrict__ Cg,
                       const double* __restrict__ Rsa = nullptr) {
    use_ra = Rsa != nullptr;
#pragma unroll
    for (int s = 0; s < S; ++s) {
#pragma unroll
      for (int a = 0; a < A; ++a) Ra[s][a] = use_ra ? Rsa[s * A + a] : 0.0;
    }
    if (use_ra) Rg = nullptr;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      R[s] = Rg ? Rg[s] : 0.0;
#pragma unroll
      for (int a = 0; a < A; ++a)
#pragma unroll
        for (int t = 0; t < S; ++t) P[a][s][t] = T[(s * A + a) * S + t];
    }
#pragma unroll
    for (int a = 0; a < A; ++a) C[a] = Cg ? Cg[a] : 0.0;
    double hs = 0.0;
#pragma unroll
    for (int t = 0; t < S; ++t) {
      double m = P[0][0][t];
#pragma unroll
      for (int a = 0; a < A; ++a)
#pragma unroll
        for (int s = 0; s < S; ++s) m = fmin(m, P[a][s][t]);
      hs += m;
    }
    k = 1.0 - hs;
  }
  // one Bellman backup: Vn[s] = max_a Q, pol[s] = first argmax
  __device__ __forceinline__ void bellman(const double (&V)[S], double lam, const ViParams& p, double (&Vn)[S],
                                          int (&pol)[S]) const {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      double best = 0.0;
      int ba = 0;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        double ev = 0.0;
#pragma unroll
        for (int t = 0; t < S; ++t) ev += P[a][s][t] * V[t];
        const double q = rew(s, a, lam) + p.gamma * ev;
        if (a == 0 || q > best) { best = q; ba = a; }
      }
      Vn[s] = best;
      pol[s] = ba;
    }
  }
  // one Bellman backup with the rewards of this lambda tabulated by the caller (same arithmetic as bellman())
  __device__ __forceinline__ void backup(const double (&V)[S], const double (&rw)[S][A], const ViParams& p, double (&Vn)[S],
                                         int (&pol)[S]) const {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      double best = 0.0;
      int ba = 0;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        double ev = 0.0;
#pragma unroll
        for (int t = 0; t < S; ++t) ev += P[a][s][t] * V[t];
        const double q = rw[s][a] + p.gamma * ev;
        if (a == 0 || q > best) { best = q; ba = a; }
      }
      Vn[s] = best;
      pol[s] = ba;
    }
  }
  // the stopping rule of one sweep (mdp.py:1375-1396); true = stop.  MODE is the stop mode as a compile-time constant: only
  // its own variation is evaluated (the run-time form computed all three per sweep, half of the loop's instructions).
  template <int MODE>
  __device__ __forceinline__ bool sweep_done(const double (&Vn)[S], const double (&V)[S], int it, int max_iter,
                                             const ViParams& p) const {
    double dmx = Vn[0] - V[0], dmn = dmx;
#pragma unroll
    for (int s = 1; s < S; ++s) {
      const double dd = Vn[s] - V[s];
      dmx = fmax(dmx, dd);
      if (MODE != RMAB_STOP_FULL) dmn = fmin(dmn, dd);
    }
    double variation = MODE == RMAB_STOP_FAST ? dmx - dmn : dmx;
    if (MODE == RMAB_STOP_ABS) variation = fmax(fabs(dmx), fabs(dmn));  // sup-norm change: true convergence
    return variation < p.thresh || (it == max_iter && MODE != RMAB_STOP_ABS) || it >= kViCap;
  }
  // ValueIteration(...).run(); returns iterations (or -1) and fills V, pol, max_iter.  The sweeps ping-pong between two
  // register sets (no copy per sweep) over rewards tabulated once per lambda; the arithmetic of a sweep is bellman()'s.
  __device__ int solve(double lam, const ViParams& p, double (&V)[S], int (&pol)[S], int& max_iter) const {
    return solve(lam, p, V, pol, max_iter, log(p.gamma * k));
  }
  // log_gk = log(gamma * k) of this arm (bound_iter_den)
  __device__ int solve(double lam, const ViParams& p, double (&V)[S], int (&pol)[S], int& max_iter, double log_gk) const {
    if (p.stop_mode == RMAB_STOP_FULL) return solve_mode<RMAB_STOP_FULL>(lam, p, V, pol, max_iter, log_gk);
    if (p.stop_mode == RMAB_STOP_FAST) return solve_mode<RMAB_STOP_FAST>(lam, p, V, pol, max_iter, log_gk);
    return solve_mode<RMAB_STOP_ABS>(lam, p, V, pol, max_iter, log_gk);
  }
  template <int MODE>
  __device__ int solve_mode(double lam, const ViParams& p, double (&V)[S], int (&pol)[S], int& max_iter, double log_gk) const {
    double rw[S][A];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
      for (int a = 0; a < A; ++a) rw[s][a] = use_ra ? Ra[s][a] : (R[s] - lam * C[a]);
    double Va[S], Vb[S];
#pragma unroll
    for (int s = 0; s < S; ++s) { Va[s] = 0.0; pol[s] = 0; }
    backup(Va, rw, p, Vb, pol);
    double mx = Vb[0], mn = Vb[0];
#pragma unroll
    for (int s = 1; s < S; ++s) { mx = fmax(mx, Vb[s]); mn = fmin(mn, Vb[s]); }
    max_iter = bound_iter_den(k, log_gk, mx - mn, p);
    if (MODE == RMAB_STOP_ABS && max_iter == -1) max_iter = kViCap;
    if (max_iter == -1) {
#pragma unroll
      for (int s = 0; s < S; ++s) V[s] = __longlong_as_double(0x7ff8000000000000ll);
      return -1;
    }
    // sweep 1 from V = 0 is the backup above (Vb, pol)
    int it = 1;
    if (!sweep_done<MODE>(Vb, Va, it, max_iter, p)) {
      while (true) {
        ++it;
        backup(Vb, rw, p, Va, pol);
        if (sweep_done<MODE>(Va, Vb, it, max_iter, p)) {
#pragma unroll
          for (int s = 0; s < S; ++s) V[s] = Va[s];
          return it;
        }
        ++it;
        backup(Va, rw, p, Vb, pol);
        if (sweep_done<MODE>(Vb, Va, it, max_iter, p)) break;
      }
    }
#pragma unroll
    for (int s = 0; s < S; ++s) V[s] = Vb[s];
    return it;
  }
  __device__ __forceinline__ double q(int s, int a, double lam, const ViParams& p, const double (&V)[S]) const {
    double ev = 0.0;
#pragma unroll
    for (int t = 0; t < S; ++t) {
      double pv = 0.0;
#pragma unroll
      for (int ss = 0; ss < S; ++ss)
#pragma unroll
        for (int aa = 0; aa < A; ++aa)
          if (ss == s && aa == a) pv = P[aa][ss][t];
      ev += pv * V[t];
    }
    return rew(s, a, lam) + p.gamma * ev;
  }
};

template <int S, int A>
__global__ void __launch_bounds__(128) vi_small_kernel(int N, int L, const double* __restrict__ T,
                                                       const double* __restrict__ R, const double* __restrict__ C,
                                                       const double* __restrict__ lambdas, int per_arm, ViParams p,
                                                       double* __restrict__ Vout, uint8_t* __restrict__ pol_out,
                                                       int32_t* __restrict__ iters, int32_t* __restrict__ max_iter_out,
                                                       double* __restrict__ Qout, const double* __restrict__ Rsa) {
  const int64_t total = (int64_t)N * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / L), l = (int)(i - (int64_t)n * L);
    SmallMdp<S, A> m;
    m.load(T + (int64_t)n * S * A * S, R ? R + (int64_t)n * S : nullptr, C, Rsa ? Rsa + (int64_t)n * S * A : nullptr);
    const double lam = Rsa ? 0.0 : (per_arm ? lambdas[i] : lambdas[l]);
    double V[S];
    int pol[S], mi;
    const int it = m.solve(lam, p, V, pol, mi);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      Vout[i * S + s] = V[s];
      if (pol_out) pol_out[i * S + s] = (uint8_t)pol[s];
    }
    if (iters) iters[i] = it;
    if (max_iter_out) max_iter_out[i] = mi;
    if (Qout)
#pragma unroll
      for (int s = 0; s < S; ++s)
#pragma unroll
        for (int a = 0; a < A; ++a) Qout[(i * S + s) * A + a] = m.q(s, a, lam, p, V);
  }
}

template <int S, int A>
__global__ void __launch_bounds__(128) whittle_small_kernel(int N, const double* __restrict__ T,
                                                            const double* __restrict__ R, const double* __restrict__ C,
                                                            ViParams p, double lam_lo, double lam_hi, int n_rounds,
                                                            double* __restrict__ index) {
  const int64_t total = (int64_t)N * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / S), st = (int)(i - (int64_t)n * S);
    SmallMdp<S, A> m;
    m.load(T + (int64_t)n * S * A * S, R + (int64_t)n * S, C);
    double lo = lam_lo, hi = lam_hi;
    const double log_gk = log(p.gamma * m.k);
    for (int r = 0; r < n_rounds; ++r) {
      const double mid = 0.5 * (lo + hi);
      double V[S];
      int pol[S], mi;
      const int it = m.solve(mid, p, V, pol, mi, log_gk);
      bool act = false;
      if (it != -1) act = (m.q(st, 1, mid, p, V) - m.q(st, 0, mid, p, V)) > 0.0;
      if (act) lo = mid; else hi = mid;
    }
    index[i] = 0.5 * (lo + hi);
  }
}

constexpr int kViLp = 4;   // probes per work item of the generic sweep

// ---- generic S: one CTA per arm, T in shared memory ----------------------------------------------
// mode 0: L given probes.  mode 1 (Whittle): L = S probes, probe l bisects the index of state l.
__global__ void __launch_bounds__(256) vi_generic_kernel(int S, int A, int L, const double* __restrict__ T,
                                                         const double* __restrict__ R, const double* __restrict__ C,
                                                         const double* __restrict__ lambdas, int per_arm, ViParams p,
                                                         int mode, double lam_lo, double lam_hi, int n_rounds,
                                                         double* __restrict__ Vout, uint8_t* __restrict__ pol_out,
                                                         int32_t* __restrict__ iters, int32_t* __restrict__ max_iter_out,
                                                         double* __restrict__ Qout, double* __restrict__ index,
                                                         const double* __restrict__ Rsa) {
  extern __shared__ __align__(16) unsigned char sraw[];
  double* sT = reinterpret_cast<double*>(sraw);        // [S][A][S]
  double* sR = sT + (size_t)S * A * S;                 // [S]
  double* sC = sR + S;                                 // [A]
  double* sV0 = sC + A;                                // [L][S]
  double* sV1 = sV0 + (size_t)L * S;                   // [L][S]
  double* sLam = sV1 + (size_t)L * S;                  // [L]
  double* sLo = sLam + L;                              // [L]
  double* sHi = sLo + L;                               // [L]
  double* sMin = sHi + L;                              // [S] column minima
  double* sRa = sMin + S;                              // [S][A] general reward (rmab_vi_general), else unused
  int* sIt = reinterpret_cast<int*>(sRa + (size_t)S * A);  // [L] iteration counters
  int* sMax = sIt + L;                                 // [L] max_iter (-1 invalid)
  int* sAct = sMax + L;                                // [L] still iterating
  uint8_t* sPol = reinterpret_cast<uint8_t*>(sAct + L);  // [L][S]
  __shared__ int s_any;
  __shared__ double s_k;
  const int n = blockIdx.x, tid = threadIdx.x;
  const double* Tn = T + (int64_t)n * S * A * S;
  for (int i = tid; i < S * A * S; i += blockDim.x) sT[i] = Tn[i];
  const bool use_ra = Rsa != nullptr;
  for (int i = tid; i < S; i += blockDim.x) sR[i] = use_ra ? 0.0 : R[(int64_t)n * S + i];
  for (int i = tid; i < A; i += blockDim.x) sC[i] = use_ra ? 0.0 : C[i];
  for (int i = tid; i < S * A; i += blockDim.x) sRa[i] = use_ra ? Rsa[(int64_t)n * S * A + i] : 0.0;
#define RMAB_REW(s_, a_, l_) (use_ra ? sRa[(s_) * A + (a_)] : (sR[s_] - sLam[l_] * sC[a_]))
  for (int l = tid; l < L; l += blockDim.x) {
    sLo[l] = lam_lo;
    sHi[l] = lam_hi;
    if (mode == 0) sLam[l] = use_ra ? 0.0 : (per_arm ? lambdas[(int64_t)n * L + l] : lambdas[l]);
  }
  __syncthreads();
  // k = 1 - sum_t min_{a,s} P[a][s,t]   (:1344-1354)
  for (int t = tid; t < S; t += blockDim.x) {
    double m = sT[t];
    for (int sa = 0; sa < S * A; ++sa) m = fmin(m, sT[(size_t)sa * S + t]);
    sMin[t] = m;
  }
  __syncthreads();
  if (tid == 0) {
    double hs = 0.0;
    for (int t = 0; t < S; ++t) hs += sMin[t];
    s_k = 1.0 - hs;
  }
  __syncthreads();
  const int rounds = mode == 1 ? n_rounds : 1;
  for (int rd = 0; rd < rounds; ++rd) {
    if (mode == 1)
      for (int l = tid; l < L; l += blockDim.x) sLam[l] = 0.5 * (sLo[l] + sHi[l]);
    __syncthreads();
    // V = 0; per-probe max_iter from the first backup (value = max_a R[s] - lam C[a])
    for (int l = tid; l < L; l += blockDim.x) {
      double mx = 0.0, mn = 0.0;
      for (int s = 0; s < S; ++s) {
        double best = RMAB_REW(s, 0, l);
        for (int a = 1; a < A; ++a) best = fmax(best, RMAB_REW(s, a, l));
        if (s == 0) { mx = best; mn = best; } else { mx = fmax(mx, best); mn = fmin(mn, best); }
      }
      int mi = bound_iter(s_k, mx - mn, p);
      if (p.stop_mode == RMAB_STOP_ABS && mi == -1) mi = kViCap;
      sMax[l] = mi;
      sIt[l] = mi == -1 ? -1 : 0;
      sAct[l] = mi != -1;
    }
    for (int i = tid; i < L * S; i += blockDim.x) { sV0[i] = 0.0; sPol[i] = 0; }
    __syncthreads();
    double* Vc = sV0;
    double* Vn = sV1;
    for (int guard = 0; guard < kViCap; ++guard) {
      if (tid == 0) s_any = 0;
      __syncthreads();
      // One work item = one state and a group of kViLp probes: a transition row element is read from shared memory once
      // for the whole group (consecutive lanes = consecutive states, so the V reads are broadcasts).  The sweep is bound by
      // shared-memory bandwidth, not by the fp64 pipe; per probe the sum over t keeps its order.
      const int LG = (L + kViLp - 1) / kViLp;
      for (int i = tid; i < LG * S; i += blockDim.x) {
        const int lg = i / S, s = i - lg * S, l0 = lg * kViLp;
        bool on[kViLp], any = false;
        int lj[kViLp];
#pragma unroll
        for (int j = 0; j < kViLp; ++j) {
          lj[j] = min(l0 + j, L - 1);
          on[j] = (l0 + j < L) && sAct[lj[j]];
          any |= on[j];
        }
        if (any) {
          double best[kViLp];
          int ba[kViLp];
          for (int a = 0; a < A; ++a) {
            const double* row = sT + ((size_t)s * A + a) * S;
            double ev[kViLp];
#pragma unroll
            for (int j = 0; j < kViLp; ++j) ev[j] = 0.0;
            for (int t = 0; t < S; ++t) {
              const double r = row[t];
#pragma unroll
              for (int j = 0; j < kViLp; ++j) ev[j] += r * Vc[(size_t)lj[j] * S + t];
            }
#pragma unroll
            for (int j = 0; j < kViLp; ++j) {
              const double q = RMAB_REW(s, a, lj[j]) + p.gamma * ev[j];
              if (a == 0 || q > best[j]) { best[j] = q; ba[j] = a; }
            }
          }
#pragma unroll
          for (int j = 0; j < kViLp; ++j)
            if (on[j]) { Vn[(size_t)lj[j] * S + s] = best[j]; sPol[(size_t)lj[j] * S + s] = (uint8_t)ba[j]; }
        }
#pragma unroll
        for (int j = 0; j < kViLp; ++j)
          if (l0 + j < L && !on[j]) Vn[(size_t)lj[j] * S + s] = Vc[(size_t)lj[j] * S + s];
      }
      __syncthreads();
      for (int l = tid; l < L; l += blockDim.x) {
        if (!sAct[l]) continue;
        double dmx = 0.0, dmn = 0.0;
        for (int s = 0; s < S; ++s) {
          const double dd = Vn[(size_t)l * S + s] - Vc[(size_t)l * S + s];
          if (s == 0) { dmx = dd; dmn = dd; } else { dmx = fmax(dmx, dd); dmn = fmin(dmn, dd); }
        }
        const int it = ++sIt[l];
        double variation = p.stop_mode == RMAB_STOP_FAST ? dmx - dmn : dmx;
        if (p.stop_mode == RMAB_STOP_ABS) variation = fmax(fabs(dmx), fabs(dmn));
        if (variation < p.thresh || (it == sMax[l] && p.stop_mode != RMAB_STOP_ABS)) sAct[l] = 0;
        else s_any = 1;
      }
      __syncthreads();
      double* tmp = Vc; Vc = Vn; Vn = tmp;
      if (!s_any) break;
      __syncthreads();
    }
    __syncthreads();
    if (mode == 1) {
      // probe l holds V for lambda = mid_l; gap at state l decides the half interval (lp_methods.py:181)
      for (int l = tid; l < L; l += blockDim.x) {
        bool act = false;
        if (sMax[l] != -1) {
          double q[2];
          for (int a = 0; a < 2; ++a) {
            const double* row = sT + ((size_t)l * A + a) * S;
            double ev = 0.0;
            for (int t = 0; t < S; ++t) ev += row[t] * Vc[(size_t)l * S + t];
            q[a] = RMAB_REW(l, a, l) + p.gamma * ev;
          }
          act = (q[1] - q[0]) > 0.0;
        }
        if (act) sLo[l] = sLam[l]; else sHi[l] = sLam[l];
      }
      __syncthreads();
    } else {
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
      for (int i = tid; i < L * S; i += blockDim.x) {
        const int l = i / S, s = i - l * S;
        const int64_t o = ((int64_t)n * L + l) * S + s;
        const bool ok = sMax[l] != -1;
        Vout[o] = ok ? Vc[i] : nan;
        if (pol_out) pol_out[o] = ok ? sPol[i] : 0;
        if (Qout)
          for (int a = 0; a < A; ++a) {
            const double* row = sT + ((size_t)s * A + a) * S;
            double ev = 0.0;
            for (int t = 0; t < S; ++t) ev += row[t] * Vc[(size_t)l * S + t];
            Qout[o * A + a] = ok ? RMAB_REW(s, a, l) + p.gamma * ev : nan;
          }
      }
      for (int l = tid; l < L; l += blockDim.x) {
        if (iters) iters[(int64_t)n * L + l] = sIt[l];
        if (max_iter_out) max_iter_out[(int64_t)n * L + l] = sMax[l];
      }
    }
  }
  if (mode == 1)
    for (int l = tid; l < L; l += blockDim.x) index[(int64_t)n * S + l] = 0.5 * (sLo[l] + sHi[l]);
}

#undef RMAB_REW

static size_t vi_generic_smem(int S, int A, int L) {
  size_t d = (size_t)S * A * S + S + A + 2 * (size_t)L * S + 3 * (size_t)L + S + (size_t)S * A;
  size_t b = d * 8 + 3 * (size_t)L * 4 + (size_t)L * S;
  return (b + 15) & ~(size_t)15;
}

static int launch_generic(int N, int S, int A, int L, const double* T, const double* R, const double* C,
                          const double* lambdas, int per_arm, const ViParams& p, int mode, double lo, double hi,
                          int rounds, double* V, uint8_t* pol, int32_t* iters, int32_t* mi, double* Q, double* index,
                          cudaStream_t st, const char* who, const double* Rsa = nullptr) {
  const size_t smem = vi_generic_smem(S, A, L);
  RMAB_REQUIRE(smem <= 227 * 1024, RMAB_E_SHAPE, "%s: S=%d A=%d L=%d needs %zu B of shared memory (> 227 KB)", who, S, A, L,
               smem);
  if (smem > 48 * 1024)
    RMAB_REQUIRE(cudaFuncSetAttribute(vi_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) ==
                     cudaSuccess,
                 RMAB_E_LAUNCH, "%s: cannot reserve %zu B smem", who, smem);
  vi_generic_kernel<<<N, 256, smem, st>>>(S, A, L, T, R, C, lambdas, per_arm, p, mode, lo, hi, rounds, V, pol, iters, mi, Q,
                                         index, Rsa);
  return check_launch(who);
}

}  // namespace rmab

using namespace rmab;

static ViParams make_params(double gamma, double eps, int stop_mode) {
  ViParams p;
  p.gamma = gamma;
  p.eps = eps;
  p.thresh = gamma < 1.0 ? eps * (1.0 - gamma) / gamma : eps;  // :1314-1318
  p.stop_mode = stop_mode;
  return p;
}

extern "C" int rmab_vi_batched(int N, int S, int A, int L, const double* T, const double* R, const double* C,
                               const double* lambdas, int lambdas_per_arm, double gamma, double eps, int stop_mode,
                               double* V, uint8_t* policy, int32_t* iters, int32_t* max_iter, double* Q,
                               rmab_stream_t stream) {
  RMAB_REQUIRE(N >= 0 && S >= 1 && S <= 256 && A >= 1 && A <= 16 && L >= 1, RMAB_E_SHAPE, "rmab_vi_batched: bad shape");
  RMAB_REQUIRE(T && R && C && lambdas && V, RMAB_E_BADARG, "rmab_vi_batched: null pointer");
  RMAB_REQUIRE(gamma > 0.0 && gamma < 1.0 && eps > 0.0, RMAB_E_BADARG, "rmab_vi_batched: need 0 < gamma < 1 and eps > 0");
  RMAB_REQUIRE(stop_mode == RMAB_STOP_FULL || stop_mode == RMAB_STOP_FAST || stop_mode == RMAB_STOP_ABS, RMAB_E_BADARG,
               "rmab_vi_batched: bad stop_mode");
  if (N == 0) return RMAB_OK;
  const ViParams p = make_params(gamma, eps, stop_mode);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)N * L;
  const int grid = (int)((total + 127) / 128 < (int64_t)kNumSMs * 32 ? (total + 127) / 128 : (int64_t)kNumSMs * 32);
  if (S == 2 && A == 2) {
    vi_small_kernel<2, 2><<<grid, 128, 0, st>>>(N, L, T, R, C, lambdas, lambdas_per_arm, p, V, policy, iters, max_iter, Q,
                                                nullptr);
    return check_launch("rmab_vi_batched<2,2>");
  }
  if (S == 3 && A == 2) {
    vi_small_kernel<3, 2><<<grid, 128, 0, st>>>(N, L, T, R, C, lambdas, lambdas_per_arm, p, V, policy, iters, max_iter, Q,
                                                nullptr);
    return check_launch("rmab_vi_batched<3,2>");
  }
  return launch_generic(N, S, A, L, T, R, C, lambdas, lambdas_per_arm, p, 0, 0.0, 0.0, 1, V, policy, iters, max_iter, Q,
                        nullptr, st, "rmab_vi_batched");
}

extern "C" int rmab_whittle_bisect(int N, int S, int A, const double* T, const double* R, const double* C, double gamma,
                                   double eps, int stop_mode, double lam_lo, double lam_hi, int n_rounds, double* index,
                                   rmab_stream_t stream) {
  RMAB_REQUIRE(N >= 0 && S >= 1 && S <= 256 && A >= 2 && A <= 16 && n_rounds >= 0, RMAB_E_SHAPE,
               "rmab_whittle_bisect: bad shape");
  RMAB_REQUIRE(T && R && C && index, RMAB_E_BADARG, "rmab_whittle_bisect: null pointer");
  RMAB_REQUIRE(gamma > 0.0 && gamma < 1.0 && eps > 0.0, RMAB_E_BADARG, "rmab_whittle_bisect: need 0 < gamma < 1, eps > 0");
  if (N == 0) return RMAB_OK;
  const ViParams p = make_params(gamma, eps, stop_mode);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)N * S;
  const int grid = (int)((total + 127) / 128 < (int64_t)kNumSMs * 32 ? (total + 127) / 128 : (int64_t)kNumSMs * 32);
  if (S == 2 && A == 2) {
    whittle_small_kernel<2, 2><<<grid, 128, 0, st>>>(N, T, R, C, p, lam_lo, lam_hi, n_rounds, index);
    return check_launch("rmab_whittle_bisect<2,2>");
  }
  if (S == 3 && A == 2) {
    whittle_small_kernel<3, 2><<<grid, 128, 0, st>>>(N, T, R, C, p, lam_lo, lam_hi, n_rounds, index);
    return check_launch("rmab_whittle_bisect<3,2>");
  }
  return launch_generic(N, S, A, S, T, R, C, nullptr, 0, p, 1, lam_lo, lam_hi, n_rounds, nullptr, nullptr, nullptr, nullptr,
                        nullptr, index, st, "rmab_whittle_bisect");
}

extern "C" int rmab_vi_general(int N, int S, int A, const double* T, const double* Rsa, double gamma, double eps,
                               int stop_mode, double* V, uint8_t* policy, int32_t* iters, int32_t* max_iter,
                               rmab_stream_t stream) {
  RMAB_REQUIRE(N >= 0 && S >= 1 && S <= 256 && A >= 1 && A <= 16, RMAB_E_SHAPE, "rmab_vi_general: bad shape");
  RMAB_REQUIRE(T && Rsa && V, RMAB_E_BADARG, "rmab_vi_general: null pointer");
  RMAB_REQUIRE(gamma > 0.0 && gamma < 1.0 && eps > 0.0, RMAB_E_BADARG, "rmab_vi_general: need 0 < gamma < 1 and eps > 0");
  RMAB_REQUIRE(stop_mode == RMAB_STOP_FULL || stop_mode == RMAB_STOP_FAST || stop_mode == RMAB_STOP_ABS, RMAB_E_BADARG,
               "rmab_vi_general: bad stop_mode");
  if (N == 0) return RMAB_OK;
  const ViParams p = make_params(gamma, eps, stop_mode);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)(((int64_t)N + 127) / 128 < (int64_t)kNumSMs * 32 ? ((int64_t)N + 127) / 128 : (int64_t)kNumSMs * 32);
  if (S == 2 && A == 2) {
    vi_small_kernel<2, 2><<<grid, 128, 0, st>>>(N, 1, T, nullptr, nullptr, nullptr, 0, p, V, policy, iters, max_iter, nullptr,
                                                Rsa);
    return check_launch("rmab_vi_general<2,2>");
  }
  if (S == 3 && A == 2) {
    vi_small_kernel<3, 2><<<grid, 128, 0, st>>>(N, 1, T, nullptr, nullptr, nullptr, 0, p, V, policy, iters, max_iter, nullptr,
                                                Rsa);
    return check_launch("rmab_vi_general<3,2>");
  }
  return launch_generic(N, S, A, 1, T, nullptr, nullptr, nullptr, 0, p, 0, 0.0, 0.0, 1, V, policy, iters, max_iter, nullptr,
                        nullptr, st, "rmab_vi_general", Rsa);
}
