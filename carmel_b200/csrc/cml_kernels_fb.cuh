// cml_kernels_fb.cuh -- forward / backward / expected-count kernels over layered-CSR trellises.
//
// Follows (does not copy) the reference's per-example E-step:
//   forward   propagate_paths_in_order over the DFS order   graehl/shared/graph.h:391-402,
//                                                           carmel/src/derivations.h:406-408
//   backward  same on the reversed graph                    derivations.h:410-412
//   counts    counts[id] += w*alpha[src]*beta[dst]*exw/P    derivations.h:439-447
// B200 mapping: one example per thread group (a warp, or a whole CTA for wide lattices); the
// group walks the example's topological levels in order, every thread owns destination states of
// the current level and PULLS over that state's incoming arcs (8-byte {src,id} records streamed
// from HBM exactly once per pass, coalesced along the arc array), state scores alpha/beta live in
// shared memory for the whole example so they never touch HBM.  Backward and the count
// accumulation are fused (one pass over the outgoing-arc CSR).
//
// Two arithmetic spaces (cml_space):
//   LOG     scores are natural logs, (+) is an online max-shifted log-sum-exp.  Same semiring as
//           the reference's logweight (graehl/shared/weight.h:737-801) up to summation order.
//   SCALED  scores are linear with a per-level power-of-two scale (alpha_hat = alpha * 2^E[level]),
//           renormalised whenever a level's maximum drifts out of a safe exponent window, so any
//           length of example keeps full mantissa precision; (+) is one FMA per arc.
#pragma once
#include "cml_common.cuh"

namespace cmlk {

template <typename Real>
struct Num;
template <>
struct Num<double> {
  static __device__ __forceinline__ double ninf() { return -CUDART_INF; }
  static __device__ __forceinline__ double ex(double x) { return exp(x); }
  static __device__ __forceinline__ double lg(double x) { return log(x); }
  static __device__ __forceinline__ int expo(double x) { return (__double2hiint(x) >> 20) & 0x7ff; }
  static constexpr int kBias = 1023;
  static constexpr int kLo = 1023 - 400;   // renormalise when the level max is below 2^-400
  static constexpr int kHi = 1023 + 400;
  static __device__ __forceinline__ double scale2(double x, int k) { return scalbn(x, k); }
};
template <>
struct Num<float> {
  static __device__ __forceinline__ float ninf() { return -CUDART_INF_F; }
  static __device__ __forceinline__ float ex(float x) { return __expf(x); }
  static __device__ __forceinline__ float lg(float x) { return __logf(x); }
  static __device__ __forceinline__ int expo(float x) { return (__float_as_int(x) >> 23) & 0xff; }
  static constexpr int kBias = 127;
  static constexpr int kLo = 127 - 24;  // keep every level's max within 2^-24 .. 2^24
  static constexpr int kHi = 127 + 24;
  static __device__ __forceinline__ float scale2(float x, int k) { return scalbnf(x, k); }
};

// Expected-count accumulation.  A slot code is either 0xFFFFFFFF (the arc feeds no trainable
// parameter), a plain slot index into counts[], or 0x80000000|h for a HOT slot: slots that occur many
// thousand times in the corpus serialise in the L2 atomic unit (measured: ~2.5 ns per RED on one
// address, 80% of the sweep's time), so they get kHotCopies replicas spread by block index and are folded
// into counts[] by k_fold_hot afterwards.
constexpr uint32_t kSlotNone = 0xFFFFFFFFu;
constexpr uint32_t kSlotHot = 0x80000000u;
constexpr uint32_t kHotCopies = 64;  // default; cml_ctx::hot_copies (a power of two, CML_HOT_COPIES) is what the buffers are sized for
struct CountSink {
  double* counts;   // [n_slots]
  double* hot;      // [copies][n_hot]
  uint32_t n_hot;
  uint32_t mask;    // copies - 1
};
__device__ __forceinline__ void count_add(const CountSink& S, uint32_t code, double v) {
#ifdef CML_DEBUG_NO_COUNTS  // profiling experiment: how long is the sweep without its REDs?
  if (v < 1e300) return;
#endif
  if (code & kSlotHot)
    atomicAdd(S.hot + (size_t)(blockIdx.x & S.mask) * S.n_hot + (code & 0x7fffffffu), v);
  else
    atomicAdd(S.counts + code, v);
}
// fold the hot replicas into the count table (one thread per hot slot)
static __global__ void k_fold_hot(uint32_t n_hot, const uint32_t* __restrict__ hot_slot, const double* __restrict__ hot,
                           double* __restrict__ counts, uint32_t copies) {
  const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= n_hot) return;
  double s = 0;
#pragma unroll 8
  for (uint32_t k = 0; k < copies; ++k) s += hot[(size_t)k * n_hot + h];
  if (s != 0.) atomicAdd(&counts[hot_slot[h]], s);
}

struct FbArgs {
  const CmlExDesc* desc;      // per example
  const uint32_t* ex_list;    // examples handled by this launch (indices into desc)
  uint32_t n_list;
  const uint32_t* lvl_off;
  const uint32_t* in_off;
  const uint2* in_arc;        // {src layered index, arc id}
  const uint32_t* out_off;
  const uint2* out_arc;       // {dst layered index, arc id}
  const void* arc_w;          // Real[n_arcs]: ln w (LOG) or w (SCALED)
  const uint32_t* arc_slot;   // [n_arcs] count slot code of every arc (see CountSink)
  CountSink sink;             // linear expected counts (fp64 RED)
  double* ex_lnp;             // [n_ex in batch] ln P_e
  void* scratch;              // GLOBAL class: Real alpha/beta slots, 2 per state
  int* scratch_lvl;           // GLOBAL class, SCALED: E/F per level (2 per level)
  uint32_t cap_states;        // shared-memory state capacity per group (WARP / CTA classes)
};

// ---------------------------------------------------------------------------------------------
// group helpers: a "group" is one warp (CTA == false) or the whole CTA (CTA == true)
// ---------------------------------------------------------------------------------------------
template <bool CTA>
__device__ __forceinline__ void group_sync() {
  if (CTA)
    __syncthreads();
  else
    __syncwarp();
}

// max over the group of a per-thread int; smax = 2 ints of shared memory (CTA only), par = parity
template <bool CTA>
__device__ __forceinline__ int group_max_and_sync(int v, int* smax, int par) {
  v = __reduce_max_sync(0xffffffffu, v);
  if (!CTA) {
    __syncwarp();
    return v;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(&smax[par], v);
  __syncthreads();
  int r = smax[par];
  if (threadIdx.x == 0) smax[par ^ 1] = INT_MIN;  // ready for the next level (ordered by the next sync)
  return r;
}

// ---------------------------------------------------------------------------------------------
// LOG space
// ---------------------------------------------------------------------------------------------
template <typename Real, bool CTA>
__device__ void fb_example_log(const FbArgs& A, const CmlExDesc& d, Real* __restrict__ al, Real* __restrict__ be,
                               int lane, int G) {
  const Real* __restrict__ w = (const Real*)A.arc_w;
  const uint32_t* __restrict__ lvl = A.lvl_off + d.lvl_base;
  const uint32_t* __restrict__ ioff = A.in_off + d.row_base;
  const uint32_t* __restrict__ ooff = A.out_off + d.row_base;
  const uint2* __restrict__ iarc = A.in_arc + d.arc_base;
  const uint2* __restrict__ oarc = A.out_arc + d.arc_base;
  const Real NI = Num<Real>::ninf();
  const uint32_t n = d.n_states, nl = d.n_levels;

  // level 0 (states without predecessors): the start state (layered index 0) has alpha = 1
  for (uint32_t s = lane; s < lvl[1]; s += G) al[s] = (s == 0) ? Real(0) : NI;
  group_sync<CTA>();
  for (uint32_t L = 1; L < nl; ++L) {
    const uint32_t s1 = lvl[L + 1];
    for (uint32_t s = lvl[L] + lane; s < s1; s += G) {
      uint32_t k = ioff[s];
      const uint32_t k1 = ioff[s + 1];
      Real m = NI, acc = 0;
      for (; k < k1; ++k) {
        const uint2 r = __ldg(&iarc[k]);
        const Real v = al[r.x] + __ldg(&w[r.y]);
        if (v > m) {
          acc = acc * Num<Real>::ex(m - v) + Real(1);
          m = v;
        } else if (v > NI) {
          acc += Num<Real>::ex(v - m);
        }
      }
      al[s] = (m > NI) ? m + Num<Real>::lg(acc) : NI;
    }
    group_sync<CTA>();
  }
  const Real lnP = al[d.fin];
  if (lane == 0) A.ex_lnp[d.ex_index] = (double)lnP;
  if (!(lnP > NI)) return;  // zero-probability example: contributes no counts (uniform branch)
  const Real cbase = (Real)d.ln_weight - lnP;

  // backward, fused with counts.  beta[fin] = 1; every other state pulls over its outgoing arcs.
  for (int L = (int)nl - 1; L >= 0; --L) {
    const uint32_t s1 = lvl[L + 1];
    for (uint32_t s = lvl[L] + lane; s < s1; s += G) {
      uint32_t k = ooff[s];
      const uint32_t k1 = ooff[s + 1];
      Real m = NI, acc = 0;
      if (s == d.fin) {
        m = 0;
        acc = 1;
      }
      const Real as = al[s] + cbase;
      for (; k < k1; ++k) {
        const uint2 r = __ldg(&oarc[k]);
        const Real v = __ldg(&w[r.y]) + be[r.x];
        if (v > m) {
          acc = acc * Num<Real>::ex(m - v) + Real(1);
          m = v;
        } else if (v > NI) {
          acc += Num<Real>::ex(v - m);
        }
        const Real lc = as + v;
        if (lc > Real(-700)) {
          const double c = (double)Num<Real>::ex(lc);
          const uint32_t slot = __ldg(&A.arc_slot[r.y]);
          if (c > 0 && slot != kSlotNone) count_add(A.sink, slot, c);
        }
      }
      be[s] = (m > NI) ? m + Num<Real>::lg(acc) : NI;
    }
    group_sync<CTA>();
  }
  (void)n;
}

// ---------------------------------------------------------------------------------------------
// SCALED space
// ---------------------------------------------------------------------------------------------
// alpha_hat[s] = alpha[s] * 2^E[level(s)] ; beta_hat[s] = beta[s] * 2^F[level(s)].
// lvl_min_src[L] / lvl_max_dst[L] (stored interleaved after the level offsets by the flattener)
// bound the levels a level's arcs reach, so the common case (no renormalisation inside that
// window) needs no per-arc exponent work at all.
template <typename Real, bool CTA>
__device__ void fb_example_scaled(const FbArgs& A, const CmlExDesc& d, Real* __restrict__ al, Real* __restrict__ be,
                                  int* __restrict__ E, int* __restrict__ F, int* smax, int lane, int G) {
  const Real* __restrict__ w = (const Real*)A.arc_w;
  const uint32_t* __restrict__ lvl = A.lvl_off + d.lvl_base;
  const uint32_t nl = d.n_levels;
  const uint32_t* __restrict__ lvl_min_src = lvl + (nl + 1);      // [nl]
  const uint32_t* __restrict__ lvl_max_dst = lvl + (nl + 1) + nl;  // [nl]
  const uint32_t* __restrict__ ioff = A.in_off + d.row_base;
  const uint32_t* __restrict__ ooff = A.out_off + d.row_base;
  const uint2* __restrict__ iarc = A.in_arc + d.arc_base;
  const uint2* __restrict__ oarc = A.out_arc + d.arc_base;

  auto level_of = [&](uint32_t s) -> uint32_t {  // binary search (slow path only)
    uint32_t lo = 0, hi = nl;
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (lvl[mid] <= s)
        lo = mid;
      else
        hi = mid;
    }
    return lo;
  };

  for (uint32_t s = lane; s < lvl[1]; s += G) al[s] = (s == 0) ? Real(1) : Real(0);
  if (lane == 0) E[0] = 0;
  if (CTA && threadIdx.x == 0) smax[0] = smax[1] = INT_MIN;
  group_sync<CTA>();
  int par = 0;
  uint32_t last_event = 0;  // highest level j with E[j] != E[j-1]
  for (uint32_t L = 1; L < nl; ++L) {
    const uint32_t s0 = lvl[L], s1 = lvl[L + 1];
    const int Eprev = E[L - 1];
    const bool uniform = (last_event <= lvl_min_src[L]);
    int mx = 0;
    for (uint32_t s = s0 + lane; s < s1; s += G) {
      uint32_t k = ioff[s];
      const uint32_t k1 = ioff[s + 1];
      Real acc = 0;
      if (uniform) {
        for (; k < k1; ++k) {
          const uint2 r = __ldg(&iarc[k]);
          acc = fma(al[r.x], __ldg(&w[r.y]), acc);
        }
      } else {
        for (; k < k1; ++k) {
          const uint2 r = __ldg(&iarc[k]);
          const int de = Eprev - E[level_of(r.x)];  // bring the source up to the previous level's scale
          acc += Num<Real>::scale2(al[r.x] * __ldg(&w[r.y]), de);
        }
      }
      al[s] = acc;
      mx = max(mx, Num<Real>::expo(acc));
    }
    mx = group_max_and_sync<CTA>(mx, smax, par);
    par ^= 1;
    int shift = 0;
    if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) shift = Num<Real>::kBias - mx;
    if (shift != 0) {
      for (uint32_t s = s0 + lane; s < s1; s += G) al[s] = Num<Real>::scale2(al[s], shift);
    }
    if (shift != 0) last_event = L;
    if (lane == 0) E[L] = Eprev + shift;
    group_sync<CTA>();
  }
  const uint32_t Lfin = level_of(d.fin);
  const Real afin = al[d.fin];
  const int Efin = E[Lfin];
  const double lnP = (afin > 0) ? log((double)afin) - (double)Efin * 0.69314718055994530942 : -CUDART_INF;
  if (lane == 0) A.ex_lnp[d.ex_index] = lnP;
  if (!(afin > 0)) return;
  const double cw = d.weight / (double)afin;

  // backward + counts
  if (CTA && threadIdx.x == 0) smax[0] = smax[1] = INT_MIN;
  group_sync<CTA>();
  par = 0;
  uint32_t last_event_b = 0xFFFFFFFFu;  // lowest level j with F[j] != F[j+1]
  for (int L = (int)nl - 1; L >= 0; --L) {
    const uint32_t s0 = lvl[L], s1 = lvl[L + 1];
    const bool last = (L == (int)nl - 1);
    const int Fnext = last ? 0 : F[L + 1];
    const bool uniform = last || (last_event_b >= lvl_max_dst[L]);
    // count scale for this level: 2^(Efin - E[L] - Fnext) * exw / alpha_hat[fin]
    const double cs = scalbn(cw, Efin - E[L] - Fnext);
    int mx = 0;
    for (uint32_t s = s0 + lane; s < s1; s += G) {
      uint32_t k = ooff[s];
      const uint32_t k1 = ooff[s + 1];
      Real acc = (s == d.fin) ? Num<Real>::scale2(Real(1), Fnext) : Real(0);
      const double as = (double)al[s] * cs;
      if (uniform) {
        for (; k < k1; ++k) {
          const uint2 r = __ldg(&oarc[k]);
          const Real t = __ldg(&w[r.y]) * be[r.x];
          acc += t;
          const double c = as * (double)t;
          const uint32_t slot = __ldg(&A.arc_slot[r.y]);
          if (c > 0 && slot != kSlotNone) count_add(A.sink, slot, c);
        }
      } else {
        for (; k < k1; ++k) {
          const uint2 r = __ldg(&oarc[k]);
          const int de = Fnext - F[level_of(r.x)];
          const Real t = Num<Real>::scale2(__ldg(&w[r.y]) * be[r.x], de);
          acc += t;
          const double c = as * (double)t;
          const uint32_t slot = __ldg(&A.arc_slot[r.y]);
          if (c > 0 && slot != kSlotNone) count_add(A.sink, slot, c);
        }
      }
      be[s] = acc;
      mx = max(mx, Num<Real>::expo(acc));
    }
    mx = group_max_and_sync<CTA>(mx, smax, par);
    par ^= 1;
    int shift = 0;
    if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) shift = Num<Real>::kBias - mx;
    if (shift != 0) {
      for (uint32_t s = s0 + lane; s < s1; s += G) be[s] = Num<Real>::scale2(be[s], shift);
    }
    if (shift != 0) last_event_b = (uint32_t)L;
    if (lane == 0) F[L] = Fnext + shift;
    group_sync<CTA>();
  }
}

// ---------------------------------------------------------------------------------------------
// kernels.  WARP: blockDim = 128 (4 examples per CTA), alpha/beta in shared memory.
//           CTA : blockDim = 256, one example per CTA, alpha/beta in shared memory.
//           GLOBAL: blockDim = 256, one example per CTA, alpha/beta in an HBM/L2 scratch.
// Shared memory per group: cap*2 Reals (+ 2*cap ints of level exponents in SCALED space).
// ---------------------------------------------------------------------------------------------
template <typename Real, bool SCALED>
static __global__ void __launch_bounds__(128) k_fb_warp(FbArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t li = blockIdx.x * 4 + warp;
  if (li >= A.n_list) return;
  const CmlExDesc d = A.desc[A.ex_list[li]];
  const size_t per = (size_t)A.cap_states * (2 * sizeof(Real) + (SCALED ? 2 * sizeof(int) : 0));
  unsigned char* base = smem + per * warp;
  Real* al = (Real*)base;
  Real* be = al + A.cap_states;
  if (SCALED) {
    int* E = (int*)(be + A.cap_states);
    int* F = E + A.cap_states;
    fb_example_scaled<Real, false>(A, d, al, be, E, F, nullptr, lane, 32);
  } else {
    fb_example_log<Real, false>(A, d, al, be, lane, 32);
  }
}

template <typename Real, bool SCALED, bool GLOBAL>
static __global__ void __launch_bounds__(256) k_fb_cta(FbArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int smax[2];
  const uint32_t li = blockIdx.x;
  if (li >= A.n_list) return;
  const CmlExDesc d = A.desc[A.ex_list[li]];
  Real *al, *be;
  int *E = nullptr, *F = nullptr;
  if (GLOBAL) {
    al = (Real*)A.scratch + 2 * d.scratch_base;
    be = al + d.n_states;
    if (SCALED) {
      E = A.scratch_lvl + 2 * d.scratch_base;  // n_levels <= n_states
      F = E + d.n_states;
    }
  } else {
    al = (Real*)smem;
    be = al + A.cap_states;
    if (SCALED) {
      E = (int*)(be + A.cap_states);
      F = E + A.cap_states;
    }
  }
  if (SCALED)
    fb_example_scaled<Real, true>(A, d, al, be, E, F, smax, threadIdx.x, blockDim.x);
  else
    fb_example_log<Real, true>(A, d, al, be, threadIdx.x, blockDim.x);
}

// Sum of ln P_e, w_e ln P_e and the zero-probability count over a batch (deterministic per block,
// one atomicAdd triple per block into the reduce buffer's scalar tail).
// ---------------------------------------------------------------------------------------------
// Lattices with a cycle (CLS_CYCLIC).  The reference does not solve the cyclic system: it walks the states once in the
// reverse post-order of its DFS (forward) and once in post-order over the reversed graph (backward), warns that
// "Forward/backward will miss some paths" (derivations.h:400-417,722-729; graph.h:241-288,391-402), and uses whatever
// alpha / beta that leaves for the counts.  To give the same numbers the walk must be the same SEQUENCE of updates, so
// one thread owns a lattice: layered index j = rank in that order (cml_add_trellises), out-lists in stored order.
// Natural-log fp64 whatever the context's space / precision (these lattices are rare and small).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double cyc_lse(double a, double b) {
  if (!(a > -CUDART_INF)) return b;
  if (!(b > -CUDART_INF)) return a;
  const double hi = fmax(a, b), d = -fabs(a - b);
  return d < -36. ? hi : hi + log1p(exp(d));  // logweight operator+ (weight.h:765-801: 36-nat cutoff)
}
template <typename Real, bool SCALED>
static __global__ void __launch_bounds__(64) k_fb_cyclic(FbArgs A, double* __restrict__ scratch) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.n_list) return;
  const CmlExDesc d = A.desc[A.ex_list[t]];
  const Real* __restrict__ w = (const Real*)A.arc_w;
  const uint32_t* __restrict__ ioff = A.in_off + d.row_base;
  const uint32_t* __restrict__ ooff = A.out_off + d.row_base;
  const uint2* __restrict__ iarc = A.in_arc + d.arc_base;
  const uint2* __restrict__ oarc = A.out_arc + d.arc_base;
  double* __restrict__ f = scratch + 2 * d.scratch_base;
  double* __restrict__ bw = f + d.n_states;
  const uint32_t n = d.n_states;
  auto lnw = [&](uint32_t id) -> double { return SCALED ? log((double)w[id]) : (double)w[id]; };
  for (uint32_t j = 0; j < n; ++j) f[j] = bw[j] = -CUDART_INF;
  f[0] = 0.;  // the start state is first in the order
  for (uint32_t j = 0; j < n; ++j) {
    const double fj = f[j];
    if (!(fj > -CUDART_INF)) continue;
    for (uint32_t k = ooff[j]; k < ooff[j + 1]; ++k) {
      const uint2 a = oarc[k];
      f[a.x] = cyc_lse(f[a.x], fj + lnw(a.y));
    }
  }
  const double P = f[d.fin];
  A.ex_lnp[d.ex_index] = P;
  if (!(P > -CUDART_INF)) return;
  bw[d.fin] = 0.;
  for (uint32_t j = n; j-- > 0;) {
    const double bj = bw[j];
    if (!(bj > -CUDART_INF)) continue;
    for (uint32_t k = ioff[j]; k < ioff[j + 1]; ++k) {
      const uint2 a = iarc[k];
      bw[a.x] = cyc_lse(bw[a.x], bj + lnw(a.y));
    }
  }
  for (uint32_t j = 0; j < n; ++j) {
    const double fj = f[j];
    if (!(fj > -CUDART_INF)) continue;
    for (uint32_t k = ooff[j]; k < ooff[j + 1]; ++k) {
      const uint2 a = oarc[k];
      const uint32_t code = A.arc_slot[a.y];
      if (code == kSlotNone) continue;
      const double c = exp(lnw(a.y) + fj + bw[a.x] - P) * d.weight;
      if (c > 0) count_add(A.sink, code, c);
    }
  }
}

static __global__ void __launch_bounds__(256) k_reduce_lnp(const double* __restrict__ ex_lnp, const double* __restrict__ ex_weight,
                                                    uint64_t n, double* __restrict__ scal) {
  double s0 = 0, s1 = 0, nz = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const double lp = ex_lnp[i];
    if (lp > -CUDART_INF) {
      s0 += lp;
      s1 += ex_weight[i] * lp;
    } else
      nz += 1;
  }
  for (int o = 16; o; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
  }
  __shared__ double sh[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sh[0][warp] = s0;
    sh[1][warp] = s1;
    sh[2][warp] = nz;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int i = 0; i < 8; ++i) {
      a += sh[0][i];
      b += sh[1][i];
      c += sh[2][i];
    }
    atomicAdd(&scal[0], a);
    atomicAdd(&scal[1], b);
    atomicAdd(&scal[2], c);
  }
}

}  // namespace cmlk
