// cml_kernels_ell.cuh -- the throughput E-step kernel: forward + backward + expected counts over
// level-sliced ELL lattices, scaled linear space.  (Same reference semantics as cml_kernels_fb.cuh:
// derivations.h:400-449, graph.h:391-402.)
//
// Why this shape (measured: profiles/r1_cipher_v1_*): the first kernel kept whole-example alpha/beta
// vectors in shared memory, which capped occupancy at 4 warps/SM and left it latency bound at 5% of
// the HBM roofline.  Here
//   * a GROUP of G = 4..32 lanes (a sub-warp) owns one example, so narrow lattices (HMM: 4 states per
//     position) still fill the warp; 256-thread CTAs hold 8..64 examples;
//   * state scores live in a small shared-memory RING indexed by (layered state index mod ring): a
//     level only ever reads sources at most `ring` indices back, so shared memory per example is a
//     few hundred bytes and occupancy is limited by registers, not by lattice size;
//   * alpha is also streamed to HBM once (coalesced) in the forward sweep and read once (own state)
//     in the backward sweep -- the "alpha written once + read once, beta on chip" traffic model;
//   * arcs of one level are stored column-major (ELL): column j holds the j-th incoming (outgoing)
//     arc of every state of the level, so lane r reads record [j][r] and a group's loads are contiguous
//     32..256-byte segments; short rows are padded with a zero-weight arc;
//   * every lane owns a row and PULLS (no atomics on state scores, deterministic sums); one
//     __syncwarp(group) per level;
//   * expected counts: the backward sweep forms c = alpha[src]*w*beta[dst]/P per arc; when every
//     column of a level feeds a single count slot (e.g. all arcs into one channel parameter) the group
//     reduces c by shuffles and issues ONE fp64 RED per column, otherwise one RED per lane.
#pragma once
#include "cml_common.cuh"
#include "cml_kernels_fb.cuh"

namespace cmlk {

struct __align__(16) EllDesc {
  uint64_t in_base;     // first record of this example in ell_in
  uint64_t out_base;    // first record in ell_out
  uint64_t meta_base;   // first per-level meta entry (n_levels entries)
  uint64_t state_base;  // first slot in the global alpha array
  uint64_t level_base;  // first level slot in the global per-level exponent scratch
  uint32_t n_states, n_levels, fin, ex_index;
  double weight;
  uint32_t fin_level, pad;
};

// per-level meta (uint4):
//   x = offset of the level's incoming ELL block      y = offset of its outgoing ELL block
//   z = layered index of the level's first state | (aggregate flag << 31)
//   w = width | D << 8 | O << 16 | min_src_delta << 24 | max_dst_delta << 28
// aggregate flag: every outgoing column of the level feeds one count slot and has no padding.
struct EllArgs {
  const EllDesc* desc;
  const uint32_t* ex_list;
  uint32_t n_list;
  const uint4* lvl_meta;
  const uint2* ell_in;   // {src layered index, arc id}
  const uint2* ell_out;  // {dst layered index, arc id}
  const void* arc_w;     // Real[n_arcs + 1] linear weights, last entry = 0 (padding arc)
  const void* arc_ws;    // {Real w; uint32 slot}[n_arcs + 1]
  double* counts;        // [n_slots]
  double* ex_lnp;
  void* alpha_g;         // Real per state
  int* lvl_exp;          // 2 ints per level: E block then F block per example
  uint32_t ring;         // shared-memory ring entries per group (power of two)
};

template <typename Real>
struct WS;
template <>
struct __align__(16) WS<double> {
  double w;
  uint32_t slot, pad;
};
template <>
struct __align__(8) WS<float> {
  float w;
  uint32_t slot;
};

template <int G>
__device__ __forceinline__ int group_max(int v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o; o >>= 1) v = max(v, __shfl_xor_sync(mask, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
  for (int o = G / 2; o; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

constexpr int kEllMaxRows = 4;  // rows per lane: level width <= G * kEllMaxRows (flattener enforces)

template <typename Real, int G>
__global__ void __launch_bounds__(256) k_fb_ell(EllArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int GPB = 256 / G;  // groups per block
  const int gib = threadIdx.x / G, lane = threadIdx.x % G;
  const uint32_t li = blockIdx.x * GPB + gib;
  if (li >= A.n_list) return;  // whole group leaves
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
  const EllDesc d = A.desc[A.ex_list[li]];
  Real* ring = (Real*)smem_raw + (size_t)gib * A.ring;
  const uint32_t M = A.ring - 1;
  const Real* __restrict__ w = (const Real*)A.arc_w;
  const WS<Real>* __restrict__ ws = (const WS<Real>*)A.arc_ws;
  const uint4* __restrict__ meta = A.lvl_meta + d.meta_base;
  const uint2* __restrict__ ein = A.ell_in + d.in_base;
  const uint2* __restrict__ eout = A.ell_out + d.out_base;
  Real* __restrict__ ag = (Real*)A.alpha_g + d.state_base;
  int* __restrict__ E = A.lvl_exp + 2 * d.level_base;
  int* __restrict__ F = E + d.n_levels;
  const uint32_t nl = d.n_levels;

  // ---------------------------------------------------------------- forward
  {
    const uint32_t W0 = __ldg(&meta[0]).w & 0xff;
    for (uint32_t r = lane; r < W0; r += G) {
      const Real v = (r == 0) ? Real(1) : Real(0);
      ring[r & M] = v;
      ag[r] = v;
    }
    if (lane == 0) E[0] = 0;
  }
  __syncwarp(gmask);
  int Eprev = 0;
  uint32_t last_event = 0;  // highest level whose scale differs from the level before it
  for (uint32_t L = 1; L < nl; ++L) {
    const uint4 m = __ldg(&meta[L]);
    const uint32_t W = m.w & 0xff, D = (m.w >> 8) & 0xff, min_src = L - ((m.w >> 24) & 0xf);
    const uint32_t s0 = m.z & 0x7fffffffu;
    const int nrow = (int)((W + G - 1) / G);
    const bool uniform = last_event <= min_src;
    Real acc[kEllMaxRows];
    int mx = 0;
#pragma unroll
    for (int i = 0; i < kEllMaxRows; ++i) {
      acc[i] = 0;
      if (i < nrow) {  // group-uniform; lanes past the width recompute the last row (result unused)
        const uint32_t rr = min((uint32_t)(lane + i * G), W - 1);
        const uint2* p = ein + m.x + rr;
        Real a = 0;
        if (uniform) {
#pragma unroll 4
          for (uint32_t j = 0; j < D; ++j) {
            const uint2 rec = __ldg(p + (size_t)j * W);
            a = fma(ring[rec.x & M], __ldg(&w[rec.y]), a);
          }
        } else {  // a source level inside the window carries another power-of-two scale
          for (uint32_t j = 0; j < D; ++j) {
            const uint2 rec = __ldg(p + (size_t)j * W);
            uint32_t ls = L - 1;
            while ((__ldg(&meta[ls]).z & 0x7fffffffu) > rec.x) --ls;
            a += Num<Real>::scale2(ring[rec.x & M] * __ldg(&w[rec.y]), Eprev - E[ls]);
          }
        }
        acc[i] = a;
        mx = max(mx, Num<Real>::expo(a));
      }
    }
    mx = group_max<G>(mx, gmask);
    int shift = 0;
    if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) shift = Num<Real>::kBias - mx;
    // No barrier between the reads above and the writes below: the ring is sized (flattener) so that
    // this level's slots never alias a source slot of the same level.
#pragma unroll
    for (int i = 0; i < kEllMaxRows; ++i) {
      const uint32_t r = lane + i * G;
      if (i < nrow && r < W) {
        const Real v = shift ? Num<Real>::scale2(acc[i], shift) : acc[i];
        ring[(s0 + r) & M] = v;
        ag[s0 + r] = v;
      }
    }
    if (shift) last_event = L;
    Eprev += shift;
    if (lane == 0) E[L] = Eprev;
    __syncwarp(gmask);
  }
  const Real afin = ag[d.fin];
  const int Efin = E[d.fin_level];
  const double lnP = (afin > 0) ? log((double)afin) - (double)Efin * 0.69314718055994530942 : -CUDART_INF;
  if (lane == 0) A.ex_lnp[d.ex_index] = lnP;
  if (!(afin > 0)) return;  // zero-probability example (group-uniform)
  const double cw = d.weight / (double)afin;

  // ---------------------------------------------------------------- backward + counts
  int Fnext = 0;
  uint32_t last_event_b = 0xFFFFFFFFu;  // lowest level whose beta scale differs from the level after it
  for (int L = (int)nl - 1; L >= 0; --L) {
    const uint4 m = __ldg(&meta[L]);
    const uint32_t W = m.w & 0xff, O = (m.w >> 16) & 0xff, max_dst = L + ((m.w >> 28) & 0xf);
    const uint32_t s0 = m.z & 0x7fffffffu;
    const bool aggregate = (m.z >> 31) != 0;
    const int nrow = (int)((W + G - 1) / G);
    const bool uniform = last_event_b >= max_dst;
    const double cs = scalbn(cw, Efin - E[L] - Fnext);
    Real acc[kEllMaxRows];
    int mx = 0;
#pragma unroll
    for (int i = 0; i < kEllMaxRows; ++i) {
      acc[i] = 0;
      if (i < nrow) {
        const uint32_t r = lane + i * G;
        const bool active = r < W;
        const uint32_t rr = min(r, W - 1);
        const uint32_t s = s0 + rr;
        const uint2* p = eout + m.y + rr;
        Real b = (s == d.fin) ? Num<Real>::scale2(Real(1), Fnext) : Real(0);
        const double as = active ? (double)ag[s] * cs : 0.;
#pragma unroll 2
        for (uint32_t j = 0; j < O; ++j) {
          const uint2 rec = __ldg(p + (size_t)j * W);
          const WS<Real> e = ws[rec.y];
          Real t = e.w * ring[rec.x & M];
          if (!uniform) {
            uint32_t ld = L + 1;
            while (ld + 1 < nl && (__ldg(&meta[ld + 1]).z & 0x7fffffffu) <= rec.x) ++ld;
            t = Num<Real>::scale2(t, Fnext - F[ld]);
          }
          b += t;
          double c = as * (double)t;
          if (aggregate) {  // the whole column feeds one slot: one RED per group
            c = group_sum<G>(c, gmask);
            if (lane == 0 && c > 0 && e.slot != 0xFFFFFFFFu) atomicAdd(&A.counts[e.slot], c);
          } else if (c > 0 && e.slot != 0xFFFFFFFFu) {
            atomicAdd(&A.counts[e.slot], c);
          }
        }
        acc[i] = b;
        if (active) mx = max(mx, Num<Real>::expo(b));
      }
    }
    mx = group_max<G>(mx, gmask);
    int shift = 0;
    if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) shift = Num<Real>::kBias - mx;
#pragma unroll
    for (int i = 0; i < kEllMaxRows; ++i) {
      const uint32_t r = lane + i * G;
      if (i < nrow && r < W) ring[(s0 + r) & M] = shift ? Num<Real>::scale2(acc[i], shift) : acc[i];
    }
    if (shift) last_event_b = (uint32_t)L;
    Fnext += shift;
    if (lane == 0) F[L] = Fnext;
    __syncwarp(gmask);
  }
}

}  // namespace cmlk
