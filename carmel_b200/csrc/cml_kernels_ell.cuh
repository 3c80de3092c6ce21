// cml_kernels_ell.cuh -- the throughput E-step kernel: forward + backward + expected counts over
// level-sliced ELL lattices, scaled linear space.  (Same reference semantics as cml_kernels_fb.cuh:
// derivations.h:400-449, graph.h:391-402.)
//
// Why this shape (measured, profiles/r1_*_v1.._v3): with whole-example alpha/beta vectors in shared
// memory the sweep ran at 4 warps/SM and 5% of the HBM roofline; with one lane per lattice state it was
// latency bound (long-scoreboard stalls, 10%): every level paid a chain of dependent HBM/L2 round trips
// (level meta -> arc records -> arc weights).  Here
//   * a GROUP of lanes owns one example: R row-lanes x C column-lanes (sub-warp groups of 4..32 lanes
//     for narrow lattices such as an HMM's 4 states per position, a whole 256-thread CTA for wide
//     dense ones such as a 27x27 cipher position); lane (r,c) owns state r of the current level and
//     columns c, c+C, .. of its arc list, so all arcs of a level are in flight at once;
//   * arcs of one level are stored column-major (ELL): column j holds the j-th incoming (outgoing) arc
//     of every state of the level; a group's loads are contiguous segments; short rows are padded with
//     a zero-weight arc;
//   * each lane streams ITS OWN records through a 4-deep cp.async (LDGSTS) pipeline into private
//     shared-memory slots, and gathers the arc weights of level L+1 into registers while level L is
//     being computed: the per-level critical path is shared-memory reads, FMAs, shuffles and one barrier;
//   * state scores live in a small shared-memory RING indexed by (layered state index mod ring); alpha
//     is also streamed to HBM once (coalesced) in the forward sweep and read once (own state) in the
//     backward sweep -- the "alpha written once + read once, beta on chip" traffic model;
//   * every row is a PULL (no atomics on state scores, deterministic sums), partial sums of the C
//     column-lanes are combined by shuffles;
//   * expected counts: the backward sweep forms c = alpha[src]*w*beta[dst]/P per arc; when every
//     column of a level feeds a single count slot (e.g. all arcs into one channel parameter) the rows are
//     reduced by shuffles and one fp64 RED per warp and column is issued, otherwise one RED per arc.
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
  uint32_t in_len, out_len;  // records of the example's incoming / outgoing stream (wide kernel: bulk-copy bounds)
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
  CountSink sink;        // expected counts (see cml_kernels_fb.cuh)
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

constexpr int kEllStages = 4;  // levels of topology in flight per lane
constexpr int kEllSlots = 4;   // records per lane per level: ceil(max(D,O)/C) <= kEllSlots, width <= R
constexpr int kEllThreads = 256;
constexpr uint32_t kEllSlotStride = kEllThreads * sizeof(uint2);  // bytes between consecutive slots of a lane

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async8(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint2 lds_u2(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}

// sum over the C column-lanes of a row (lanes t, t^1, .. within a C-aligned block)
template <int C, typename T>
__device__ __forceinline__ T col_sum(T v, unsigned mask) {
#pragma unroll
  for (int o = C / 2; o; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
// reduce over the row-lanes of a group that live in the same warp (lanes differing in bits >= log2 C)
template <int C, int GW, typename T, typename Op>
__device__ __forceinline__ T row_reduce(T v, unsigned mask, Op op) {
#pragma unroll
  for (int o = GW / 2; o >= C; o >>= 1) v = op(v, __shfl_xor_sync(mask, v, o));
  return v;
}

// R row-lanes x C column-lanes per example (level width <= R).  CTA: R*C == 256, one example per block.
template <typename Real, int R, int C, bool CTA>
static __global__ void __launch_bounds__(kEllThreads) k_fb_ell(EllArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int smax[3];
  constexpr int G = R * C;
  constexpr int GW = (G < 32) ? G : 32;           // lanes of the group inside one warp
  constexpr int GPB = CTA ? 1 : kEllThreads / G;  // groups per block
  const int t = CTA ? threadIdx.x : threadIdx.x % G;  // lane within the group
  const int gib = CTA ? 0 : threadIdx.x / G;
  const uint32_t li = blockIdx.x * GPB + gib;
  if (li >= A.n_list) return;  // whole group leaves (CTA: whole block)
  const unsigned gmask = (GW == 32) ? 0xffffffffu : (((1u << GW) - 1u) << (((threadIdx.x & 31) / GW) * GW));
  const uint32_t c = t % C, r = t / C;
  const EllDesc d = A.desc[A.ex_list[li]];
  // shared memory: per-thread record slots [stage][slot][thread], then the rings
  const uint32_t sbase = smem_u32(smem_raw) + threadIdx.x * (uint32_t)sizeof(uint2);
  Real* ring = (Real*)(smem_raw + (size_t)kEllStages * kEllSlots * kEllSlotStride) + (size_t)gib * A.ring;
  const uint32_t M = A.ring - 1;
  const Real* __restrict__ w = (const Real*)A.arc_w;
  const WS<Real>* __restrict__ ws = (const WS<Real>*)A.arc_ws;
  const uint4* __restrict__ meta = A.lvl_meta + d.meta_base;
  const uint2* __restrict__ ein = A.ell_in + d.in_base;
  const uint2* __restrict__ eout = A.ell_out + d.out_base;
  Real* __restrict__ ag = (Real*)A.alpha_g + d.state_base;
  int* __restrict__ E = A.lvl_exp + 2 * d.level_base;
  int* __restrict__ F = E + d.n_levels;
  const int nl = (int)d.n_levels;
  auto group_sync = [&]() {
    if (CTA)
      __syncthreads();
    else
      __syncwarp(gmask);
  };
  // queue this lane's records of one level block (width W, `deg` columns starting at `base`) into stage st
  auto issue = [&](const uint2* base, uint32_t W, uint32_t deg, int st) {
    if (r < W) {
      const uint2* p = base + c * W + r;
      uint32_t dst = sbase + (uint32_t)st * (kEllSlots * kEllSlotStride);
#pragma unroll
      for (int k = 0; k < kEllSlots; ++k) {
        if (c + k * C < deg) cp_async8(dst, p);
        p += C * W;
        dst += kEllSlotStride;
      }
    }
  };
  // registers of one level: the lane's (<= kEllSlots) arcs with their weights
  uint32_t idx[kEllSlots];   // layered index of the other endpoint
  Real wv[kEllSlots];        // arc weight (0 for absent slots)
  uint32_t sl[kEllSlots];    // count slot (backward only)
  uint32_t nidx[kEllSlots];
  Real nwv[kEllSlots];
  uint32_t nsl[kEllSlots];
  uint4 mcur, mnext;

  // ================================================================ forward
  if (t == 0) {
    ring[0] = Real(1);
    ag[0] = Real(1);
    E[0] = 0;
  }
  if (CTA && threadIdx.x == 0) smax[0] = smax[1] = smax[2] = 0;
#pragma unroll
  for (int s = 0; s < kEllStages; ++s) {
    const int Lp = 1 + s;
    if (Lp < nl) {
      const uint4 mp = __ldg(&meta[Lp]);
      issue(ein + mp.x, mp.w & 0xff, (mp.w >> 8) & 0xff, Lp % kEllStages);
    }
    cp_async_commit();
  }
  // gather level 1 (records, then weights) into the "next" registers
  auto gather_in = [&](int Lg) {  // records of level Lg have landed
    mnext = __ldg(&meta[Lg]);
    const uint32_t Wg = mnext.w & 0xff, Dg = (mnext.w >> 8) & 0xff;
    uint32_t src = sbase + (uint32_t)(Lg % kEllStages) * (kEllSlots * kEllSlotStride);
#pragma unroll
    for (int k = 0; k < kEllSlots; ++k) {
      nidx[k] = 0;
      nwv[k] = 0;
      if (r < Wg && c + k * C < Dg) {
        const uint2 rec = lds_u2(src);
        nidx[k] = rec.x;
        nwv[k] = __ldg(&w[rec.y]);
      }
      src += kEllSlotStride;
    }
  };
  if (nl > 1) {
    cp_async_wait<kEllStages - 1>();
    gather_in(1);
  }
  group_sync();
  int Eprev = 0, par = 0;
  int last_event = 0;  // highest level whose scale differs from the level before it
  for (int L = 1; L < nl; ++L) {
    mcur = mnext;
#pragma unroll
    for (int k = 0; k < kEllSlots; ++k) {
      idx[k] = nidx[k];
      wv[k] = nwv[k];
    }
    // (a) topology of level L+1 is in shared memory by now: start its weight gathers
    if (L + 1 < nl) {
      cp_async_wait<kEllStages - 2>();
      gather_in(L + 1);
    }
    // (b) refill the stage level L used with level L + kEllStages
    if (L + kEllStages < nl) {
      const uint4 mp = __ldg(&meta[L + kEllStages]);
      issue(ein + mp.x, mp.w & 0xff, (mp.w >> 8) & 0xff, L % kEllStages);
    }
    cp_async_commit();
    // (c) level L
    const uint32_t W = mcur.w & 0xff, s0 = mcur.z & 0x7fffffffu;
    const int min_src = L - (int)((mcur.w >> 24) & 0xf);
    Real a = 0;
    if (last_event <= min_src) {
#pragma unroll
      for (int k = 0; k < kEllSlots; ++k) a = fma(ring[idx[k] & M], wv[k], a);
    } else {  // a source level inside the window carries another power-of-two scale
#pragma unroll
      for (int k = 0; k < kEllSlots; ++k) {
        if (wv[k] != Real(0)) {
          int ls = L - 1;
          while ((__ldg(&meta[ls]).z & 0x7fffffffu) > idx[k]) --ls;
          // (E[L-1] may not be visible yet to every lane of a CTA group: use the register copy)
          a += Num<Real>::scale2(ring[idx[k] & M] * wv[k], Eprev - (ls == L - 1 ? Eprev : E[ls]));
        }
      }
    }
    a = col_sum<C>(a, gmask);
    int mx = (r < W) ? Num<Real>::expo(a) : 0;
    int shift = 0;
    if (!CTA) {
      mx = row_reduce<1, GW>(mx, gmask, [](int x, int y) { return max(x, y); });
      if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) {
        shift = Num<Real>::kBias - mx;
        a = Num<Real>::scale2(a, shift);
      }
    }
    // The ring is sized (flattener) so that this level's slots never alias a source slot of the level.
    if (c == 0 && r < W) {
      ring[(s0 + r) & M] = a;
      ag[s0 + r] = a;
    }
    if (CTA) {  // block-wide maximum; renormalise after the fact (rare)
      mx = row_reduce<1, 32>(mx, 0xffffffffu, [](int x, int y) { return max(x, y); });
      if ((threadIdx.x & 31) == 0 && mx) atomicMax(&smax[par], mx);
      __syncthreads();
      mx = smax[par];
      if (threadIdx.x == 0) smax[(par + 2) % 3] = 0;  // last used one level ago, next used two levels ahead
      par = (par + 1) % 3;
      if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) {
        shift = Num<Real>::kBias - mx;
        if (c == 0 && r < W) {
          a = Num<Real>::scale2(a, shift);
          ring[(s0 + r) & M] = a;
          ag[s0 + r] = a;
        }
        __syncthreads();
      }
    } else {
      __syncwarp(gmask);
    }
    if (shift) last_event = L;
    Eprev += shift;
    if (t == 0) E[L] = Eprev;
  }
  cp_async_wait<0>();
  group_sync();
  const Real afin = ag[d.fin];
  const int Efin = E[d.fin_level];
  const double lnP = (afin > 0) ? log((double)afin) - (double)Efin * 0.69314718055994530942 : -CUDART_INF;
  if (t == 0) A.ex_lnp[d.ex_index] = lnP;
  if (!(afin > 0)) return;  // zero-probability example (group-uniform)
  const double cw = d.weight / (double)afin;
  group_sync();

  // ================================================================ backward + counts
  if (CTA && threadIdx.x == 0) smax[0] = smax[1] = smax[2] = 0;
  // iteration `it` handles level nl-1-it and uses stage it % kEllStages
#pragma unroll
  for (int s = 0; s < kEllStages; ++s) {
    const int Lp = nl - 1 - s;
    if (Lp >= 0) {
      const uint4 mp = __ldg(&meta[Lp]);
      issue(eout + mp.y, mp.w & 0xff, (mp.w >> 16) & 0xff, s);
    }
    cp_async_commit();
  }
  auto gather_out = [&](int Lg, int st) {
    mnext = __ldg(&meta[Lg]);
    const uint32_t Wg = mnext.w & 0xff, Og = (mnext.w >> 16) & 0xff;
    uint32_t src = sbase + (uint32_t)st * (kEllSlots * kEllSlotStride);
#pragma unroll
    for (int k = 0; k < kEllSlots; ++k) {
      nidx[k] = 0;
      nwv[k] = 0;
      nsl[k] = 0xFFFFFFFFu;
      if (r < Wg && c + k * C < Og) {
        const uint2 rec = lds_u2(src);
        const WS<Real> e = ws[rec.y];
        nidx[k] = rec.x;
        nwv[k] = e.w;
        nsl[k] = e.slot;
      }
      src += kEllSlotStride;
    }
  };
  cp_async_wait<kEllStages - 1>();
  gather_out(nl - 1, 0);
  group_sync();
  int Fnext = 0;
  par = 0;
  int last_event_b = 0x7fffffff;  // lowest level whose beta scale differs from the level after it
  for (int L = nl - 1, it = 0; L >= 0; --L, ++it) {
    mcur = mnext;
#pragma unroll
    for (int k = 0; k < kEllSlots; ++k) {
      idx[k] = nidx[k];
      wv[k] = nwv[k];
      sl[k] = nsl[k];
    }
    if (L - 1 >= 0) {
      cp_async_wait<kEllStages - 2>();
      gather_out(L - 1, (it + 1) % kEllStages);
    }
    if (L - kEllStages >= 0) {
      const uint4 mp = __ldg(&meta[L - kEllStages]);
      issue(eout + mp.y, mp.w & 0xff, (mp.w >> 16) & 0xff, it % kEllStages);
    }
    cp_async_commit();
    const uint32_t W = mcur.w & 0xff, s0 = mcur.z & 0x7fffffffu;
    const int max_dst = L + (int)((mcur.w >> 28) & 0xf);
    const bool aggregate = (mcur.z >> 31) != 0;
    const bool uniform = last_event_b >= max_dst;
    const bool active = r < W;
    const uint32_t s = s0 + (active ? r : 0);
    const int de = Efin - E[L] - Fnext;
    const double as = active ? (double)ag[s] * (de ? scalbn(cw, de) : cw) : 0.;
    Real b = 0;
#pragma unroll
    for (int k = 0; k < kEllSlots; ++k) {
      Real tt = wv[k] * ring[idx[k] & M];
      if (!uniform && wv[k] != Real(0)) {
        int ld = L + 1;
        while (ld + 1 < nl && (__ldg(&meta[ld + 1]).z & 0x7fffffffu) <= idx[k]) ++ld;
        tt = Num<Real>::scale2(tt, Fnext - (ld == L + 1 ? Fnext : F[ld]));
      }
      b += tt;
      double cval = as * (double)tt;
      if (aggregate) {  // the whole column feeds one slot: one RED per warp and column
        if (k * C < ((mcur.w >> 16) & 0xff)) {  // group-uniform: column block k exists
          cval = row_reduce<C, GW>(cval, gmask, [](double x, double y) { return x + y; });
          if (((threadIdx.x & (GW - 1)) / C) == 0 && cval > 0 && sl[k] != kSlotNone) count_add(A.sink, sl[k], cval);
        }
      } else if (cval > 0 && sl[k] != 0xFFFFFFFFu) {
        count_add(A.sink, sl[k], cval);
      }
    }
    b = col_sum<C>(b, gmask);
    if (active && s == d.fin) b += Num<Real>::scale2(Real(1), Fnext);
    int mx = active ? Num<Real>::expo(b) : 0;
    int shift = 0;
    if (!CTA) {
      mx = row_reduce<1, GW>(mx, gmask, [](int x, int y) { return max(x, y); });
      if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) {
        shift = Num<Real>::kBias - mx;
        b = Num<Real>::scale2(b, shift);
      }
    }
    if (c == 0 && active) ring[s & M] = b;
    if (CTA) {
      mx = row_reduce<1, 32>(mx, 0xffffffffu, [](int x, int y) { return max(x, y); });
      if ((threadIdx.x & 31) == 0 && mx) atomicMax(&smax[par], mx);
      __syncthreads();
      mx = smax[par];
      if (threadIdx.x == 0) smax[(par + 2) % 3] = 0;
      par = (par + 1) % 3;
      if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) {
        shift = Num<Real>::kBias - mx;
        if (c == 0 && active) ring[s & M] = Num<Real>::scale2(b, shift);
        __syncthreads();
      }
    } else {
      __syncwarp(gmask);
    }
    if (shift) last_event_b = L;
    Fnext += shift;
    if (t == 0) F[L] = Fnext;
  }
  cp_async_wait<0>();
}

}  // namespace cmlk
