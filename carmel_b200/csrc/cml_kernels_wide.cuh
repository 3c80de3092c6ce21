// cml_kernels_wide.cuh -- E-step kernel for WIDE lattices (9..32 states per topological level: a cipher's 27-state
// positions, a tagger's tag sets): one lattice per WARP, lane = state of the current level, persistent warps.
// (Same reference semantics as the other sweeps: derivations.h:400-449 compute_fb / collect_counts,
// graph.h:391-402 propagate_paths_in_order.)
//
// Why (profiles/r1_cipher_v2_k_fb_ell_f64.txt): k_fb_ell<32,8,CTA> ran one lattice per 256-thread CTA
// with a block barrier per level, 8-byte LDGSTS per record and a dependent L2 gather of the arc weight per record, then
// one fp64 RED per arc: 0.34 waves, long-scoreboard 17.7 cycles per issue, 21% of the HBM roofline.  Here
//   * the level slabs of a lattice (column-major ELL blocks, cml_device.cu) are one contiguous stream per sweep; lane 0
//     pulls it through shared memory in 2 KB chunks with cp.async.bulk (TMA, SASS UBLKCP) completing on per-slot
//     mbarriers, kWideStages chunks in flight per warp -- no per-record load instructions, no registers held across the
//     memory latency;
//   * arc weights are FACTORED (cml_device.cu "arc classes"): a record carries the class of the arc's rest part, a state
//     carries the class of the part shared by all its incoming arcs; both class tables (weight + count slot) sit in shared
//     memory when they fit (cipher: 1.5k classes), so the per-arc work is LDS record, LDS score, LDS weight, FMA;
//   * the state part's expected count is the state posterior: ONE accumulation per state instead of one per arc (for an
//     LM o channel cascade with a locked LM the sweep issues no per-arc atomics at all);
//   * a warp owns its lattice: levels are separated by __syncwarp only; state scores of recent levels live in a per-warp
//     shared-memory ring; alpha is written once and read once (coalesced), beta never leaves the SM.
// Scaled linear space with per-level power-of-two exponents exactly as k_fb_ell (E[] forward, F[] backward).
#pragma once
#include "cml_common.cuh"
#include "cml_kernels_ell.cuh"
#include "cml_kernels_fb.cuh"

namespace cmlk {

constexpr int kWideStages = 4;          // chunks in flight per warp
constexpr uint32_t kWideChunkRec = 256; // records per chunk (2 KB)
constexpr uint32_t kWideChunkBytes = kWideChunkRec * 8;

struct WideArgs {
  const EllDesc* desc;
  const uint32_t* ex_list;   // indices into desc, handled by this launch
  uint32_t n_list;
  const uint4* lvl_meta;
  const uint2* ell_in;       // {src layered index, arc class id}
  const uint2* ell_out;      // {dst layered index, arc class id}
  const uint32_t* st_vcls;   // [ELL state slot] state class id (0 = none)
  const void* a_w;           // Real[n_a]: arc class weights (entry 0 = 0: padding)
  const uint32_t* a_slot;    // [n_a] count slot codes
  const void* v_w;           // Real[n_v]: state class weights (entry 0 = 1: no state part)
  const uint32_t* v_slot;
  uint32_t n_a, n_v;
  int any_a_slot;            // some arc class has a count slot (else the sweep skips the per-arc count work)
  CountSink sink;
  double* ex_lnp;
  void* alpha_g;
  int* lvl_exp;
  uint32_t ring;             // ring entries per warp (power of two)
  uint32_t warps_per_cta;
  int no_counts;
};

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x; SASS: SYNCS.*, UBLKCP) --------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// One warp's view of a record stream that arrives in chunks.  Chunk c of the current stream has global sequence number
// seq0 + c (or seq0 + (n_chunks-1-c) when the stream is consumed backwards); sequence number q lives in slot q % stages
// and completes phase (q / stages) & 1 of that slot's mbarrier.
struct WidePipe {
  uint32_t bar0, buf0;       // shared-memory addresses: mbarriers (8 B each), chunk buffers
  uint32_t seq;              // sequence number of the next chunk to issue (monotone over the warp's lifetime)
  uint32_t seq0;             // sequence number of the current stream's first consumed chunk
  const uint2* src;          // current stream
  uint32_t n_rec, n_chunks;  // records (even), chunks
  uint32_t issued, waited, freed;  // counts in consumption order
  bool reverse;
};

__device__ __forceinline__ uint32_t wide_chunk_of(const WidePipe& P, uint32_t ord) {  // consumption ordinal -> chunk index
  return P.reverse ? P.n_chunks - 1 - ord : ord;
}
__device__ __forceinline__ void wide_issue(WidePipe& P, int lane) {  // issue consumption ordinal P.issued
  if (lane == 0) {
    const uint32_t c = wide_chunk_of(P, P.issued);
    const uint32_t first = c * kWideChunkRec;
    const uint32_t bytes = min(kWideChunkRec, P.n_rec - first) * 8u;
    const uint32_t slot = P.seq % kWideStages;
    const uint32_t bar = P.bar0 + slot * 8u;
    fence_proxy_async();  // the slot's previous contents were read through the generic proxy
    mbar_expect_tx(bar, bytes);
    bulk_g2s(P.buf0 + slot * kWideChunkBytes, P.src + first, bytes, bar);
  }
  ++P.seq;
  ++P.issued;
}
__device__ __forceinline__ void wide_open(WidePipe& P, const uint2* src, uint32_t n_rec, bool reverse, int lane) {
  P.src = src;
  P.n_rec = n_rec;
  P.n_chunks = (n_rec + kWideChunkRec - 1) / kWideChunkRec;
  P.reverse = reverse;
  P.seq0 = P.seq;
  P.issued = P.waited = P.freed = 0;
  __syncwarp();  // every lane is done with the previous stream's buffers
  while (P.issued < P.n_chunks && P.issued < (uint32_t)kWideStages) wide_issue(P, lane);
}
// make records [lo, hi] (stream indices, lo <= hi) readable; chunks wholly behind the window are refilled.
__device__ __forceinline__ void wide_need(WidePipe& P, uint32_t lo, uint32_t hi, int lane) {
  uint32_t o_first, o_last;  // consumption ordinals of the first / last chunk the window touches
  if (!P.reverse) {
    o_first = lo / kWideChunkRec;
    o_last = hi / kWideChunkRec;
  } else {
    o_first = P.n_chunks - 1 - hi / kWideChunkRec;
    o_last = P.n_chunks - 1 - lo / kWideChunkRec;
  }
  if (P.freed < o_first) {
    __syncwarp();  // all lanes have consumed the chunks before the window
    while (P.freed < o_first) {
      ++P.freed;
      if (P.issued < P.n_chunks) wide_issue(P, lane);
    }
  }
  while (P.waited <= o_last) {
    const uint32_t q = P.seq0 + P.waited;
    mbar_wait(P.bar0 + (q % kWideStages) * 8u, (q / kWideStages) & 1u);
    ++P.waited;
  }
}
__device__ __forceinline__ uint2 wide_rec(const WidePipe& P, uint32_t idx) {  // record idx of the current stream
  const uint32_t c = idx / kWideChunkRec;
  const uint32_t q = P.seq0 + (P.reverse ? P.n_chunks - 1 - c : c);
  return lds_u2(P.buf0 + (q % kWideStages) * kWideChunkBytes + (idx % kWideChunkRec) * 8u);
}

template <typename Real, bool TBL>  // TBL: both class tables staged in shared memory
static __global__ void __launch_bounds__(512) k_fb_wide(WideArgs A) {
  extern __shared__ __align__(128) unsigned char smem_wide[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  // ---- shared memory: [class tables] | per warp: mbarriers (64 B) | chunk buffers | ring
  size_t tbl_bytes = 0;
  const Real* aw = (const Real*)A.a_w;
  const uint32_t* asl = A.a_slot;
  const Real* vw = (const Real*)A.v_w;
  const uint32_t* vsl = A.v_slot;
  if (TBL) {
    Real* s_aw = (Real*)smem_wide;
    Real* s_vw = s_aw + A.n_a;
    uint32_t* s_asl = (uint32_t*)(s_vw + A.n_v);
    uint32_t* s_vsl = s_asl + A.n_a;
    for (uint32_t i = threadIdx.x; i < A.n_a; i += blockDim.x) {
      s_aw[i] = aw[i];
      s_asl[i] = asl[i];
    }
    for (uint32_t i = threadIdx.x; i < A.n_v; i += blockDim.x) {
      s_vw[i] = vw[i];
      s_vsl[i] = vsl[i];
    }
    aw = s_aw;
    vw = s_vw;
    asl = s_asl;
    vsl = s_vsl;
    tbl_bytes = ((size_t)(A.n_a + A.n_v) * (sizeof(Real) + 4) + 127) & ~(size_t)127;
  }
  const size_t per_warp = 64 + (size_t)kWideStages * kWideChunkBytes + (size_t)A.ring * sizeof(Real);
  unsigned char* wbase = smem_wide + tbl_bytes + (size_t)wib * per_warp;
  WidePipe P;
  P.bar0 = smem_u32(wbase);
  P.buf0 = P.bar0 + 64;
  P.seq = 0;
  Real* ring = (Real*)(wbase + 64 + (size_t)kWideStages * kWideChunkBytes);
  const uint32_t M = A.ring - 1;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kWideStages; ++k) mbar_init(P.bar0 + k * 8u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();  // tables staged, barriers initialised

  const uint32_t gw = blockIdx.x * A.warps_per_cta + wib, nw = gridDim.x * A.warps_per_cta;
  for (uint32_t li = gw; li < A.n_list; li += nw) {
    const EllDesc d = A.desc[A.ex_list[li]];
    const uint4* __restrict__ meta = A.lvl_meta + d.meta_base;
    Real* __restrict__ ag = (Real*)A.alpha_g + d.state_base;
    const uint32_t* __restrict__ svc = A.st_vcls + d.state_base;
    int* __restrict__ E = A.lvl_exp + 2 * d.level_base;
    int* __restrict__ F = E + d.n_levels;
    const int nl = (int)d.n_levels;

    // ================================================================ forward
    wide_open(P, A.ell_in + d.in_base, d.in_len, false, lane);
    if (lane == 0) {
      ring[0] = Real(1);
      ag[0] = Real(1);
      E[0] = 0;
    }
    __syncwarp();
    int Eprev = 0, last_event = 0;
    uint4 mnext = nl > 1 ? __ldg(&meta[1]) : make_uint4(0, 0, 0, 0);
    uint32_t vnext = 0;
    if (nl > 1 && (uint32_t)lane < (mnext.w & 0xff)) vnext = __ldg(&svc[(mnext.z & 0x7fffffffu) + lane]);
    for (int L = 1; L < nl; ++L) {
      const uint4 m = mnext;
      const uint32_t vc = vnext;
      if (L + 1 < nl) {  // next level's meta and state classes (global loads off the critical path)
        mnext = __ldg(&meta[L + 1]);
        vnext = 0;
        if ((uint32_t)lane < (mnext.w & 0xff)) vnext = __ldg(&svc[(mnext.z & 0x7fffffffu) + lane]);
      }
      const uint32_t W = m.w & 0xff, D = (m.w >> 8) & 0xff, s0 = m.z & 0x7fffffffu;
      const int min_src = L - (int)((m.w >> 24) & 0xf);
      const bool act = (uint32_t)lane < W;
      const uint32_t lw = act ? lane : 0;  // idle lanes read a valid record and drop the result
      Real a0 = 0, a1 = 0;
      if (D) {
        if (last_event <= min_src) {
          uint32_t k = 0;
          for (; k + 4 <= D; k += 4) {
            wide_need(P, m.x + k * W, m.x + (k + 4) * W - 1, lane);
            const uint2 r0 = wide_rec(P, m.x + k * W + lw), r1 = wide_rec(P, m.x + (k + 1) * W + lw);
            const uint2 r2 = wide_rec(P, m.x + (k + 2) * W + lw), r3 = wide_rec(P, m.x + (k + 3) * W + lw);
            const Real x0 = ring[r0.x & M], x1 = ring[r1.x & M], x2 = ring[r2.x & M], x3 = ring[r3.x & M];
            const Real w0 = aw[r0.y], w1 = aw[r1.y], w2 = aw[r2.y], w3 = aw[r3.y];
            a0 = fma(x0, w0, a0);
            a1 = fma(x1, w1, a1);
            a0 = fma(x2, w2, a0);
            a1 = fma(x3, w3, a1);
          }
          for (; k < D; ++k) {
            wide_need(P, m.x + k * W, m.x + (k + 1) * W - 1, lane);
            const uint2 r0 = wide_rec(P, m.x + k * W + lw);
            a0 = fma(ring[r0.x & M], aw[r0.y], a0);
          }
        } else {  // a source level inside the window carries another power-of-two scale (rare)
          for (uint32_t k = 0; k < D; ++k) {
            wide_need(P, m.x + k * W, m.x + (k + 1) * W - 1, lane);
            const uint2 r0 = wide_rec(P, m.x + k * W + lw);
            const Real wv = aw[r0.y];
            if (wv != Real(0)) {
              int ls = L - 1;
              while ((__ldg(&meta[ls]).z & 0x7fffffffu) > r0.x) --ls;
              a0 += Num<Real>::scale2(ring[r0.x & M] * wv, Eprev - (ls == L - 1 ? Eprev : E[ls]));
            }
          }
        }
      }
      Real a = act ? (a0 + a1) * vw[vc] : Real(0);
      int mx = __reduce_max_sync(0xffffffffu, act ? Num<Real>::expo(a) : 0);
      int shift = 0;
      if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) {
        shift = Num<Real>::kBias - mx;
        a = Num<Real>::scale2(a, shift);
      }
      // (the flattener sizes the ring so that this level's slots never alias a source slot of the level)
      if (act) {
        ring[(s0 + lane) & M] = a;
        ag[s0 + lane] = a;
      }
      if (shift) last_event = L;
      Eprev += shift;
      if (lane == 0) E[L] = Eprev;
      __syncwarp();
    }
    const Real afin = ag[d.fin];  // (written by this warp; ordered by the __syncwarp that closes every level)
    const int Efin = E[d.fin_level];
    const double lnP = (afin > 0) ? log((double)afin) - (double)Efin * 0.69314718055994530942 : -CUDART_INF;
    if (lane == 0) A.ex_lnp[d.ex_index] = lnP;
    if (!(afin > 0)) continue;  // zero-probability example (warp-uniform)
    const double cw = d.weight / (double)afin;

    // ================================================================ backward + counts
    wide_open(P, A.ell_out + d.out_base, d.out_len, true, lane);
    int Fnext = 0, last_event_b = 0x7fffffff;
    mnext = __ldg(&meta[nl - 1]);
    vnext = 0;
    Real anext = 0;
    if ((uint32_t)lane < (mnext.w & 0xff)) {
      vnext = __ldg(&svc[(mnext.z & 0x7fffffffu) + lane]);
      anext = ag[(mnext.z & 0x7fffffffu) + lane];
    }
    for (int L = nl - 1; L >= 0; --L) {
      const uint4 m = mnext;
      const uint32_t vc = vnext;
      const Real al = anext;
      if (L >= 1) {
        mnext = __ldg(&meta[L - 1]);
        vnext = 0;
        anext = 0;
        if ((uint32_t)lane < (mnext.w & 0xff)) {
          vnext = __ldg(&svc[(mnext.z & 0x7fffffffu) + lane]);
          anext = ag[(mnext.z & 0x7fffffffu) + lane];
        }
      }
      const uint32_t W = m.w & 0xff, O = (m.w >> 16) & 0xff, s0 = m.z & 0x7fffffffu;
      const int max_dst = L + (int)((m.w >> 28) & 0xf);
      const bool uniform = last_event_b >= max_dst;
      const bool act = (uint32_t)lane < W;
      const uint32_t lw = act ? lane : 0;
      const uint32_t s = s0 + lw;
      const int de = Efin - E[L] - Fnext;
      const double as = act ? (double)al * (de ? scalbn(cw, de) : cw) : 0.;
      Real b0 = 0, b1 = 0;
      if (O) {
        if (uniform && !A.any_a_slot) {  // no per-arc count work at all
          // (rows are taken last to first: the outgoing stream is consumed strictly backwards)
          uint32_t k = O;
          for (; k >= 4; k -= 4) {
            wide_need(P, m.y + (k - 4) * W, m.y + k * W - 1, lane);
            const uint2 r0 = wide_rec(P, m.y + (k - 1) * W + lw), r1 = wide_rec(P, m.y + (k - 2) * W + lw);
            const uint2 r2 = wide_rec(P, m.y + (k - 3) * W + lw), r3 = wide_rec(P, m.y + (k - 4) * W + lw);
            const Real x0 = ring[r0.x & M], x1 = ring[r1.x & M], x2 = ring[r2.x & M], x3 = ring[r3.x & M];
            const Real w0 = aw[r0.y], w1 = aw[r1.y], w2 = aw[r2.y], w3 = aw[r3.y];
            b0 = fma(x0, w0, b0);
            b1 = fma(x1, w1, b1);
            b0 = fma(x2, w2, b0);
            b1 = fma(x3, w3, b1);
          }
          for (; k > 0; --k) {
            wide_need(P, m.y + (k - 1) * W, m.y + k * W - 1, lane);
            const uint2 r0 = wide_rec(P, m.y + (k - 1) * W + lw);
            b0 = fma(ring[r0.x & M], aw[r0.y], b0);
          }
        } else {
          for (uint32_t k = O; k-- > 0;) {
            wide_need(P, m.y + k * W, m.y + (k + 1) * W - 1, lane);
            const uint2 r0 = wide_rec(P, m.y + k * W + lw);
            const Real wv = aw[r0.y];
            Real tt = wv * ring[r0.x & M];
            if (!uniform && wv != Real(0)) {
              int ld = L + 1;
              while (ld + 1 < nl && (__ldg(&meta[ld + 1]).z & 0x7fffffffu) <= r0.x) ++ld;
              tt = Num<Real>::scale2(tt, Fnext - (ld == L + 1 ? Fnext : F[ld]));
            }
            b0 += tt;
            if (A.any_a_slot && act && !A.no_counts) {
              const uint32_t sc = asl[r0.y];
              const double cval = as * (double)tt;
              if (sc != kSlotNone && cval > 0) count_add(A.sink, sc, cval);
            }
          }
        }
      }
      Real b = b0 + b1;
      if (act && s == d.fin) b += Num<Real>::scale2(Real(1), Fnext);
      if (!act) b = 0;
      int mx = __reduce_max_sync(0xffffffffu, act ? Num<Real>::expo(b) : 0);
      int shift = 0;
      if (mx != 0 && (mx < Num<Real>::kLo || mx > Num<Real>::kHi)) {
        shift = Num<Real>::kBias - mx;
        b = Num<Real>::scale2(b, shift);
      }
      if (act) {
        // state posterior = expected count of the state's class: alpha * beta / P (scales: E[L], Fnext + shift)
        const uint32_t sc = vsl[vc];
        if (sc != kSlotNone && !A.no_counts) {
          const double g = as * (double)b * (shift ? exp2((double)-shift) : 1.);
          if (g > 0) count_add(A.sink, sc, g);
        }
        ring[s & M] = b * vw[vc];  // what predecessors multiply by: V[state part] * beta
      }
      if (shift) last_event_b = L;
      Fnext += shift;
      if (lane == 0) F[L] = Fnext;
      __syncwarp();
    }
  }
}

// class weight tables: w[i] = product of the class's parameters (linear domain); entry 0 is the padding entry
template <typename Real>
static __global__ void k_class_weights(uint32_t n, const uint32_t* __restrict__ off, const uint32_t* __restrict__ param,
                                       const double* __restrict__ ln_w, Real pad_value, Real* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) {
    out[0] = pad_value;
    return;
  }
  double s = 0;
  for (uint32_t k = off[i], e = off[i + 1]; k < e; ++k) s += ln_w[param[k]];
  out[i] = (Real)exp(s);
}

}  // namespace cmlk
