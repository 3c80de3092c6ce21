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
constexpr int kWidePrefetch = 12;       // chunks ahead of the sweep that are requested into L2

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
  int l2_prefetch;           // chunks ahead of the sweep requested into L2 (0 = off)
  int generic_addr;          // experiment: generic-pointer addressing in the inner loops
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
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// One warp's view of a record stream that arrives in 2 KB chunks.  Chunk c of a stream always lives in slot c % stages
// (so record idx sits at byte ((idx mod stages*256) * 8) of the buffer, whichever way the stream is consumed); every
// issued chunk is waited for exactly once, so a slot's mbarrier phase parity is one bit that flips at each wait.
// All fields are warp-uniform.
struct WidePipe {
  uint32_t bar0, buf0;       // shared-memory addresses: mbarriers (8 B each), chunk buffers
  uint32_t wpar;             // bit s: parity to wait for on slot s
  const uint2* src;          // current stream
  uint32_t n_rec, n_chunks;  // records (even), chunks
  uint32_t issued, waited, freed;  // chunk counts in consumption order
  // forward: records < wait_lim are readable; chunk `freed` may be refilled once the window starts at >= free_lim.
  // reverse: records >= wait_lim are readable; the oldest chunk may be refilled once the window ends below free_lim.
  uint32_t wait_lim, free_lim;
  uint32_t pf;               // L2 prefetch distance in chunks (0 = off)
};

template <bool REV, int ST = kWideStages>
__device__ __forceinline__ void wide_issue(WidePipe& P, bool leader) {  // issue the next chunk in consumption order
  if (leader) {
    const uint32_t c = REV ? P.n_chunks - 1 - P.issued : P.issued;
    const uint32_t first = c * kWideChunkRec;
    const uint32_t bytes = min(kWideChunkRec, P.n_rec - first) * 8u;
    const uint32_t slot = c % ST;
    const uint32_t bar = P.bar0 + slot * 8u;
    fence_proxy_async();  // the slot's previous contents were read through the generic proxy
    mbar_expect_tx(bar, bytes);
    bulk_g2s(P.buf0 + slot * kWideChunkBytes, P.src + first, bytes, bar);
    // pull the chunk kWidePrefetch positions further along the sweep into L2: the shared-memory ring only covers the
    // L2 latency, the DRAM latency is covered here (no shared memory needed for the bytes in flight)
    const uint32_t ahead = P.issued + P.pf;
    if (P.pf && ahead < P.n_chunks) {
      const uint32_t ca = REV ? P.n_chunks - 1 - ahead : ahead;
      bulk_prefetch_l2(P.src + ca * kWideChunkRec, min(kWideChunkRec, P.n_rec - ca * kWideChunkRec) * 8u);
    }
  }
  ++P.issued;
}
template <bool REV, int ST = kWideStages>
__device__ __forceinline__ void wide_open(WidePipe& P, const uint2* src, uint32_t n_rec, bool leader) {
  P.src = src;
  P.n_rec = n_rec;
  P.n_chunks = (n_rec + kWideChunkRec - 1) / kWideChunkRec;
  P.issued = P.waited = P.freed = 0;
  P.wait_lim = REV ? P.n_chunks * kWideChunkRec : 0u;
  P.free_lim = REV ? (P.n_chunks ? (P.n_chunks - 1) * kWideChunkRec : 0u) : kWideChunkRec;
  __syncwarp();  // every lane is done with the previous stream's buffers (all of its chunks have been waited for)
  if (leader)  // L2 prefetch of the chunks between the shared-memory ring and the steady-state prefetch distance
    for (uint32_t a = ST; a < P.n_chunks && a < P.pf; ++a) {
      const uint32_t ca = REV ? P.n_chunks - 1 - a : a;
      bulk_prefetch_l2(src + ca * kWideChunkRec, min(kWideChunkRec, n_rec - ca * kWideChunkRec) * 8u);
    }
  while (P.issued < P.n_chunks && P.issued < (uint32_t)ST) wide_issue<REV, ST>(P, leader);
}
// make records [lo, hi] (lo <= hi, at most two chunks) readable; chunks wholly behind the window are refilled.
// The common case (window inside the chunks already waited for) costs two compares.
template <bool REV, int ST = kWideStages>
__device__ __forceinline__ void wide_need(WidePipe& P, uint32_t lo, uint32_t hi, bool leader) {
  if (!REV) {
    if (lo >= P.free_lim) {
      __syncwarp();  // all lanes have consumed the chunks before the window
      do {
        ++P.freed;
        P.free_lim += kWideChunkRec;
        if (P.issued < P.n_chunks) wide_issue<REV, ST>(P, leader);
      } while (lo >= P.free_lim);
    }
    while (hi >= P.wait_lim) {
      const uint32_t slot = P.waited % ST;
      mbar_wait(P.bar0 + slot * 8u, (P.wpar >> slot) & 1u);
      P.wpar ^= 1u << slot;
      ++P.waited;
      P.wait_lim += kWideChunkRec;
    }
  } else {
    if (hi < P.free_lim) {
      __syncwarp();
      do {
        ++P.freed;
        P.free_lim -= kWideChunkRec;  // (reaches 0 with the last chunk: `hi < 0` never holds again)
        if (P.issued < P.n_chunks) wide_issue<REV, ST>(P, leader);
      } while (hi < P.free_lim);
    }
    while (lo < P.wait_lim) {
      const uint32_t slot = (P.n_chunks - 1 - P.waited) % ST;
      mbar_wait(P.bar0 + slot * 8u, (P.wpar >> slot) & 1u);
      P.wpar ^= 1u << slot;
      ++P.waited;
      P.wait_lim -= kWideChunkRec;
    }
  }
}
// shared-memory accessors on 32-bit shared addresses (no generic-address conversion in the inner loops)
template <typename Real>
__device__ __forceinline__ Real lds_real(uint32_t a);
template <>
__device__ __forceinline__ double lds_real<double>(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(a));
  return v;
}
template <>
__device__ __forceinline__ float lds_real<float>(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(a));
  return v;
}
template <int ST = kWideStages>
__device__ __forceinline__ uint2 wide_rec(const WidePipe& P, uint32_t idx) {  // record idx of the current stream
  return lds_u2(P.buf0 + ((idx & (ST * kWideChunkRec - 1)) << 3));
}

template <typename Real, bool TBL>  // TBL: both class tables staged in shared memory
static __global__ void __launch_bounds__(512) k_fb_wide(WideArgs A) {
  extern __shared__ __align__(128) unsigned char smem_wide[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const bool leader = lane == 0;
  // ---- shared memory (offsets from the first 8 KB boundary, so that a warp's chunk buffer and its ring are aligned to
  //      their own sizes and "base | (offset & mask)" is one LOP3):
  //      chunk buffers [warps][8 KB] | mbarriers [warps][64 B] | rings [warps][ring] | class tables
  constexpr uint32_t kBuf = kWideStages * kWideChunkBytes;
  constexpr int LR = sizeof(Real) == 8 ? 3 : 2;
  const uint32_t ring_bytes = A.ring << LR;
  const uint32_t s0base = (smem_u32(smem_wide) + (kBuf - 1)) & ~(kBuf - 1);
  unsigned char* g0base = smem_wide + (s0base - smem_u32(smem_wide));
  const uint32_t off_bar = A.warps_per_cta * kBuf;
  const uint32_t off_ring = (off_bar + A.warps_per_cta * 64u + ring_bytes - 1) & ~(ring_bytes - 1);
  const uint32_t off_tbl = off_ring + A.warps_per_cta * ring_bytes;
  const Real* aw = (const Real*)A.a_w;
  const uint32_t* asl = A.a_slot;
  const Real* vw = (const Real*)A.v_w;
  const uint32_t* vsl = A.v_slot;
  uint32_t aw_s = 0;  // shared address of the arc-class weights (TBL)
  if (TBL) {
    Real* s_aw = (Real*)(g0base + off_tbl);
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
    aw_s = s0base + off_tbl;
  }
  WidePipe P;
  P.buf0 = s0base + (uint32_t)wib * kBuf;
  P.bar0 = s0base + off_bar + (uint32_t)wib * 64u;
  P.wpar = 0;
  P.pf = (uint32_t)A.l2_prefetch;
  const uint32_t ring_s = s0base + off_ring + (uint32_t)wib * ring_bytes;
  Real* ring = (Real*)(g0base + off_ring + (size_t)wib * ring_bytes);
  const uint32_t M = A.ring - 1;
  const uint32_t M8 = M << LR;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kWideStages; ++k) mbar_init(P.bar0 + k * 8u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();  // tables staged, barriers initialised
  // keep the address bases in registers (ptxas would otherwise rematerialise them from the kernel parameters inside the
  // inner loops: ~20 of 60 instructions per 4-row block)
  uint32_t buf_s = P.buf0, ring_m8 = M8, ring_b = ring_s, aw_b = aw_s;
  asm volatile("" : "+r"(buf_s), "+r"(ring_m8), "+r"(ring_b), "+r"(aw_b));
  // record r of the current stream (byte offset o = 8*r): buffer | (o & (8 KB - 1)); ring entry of state x: ring | ((x << LR) & M8)
  const bool ga = A.generic_addr != 0;
  const uint2* cbuf = (const uint2*)(g0base + (size_t)wib * kBuf);
  auto rec_at = [&](uint32_t byte_off) -> uint2 {
    return ga ? cbuf[(byte_off >> 3) & (kWideStages * kWideChunkRec - 1)] : lds_u2(buf_s | (byte_off & (kBuf - 1)));
  };
  auto ring_at = [&](uint32_t x) -> Real { return ga ? ring[x & M] : lds_real<Real>(ring_b | ((x << LR) & ring_m8)); };
  auto aw_at = [&](uint32_t c) -> Real { return (TBL && !ga) ? lds_real<Real>(aw_b + (c << LR)) : aw[c]; };

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
    wide_open<false>(P, A.ell_in + d.in_base, d.in_len, leader);
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
          uint32_t k = 0, base = m.x;  // base: first record of row k (warp-uniform)
          const uint32_t W4 = 4 * W, Wb = W << 3;
          uint32_t o = (base + lw) << 3;  // byte offset of this lane's record in row k
          for (; k + 4 <= D; k += 4, base += W4, o += 4 * Wb) {
            wide_need<false>(P, base, base + W4 - 1, leader);
            const uint2 r0 = rec_at(o), r1 = rec_at(o + Wb), r2 = rec_at(o + 2 * Wb), r3 = rec_at(o + 3 * Wb);
            const Real x0 = ring_at(r0.x), x1 = ring_at(r1.x), x2 = ring_at(r2.x), x3 = ring_at(r3.x);
            const Real w0 = aw_at(r0.y), w1 = aw_at(r1.y), w2 = aw_at(r2.y), w3 = aw_at(r3.y);
            a0 = fma(x0, w0, a0);
            a1 = fma(x1, w1, a1);
            a0 = fma(x2, w2, a0);
            a1 = fma(x3, w3, a1);
          }
          if (k < D) {  // 1..3 rows left: absent rows read record {state 0, padding class} (weight 0)
            const uint32_t rem = D - k;
            wide_need<false>(P, base, base + rem * W - 1, leader);
            const uint2 pad = make_uint2(0u, 0u);
            const uint2 r0 = rec_at(o), r1 = rem > 1 ? rec_at(o + Wb) : pad, r2 = rem > 2 ? rec_at(o + 2 * Wb) : pad;
            const Real x0 = ring_at(r0.x), x1 = ring_at(r1.x), x2 = ring_at(r2.x);
            const Real w0 = aw_at(r0.y), w1 = aw_at(r1.y), w2 = aw_at(r2.y);
            a0 = fma(x0, w0, a0);
            a1 = fma(x1, w1, a1);
            a0 = fma(x2, w2, a0);
          }
        } else {  // a source level inside the window carries another power-of-two scale (rare)
          for (uint32_t k = 0; k < D; ++k) {
            wide_need<false>(P, m.x + k * W, m.x + (k + 1) * W - 1, leader);
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
    wide_open<true>(P, A.ell_out + d.out_base, d.out_len, leader);
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
          uint32_t k = O, base = m.y + O * W;  // base: one past the last record of row k-1 (warp-uniform)
          const uint32_t W4 = 4 * W, Wb = W << 3;
          uint32_t o = (base + lw) << 3;
          for (; k >= 4; k -= 4) {
            base -= W4;
            o -= 4 * Wb;
            wide_need<true>(P, base, base + W4 - 1, leader);
            const uint2 r0 = rec_at(o + 3 * Wb), r1 = rec_at(o + 2 * Wb), r2 = rec_at(o + Wb), r3 = rec_at(o);
            const Real x0 = ring_at(r0.x), x1 = ring_at(r1.x), x2 = ring_at(r2.x), x3 = ring_at(r3.x);
            const Real w0 = aw_at(r0.y), w1 = aw_at(r1.y), w2 = aw_at(r2.y), w3 = aw_at(r3.y);
            b0 = fma(x0, w0, b0);
            b1 = fma(x1, w1, b1);
            b0 = fma(x2, w2, b0);
            b1 = fma(x3, w3, b1);
          }
          if (k > 0) {  // rows k-1 .. 0
            base -= k * W;
            o -= k * Wb;
            wide_need<true>(P, base, base + k * W - 1, leader);
            const uint2 pad = make_uint2(0u, 0u);
            const uint2 r0 = rec_at(o), r1 = k > 1 ? rec_at(o + Wb) : pad, r2 = k > 2 ? rec_at(o + 2 * Wb) : pad;
            const Real x0 = ring_at(r0.x), x1 = ring_at(r1.x), x2 = ring_at(r2.x);
            const Real w0 = aw_at(r0.y), w1 = aw_at(r1.y), w2 = aw_at(r2.y);
            b0 = fma(x0, w0, b0);
            b1 = fma(x1, w1, b1);
            b0 = fma(x2, w2, b0);
          }
        } else {
          for (uint32_t k = O; k-- > 0;) {
            wide_need<true>(P, m.y + k * W, m.y + (k + 1) * W - 1, leader);
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
