// cml_kernels_lane.cuh -- E-step kernel for corpora of MANY NARROW lattices (HMM / tagging style: a few states
// per position, every arc from one topological level to the next): one lattice per LANE.
// (Same reference semantics as the other sweeps: derivations.h:400-449, graph.h:391-402.)
//
// Why (profiles/r1c_hmm_*): with a group of 16 lanes per lattice (k_fb_ell 4x4) a 16-arc level costs every lane
// ~60 instructions of shuffles, exponent bookkeeping and synchronisation for ONE fused multiply-add: 4.5 warp
// instructions per arc, issue slots 25% busy, long-scoreboard stalls on the dependent record -> weight -> score
// chain, 19% of the HBM roofline.  Here a lane walks its own lattice arc by arc:
//   * 32 lattices of similar size form a TILE; their arc streams are stored transposed, fw[row][lane], so one
//     warp load fetches the next record of all 32 lattices (256 contiguous bytes);
//   * streams are aligned by state ordinal: the k-th state of every lattice in the tile gets the same number of
//     rows (the tile maximum of its degree; short lists are padded with a zero-weight arc), so "end of state"
//     is warp-uniform, the state counter is uniform and the alpha rows alpha[state][lane] are written and read
//     fully coalesced;
//   * state scores of the last two levels live in a per-lane shared-memory ring (column = lane: conflict free,
//     private to the lane, so there is no synchronisation of any kind in the sweep);
//   * records are prefetched two chunks ahead and weights one chunk ahead in registers; the only per-arc work is
//     1 coalesced LDG.64 + 1 gather + 1 LDS + 1 FMA (forward), plus the count RED (backward);
//   * power-of-two rescaling is per lane and per level: a level whose maximum leaves the exponent window changes
//     the scale of the NEXT level (scores already stored are never rewritten).
#pragma once
#include "cml_common.cuh"
#include "cml_kernels_ell.cuh"
#include "cml_kernels_fb.cuh"

namespace cmlk {

struct __align__(16) LaneTile {
  uint64_t fw_base;  // first forward record (uint2 units); row r of lane l at fw_base + r*32 + l
  uint64_t bw_base;
  uint64_t st_base;  // first alpha slot; state ordinal s of lane l at st_base + s*32 + l
  uint64_t lv_base;  // first level-exponent slot; level L of lane l at lv_base + L*32 + l
  uint32_t rows_f, rows_b;  // multiples of kLaneU
  uint32_t n_states, n_levels;  // tile maxima
};

struct LaneArgs {
  const LaneTile* tile;
  uint32_t n_tiles;
  const uint2* fw;        // {source layered index | flags, internal arc id}, states 1..n-1 ascending
  const uint2* bw;        // {destination layered index | flags, internal arc id}, states n-1..0 descending
  const uint32_t* ex;     // [tile*32+lane] example index in the batch, 0xFFFFFFFF = empty lane
  const uint32_t* fin;    // [tile*32+lane] layered index of the goal state (0xFFFFFFFF for empty lanes)
  const uint32_t* nlev;   // [tile*32+lane] levels of the lane's lattice
  const double* weight;   // [tile*32+lane]
  const uint32_t* vcls;   // [state ordinal][lane] state class id (0 = none), indexed like alpha
  const void* a_w;        // Real[n_a]: arc class weights (entry 0 = 0: padding)
  const uint32_t* a_slot; // [n_a] count slot codes
  const void* v_w;        // Real[n_v]: state class weights (entry 0 = 1: no state part)
  const uint32_t* v_slot;
  uint32_t n_a, n_v;
  CountSink sink;
  double* ex_lnp;
  void* alpha;            // Real
  int* lvle;              // scale exponent of every level
  int no_counts;          // profiling: skip the count REDs
};

constexpr uint32_t kLaneLast = 0x80000000u, kLaneLevelEnd = 0x40000000u;
constexpr int kLaneRing = 16;  // ring entries per lane: two levels of width <= 8
constexpr int kLaneU = 4;      // rows per software-pipeline chunk
constexpr int kLaneWarps = 8;

// TA / TV: the arc-class / state-class tables (weight + slot code) are staged in shared memory.  Arc weights are
// FACTORED (cml_device.cu "arc classes"): w(arc) = U[record's class] * V[class of the destination state]; the state
// part's expected count is the state posterior, accumulated once per state.
// MINB: resident blocks per SM the register allocation is capped for (2 = uncapped, ~86 registers; 3 caps at 80).
template <typename Real, bool TA, bool TV, int MINB>
static __global__ void __launch_bounds__(kLaneWarps * 32, MINB) k_fb_lane(LaneArgs A) {
  extern __shared__ __align__(16) unsigned char smem_lane[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t tile = blockIdx.x * kLaneWarps + wib;
  const Real* __restrict__ w = reinterpret_cast<const Real*>(A.a_w);
  const uint32_t* __restrict__ wsl = A.a_slot;
  const Real* __restrict__ vwt = reinterpret_cast<const Real*>(A.v_w);
  const uint32_t* __restrict__ vsl = A.v_slot;
  unsigned char* sp = smem_lane + (size_t)kLaneWarps * kLaneRing * 32 * sizeof(Real);
  if (TA) {
    Real* s_w = reinterpret_cast<Real*>(sp);
    uint32_t* s_sl = reinterpret_cast<uint32_t*>(s_w + A.n_a);
    for (uint32_t i = threadIdx.x; i < A.n_a; i += blockDim.x) {
      s_w[i] = w[i];
      s_sl[i] = wsl[i];
    }
    w = s_w;
    wsl = s_sl;
    sp += (((size_t)A.n_a * (sizeof(Real) + 4)) + 15) & ~(size_t)15;
  }
  if (TV) {
    Real* s_w = reinterpret_cast<Real*>(sp);
    uint32_t* s_sl = reinterpret_cast<uint32_t*>(s_w + A.n_v);
    for (uint32_t i = threadIdx.x; i < A.n_v; i += blockDim.x) {
      s_w[i] = vwt[i];
      s_sl[i] = vsl[i];
    }
    vwt = s_w;
    vsl = s_sl;
  }
  if (TA || TV) __syncthreads();
  if (tile >= A.n_tiles) return;
  Real* ring = reinterpret_cast<Real*>(smem_lane) + (size_t)wib * kLaneRing * 32 + lane;
#pragma unroll
  for (int k = 0; k < kLaneRing; ++k) ring[k * 32] = Real(0);
  const LaneTile T = A.tile[tile];
  const uint32_t li = tile * 32 + lane;
  const uint32_t ex = A.ex[li], fin = A.fin[li];
  Real* __restrict__ ag = reinterpret_cast<Real*>(A.alpha) + T.st_base + lane;
  const uint32_t* __restrict__ vcl = A.vcls + T.st_base + lane;
  int* __restrict__ lve = A.lvle + T.lv_base + lane;

  // ================================================================ forward
  Real afin = 0;
  int Efin = 0;
  {
    const uint2* __restrict__ fw = A.fw + T.fw_base + lane;
    ring[0] = Real(1);  // the start state: layered index 0, scale 0
    ag[0] = Real(1);
    lve[0] = 0;
    if (fin == 0) afin = Real(1);  // a one-state lattice (empty strings)
    Real acc = 0, pend = Real(1);
    int Ew = 0, mxe = 0;
    uint32_t s = 1, lvl = 1;
    uint2 r0[kLaneU], r1[kLaneU];
    Real w0[kLaneU];
    const uint32_t R = T.rows_f;
#pragma unroll
    for (int k = 0; k < kLaneU; ++k) r0[k] = fw[(size_t)k * 32];
#pragma unroll
    for (int k = 0; k < kLaneU; ++k) w0[k] = w[r0[k].y];
#pragma unroll
    for (int k = 0; k < kLaneU; ++k) r1[k] = fw[(size_t)(kLaneU + k) * 32];
    // state parts: class of state s+1 two closes ahead, its weight one close ahead (the array has a padded tail)
    uint32_t vc_next = vcl[(size_t)2 * 32];
    Real vw_cur = vwt[vcl[(size_t)1 * 32]];
    for (uint32_t r = 0; r < R; r += kLaneU) {
      Real w1[kLaneU];
      uint2 r2[kLaneU];
#pragma unroll
      for (int k = 0; k < kLaneU; ++k) w1[k] = w[r1[k].y];
#pragma unroll
      for (int k = 0; k < kLaneU; ++k) r2[k] = fw[(size_t)(r + 2 * kLaneU + k) * 32];  // (the array has a padded tail)
#pragma unroll
      for (int k = 0; k < kLaneU; ++k) {
        const uint32_t x = r0[k].x;
        acc = fma(ring[(x & (kLaneRing - 1)) * 32], w0[k], acc);
        if (x & kLaneLast) {  // warp-uniform: the row closes state ordinal s
          const Real a = acc * pend * vw_cur;
          vw_cur = vwt[vc_next];
          vc_next = vcl[(size_t)min(s + 2, T.n_states + 1) * 32];
          acc = 0;
          ring[(s & (kLaneRing - 1)) * 32] = a;
          ag[(size_t)s * 32] = a;
          mxe = max(mxe, Num<Real>::expo(a));
          if (s == fin) {
            afin = a;
            Efin = Ew;
          }
          ++s;
          if (x & kLaneLevelEnd) {  // per lane: the lane's level `lvl` is complete
            lve[(size_t)lvl * 32] = Ew;
            int shift = 0;
            if (mxe != 0 && (mxe < Num<Real>::kLo || mxe > Num<Real>::kHi)) shift = Num<Real>::kBias - mxe;
            pend = Num<Real>::scale2(Real(1), shift);
            Ew += shift;
            mxe = 0;
            ++lvl;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kLaneU; ++k) {
        r0[k] = r1[k];
        w0[k] = w1[k];
        r1[k] = r2[k];
      }
    }
  }
  const bool live = ex != 0xFFFFFFFFu;
  if (live) A.ex_lnp[ex] = (afin > 0) ? log((double)afin) - (double)Efin * 0.69314718055994530942 : -CUDART_INF;
  const double cw = (live && afin > 0) ? A.weight[li] / (double)afin : 0.;

  // ================================================================ backward + counts
  {
    const uint2* __restrict__ bw = A.bw + T.bw_base + lane;
    Real bacc = 0, pendb = Real(1);
    int Fr = 0, Fw = 0, mxe = 0;
    uint32_t s = T.n_states - 1;
    int lvl = live ? (int)A.nlev[li] - 1 : 0;
    int E_cur = live ? lve[(size_t)lvl * 32] : 0;
    int E_next = (live && lvl >= 1) ? lve[(size_t)(lvl - 1) * 32] : 0;
    double cwE = cw * exp2((double)(Efin - E_cur));
    Real a_cur = ag[(size_t)s * 32];
    Real a_next = s >= 1 ? ag[(size_t)(s - 1) * 32] : Real(0);
    double as = (double)a_cur * cwE;
    uint2 r0[kLaneU], r1[kLaneU];
    Real e0[kLaneU];
    uint32_t q0[kLaneU];
    const uint32_t R = T.rows_b;
#pragma unroll
    for (int k = 0; k < kLaneU; ++k) r0[k] = bw[(size_t)k * 32];
#pragma unroll
    for (int k = 0; k < kLaneU; ++k) {
      e0[k] = w[r0[k].y];
      q0[k] = wsl[r0[k].y];
    }
#pragma unroll
    for (int k = 0; k < kLaneU; ++k) r1[k] = bw[(size_t)(kLaneU + k) * 32];
    // state parts: class of state s-1 two closes ahead, weight / slot of the closing state one close ahead
    uint32_t vc_next = s >= 1 ? vcl[(size_t)(s - 1) * 32] : 0u;
    Real vw_cur;
    uint32_t vs_cur;
    {
      const uint32_t vc = vcl[(size_t)s * 32];
      vw_cur = vwt[vc];
      vs_cur = vsl[vc];
    }
    for (uint32_t r = 0; r < R; r += kLaneU) {
      Real e1[kLaneU];
      uint32_t q1[kLaneU];
      uint2 r2[kLaneU];
#pragma unroll
      for (int k = 0; k < kLaneU; ++k) {
        e1[k] = w[r1[k].y];
        q1[k] = wsl[r1[k].y];
      }
#pragma unroll
      for (int k = 0; k < kLaneU; ++k) r2[k] = bw[(size_t)(r + 2 * kLaneU + k) * 32];
#pragma unroll
      for (int k = 0; k < kLaneU; ++k) {
        const uint32_t x = r0[k].x;
        const Real tt = e0[k] * ring[(x & (kLaneRing - 1)) * 32];
        bacc += tt;
        const double cval = as * (double)tt;
        if (cval > 0 && q0[k] != kSlotNone && !A.no_counts) count_add(A.sink, q0[k], cval);
        if (x & kLaneLast) {  // warp-uniform
          const Real b = (s == fin) ? Real(1) : bacc * pendb;
          {  // the state part: its expected count is the state posterior; predecessors see V * beta
            const double g = (s == fin) ? as : as * (double)bacc;
            if (vs_cur != kSlotNone && g > 0 && !A.no_counts) count_add(A.sink, vs_cur, g);
          }
          bacc = 0;
          ring[(s & (kLaneRing - 1)) * 32] = b * vw_cur;
          vw_cur = vwt[vc_next];
          vs_cur = vsl[vc_next];
          vc_next = (s >= 2 && s != 0xFFFFFFFFu) ? vcl[(size_t)(s - 2) * 32] : 0u;
          mxe = max(mxe, Num<Real>::expo(b));
          if (x & kLaneLevelEnd) {  // per lane: the lane's level `lvl` is complete (its lowest state was s)
            int shift = 0;
            if (mxe != 0 && (mxe < Num<Real>::kLo || mxe > Num<Real>::kHi)) shift = Num<Real>::kBias - mxe;
            Fr = Fw;
            pendb = Num<Real>::scale2(Real(1), shift);
            Fw = Fr + shift;
            mxe = 0;
            --lvl;
            E_cur = E_next;
            E_next = lvl >= 1 ? lve[(size_t)(lvl - 1) * 32] : 0;
            cwE = cw * exp2((double)(Efin - E_cur - Fr));
          }
          --s;  // (wraps after state 0: the sweep is over)
          a_cur = a_next;
          a_next = (s >= 1 && s != 0xFFFFFFFFu) ? ag[(size_t)(s - 1) * 32] : Real(0);
          as = (double)a_cur * cwE;
        }
      }
#pragma unroll
      for (int k = 0; k < kLaneU; ++k) {
        r0[k] = r1[k];
        e0[k] = e1[k];
        q0[k] = q1[k];
        r1[k] = r2[k];
      }
    }
  }
}

}  // namespace cmlk
