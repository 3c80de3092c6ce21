// cml_dense.cu -- the dense-state E-step (K7): forward / backward / expected counts over position-synchronous
// lattices that are never materialised.
//
// Reference semantics (restated, not copied): the same per-example computation as the lattice kernels --
//   derivations::compute / derive      carmel/src/derivations.h:479-513,640-704   (lattice state = (position, WFST state))
//   compute_fb                         carmel/src/derivations.h:400-417
//   collect_counts                     carmel/src/derivations.h:432-449
//   cascade_parameters chains          carmel/src/cascade.h:426-433 (arc weight = product of its chain)
// -- for a transducer whose every arc consumes exactly one observed symbol.  Then the lattice of a sequence
// o_1..o_n has the states (t, s) and between positions t-1 and t exactly the arcs (i -> j, o_t); when the chain
// of every arc splits into a part that depends on (i, j) only and a part that depends on (j, o) only,
//     w(i -> j, o) = T[i][j] * E[j][o],
// one position step of a BATCH of sequences is the dense product  alpha_t = (alpha_{t-1} . T) o E[:, o_t]
// ([batch x S] . [S x S], then the emission column of each row's observed symbol).  Lattice states the
// reference would prune (unreachable or dead) simply carry alpha = 0 or beta = 0, so likelihoods and expected
// counts are those of the pruned lattice.
//
// (A variant with a PAIR of warps per sequence -- forward and backward recurrences concurrently, counts in a
// position-parallel pass -- was measured and dropped: 126 registers per thread cap the SM at 16 warps either way, so
// twice the warps per sequence meant half the resident sequences: 69 us against 59 us.)
// B200 mapping.  S <= 32: one lane per WFST state, one warp per sequence (a warp walks several sequences, a CTA
// has 8 warps).  T lives in registers (a column per lane in the forward sweep, a row per lane in the backward
// sweep); the state vector of the current position is exchanged through a 2 x 32 shared-memory buffer read back
// as 128-bit broadcasts, so a position step costs 32 FMAs + 8 LDS.128 + one redux.sync per lane and touches HBM
// only for the alpha row (written once, read once) and 2 bytes of symbol.  Scores are linear and renormalised by
// an exact power of two EVERY step (redux.sync.max on the bit patterns), exponents accumulated per row.
// Expected counts: gamma_t(j) goes to the E cell (j, o_t), xi_t(i, j) to the T cell (i, j) (kept only when some
// transition parameter is trainable: 32 more FMAs per step on the same shared-memory reads).  Small alphabets
// keep per-warp private gamma tables in shared memory (no atomics in the sweep, one RED per cell and CTA at the
// end); large alphabets issue fp64 REDs on the cell's count slot directly.
//
// The M-step is the lattice path's (cml_maximize): cml_add_sequences re-declares the count slots as the
// trainable T / E cells with their parameter chains.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>

#include "cml_ctx.cuh"
#include "cml_kernels_model.cuh"

namespace {

constexpr int kDW = 8;    // warps per CTA
constexpr int kDS = 32;   // padded state count (lanes)
constexpr uint32_t kNone = 0xFFFFFFFFu;

template <typename Real>
struct DN;
template <>
struct DN<float> {
  static __device__ __forceinline__ int bits(float x) { return __float_as_int(x); }  // monotone for x >= 0
  static __device__ __forceinline__ int expo(int b) { return min(max(((b >> 23) & 0xff) - 127, -126), 126); }
  static __device__ __forceinline__ float pow2(int k) { return __int_as_float((127 + k) << 23); }
};
template <>
struct DN<double> {
  static __device__ __forceinline__ int bits(double x) { return __double2hiint(x); }
  static __device__ __forceinline__ int expo(int b) { return min(max(((b >> 20) & 0x7ff) - 1023, -1022), 1022); }
  static __device__ __forceinline__ double pow2(int k) { return __hiloint2double((1023 + k) << 20, 0); }
};
// 2^k as a double for any k (0 below the normal range: such a posterior is < 1e-308)
__device__ __forceinline__ double pow2d(int k) {
  if (k < -1022) return 0.;
  return __hiloint2double((1023 + min(k, 1023)) << 20, 0);
}
// renormalise a warp's state vector by the power of two of its largest entry
template <typename Real>
__device__ __forceinline__ void renorm(Real& a, int& E) {
  const int mx = __reduce_max_sync(0xffffffffu, DN<Real>::bits(a));
  if (mx > 0) {
    const int e = DN<Real>::expo(mx);
    a *= DN<Real>::pow2(-e);
    E += e;
  }
}
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4*>(p);
  v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
__device__ __forceinline__ void ld4(const double* p, double (&v)[4]) {
  const double2 q0 = *reinterpret_cast<const double2*>(p), q1 = *reinterpret_cast<const double2*>(p + 2);
  v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y;
}

struct DenseArgs {
  const uint64_t* seq_off;
  const uint16_t* sym;
  const double* seq_weight;
  uint32_t n_seq, n_sym, start, fin;
  const void* T;             // Real T[32][32] (row = source state), zero padded
  const void* Et;            // Real Et[n_sym][32] (E[j][o] at o*32+j), zero padded
  const uint32_t* cell_slot; // [1024 T cells | n_sym*32 E cells]: count slot or kNone
  double* counts;            // the reduce buffer's count slots
  double* ex_lnp;
  void* alpha_g;
  int* exp_g;
};

// K6 for the dense view: T / E cell value = product of the cell's parameter chain (0 for absent cells)
template <typename Real>
__global__ void k_dense_tables(uint32_t n_cells, const uint32_t* __restrict__ cell_off,
                               const uint32_t* __restrict__ cell_param, const unsigned char* __restrict__ exists,
                               const double* __restrict__ ln_w, Real* __restrict__ out) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  double s = -CUDART_INF;
  if (exists[c]) {
    s = 0;
    for (uint32_t k = cell_off[c], e = cell_off[c + 1]; k < e; ++k) s += ln_w[cell_param[k]];
  }
  out[c] = (Real)exp(s);
}

// shared memory: Real buf[kDW][2][32] | Real Ts[32][32] | Real Tt[32][32] | (GSM: Real Es[n_sym][32]) |
//                (GSM: double gam[kDW][n_sym][32]) | (XI: double xis[32][32], transposed)
template <typename Real>
static size_t dense_smem(uint32_t n_sym, bool xi, bool gsm) {
  size_t b = (size_t)kDW * 64 * sizeof(Real) + 2 * 1024 * sizeof(Real);
  if (gsm) b += (size_t)n_sym * 32 * sizeof(Real);
  b = (b + 15) & ~(size_t)15;
  if (gsm) b += (size_t)kDW * n_sym * 32 * sizeof(double);
  if (xi) b += 1024 * sizeof(double);
  return b;
}

template <typename Real, bool XI, bool GSM>
__global__ void __launch_bounds__(kDW * 32) k_fb_dense(DenseArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t n_sym = A.n_sym;
  Real* buf = reinterpret_cast<Real*>(smem) + w * 64;
  Real* Ts = reinterpret_cast<Real*>(smem) + kDW * 64;
  Real* Tt = Ts + 1024;
  size_t off = ((size_t)kDW * 64 + 2048) * sizeof(Real);
  const Real* Et = reinterpret_cast<const Real*>(A.Et);
  if (GSM) {
    Real* Es = reinterpret_cast<Real*>(smem + off);
    for (uint32_t i = threadIdx.x; i < n_sym * 32; i += blockDim.x) Es[i] = Et[i];
    Et = Es;
    off += (size_t)n_sym * 32 * sizeof(Real);
  }
  off = (off + 15) & ~(size_t)15;
  double* gam_all = reinterpret_cast<double*>(smem + off);
  double* gam = gam_all + (size_t)w * n_sym * 32;
  if (GSM) {
    for (uint32_t i = threadIdx.x; i < kDW * n_sym * 32; i += blockDim.x) gam_all[i] = 0.;
    off += (size_t)kDW * n_sym * 32 * sizeof(double);
  }
  double* xis = reinterpret_cast<double*>(smem + off);
  {
    const Real* Tg = reinterpret_cast<const Real*>(A.T);
    for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) {
      const Real v = Tg[i];
      Ts[i] = v;
      Tt[(i & 31) * 32 + (i >> 5)] = v;
      if (XI) xis[i] = 0.;
    }
  }
  __syncthreads();
  const uint32_t* __restrict__ e_slot = A.cell_slot + 1024;
  Real* __restrict__ ag = reinterpret_cast<Real*>(A.alpha_g);

  for (uint64_t e = (uint64_t)blockIdx.x * kDW + w; e < A.n_seq; e += (uint64_t)gridDim.x * kDW) {
    const uint64_t base = A.seq_off[e];
    const uint32_t n = (uint32_t)(A.seq_off[e + 1] - base);
    const uint64_t r0 = base + e;  // first alpha row of this sequence (n + 1 rows)
    const uint16_t* __restrict__ sy = A.sym + base;
    // ---------------------------------------------------------------- forward
    Real a = (lane == (int)A.start) ? Real(1) : Real(0);
    int Ea = 0;
    ag[r0 * 32 + lane] = a;
    if (lane == 0) A.exp_g[r0] = 0;
    {
      Real Tcol[32];  // T[i][lane]
#pragma unroll
      for (int i = 0; i < 32; ++i) Tcol[i] = Ts[i * 32 + lane];
      uint32_t symv = ((uint32_t)lane < n) ? sy[lane] : 0u;
      uint32_t symn = (32u + lane < n) ? sy[32 + lane] : 0u;
      Real ev = Et[__shfl_sync(0xffffffffu, symv, 0) * 32 + lane];
      int p = 0;
      for (uint32_t t = 0; t < n; ++t) {
        buf[p * 32 + lane] = a;
        __syncwarp();
        // emission column of the next step (off the critical path)
        Real evn = 0;
        if (t + 1 < n) {
          if (((t + 1) & 31) == 0) {
            symv = symn;
            symn = (t + 33 + lane < n) ? sy[t + 33 + lane] : 0u;
          }
          evn = Et[__shfl_sync(0xffffffffu, symv, (t + 1) & 31) * 32 + lane];
        }
        Real acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          Real v[4];
          ld4(buf + p * 32 + i, v);
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[k] = fma(v[k], Tcol[i + k], acc[k]);
        }
        a = ((acc[0] + acc[1]) + (acc[2] + acc[3])) * ev;
        renorm<Real>(a, Ea);
        ag[(r0 + t + 1) * 32 + lane] = a;
        if (lane == 0) A.exp_g[r0 + t + 1] = Ea;
        ev = evn;
        p ^= 1;
      }
    }
    const Real afin = __shfl_sync(0xffffffffu, a, A.fin);
    const int EaN = Ea;
    if (lane == 0)
      A.ex_lnp[e] = (afin > 0) ? log((double)afin) + (double)EaN * 0.69314718055994530942 : -CUDART_INF;
    if (!(afin > 0)) continue;  // zero-probability sequence: no counts (warp-uniform)
    const double cw = A.seq_weight[e] / (double)afin;
    __syncwarp();
    // ---------------------------------------------------------------- backward + counts
    {
      Real Trow[32];  // T[lane][j]
#pragma unroll
      for (int j = 0; j < 32; ++j) Trow[j] = Tt[j * 32 + lane];
      Real xi[XI ? 32 : 1];
      if (XI) {
#pragma unroll
        for (int j = 0; j < 32; ++j) xi[j] = 0;
      }
      Real b = (lane == (int)A.fin) ? Real(1) : Real(0);
      int Eb = 0;
      Real a1 = a;  // alpha row n
      int Ea1 = EaN;
      // symbols in reverse: reverse index u = n-1-t
      uint32_t symv = ((uint32_t)lane < n) ? sy[n - 1 - lane] : 0u;
      uint32_t symn = (32u + lane < n) ? sy[n - 33 - lane] : 0u;
      uint32_t o = __shfl_sync(0xffffffffu, symv, 0);
      Real ev = Et[o * 32 + lane];
      int p = 0;
      for (uint32_t u = 0; u < n; ++u) {
        const uint32_t t = n - 1 - u;
        // alpha row t and its exponent (needed at the end of this step and in the next one)
        const Real a0 = ag[(r0 + t) * 32 + lane];
        const int Ea0 = A.exp_g[r0 + t];
        const Real bt = b * ev;  // E[j][o_t] * beta_{t+1}[j]
        // posterior of lattice state (t+1, lane) -> E cell (lane, o_t)
        const double gamma = (double)a1 * (double)b * (cw * pow2d(Ea1 + Eb - EaN));
        if (GSM) {
          gam[o * 32 + lane] += gamma;
        } else if (gamma > 0) {
          const uint32_t sl = e_slot[o * 32 + lane];
          if (sl != kNone) atomicAdd(A.counts + sl, gamma);
        }
        buf[p * 32 + lane] = bt;
        __syncwarp();
        Real evn = 0;
        uint32_t on = 0;
        if (u + 1 < n) {
          if (((u + 1) & 31) == 0) {
            symv = symn;
            symn = (u + 33 + lane < n) ? sy[n - 1 - (u + 33 + lane)] : 0u;
          }
          on = __shfl_sync(0xffffffffu, symv, (u + 1) & 31);
          evn = Et[on * 32 + lane];
        }
        Real x0 = 0;
        if (XI) x0 = (Real)((double)a0 * (cw * pow2d(Ea0 + Eb - EaN)));
        Real acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          Real v[4];
          ld4(buf + p * 32 + j, v);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc[k] = fma(v[k], Trow[j + k], acc[k]);
            if (XI) xi[j + k] = fma(x0, v[k], xi[j + k]);
          }
        }
        b = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        renorm<Real>(b, Eb);
        a1 = a0;
        Ea1 = Ea0;
        ev = evn;
        o = on;
        p ^= 1;
        if (XI && ((u & 255) == 255 || u + 1 == n)) {  // fold the xi partial sums (fp64, per CTA)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const double v = (double)xi[j] * (double)Trow[j];
            if (v != 0.) atomicAdd(&xis[j * 32 + lane], v);
            xi[j] = 0;
          }
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (GSM) {
    for (uint32_t c = threadIdx.x; c < n_sym * 32; c += blockDim.x) {
      const uint32_t sl = e_slot[c];
      if (sl == kNone) continue;
      double s = 0;
#pragma unroll
      for (int k = 0; k < kDW; ++k) s += gam_all[(size_t)k * n_sym * 32 + c];
      if (s != 0.) atomicAdd(A.counts + sl, s);
    }
  }
  if (XI) {
    for (uint32_t c = threadIdx.x; c < 1024; c += blockDim.x) {  // c = j*32 + i (transposed)
      const uint32_t sl = A.cell_slot[(c & 31) * 32 + (c >> 5)];
      const double s = xis[c];
      if (sl != kNone && s != 0.) atomicAdd(A.counts + sl, s);
    }
  }
}

// =====================================================================================================
// Tensor-core variant of the dense-state sweep for LARGE batches in fp32 mode (north_star: the position step of a
// batch of sequences as a dense GEMM in 3xTF32).  One warp carries 16 sequences (the M dimension of
// mma.sync.m16n8k8): per position  C[16 x 32] = A[16 x 32] . T[32 x 32]  as 4 x 4 tiles of m16n8k8 TF32 MMAs, each
// product issued three times (A_hi T_hi + A_lo T_hi + A_hi T_lo, fp32 accumulate) so the result keeps fp32
// accuracy; the 16 x 32 state matrix goes through a padded shared-memory tile between the accumulator layout and
// the A-operand layout; emission columns, per-row power-of-two renormalisation and the row stores stay in
// registers.  Forward and backward are two launches of the same kernel (B operand = T or T^T, emission applied
// after / before the product); the expected counts are a third, position-parallel launch (one warp per sequence,
// the pair's private gamma table) -- 16 sequences per warp would collide on the count cells.
// Used for fp32 contexts with >= 16,384 sequences (CML_DENSE_TC=1 forces it, =0 forbids it):
// with the 2,000 lines of the bench corpus there would be 125 warps for 148 SMs.  This is the legacy warp-level MMA
// path, not tcgen05: a 16-row tile per warp keeps the serial chain of a sequence inside one warp's registers, where
// a 128-row tcgen05 tile would need a TMEM -> register -> shared-memory round trip per position.
// =====================================================================================================
struct TcArgs {
  const uint64_t* seq_off;
  const uint16_t* sym;
  const double* seq_weight;
  uint32_t n_seq, n_groups, n_sym, start, fin;
  const uint32_t* order;  // [n_groups * 16] sequence of every row (longest first), 0xFFFFFFFF = empty row
  const float* T;         // [32][32]
  const float* Et;        // [n_sym][32]
  float* rows;            // alpha_g (forward) / beta_g (backward)
  int* exps;              // exp_g / bexp_g
  const float* arows;     // counts: alpha rows
  const int* aexps;
  double* ex_lnp;
  double* afin_g;         // per sequence: alpha_n[fin] (renormalised) and its exponent
  int* ean_g;
  const uint32_t* cell_slot;
  double* counts;
};
constexpr int kTcWarps = 4, kTcStride = 36;  // row stride of the shared state tile (conflict-free fragment loads)

__device__ __forceinline__ void tf32_split(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// XI (backward launch only): the transition counts on the tensor cores as well.  xi[i][j] = T[i][j] * sum over
// sequences b and positions t of (alpha_t[b][i] * scale_b) * betatilde_{t+1}[b][j] is, per position, the product
// [32 x 16] . [16 x 32] (K = the 16 sequences of the warp): 2 x 4 output tiles x 2 k-steps of m16n8k8, in 3xTF32,
// accumulated in 32 fp32 registers per lane and folded into an fp64 table every 64 positions -- no atomics per
// position at all (the FMA kernels pay one shared-memory CAS loop or 32 register FMAs per lane for this).
template <bool BWD, bool XI>
__global__ void __launch_bounds__(kTcWarps * 32) k_dense_tc(TcArgs A) {
  extern __shared__ __align__(16) float smem_tc[];
  __shared__ double xis[XI ? 1024 : 1];
  float* Es = smem_tc;
  for (uint32_t i = threadIdx.x; i < A.n_sym * 32; i += blockDim.x) Es[i] = A.Et[i];
  if (XI)
    for (uint32_t i = threadIdx.x; i < 1024; i += blockDim.x) xis[i] = 0.;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t grp = blockIdx.x * kTcWarps + warp;
  if (grp < A.n_groups) {
  float* X = smem_tc + A.n_sym * 32 + warp * (XI ? 3 : 1) * 16 * kTcStride;
  float* Y = X + 16 * kTcStride;  // XI: alpha_t rows, scaled     [sequence][state]
  float* Z = Y + 16 * kTcStride;  // XI: betatilde_{t+1} rows      [sequence][state]
  const int g = lane >> 2, tq = lane & 3;
  const unsigned qmask = 0xFu << (lane & ~3);  // the quad that shares this lane's two rows
  // this lane's rows: g and g + 8
  uint32_t e[2], n[2];
  uint64_t base[2], r0[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    e[h] = A.order[grp * 16 + g + 8 * h];
    const bool ok = e[h] != 0xFFFFFFFFu;
    base[h] = ok ? A.seq_off[e[h]] : 0;
    n[h] = ok ? (uint32_t)(A.seq_off[e[h] + 1] - base[h]) : 0u;
    r0[h] = base[h] + (ok ? e[h] : 0);
  }
  const uint32_t nmax = __shfl_sync(0xffffffffu, n[0], 0);  // row 0 of the group is its longest sequence
  double cwr[2] = {0., 0.};  // XI: weight / P of this lane's rows, and the exponent of P
  int eanr[2] = {0, 0};
  float cxi[2][4][4];        // XI: accumulators of the 2 x 4 output tiles
  if (XI) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (e[h] != 0xFFFFFFFFu) {
        const double af = A.afin_g[e[h]];
        cwr[h] = af > 0 ? A.seq_weight[e[h]] / af : 0.;
        eanr[h] = A.ean_g[e[h]];
      }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) cxi[mt][nt][k] = 0.f;
  }
  auto fold_xi = [&]() {  // accumulators -> the CTA's fp64 table (times T), then cleared
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = 16 * mt + g + 8 * (k >> 1), j = 8 * nt + 2 * tq + (k & 1);
          const double v = (double)cxi[mt][nt][k] * (double)A.T[i * 32 + j];
          if (v != 0.) atomicAdd(&xis[XI ? i * 32 + j : 0], v);
          cxi[mt][nt][k] = 0.f;
        }
  };
  // B operand: T (forward: B[k][n] = T[k][n]) or T^T (backward: B[k = j][n = i] = T[i][j]), split once
  uint32_t bh[4][4][2], bl[4][4][2];
#pragma unroll
  for (int kt = 0; kt < 4; ++kt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int k0 = 8 * kt + tq, k1 = k0 + 4, nn = 8 * nt + g;
      const float v0 = BWD ? A.T[nn * 32 + k0] : A.T[k0 * 32 + nn];
      const float v1 = BWD ? A.T[nn * 32 + k1] : A.T[k1 * 32 + nn];
      tf32_split(v0, bh[kt][nt][0], bl[kt][nt][0]);
      tf32_split(v1, bh[kt][nt][1], bl[kt][nt][1]);
    }
  int E[2] = {0, 0};  // cumulative exponent of each row
  // one-hot start (forward) / final (backward) rows, in the accumulator layout: cols 8nt + 2tq, +1
  auto put_onehot = [&](int h, uint32_t hot, uint64_t hbm_row) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const uint32_t c0 = 8 * nt + 2 * tq;
      const float2 v = make_float2(c0 == hot ? 1.f : 0.f, c0 + 1 == hot ? 1.f : 0.f);
      *reinterpret_cast<float2*>(&X[(g + 8 * h) * kTcStride + c0]) = v;
      if (e[h] != 0xFFFFFFFFu) *reinterpret_cast<float2*>(&A.rows[hbm_row * 32 + c0]) = v;
    }
    if (e[h] != 0xFFFFFFFFu && tq == 0) A.exps[hbm_row] = 0;
    E[h] = 0;
  };
  if (!BWD) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      put_onehot(h, A.start, r0[h]);
      if (e[h] != 0xFFFFFFFFu && n[h] == 0 && tq == 0) {  // empty sequence: P = [start == final]
        A.ex_lnp[e[h]] = A.start == A.fin ? 0. : -CUDART_INF;
        A.afin_g[e[h]] = A.start == A.fin ? 1. : 0.;
        A.ean_g[e[h]] = 0;
      }
    }
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // rows shorter than the group start later; clear them, start the longest now
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<float2*>(&X[(g + 8 * h) * kTcStride + 8 * nt + 2 * tq]) = make_float2(0.f, 0.f);
      if (n[h] == nmax || n[h] == 0) put_onehot(h, A.fin, r0[h] + n[h]);
    }
  }
  __syncwarp();
  for (uint32_t s = 0; s < nmax; ++s) {
    const uint32_t t = BWD ? nmax - 1 - s : s;  // link between positions t and t+1, symbol o_t
    uint32_t o[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) o[h] = t < n[h] ? A.sym[base[h] + t] : 0u;
    if (BWD) {  // rows whose sequence ends at t+1 start here with beta = one-hot(final)
      bool any = false;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (n[h] == t + 1 && n[h] != nmax) {
          put_onehot(h, A.fin, r0[h] + n[h]);
          any = true;
        }
      if (__any_sync(0xffffffffu, any)) __syncwarp();
    }
    if (BWD && XI) {
      // stage betatilde_{t+1} (this lane's elements of the state tile times the emission column) and the scaled
      // alpha_t rows, then one [32 x 16] . [16 x 32] product into the xi accumulators
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool on = e[h] != 0xFFFFFFFFu && t < n[h] && cwr[h] > 0.;
        float sc = 0.f;
        if (on) sc = (float)(cwr[h] * pow2d(A.aexps[r0[h] + t] + E[h] - eanr[h]));
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int c0 = 8 * nt + 2 * tq;
          const float2 xv = *reinterpret_cast<const float2*>(&X[(g + 8 * h) * kTcStride + c0]);
          const float2 ev = *reinterpret_cast<const float2*>(&Es[o[h] * 32 + c0]);
          *reinterpret_cast<float2*>(&Z[(g + 8 * h) * kTcStride + c0]) = on ? make_float2(xv.x * ev.x, xv.y * ev.y) : make_float2(0.f, 0.f);
          float2 av = make_float2(0.f, 0.f);
          if (on) {
            av = *reinterpret_cast<const float2*>(&A.arows[(r0[h] + t) * 32 + c0]);
            av.x *= sc;
            av.y *= sc;
          }
          *reinterpret_cast<float2*>(&Y[(g + 8 * h) * kTcStride + c0]) = av;
        }
      }
      __syncwarp();
      uint32_t zh[2][4][2], zl[2][4][2];  // B' = betatilde: [k = sequence][n = j]
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          tf32_split(Z[(8 * ks + tq) * kTcStride + 8 * nt + g], zh[ks][nt][0], zl[ks][nt][0]);
          tf32_split(Z[(8 * ks + tq + 4) * kTcStride + 8 * nt + g], zh[ks][nt][1], zl[ks][nt][1]);
        }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t yh[4], yl[4];  // A' = alpha^T: [m = i][k = sequence]
          tf32_split(Y[(8 * ks + tq) * kTcStride + 16 * mt + g], yh[0], yl[0]);
          tf32_split(Y[(8 * ks + tq) * kTcStride + 16 * mt + g + 8], yh[1], yl[1]);
          tf32_split(Y[(8 * ks + tq + 4) * kTcStride + 16 * mt + g], yh[2], yl[2]);
          tf32_split(Y[(8 * ks + tq + 4) * kTcStride + 16 * mt + g + 8], yh[3], yl[3]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            mma_tf32(cxi[mt][nt], yl, zh[ks][nt]);
            mma_tf32(cxi[mt][nt], yh, zl[ks][nt]);
            mma_tf32(cxi[mt][nt], yh, zh[ks][nt]);
          }
        }
      if ((s & 63) == 63) fold_xi();
      __syncwarp();
    }
    // A operand from the shared state tile (backward: times the emission column of each row's symbol)
    uint32_t ah[4][4], al[4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const int k0 = 8 * kt + tq, k1 = k0 + 4;
      float a0 = X[g * kTcStride + k0], a1 = X[(g + 8) * kTcStride + k0];
      float a2 = X[g * kTcStride + k1], a3 = X[(g + 8) * kTcStride + k1];
      if (BWD) {
        a0 *= Es[o[0] * 32 + k0];
        a1 *= Es[o[1] * 32 + k0];
        a2 *= Es[o[0] * 32 + k1];
        a3 *= Es[o[1] * 32 + k1];
      }
      tf32_split(a0, ah[kt][0], al[kt][0]);
      tf32_split(a1, ah[kt][1], al[kt][1]);
      tf32_split(a2, ah[kt][2], al[kt][2]);
      tf32_split(a3, ah[kt][3], al[kt][3]);
    }
    __syncwarp();  // everybody has read the tile: it can be overwritten below
    float c[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int k = 0; k < 4; ++k) c[nt][k] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        mma_tf32(c[nt], al[kt], bh[kt][nt]);
        mma_tf32(c[nt], ah[kt], bl[kt][nt]);
        mma_tf32(c[nt], ah[kt], bh[kt][nt]);
      }
    }
    // epilogue per row: (forward: emission column), power-of-two renormalisation, stores
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[8];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        v[2 * nt] = c[nt][2 * h];
        v[2 * nt + 1] = c[nt][2 * h + 1];
        if (!BWD) {
          const float2 ev = *reinterpret_cast<const float2*>(&Es[o[h] * 32 + 8 * nt + 2 * tq]);
          v[2 * nt] *= ev.x;
          v[2 * nt + 1] *= ev.y;
        }
      }
      int mx = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) mx = max(mx, __float_as_int(v[k]));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      if (mx > 0) {
        const int ex = DN<float>::expo(mx);
        const float f = DN<float>::pow2(-ex);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= f;
        E[h] += ex;
      }
      const bool active = e[h] != 0xFFFFFFFFu && t < n[h];  // this row really has a position t
      const uint64_t hrow = r0[h] + (BWD ? t : t + 1);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float2 w2 = make_float2(v[2 * nt], v[2 * nt + 1]);
        *reinterpret_cast<float2*>(&X[(g + 8 * h) * kTcStride + 8 * nt + 2 * tq]) = w2;
        if (active) *reinterpret_cast<float2*>(&A.rows[hrow * 32 + 8 * nt + 2 * tq]) = w2;
      }
      if (active && tq == 0) A.exps[hrow] = E[h];
      if (!BWD && e[h] != 0xFFFFFFFFu && t + 1 == n[h]) {  // the row's last position: P = alpha_n[final] (quad-uniform branch)
        const int ft = A.fin >> 3, fc = A.fin & 7;
        float mine = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          if (nt == ft) mine = (fc & 1) ? v[2 * nt + 1] : v[2 * nt];
        const float afin = __shfl_sync(qmask, mine, (lane & ~3) | (fc >> 1));
        if (tq == 0) {
          A.ex_lnp[e[h]] = (afin > 0) ? log((double)afin) + (double)E[h] * 0.69314718055994530942 : -CUDART_INF;
          A.afin_g[e[h]] = (double)afin;
          A.ean_g[e[h]] = E[h];
        }
      }
    }
    __syncwarp();  // the new tile is complete before the next step's fragment loads
  }
  if (BWD && XI) fold_xi();
  }  // grp < n_groups
  if (BWD && XI) {
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < 1024; c += blockDim.x) {
      const uint32_t sl = A.cell_slot[c];
      const double v = xis[XI ? c : 0];
      if (sl != kNone && v != 0.) atomicAdd(A.counts + sl, v);
    }
  }
}

// expected counts of the tensor-core sweeps: position-parallel, one warp per sequence, lane = state
__global__ void __launch_bounds__(kDW * 32) k_dense_tc_counts(TcArgs A) {
  extern __shared__ __align__(16) unsigned char smem_tcc[];
  double* gam_all = reinterpret_cast<double*>(smem_tcc);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t n_sym = A.n_sym;
  for (uint32_t i = threadIdx.x; i < kDW * n_sym * 32; i += blockDim.x) gam_all[i] = 0.;
  __syncthreads();
  double* gam = gam_all + (size_t)w * n_sym * 32;
  for (uint64_t e = (uint64_t)blockIdx.x * kDW + w; e < A.n_seq; e += (uint64_t)gridDim.x * kDW) {
    const double afin = A.afin_g[e];
    if (!(afin > 0)) continue;
    const int EaN = A.ean_g[e];
    const double cw = A.seq_weight[e] / afin;
    const uint64_t base = A.seq_off[e];
    const uint32_t n = (uint32_t)(A.seq_off[e + 1] - base);
    const uint64_t r0 = base + e;
    // four positions per round: all loads first (the kernel is a pure stream of alpha / beta rows), then the
    // dependent shared-memory updates
    for (uint32_t t0 = 0; t0 < n; t0 += 4) {
      float a1[4], b1[4];
      int de[4];
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t t = min(t0 + k, n - 1);
        o[k] = A.sym[base + t];
        a1[k] = A.arows[(r0 + t + 1) * 32 + lane];
        b1[k] = A.rows[(r0 + t + 1) * 32 + lane];
        de[k] = A.aexps[r0 + t + 1] + A.exps[r0 + t + 1] - EaN;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (t0 + k < n) gam[o[k] * 32 + lane] += (double)a1[k] * (double)b1[k] * (cw * pow2d(de[k]));
    }
  }
  __syncthreads();
  const uint32_t* __restrict__ e_slot = A.cell_slot + 1024;
  for (uint32_t c = threadIdx.x; c < n_sym * 32; c += blockDim.x) {
    const uint32_t sl = e_slot[c];
    if (sl == kNone) continue;
    double s = 0;
#pragma unroll
    for (int k = 0; k < kDW; ++k) s += gam_all[(size_t)k * n_sym * 32 + c];
    if (s != 0.) atomicAdd(A.counts + sl, s);
  }
}

static int launch_dense_tc(cml_ctx* ctx) {
  DenseState& D = *ctx->dense;
  cudaStream_t s = ctx->stream;
  k_dense_tables<float><<<cdiv(D.n_cells, 256), 256, 0, s>>>(D.n_cells, D.cell_off.p, D.cell_param.p, D.cell_exists.p,
                                                            ctx->ln_w.p, reinterpret_cast<float*>(D.tables.p));
  ++ctx->launches;
  CML_CUDA(cudaMemsetAsync(ctx->reduce, 0, ctx->reduce_n * sizeof(double), s));
  TcArgs A;
  A.seq_off = D.seq_off.p;
  A.sym = D.sym.p;
  A.seq_weight = D.seq_weight.p;
  A.n_seq = (uint32_t)D.n_seq;
  A.n_groups = D.tc_groups;
  A.n_sym = D.n_sym;
  A.start = D.start;
  A.fin = D.fin;
  A.order = D.tc_order.p;
  A.T = reinterpret_cast<const float*>(D.tables.p);
  A.Et = A.T + 1024;
  A.ex_lnp = D.ex_lnp.p;
  A.afin_g = D.tc_afin.p;
  A.ean_g = D.tc_ean.p;
  A.cell_slot = D.cell_slot.p;
  A.counts = ctx->reduce;
  A.arows = reinterpret_cast<const float*>(D.alpha_g.p);
  A.aexps = D.exp_g.p;
  const bool xi = D.n_t_slots > 0;
  const size_t smem = ((size_t)D.n_sym * 32 + (size_t)kTcWarps * 16 * kTcStride) * sizeof(float);
  const size_t smem_b = ((size_t)D.n_sym * 32 + (size_t)kTcWarps * (xi ? 3 : 1) * 16 * kTcStride) * sizeof(float);
  const size_t csmem = (size_t)kDW * D.n_sym * 32 * sizeof(double);
  CML_CUDA(cudaFuncSetAttribute(k_dense_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CML_CUDA(cudaFuncSetAttribute(k_dense_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  CML_CUDA(cudaFuncSetAttribute(k_dense_tc<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  CML_CUDA(cudaFuncSetAttribute(k_dense_tc_counts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
  if (!D.ev0) {
    CML_CUDA(cudaEventCreate(&D.ev0));
    CML_CUDA(cudaEventCreate(&D.ev1));
  }
  if (!ctx->capturing) CML_CUDA(cudaEventRecord(D.ev0, s));
  if (D.tc_groups) {
    A.rows = reinterpret_cast<float*>(D.alpha_g.p);
    A.exps = D.exp_g.p;
    k_dense_tc<false, false><<<cdiv(D.tc_groups, kTcWarps), kTcWarps * 32, smem, s>>>(A);
    A.rows = reinterpret_cast<float*>(D.beta_g.p);
    A.exps = D.bexp_g.p;
    if (xi)
      k_dense_tc<true, true><<<cdiv(D.tc_groups, kTcWarps), kTcWarps * 32, smem_b, s>>>(A);
    else
      k_dense_tc<true, false><<<cdiv(D.tc_groups, kTcWarps), kTcWarps * 32, smem_b, s>>>(A);
    k_dense_tc_counts<<<std::max(1u, std::min<unsigned>(cdiv(D.n_seq, kDW), 4u * ctx->sm_count)), kDW * 32, csmem, s>>>(A);
    ctx->launches += 3;
  }
  if (!ctx->capturing) CML_CUDA(cudaEventRecord(D.ev1, s));
  if (D.n_seq) {
    cmlk::k_reduce_lnp<<<std::min<unsigned>(cdiv(D.n_seq, 256), 4 * ctx->sm_count), 256, 0, s>>>(
        D.ex_lnp.p, D.seq_weight.p, D.n_seq, ctx->reduce + ctx->n_slots);
    ++ctx->launches;
  }
  CML_CUDA(cudaGetLastError());
  return CML_OK;
}

template <typename Real>
int launch_dense(cml_ctx* ctx) {
  DenseState& D = *ctx->dense;
  cudaStream_t s = ctx->stream;
  k_dense_tables<Real><<<cdiv(D.n_cells, 256), 256, 0, s>>>(D.n_cells, D.cell_off.p, D.cell_param.p, D.cell_exists.p,
                                                           ctx->ln_w.p, reinterpret_cast<Real*>(D.tables.p));
  ++ctx->launches;
  CML_CUDA(cudaMemsetAsync(ctx->reduce, 0, ctx->reduce_n * sizeof(double), s));
  DenseArgs A;
  A.seq_off = D.seq_off.p;
  A.sym = D.sym.p;
  A.seq_weight = D.seq_weight.p;
  A.n_seq = (uint32_t)D.n_seq;
  A.n_sym = D.n_sym;
  A.start = D.start;
  A.fin = D.fin;
  A.T = D.tables.p;
  A.Et = D.tables.p + 1024 * sizeof(Real);
  A.cell_slot = D.cell_slot.p;
  A.counts = ctx->reduce;
  A.ex_lnp = D.ex_lnp.p;
  A.alpha_g = D.alpha_g.p;
  A.exp_g = D.exp_g.p;
  const bool xi = D.n_t_slots > 0;
  const bool gsm = dense_smem<Real>(D.n_sym, xi, true) <= 100 * 1024;
  const size_t smem = dense_smem<Real>(D.n_sym, xi, gsm);
  void (*kern)(DenseArgs) = xi ? (gsm ? k_fb_dense<Real, true, true> : k_fb_dense<Real, true, false>)
                               : (gsm ? k_fb_dense<Real, false, true> : k_fb_dense<Real, false, false>);
  CML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  CML_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDW * 32, smem));
  const unsigned grid = std::max(1u, std::min<unsigned>(cdiv(D.n_seq, kDW), (unsigned)(std::max(per_sm, 1) * ctx->sm_count)));
  if (!D.ev0) {
    CML_CUDA(cudaEventCreate(&D.ev0));
    CML_CUDA(cudaEventCreate(&D.ev1));
  }
  if (!ctx->capturing) CML_CUDA(cudaEventRecord(D.ev0, s));
  kern<<<grid, kDW * 32, smem, s>>>(A);
  ++ctx->launches;
  if (!ctx->capturing) CML_CUDA(cudaEventRecord(D.ev1, s));
  if (D.n_seq) {
    cmlk::k_reduce_lnp<<<std::min<unsigned>(cdiv(D.n_seq, 256), 4 * ctx->sm_count), 256, 0, s>>>(
        D.ex_lnp.p, D.seq_weight.p, D.n_seq, ctx->reduce + ctx->n_slots);
    ++ctx->launches;
  }
  CML_CUDA(cudaGetLastError());
  return CML_OK;
}

// =====================================================================================================
// Sparse-emission variant (k_fb_sparse): HMM / tagging style models where a symbol is emitted by only a few
// states (a word has <= 8 possible tags) and sequences are many and short.  One sequence per LANE (tiles of
// 32 sequences of similar length, symbol streams stored transposed sym[t][lane]); a lane keeps the alpha values
// of the <= K active states of the current position in registers, the transition matrix T (S <= 64) and the
// final weights live in shared memory, the emission row of a symbol is ONE gather of K values.  A position
// costs K*K shared-memory reads + FMAs and touches HBM for the symbol, K alpha values (written once, read once)
// and an exponent -- the lattice (K*K 16-byte arc records per position) is never streamed.
// epsilon arcs into a final state without outgoing arcs are final weights phi[s] (tagging.fsa style).
// =====================================================================================================
constexpr int kSW = 8;       // warps per CTA (sparse kernel)
constexpr int kSXi = 4;      // transition-count tables per CTA when they fit: one per pair of warps
struct SparseArgs {
  const SparseTile* tile;
  uint32_t n_tiles;
  const uint16_t* sym;
  const uint32_t* len;        // [tile*32+lane]
  const uint32_t* seq;        // [tile*32+lane] sequence index (0xFFFFFFFF = empty lane)
  const double* weight;       // [tile*32+lane]
  uint32_t S, SP, start, fin;
  const void* T;              // Real[SP*SP]
  const void* Ev;             // Real[n_sym*K]
  const void* Phi;            // Real[SP] final weights of the epsilon arcs (0 where absent)
  const unsigned char* e_state;  // [n_sym*K] state of every emission entry (SP-1, the zero state, for padding)
  const uint32_t* e_code;     // [n_sym*K] count-slot code of every emission entry (CountSink codes)
  const uint32_t* t_slot;     // [SP*SP] slot or kNone
  const uint32_t* f_slot;     // [SP]
  cmlk::CountSink sink;
  double* ex_lnp;
  void* alpha;
  int* exps;
  int has_phi;                // epsilon final arcs exist (otherwise P = alpha_n[fin])
  uint32_t xi_tables;         // shared-memory transition-count tables per CTA: kSW (one per warp) or 1
};

template <typename Real, int K>
struct RowVec {
  Real v[K];
};
template <typename Real, int K>
__device__ __forceinline__ void load_row(const Real* p, Real (&v)[K]) {  // K consecutive Reals, 16-byte aligned
  if (sizeof(Real) * K % 16 == 0) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 raw[sizeof(Real) * K / 16];
#pragma unroll
    for (int i = 0; i < (int)(sizeof(Real) * K / 16); ++i) raw[i] = __ldg(q + i);
    memcpy(v, raw, sizeof(Real) * K);
  } else {
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = __ldg(p + i);
  }
}
template <int K>
__device__ __forceinline__ void load_states(const unsigned char* p, uint32_t (&st)[K]) {
  if (K == 4) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
#pragma unroll
    for (int i = 0; i < K; ++i) st[i] = (w >> (8 * i)) & 0xff;
  } else {
    const uint2 w = __ldg(reinterpret_cast<const uint2*>(p));
#pragma unroll
    for (int i = 0; i < K; ++i) st[i] = ((i < 4 ? w.x : w.y) >> (8 * (i & 3))) & 0xff;
  }
}
template <typename Real, int K>
__device__ __forceinline__ void lane_renorm(Real (&a)[K], int& E) {
  int mx = 0;
#pragma unroll
  for (int c = 0; c < K; ++c) mx = max(mx, DN<Real>::bits(a[c]));
  if (mx > 0) {
    const int e = DN<Real>::expo(mx);
    const Real f = DN<Real>::pow2(-e);
#pragma unroll
    for (int c = 0; c < K; ++c) a[c] *= f;
    E += e;
  }
}


template <typename Real, int K, bool XI>
__global__ void __launch_bounds__(kSW * 32) k_fb_sparse(SparseArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t SP = A.SP;
  Real* Ts = reinterpret_cast<Real*>(smem);
  Real* Ph = Ts + SP * SP;
  // expected transition counts: a few tables per CTA when they fit (fp64 shared-memory atomics are CAS loops on
  // sm_100a; with one CTA-wide table they were 75% of the kernel: profiles/r1e_hmm_k_fb_sparse_f64.txt)
  double* xis_all = reinterpret_cast<double*>(smem + (((size_t)SP * SP + SP) * sizeof(Real) + 15) / 16 * 16);
  const uint32_t n_xi = XI ? A.xi_tables : 0;
  double* phc = xis_all + (size_t)n_xi * SP * SP;
  double* xis = xis_all + (size_t)(n_xi > 1 ? ((threadIdx.x >> 5) % n_xi) : 0) * SP * SP;
  // per-lane staging columns of the xi update (private to the lane: no synchronisation)
  Real* xst = reinterpret_cast<Real*>(phc + SP) + (size_t)(threadIdx.x >> 5) * 2 * K * 32 + (threadIdx.x & 31);
  Real* bst = xst + K * 32;
  for (uint32_t i = threadIdx.x; i < SP * SP; i += blockDim.x) Ts[i] = reinterpret_cast<const Real*>(A.T)[i];
  for (uint32_t i = threadIdx.x; i < n_xi * SP * SP; i += blockDim.x) xis_all[i] = 0.;
  for (uint32_t i = threadIdx.x; i < SP; i += blockDim.x) {
    Ph[i] = reinterpret_cast<const Real*>(A.Phi)[i];
    phc[i] = 0.;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t tile = blockIdx.x * kSW + (threadIdx.x >> 5);
  if (tile < A.n_tiles) {
    const SparseTile T = A.tile[tile];
    const uint32_t li = tile * 32 + lane;
    const uint32_t n = A.len[li], seq = A.seq[li];
    const bool live = seq != 0xFFFFFFFFu;
    const uint16_t* __restrict__ sy = A.sym + T.sym_base + lane;
    Real* __restrict__ ag = reinterpret_cast<Real*>(A.alpha) + T.row_base * K * 32 + lane;
    int* __restrict__ ex = A.exps + T.row_base * 32 + lane;
    const Real* __restrict__ Ev = reinterpret_cast<const Real*>(A.Ev);
    // ------------------------------------------------------------ forward
    Real a[K];
    uint32_t st[K];
#pragma unroll
    for (int c = 0; c < K; ++c) {
      a[c] = 0;
      st[c] = SP - 1;  // the padding state: its row and column of T and its final weight are zero
    }
    a[0] = live ? Real(1) : Real(0);
    st[0] = A.start;
    int Ea = 0;
#pragma unroll
    for (int c = 0; c < K; ++c) ag[(size_t)c * 32] = a[c];
    ex[0] = 0;
    uint32_t o_next = T.n_max ? sy[0] : 0;
    for (uint32_t t = 0; t < T.n_max; ++t) {
      const uint32_t o = o_next;
      if (t + 1 < T.n_max) o_next = sy[(size_t)(t + 1) * 32];
      if (t < n) {
        uint32_t es[K];
        Real ev[K], nw[K];
        load_states<K>(A.e_state + (size_t)o * K, es);
        load_row<Real, K>(Ev + (size_t)o * K, ev);
#pragma unroll
        for (int c = 0; c < K; ++c) {
          Real acc = 0;
#pragma unroll
          for (int d = 0; d < K; ++d) acc = fma(a[d], Ts[st[d] * SP + es[c]], acc);
          nw[c] = acc * ev[c];
        }
#pragma unroll
        for (int c = 0; c < K; ++c) {
          a[c] = nw[c];
          st[c] = es[c];
        }
        lane_renorm<Real, K>(a, Ea);
#pragma unroll
        for (int c = 0; c < K; ++c) ag[((size_t)(t + 1) * K + c) * 32] = a[c];
        ex[(size_t)(t + 1) * 32] = Ea;
      }
    }
    // probability: final weights of the active states (or the final state itself)
    Real b[K];
    Real pfin = 0;
#pragma unroll
    for (int c = 0; c < K; ++c) {
      b[c] = A.has_phi ? Ph[st[c]] : (st[c] == A.fin ? Real(1) : Real(0));
      pfin = fma(a[c], b[c], pfin);
    }
    const int EaN = Ea;
    if (live) A.ex_lnp[seq] = (pfin > 0) ? log((double)pfin) + (double)EaN * 0.69314718055994530942 : -CUDART_INF;
    const double cw = (live && pfin > 0) ? A.weight[li] / (double)pfin : 0.;
    // ------------------------------------------------------------ backward + counts
    if (cw > 0) {
      if (A.has_phi) {  // counts of the epsilon final arcs: gamma_n(s) * phi[s] / P
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const double g = (double)a[c] * (double)b[c] * cw;
          if (g > 0) atomicAdd(&phc[st[c]], g);
        }
      }
      int Eb = 0;
      Real a1[K];
#pragma unroll
      for (int c = 0; c < K; ++c) a1[c] = a[c];
      int Ea1 = EaN;
      // emission row of the symbol that leads INTO position t+1 = the active states of position t+1
      uint32_t es[K];
      Real ev[K];
      uint32_t o = n ? sy[(size_t)(n - 1) * 32] : 0;
#pragma unroll
      for (int c = 0; c < K; ++c) es[c] = st[c];
      if (n) load_row<Real, K>(Ev + (size_t)o * K, ev);
      for (uint32_t t = n; t-- > 0;) {
        // active states of position t: emission row of symbol t-1 (or the start state)
        uint32_t ps[K];
        uint32_t op = 0;
        if (t > 0) {
          op = sy[(size_t)(t - 1) * 32];
          load_states<K>(A.e_state + (size_t)op * K, ps);
        } else {
#pragma unroll
          for (int c = 0; c < K; ++c) ps[c] = SP - 1;
          ps[0] = A.start;
        }
        Real a0[K];
#pragma unroll
        for (int c = 0; c < K; ++c) a0[c] = ag[((size_t)t * K + c) * 32];
        const int Ea0 = ex[(size_t)t * 32];
        // gamma of position t+1 -> emission cells of symbol o
        uint32_t code[K];
        {
          const uint4* q = reinterpret_cast<const uint4*>(A.e_code + (size_t)o * K);
          uint4 r0 = __ldg(q);
          code[0] = r0.x; code[1] = r0.y; code[2] = r0.z; code[3] = r0.w;
          if (K == 8) {
            uint4 r1 = __ldg(q + 1);
            code[4 % K] = r1.x; code[5 % K] = r1.y; code[6 % K] = r1.z; code[7 % K] = r1.w;
          }
        }
        const double g1 = cw * pow2d(Ea1 + Eb - EaN);
        Real bt[K];
#pragma unroll
        for (int c = 0; c < K; ++c) {
          const double gamma = (double)a1[c] * (double)b[c] * g1;
          if (gamma > 0 && code[c] != kNone) cmlk::count_add(A.sink, code[c], gamma);
          bt[c] = b[c] * ev[c];
        }
        const double g0 = cw * pow2d(Ea0 + Eb - EaN);
        Real nb[K];
#pragma unroll
        for (int d = 0; d < K; ++d) {
          Real acc = 0;
#pragma unroll
          for (int c = 0; c < K; ++c) acc = fma(Ts[ps[d] * SP + es[c]], bt[c], acc);
          nb[d] = acc;
        }
        if (XI) {
          // xi_t(i,j) = alpha_t[i] T[i][j] E[j][o] beta_{t+1}[j] / P for the K x K active pairs.  The lanes of a warp
          // share the warp's count table, and frequent tags sit at the same list position in most lanes, so the pairs
          // are visited in a lane-rotated order: simultaneous updates of one cell (a CAS retry each) become rare.
          uint64_t pp = 0, ee = 0;
#pragma unroll
          for (int c = 0; c < K; ++c) {
            xst[c * 32] = (Real)((double)a0[c] * g0);
            bst[c * 32] = bt[c];
            pp |= (uint64_t)ps[c] << (8 * c);
            ee |= (uint64_t)es[c] << (8 * c);
          }
#pragma unroll 4
          for (int k = 0; k < K * K; ++k) {
            const int idx = (k + lane) & (K * K - 1);
            const int d = idx / K, c = idx % K;
            const uint32_t cell = (uint32_t)((pp >> (8 * d)) & 0xff) * SP + (uint32_t)((ee >> (8 * c)) & 0xff);
            const double xv = (double)xst[d * 32] * (double)Ts[cell] * (double)bst[c * 32];
            if (xv > 0) atomicAdd(&xis[cell], xv);
          }
        }
#pragma unroll
        for (int c = 0; c < K; ++c) {
          b[c] = nb[c];
          a1[c] = a0[c];
          es[c] = ps[c];
        }
        lane_renorm<Real, K>(b, Eb);
        Ea1 = Ea0;
        o = op;
        if (t > 0) load_row<Real, K>(Ev + (size_t)o * K, ev);
      }
    }
  }
  __syncthreads();
  if constexpr (XI)
    for (uint32_t c = threadIdx.x; c < SP * SP; c += blockDim.x) {
      const uint32_t sl = A.t_slot[c];
      if (sl == kNone) continue;
      double v = 0;
      for (uint32_t k = 0; k < n_xi; ++k) v += xis_all[(size_t)k * SP * SP + c];
      if (v != 0.) atomicAdd(A.sink.counts + sl, v);
    }
  for (uint32_t c = threadIdx.x; c < SP; c += blockDim.x) {
    const uint32_t sl = A.f_slot[c];
    if (sl != kNone && phc[c] != 0.) atomicAdd(A.sink.counts + sl, phc[c]);
  }
}

template <typename Real>
static size_t sparse_smem(uint32_t SP, uint32_t xi_tables, uint32_t K) {
  size_t b = (((size_t)SP * SP + SP) * sizeof(Real) + 15) / 16 * 16;
  b += ((size_t)xi_tables * SP * SP + SP) * sizeof(double);
  if (xi_tables) b += (size_t)kSW * 2 * K * 32 * sizeof(Real);
  return b;
}

template <typename Real>
int launch_sparse(cml_ctx* ctx) {
  DenseState& D = *ctx->dense;
  cudaStream_t s = ctx->stream;
  k_dense_tables<Real><<<cdiv(D.n_cells, 256), 256, 0, s>>>(D.n_cells, D.cell_off.p, D.cell_param.p, D.cell_exists.p,
                                                           ctx->ln_w.p, reinterpret_cast<Real*>(D.tables.p));
  ++ctx->launches;
  CML_CUDA(cudaMemsetAsync(ctx->reduce, 0, ctx->reduce_n * sizeof(double), s));
  if (ctx->n_hot)
    CML_CUDA(cudaMemsetAsync(ctx->hot_counts.p, 0, (size_t)ctx->n_hot * ctx->hot_copies * sizeof(double), s));
  const size_t rs = sizeof(Real);
  const uint32_t SP = D.SP, K = D.K;
  SparseArgs A;
  A.tile = D.stile.p;
  A.n_tiles = D.n_tiles;
  A.sym = D.sym.p;
  A.len = D.lane_len.p;
  A.seq = D.lane_seq.p;
  A.weight = D.lane_weight.p;
  A.S = D.S;
  A.SP = SP;
  A.start = D.start;
  A.fin = D.fin;
  A.T = D.tables.p;
  A.Ev = D.tables.p + (size_t)D.nT * rs;
  A.Phi = D.tables.p + ((size_t)D.nT + (size_t)D.n_sym * K) * rs;
  A.e_state = D.e_state.p;
  A.e_code = D.e_code.p;
  A.t_slot = D.cell_slot.p;
  A.f_slot = D.cell_slot.p + (size_t)D.nT + (size_t)D.n_sym * K;
  A.sink = cmlk::CountSink{ctx->reduce, ctx->hot_counts.p, ctx->n_hot, ctx->hot_copies - 1};
  A.ex_lnp = D.ex_lnp.p;
  A.alpha = D.alpha_g.p;
  A.exps = D.exp_g.p;
  A.has_phi = D.has_phi;
  const bool xi = D.n_t_slots > 0;
  uint32_t want_xi = kSXi;
  if (const char* e = getenv("CML_SPARSE_XI_TABLES")) want_xi = std::max(1, std::min(8, atoi(e)));  // tuning knob
  A.xi_tables = !xi ? 0u : (sparse_smem<Real>(SP, want_xi, K) <= 72 * 1024 ? want_xi : 1u);
  const size_t smem = sparse_smem<Real>(SP, A.xi_tables, K);
  void (*kern)(SparseArgs) =
      K == 4 ? (xi ? k_fb_sparse<Real, 4, true> : k_fb_sparse<Real, 4, false>)
             : (xi ? k_fb_sparse<Real, 8, true> : k_fb_sparse<Real, 8, false>);
  CML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (!D.ev0) {
    CML_CUDA(cudaEventCreate(&D.ev0));
    CML_CUDA(cudaEventCreate(&D.ev1));
  }
  if (!ctx->capturing) CML_CUDA(cudaEventRecord(D.ev0, s));
  if (D.n_tiles) {
    kern<<<cdiv(D.n_tiles, kSW), kSW * 32, smem, s>>>(A);
    ++ctx->launches;
  }
  if (!ctx->capturing) CML_CUDA(cudaEventRecord(D.ev1, s));
  if (ctx->n_hot) {
    cmlk::k_fold_hot<<<cdiv(ctx->n_hot, 256), 256, 0, s>>>(ctx->n_hot, ctx->hot_slot.p, ctx->hot_counts.p, ctx->reduce, ctx->hot_copies);
    ++ctx->launches;
  }
  if (D.n_seq) {
    cmlk::k_reduce_lnp<<<std::min<unsigned>(cdiv(D.n_seq, 256), 4 * ctx->sm_count), 256, 0, s>>>(
        D.ex_lnp.p, D.seq_weight.p, D.n_seq, ctx->reduce + ctx->n_slots);
    ++ctx->launches;
  }
  CML_CUDA(cudaGetLastError());
  return CML_OK;
}

// multiset helpers on sorted vectors
std::vector<uint32_t> ms_intersect(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
  std::vector<uint32_t> r;
  std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r));
  return r;
}
std::vector<uint32_t> ms_minus(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
  std::vector<uint32_t> r;
  std::set_difference(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(r));
  return r;
}

}  // namespace

int cml_dense_estimate_launch(cml_ctx* ctx) {
  if (ctx->dense->tc) return launch_dense_tc(ctx);
  if (ctx->dense->sparse) return ctx->precision == 64 ? launch_sparse<double>(ctx) : launch_sparse<float>(ctx);
  return ctx->precision == 64 ? launch_dense<double>(ctx) : launch_dense<float>(ctx);
}

extern "C" int cml_add_sequences(cml_ctx* ctx, const cml_dense_view* v, const cml_sequence_batch* b) {
  if (!ctx || !v || !b) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model, CML_ERR_STATE, "cml_set_model first");
  CML_REQUIRE(ctx->batches.empty() && !ctx->dense, CML_ERR_STATE,
              "cml_add_sequences needs a context without resident lattices or sequences (cml_set_model resets it)");
  CML_REQUIRE(ctx->space == CML_SPACE_SCALED, CML_ERR_STATE, "dense-state sequences need a CML_SPACE_SCALED context");
  CML_REQUIRE(v->arc_src && v->arc_dst && v->arc_sym && b->seq_off && (b->sym || b->seq_off[b->n_seq] == 0),
              CML_ERR_ARG, "null array");
  CML_REQUIRE(b->n_seq < 0xFFFFFFFFull, CML_ERR_ARG, "too many sequences");
  CML_REQUIRE(v->n_symbols > 0 && v->n_symbols <= 65535, CML_ERR_ARG, "n_symbols must be in 1..65535");
  CML_REQUIRE(v->start < v->n_states && v->final_state < v->n_states, CML_ERR_ARG, "start / final state out of range");
  const uint32_t S = v->n_states, V = v->n_symbols, nA = ctx->n_arcs;
  if (S > 64) {
    ctx->err = "dense-state path needs n_states <= 64";
    return CML_ERR_NOT_DENSE;
  }
  bool has_phi = false;
  for (uint32_t a = 0; a < nA; ++a) {
    CML_REQUIRE(v->arc_src[a] < S && v->arc_dst[a] < S && (v->arc_sym[a] < V || v->arc_sym[a] == CML_DENSE_EPS),
                CML_ERR_ARG, "arc triple out of range");
    if (v->arc_sym[a] == CML_DENSE_EPS) has_phi = true;
  }
  if (has_phi)  // epsilon arcs are final weights: only into the final state, which must be a sink
    for (uint32_t a = 0; a < nA; ++a) {
      const bool eps = v->arc_sym[a] == CML_DENSE_EPS;
      if ((eps && v->arc_dst[a] != v->final_state) || v->arc_src[a] == v->final_state ||
          (!eps && v->arc_dst[a] == v->final_state)) {
        ctx->err = "epsilon arcs are only supported as final weights (into a final state that is reached by nothing else "
                   "and has no outgoing arcs)";
        return CML_ERR_NOT_DENSE;
      }
    }
  const uint64_t n_pos = b->seq_off[b->n_seq];
  for (uint64_t e = 0; e < b->n_seq; ++e)
    CML_REQUIRE(b->seq_off[e] <= b->seq_off[e + 1] && b->seq_off[e + 1] - b->seq_off[e] < 0x7FFFFFFFull, CML_ERR_ARG,
                "seq_off not monotone");
  for (uint64_t i = 0; i < n_pos; ++i) CML_REQUIRE(b->sym[i] < V, CML_ERR_ARG, "sequence symbol out of range");

  // ---- which kernel: emission rows of <= 8 states -> lane-per-sequence sparse kernel; else the warp-per-sequence
  //      dense kernel (S <= 32, no final weights)
  std::vector<std::vector<uint32_t>> allowed(V);  // states that can emit a symbol, ascending
  {
    std::vector<uint64_t> mask(V, 0);
    for (uint32_t a = 0; a < nA; ++a)
      if (v->arc_sym[a] != CML_DENSE_EPS) mask[v->arc_sym[a]] |= 1ull << v->arc_dst[a];
    for (uint32_t o = 0; o < V; ++o)
      for (uint32_t j = 0; j < S; ++j)
        if ((mask[o] >> j) & 1) allowed[o].push_back(j);
  }
  uint32_t kmax = 1;
  for (auto const& al : allowed) kmax = std::max<uint32_t>(kmax, (uint32_t)al.size());
  const bool can_dense = S <= (uint32_t)kDS && !has_phi;
  const bool can_sparse = kmax <= 8 && S <= 63;  // (one spare state index is the zero padding state)
  uint64_t sparse_min = 4096;
  if (const char* e = getenv("CML_SPARSE_MIN_SEQ")) sparse_min = (uint64_t)atoll(e);
  const bool sparse = can_sparse && (!can_dense || (b->n_seq >= sparse_min && 4 * kmax <= S));
  if (!sparse && !can_dense) {
    ctx->err = "no dense-state kernel for this shape (more than 32 states or final weights, and symbols emitted by more "
               "than 8 states)";
    return CML_ERR_NOT_DENSE;
  }
  const uint32_t K = sparse ? (kmax <= 4 ? 4u : 8u) : 0u;
  const uint32_t SP = sparse ? S + 1 : (uint32_t)kDS;  // sparse: row stride of T; index S is the zero padding state

  // ---- factorisation: chain(a) = A(i,j) + B(j,o) as multisets of parameter ids --------------------------------
  auto chain_of = [&](uint32_t a) {
    std::vector<uint32_t> c;
    if (ctx->trivial)
      c.push_back(a);
    else
      c.assign(ctx->h_chain_param.begin() + ctx->h_chain_off[a], ctx->h_chain_param.begin() + ctx->h_chain_off[a + 1]);
    std::sort(c.begin(), c.end());
    return c;
  };
  const uint32_t nT = (SP * SP + 3u) & ~3u;  // (keeps the emission table 16-byte aligned behind T)
  const uint32_t nE = sparse ? V * K : V * (uint32_t)kDS, nF = sparse ? SP : 0, n_cells = nT + nE + nF;
  auto is_eps = [&](uint32_t a) { return v->arc_sym[a] == CML_DENSE_EPS; };
  auto tcell = [&](uint32_t a) { return v->arc_src[a] * SP + v->arc_dst[a]; };
  auto ecell = [&](uint32_t a) {
    const uint32_t o = v->arc_sym[a], j = v->arc_dst[a];
    if (!sparse) return nT + o * (uint32_t)kDS + j;
    const auto& al = allowed[o];
    return nT + o * K + (uint32_t)(std::lower_bound(al.begin(), al.end(), j) - al.begin());
  };
  auto fcell = [&](uint32_t a) { return nT + nE + v->arc_src[a]; };
  std::vector<std::vector<uint32_t>> cell_chain(n_cells);
  std::vector<unsigned char> exists(n_cells, 0);
  std::vector<std::vector<uint32_t>> chains(nA);
  {
    std::map<uint64_t, uint32_t> seen;  // (i, j, o) must be unique
    for (uint32_t a = 0; a < nA; ++a) {
      chains[a] = chain_of(a);
      const uint64_t key = ((uint64_t)tcell(a) << 32) | v->arc_sym[a];
      if (!seen.emplace(key, a).second) {
        ctx->err = "parallel arcs with the same (source, destination, symbol): no dense view";
        return CML_ERR_NOT_DENSE;
      }
    }
  }
  uint32_t n_sym_arcs = 0;
  for (uint32_t a = 0; a < nA; ++a) {
    if (is_eps(a)) {  // a final weight is its own cell
      cell_chain[fcell(a)] = chains[a];
      exists[fcell(a)] = 1;
      continue;
    }
    ++n_sym_arcs;
    const uint32_t c = ecell(a);  // B(j,o) = intersection over the sources i
    cell_chain[c] = exists[c] ? ms_intersect(cell_chain[c], chains[a]) : chains[a];
    exists[c] = 1;
  }
  std::vector<std::vector<uint32_t>> rest(nA);
  for (uint32_t a = 0; a < nA; ++a) {  // A(i,j) = intersection over the symbols o of what B leaves
    if (is_eps(a)) continue;
    rest[a] = ms_minus(chains[a], cell_chain[ecell(a)]);
    const uint32_t c = tcell(a);
    cell_chain[c] = exists[c] ? ms_intersect(cell_chain[c], rest[a]) : rest[a];
    exists[c] = 1;
  }
  for (uint32_t a = 0; a < nA; ++a)
    if (!is_eps(a) && rest[a] != cell_chain[tcell(a)]) {
      ctx->err = "arc chains do not factor into (source,destination) x (destination,symbol) parts";
      return CML_ERR_NOT_DENSE;
    }
  {  // completeness: every (i,j) x (j,o) combination must be an arc, or the dense product would invent arcs
    std::vector<uint32_t> n_in(S, 0), n_o(S, 0);
    for (uint32_t i = 0; i < S; ++i)
      for (uint32_t j = 0; j < S; ++j) n_in[j] += exists[i * SP + j];
    for (uint32_t o = 0; o < V; ++o)
      for (uint32_t j : allowed[o]) ++n_o[j];
    uint64_t tot = 0;
    for (uint32_t j = 0; j < S; ++j) tot += (uint64_t)n_in[j] * n_o[j];
    if (tot != n_sym_arcs) {
      ctx->err = "the arc table is not the full product of its transition and emission supports";
      return CML_ERR_NOT_DENSE;
    }
  }

  // ---- count slots := trainable cells; priors of an arc are shared out to the cells of its parameters -----------
  std::vector<uint32_t> cell_off(n_cells + 1, 0), cell_param, cell_slot(n_cells, kNone), slot_off{0}, slot_param;
  std::vector<double> slot_prior;
  uint32_t n_slots = 0, n_t_slots = 0, n_e_slots = 0;
  for (uint32_t c = 0; c < n_cells; ++c) {
    cell_param.insert(cell_param.end(), cell_chain[c].begin(), cell_chain[c].end());
    cell_off[c + 1] = (uint32_t)cell_param.size();
    if (!exists[c]) continue;
    std::vector<uint32_t> un;
    for (uint32_t p : cell_chain[c])
      if (ctx->h_param_tie[p] != CML_LOCKED_GROUP) un.push_back(p);
    if (un.empty()) continue;
    cell_slot[c] = n_slots++;
    (c < nT ? n_t_slots : n_e_slots)++;
    slot_param.insert(slot_param.end(), un.begin(), un.end());
    slot_off.push_back((uint32_t)slot_param.size());
    slot_prior.push_back(0.);
  }
  const bool have_prior = !ctx->h_arc_prior.empty();
  if (have_prior)
    for (uint32_t a = 0; a < nA; ++a) {
      if (is_eps(a)) {
        if (cell_slot[fcell(a)] != kNone) slot_prior[cell_slot[fcell(a)]] += ctx->h_arc_prior[a];
        continue;
      }
      if (cell_slot[tcell(a)] != kNone) slot_prior[cell_slot[tcell(a)]] += ctx->h_arc_prior[a];
      if (cell_slot[ecell(a)] != kNone) slot_prior[cell_slot[ecell(a)]] += ctx->h_arc_prior[a];
    }
  if (n_slots == 0) {
    n_slots = 1;
    slot_off.push_back(0);
    slot_prior.push_back(0.);
  }

  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  std::unique_ptr<DenseState> D(new DenseState());
  D->S = S;
  D->SP = SP;
  D->nT = nT;
  D->K = K;
  D->sparse = sparse;
  D->has_phi = has_phi ? 1 : 0;
  D->n_sym = V;
  D->start = v->start;
  D->fin = v->final_state;
  D->n_seq = b->n_seq;
  D->n_pos = n_pos;
  D->n_cells = n_cells;
  D->n_t_slots = n_t_slots;
  D->n_e_slots = n_e_slots;
  const size_t rs = ctx->precision / 8;
  std::vector<double> wts(std::max<uint64_t>(1, b->n_seq), 1.);
  if (b->seq_weight) std::copy(b->seq_weight, b->seq_weight + b->n_seq, wts.begin());
  CML_CUDA(D->cell_off.upload(cell_off.data(), cell_off.size(), s));
  CML_CUDA(D->cell_param.upload(cell_param.data(), cell_param.size(), s));
  CML_CUDA(D->cell_slot.upload(cell_slot.data(), cell_slot.size(), s));
  CML_CUDA(D->cell_exists.upload(exists.data(), exists.size(), s));
  CML_CUDA(D->tables.alloc((size_t)n_cells * rs));
  CML_CUDA(D->seq_weight.upload(wts.data(), wts.size(), s));
  CML_CUDA(D->ex_lnp.alloc(std::max<uint64_t>(1, b->n_seq)));
  std::vector<uint32_t> hot, e_code;
  if (!sparse) {
    std::vector<uint16_t> sym16(std::max<uint64_t>(1, n_pos));
    for (uint64_t i = 0; i < n_pos; ++i) sym16[i] = (uint16_t)b->sym[i];
    CML_CUDA(D->seq_off.upload(b->seq_off, b->n_seq + 1, s));
    CML_CUDA(D->sym.upload(sym16.data(), sym16.size(), s));
    CML_CUDA(D->alpha_g.alloc((size_t)(n_pos + b->n_seq) * kDS * rs));
    CML_CUDA(D->exp_g.alloc((size_t)(n_pos + b->n_seq)));
    // tensor-core sweeps (3xTF32): fp32, locked transitions, small alphabet, many sequences
    uint64_t tc_min = 16384;
    if (const char* ev = getenv("CML_DENSE_TC")) tc_min = atoi(ev) > 0 ? 0 : ~0ull;
    if (ctx->precision == 32 && b->n_seq >= tc_min && b->n_seq > 0 &&
        (size_t)kDW * V * 32 * sizeof(double) <= 100 * 1024) {
      std::vector<uint32_t> order(b->n_seq);
      for (uint32_t e = 0; e < b->n_seq; ++e) order[e] = e;
      std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        return b->seq_off[x + 1] - b->seq_off[x] > b->seq_off[y + 1] - b->seq_off[y];
      });
      D->tc_groups = (uint32_t)((b->n_seq + 15) / 16);
      order.resize((size_t)D->tc_groups * 16, 0xFFFFFFFFu);
      CML_CUDA(D->tc_order.upload(order.data(), order.size(), s));
      CML_CUDA(D->beta_g.alloc((size_t)(n_pos + b->n_seq) * kDS * rs));
      CML_CUDA(D->bexp_g.alloc((size_t)(n_pos + b->n_seq)));
      CML_CUDA(D->tc_afin.alloc(b->n_seq));
      CML_CUDA(D->tc_ean.alloc(b->n_seq));
      D->tc = true;
    }
  } else {
    // tiles of 32 sequences of similar length, symbols transposed
    std::vector<uint32_t> order(b->n_seq);
    for (uint32_t e = 0; e < b->n_seq; ++e) order[e] = e;
    auto len_of = [&](uint32_t e) { return (uint32_t)(b->seq_off[e + 1] - b->seq_off[e]); };
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return len_of(x) > len_of(y); });
    const uint32_t n_tiles = (uint32_t)((b->n_seq + 31) / 32);
    std::vector<SparseTile> tiles(n_tiles);
    std::vector<uint32_t> l_len((size_t)n_tiles * 32, 0), l_seq((size_t)n_tiles * 32, 0xFFFFFFFFu);
    std::vector<double> l_w((size_t)n_tiles * 32, 0.);
    uint64_t sym_rows = 0, a_rows = 0;
    for (uint32_t t = 0; t < n_tiles; ++t) {
      const uint32_t nmax = len_of(order[(size_t)t * 32]);
      tiles[t].sym_base = sym_rows * 32;
      tiles[t].row_base = a_rows;
      tiles[t].n_max = nmax;
      tiles[t].pad = 0;
      sym_rows += nmax;
      a_rows += nmax + 1;
    }
    std::vector<uint16_t> symt(std::max<uint64_t>(1, sym_rows * 32), 0);
    for (uint32_t t = 0; t < n_tiles; ++t)
      for (uint32_t l = 0; l < 32 && (size_t)t * 32 + l < b->n_seq; ++l) {
        const uint32_t e = order[(size_t)t * 32 + l], n = len_of(e);
        l_len[(size_t)t * 32 + l] = n;
        l_seq[(size_t)t * 32 + l] = e;
        l_w[(size_t)t * 32 + l] = wts[e];
        for (uint32_t k = 0; k < n; ++k) symt[tiles[t].sym_base + (size_t)k * 32 + l] = (uint16_t)b->sym[b->seq_off[e] + k];
      }
    D->n_tiles = n_tiles;
    CML_CUDA(D->stile.upload(tiles.data(), tiles.size(), s));
    CML_CUDA(D->sym.upload(symt.data(), symt.size(), s));
    CML_CUDA(D->lane_len.upload(l_len.data(), l_len.size(), s));
    CML_CUDA(D->lane_seq.upload(l_seq.data(), l_seq.size(), s));
    CML_CUDA(D->lane_weight.upload(l_w.data(), l_w.size(), s));
    CML_CUDA(D->alpha_g.alloc(std::max<uint64_t>(1, a_rows) * K * 32 * rs));
    CML_CUDA(D->exp_g.alloc(std::max<uint64_t>(1, a_rows) * 32));
    // emission rows: state ids, count-slot codes (hot slots replicated: CountSink)
    std::vector<unsigned char> e_state((size_t)V * K, (unsigned char)(SP - 1));  // padding -> the zero state
    for (uint32_t o = 0; o < V; ++o)
      for (size_t c = 0; c < allowed[o].size(); ++c) e_state[(size_t)o * K + c] = (unsigned char)allowed[o][c];
    std::vector<uint64_t> sym_occ(V, 0), slot_occ(n_slots, 0);
    for (uint64_t i = 0; i < n_pos; ++i) ++sym_occ[b->sym[i]];
    for (uint32_t o = 0; o < V; ++o)
      for (uint32_t c = 0; c < K; ++c)
        if (cell_slot[nT + o * K + c] != kNone) slot_occ[cell_slot[nT + o * K + c]] += sym_occ[o];
    std::vector<uint32_t> hot_index(n_slots, kNone);
    for (uint32_t sl = 0; sl < n_slots; ++sl)
      if (slot_occ[sl] >= 4096) {
        hot_index[sl] = (uint32_t)hot.size();
        hot.push_back(sl);
      }
    e_code.assign((size_t)V * K, kNone);
    for (size_t c = 0; c < e_code.size(); ++c) {
      const uint32_t sl = cell_slot[nT + c];
      if (sl != kNone) e_code[c] = hot_index[sl] != kNone ? (cmlk::kSlotHot | hot_index[sl]) : sl;
    }
    CML_CUDA(D->e_state.upload(e_state.data(), e_state.size(), s));
    CML_CUDA(D->e_code.upload(e_code.data(), e_code.size(), s));
  }
  // the M-step's view of the count slots
  CML_CUDA(ctx->slot_off.upload(slot_off.data(), slot_off.size(), s));
  CML_CUDA(ctx->slot_param.upload(slot_param.data(), slot_param.size(), s));
  ctx->have_prior = have_prior;
  if (have_prior) CML_CUDA(ctx->slot_prior.upload(slot_prior.data(), slot_prior.size(), s));
  CML_CUDA(ctx->reduce_own.alloc((size_t)n_slots + 3));
  ctx->reduce = ctx->reduce_own.p;
  ctx->reduce_n = (uint64_t)n_slots + 3;
  CML_CUDA(cudaMemsetAsync(ctx->reduce, 0, ctx->reduce_n * sizeof(double), s));
  ctx->n_hot = (uint32_t)hot.size();
  CML_CUDA(ctx->hot_slot.upload(hot.data(), hot.size(), s));
  CML_CUDA(ctx->hot_counts.alloc(std::max<size_t>(1, (size_t)ctx->n_hot * ctx->hot_copies)));
  CML_CUDA(cudaStreamSynchronize(s));
  ctx->n_slots = n_slots;
  ctx->slots_are_arcs = false;
  ctx->hot_dirty = false;
  ctx->slot_occ.assign(n_slots, 0);
  ctx->dense = std::move(D);
  return CML_OK;
}

extern "C" int cml_dense_stats(cml_ctx* ctx, uint64_t* n_seq, uint64_t* n_positions, uint32_t* n_t_slots,
                               uint32_t* n_e_slots) {
  if (!ctx) return CML_ERR_ARG;
  const DenseState* D = ctx->dense.get();
  if (n_seq) *n_seq = D ? D->n_seq : 0;
  if (n_positions) *n_positions = D ? D->n_pos : 0;
  if (n_t_slots) *n_t_slots = D ? D->n_t_slots : 0;
  if (n_e_slots) *n_e_slots = D ? D->n_e_slots : 0;
  return CML_OK;
}

extern "C" int cml_dense_kernel(cml_ctx* ctx, int* sparse, uint32_t* k, uint32_t* n_states) {
  if (!ctx) return CML_ERR_ARG;
  const DenseState* D = ctx->dense.get();
  if (sparse) *sparse = D ? (D->sparse ? 1 : (D->tc ? 2 : 0)) : -1;
  if (k) *k = D ? D->K : 0;
  if (n_states) *n_states = D ? D->S : 0;
  return CML_OK;
}
