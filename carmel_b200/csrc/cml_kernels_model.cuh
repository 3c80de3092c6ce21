// cml_kernels_model.cuh -- parameter-table kernels: cascade chain products (K6), chain count
// distribution (K4) and the normalisation M-step (K5).  All fp64, log domain, so the M-step has the
// reference's dynamic range (weights like e^-900 survive).
//
// Reference semantics restated here:
//   cascade_parameters::calculate_chain_weights / update   carmel/src/cascade.h:426-433,466-479
//   for_arcs::prep_new_weights                               carmel/src/train.cc:134-153
//   cascade_parameters::distribute_counts                    carmel/src/cascade.h:286-325
//   WFST::normalize (locked '!' and tied '!N' arcs)          carmel/src/fst.cc:86-244
//   for_arcs::overrelax / max_change                         carmel/src/train.cc:157-182
#pragma once
#include "cml_common.cuh"

namespace cmlk {

__device__ __forceinline__ double lse2(double a, double b) {  // ln(e^a + e^b), -inf safe
  if (!(a > -CUDART_INF)) return b;
  if (!(b > -CUDART_INF)) return a;
  const double m = fmax(a, b);
  return m + log1p(exp(-fabs(a - b)));
}
// ln(e^a - e^b) clamped at zero like logweight operator- (weight.h:803-830)
__device__ __forceinline__ double lsub(double a, double b) {
  if (!(b > -CUDART_INF)) return a;
  const double rd = b - a;
  if (rd >= 0) return -CUDART_INF;
  if (rd < -36.) return a;
  return a + log1p(-exp(rd));
}
// warp-wide max-shifted log-sum-exp of one value per lane
__device__ __forceinline__ double warp_lse(double v) {
  double m = v;
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (!(m > -CUDART_INF)) return -CUDART_INF;
  double s = (v > -CUDART_INF) ? exp(v - m) : 0.;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return m + log(s);
}
__device__ __forceinline__ double warp_max(double m) {
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  return m;
}

// K6: arc weight = product of its chain's parameters.  out_real = ln w (LOG) or w (SCALED) as Real;
// out_ws pairs the weight with the arc's count slot (one gather in the backward sweep).  Entry n_arcs
// is the zero-weight padding arc of the ELL layout.
template <typename Real>
struct WSOut;
template <>
struct __align__(16) WSOut<double> {
  double w;
  uint32_t slot, pad;
};
template <>
struct __align__(8) WSOut<float> {
  float w;
  uint32_t slot;
};
template <typename Real, bool SCALED, typename WSReal>
static __global__ void k_arc_weights(uint32_t n_arcs, const uint32_t* __restrict__ chain_off,
                              const uint32_t* __restrict__ chain_param, const double* __restrict__ ln_w,
                              const uint32_t* __restrict__ slot_code, const uint32_t* __restrict__ perm,
                              double* __restrict__ arc_lnw, Real* __restrict__ out_real, WSReal* __restrict__ out_ws) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a > n_arcs) return;
  double s;
  uint32_t slot = 0xFFFFFFFFu;
  if (a == n_arcs) {
    s = -CUDART_INF;
  } else {
    if (chain_off) {
      s = 0;
      for (uint32_t k = chain_off[a], e = chain_off[a + 1]; k < e; ++k) s += ln_w[chain_param[k]];
    } else
      s = ln_w[a];
    arc_lnw[a] = s;
  }
  const uint32_t ia = perm[a];  // internal (locality-ordered) arc id; slot codes are stored in that order
  slot = (a == n_arcs) ? 0xFFFFFFFFu : slot_code[ia];
  const Real v = SCALED ? (Real)exp(s) : (Real)s;
  out_real[ia] = v;
  out_ws[ia].w = v;
  out_ws[ia].slot = slot;
}

// K4 + prep_new_weights: add (count + prior) of every count slot to each parameter of the slot's
// (unlocked) chain.  Trivial cascade: slot == arc == parameter, acc[p] = counts[p] + prior[p].
static __global__ void k_param_acc(uint32_t n_slots, const uint32_t* __restrict__ slot_off,
                            const uint32_t* __restrict__ slot_param, const double* __restrict__ counts,
                            const double* __restrict__ prior, const uint32_t* __restrict__ param_tie,
                            double* __restrict__ acc) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_slots) return;
  const double v = counts[a] + (prior ? prior[a] : 0.);
  if (slot_off) {
    for (uint32_t k = slot_off[a], e = slot_off[a + 1]; k < e; ++k) {
      const uint32_t p = slot_param[k];
      if (param_tie[p] != CML_LOCKED_GROUP && v != 0.) atomicAdd(&acc[p], v);
    }
  } else
    acc[a] = v;
}

// unnormalised new weights u[p]: ln(acc) for trainable parameters, the old weight for locked
// parameters and for members of NONE-normalised transducers (cascade.h:339-350 save/load_none).
static __global__ void k_unnorm(uint32_t n_params, const double* __restrict__ acc, const double* __restrict__ ln_w,
                         const uint32_t* __restrict__ param_tie, const uint32_t* __restrict__ param_group,
                         double* __restrict__ u, double* __restrict__ old) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_params) return;
  const double w = ln_w[p];
  old[p] = w;
  const bool keep = (param_tie[p] == CML_LOCKED_GROUP) || (param_group[p] == CML_NO_GROUP);
  u[p] = keep ? w : (acc[p] > 0 ? log(acc[p]) : -CUDART_INF);
}

// log-sum-exp / max over the TPG threads that share one normalisation group (TPG = 32: a warp; TPG = 256: the block)
template <int TPG>
__device__ __forceinline__ double group_lse(double v, double* sh) {
  v = warp_lse(v);
  if (TPG == 32) return v;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // (sh may still be read from a previous call)
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = -CUDART_INF;
#pragma unroll
  for (int k = 0; k < TPG / 32; ++k) r = lse2(r, sh[k]);
  return r;
}

// normalize pass 1 (fst.cc:115-133): TPG threads per normalisation group (a warp, or a whole block when some group
// has thousands of members: a lexicon's conditional group of one frequent tag).  Adds the group's additive
// prior to EVERY member (the reference adds it to locked arcs' stored weight too), then sums.
template <int TPG>
static __global__ void __launch_bounds__(256) k_norm_sums(uint32_t n_groups, const uint32_t* __restrict__ group_off,
                            const uint32_t* __restrict__ group_members, const double* __restrict__ group_add,
                            const uint32_t* __restrict__ param_tie, double* __restrict__ u,
                            double* __restrict__ gsum, double* __restrict__ glocked) {
  __shared__ double sh[8];
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) / TPG;
  const int t = threadIdx.x % TPG;
  const bool live = g < n_groups;  // (TPG == 256: block-uniform)
  if (TPG == 32 && !live) return;
  const double addc = (live && group_add) ? group_add[g] : -CUDART_INF;
  double s = -CUDART_INF, l = -CUDART_INF;
  if (live)
    for (uint32_t k = group_off[g] + t, e = group_off[g + 1]; k < e; k += TPG) {
      const uint32_t p = group_members[k];
      const double v = lse2(u[p], addc);
      u[p] = v;
      if (param_tie[p] == CML_LOCKED_GROUP)
        l = lse2(l, v);
      else
        s = lse2(s, v);
    }
  s = group_lse<TPG>(s, sh);
  l = group_lse<TPG>(l, sh);
  if (live && t == 0) {
    gsum[g] = s;
    glocked[g] = l;
  }
}

// tie-group totals (fst.cc:134-152): one warp per tie id.
static __global__ void k_tie_totals(uint32_t n_ties, const uint32_t* __restrict__ tie_off,
                             const uint32_t* __restrict__ tie_members, const uint32_t* __restrict__ param_group,
                             const double* __restrict__ u, const double* __restrict__ gsum,
                             const double* __restrict__ glocked, double* __restrict__ tie_arc_total,
                             double* __restrict__ tie_state_total, double* __restrict__ tie_max_locked) {
  const uint32_t t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= n_ties) return;
  double at = -CUDART_INF, st = -CUDART_INF, ml = -CUDART_INF;
  for (uint32_t k = tie_off[t] + lane, e = tie_off[t + 1]; k < e; k += 32) {
    const uint32_t p = tie_members[k];
    const uint32_t g = param_group[p];
    if (g == CML_NO_GROUP) continue;
    at = lse2(at, u[p]);
    st = lse2(st, gsum[g]);
    ml = fmax(ml, glocked[g]);
  }
  at = warp_lse(at);
  st = warp_lse(st);
  ml = warp_max(ml);
  if (lane == 0) {
    tie_arc_total[t] = at;
    tie_state_total[t] = st;
    tie_max_locked[t] = ml;
  }
}

// normalize pass 2 (fst.cc:160-229): TPG threads per group; writes the new ln weights.
template <int TPG>
static __global__ void __launch_bounds__(256) k_norm_assign(uint32_t n_groups, const uint32_t* __restrict__ group_off,
                              const uint32_t* __restrict__ group_members, const uint32_t* __restrict__ param_tie,
                              const double* __restrict__ u, const double* __restrict__ tie_arc_total,
                              const double* __restrict__ tie_state_total, const double* __restrict__ tie_max_locked,
                              double* __restrict__ ln_w) {
  __shared__ double sh[8];
  const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) / TPG;
  const int t = threadIdx.x % TPG;
  const bool live = g < n_groups;
  if (TPG == 32 && !live) return;
  double normal_sum = -CUDART_INF, reserved = -CUDART_INF;
  const uint32_t k0 = live ? group_off[g] : 0, k1 = live ? group_off[g + 1] : 0;
  for (uint32_t k = k0 + t; k < k1; k += TPG) {
    const uint32_t p = group_members[k];
    const uint32_t tie = param_tie[p];
    if (tie == CML_NO_GROUP) {
      normal_sum = lse2(normal_sum, u[p]);
    } else if (tie == CML_LOCKED_GROUP) {
      reserved = lse2(reserved, u[p]);
      ln_w[p] = u[p];
    } else {
      const uint32_t ti = tie - 1;
      double groupNorm = tie_state_total[ti];
      const double gmax = tie_max_locked[ti];
      double nw = -CUDART_INF;
      if (!(gmax > 0.)) {
        if (gmax > -CUDART_INF) groupNorm -= lsub(0., gmax);
        const double groupTotal = tie_arc_total[ti];
        if (groupTotal > -CUDART_INF) {
          nw = groupTotal - groupNorm;
          reserved = lse2(reserved, nw);
        }
      }
      ln_w[p] = nw;
    }
  }
  normal_sum = group_lse<TPG>(normal_sum, sh);
  reserved = group_lse<TPG>(reserved, sh);
  const double fraction_remain = lsub(0., reserved);
  const bool give = (fraction_remain > -CUDART_INF) && (normal_sum > -CUDART_INF);
  for (uint32_t k = k0 + t; k < k1; k += TPG) {
    const uint32_t p = group_members[k];
    if (param_tie[p] == CML_NO_GROUP) ln_w[p] = give ? fraction_remain + u[p] - normal_sum : -CUDART_INF;
  }
}

// parameters outside every normalisation group keep their unnormalised value (NONE method)
static __global__ void k_copy_ungrouped(uint32_t n_params, const uint32_t* __restrict__ param_group,
                                 const double* __restrict__ u, double* __restrict__ ln_w) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_params && param_group[p] == CML_NO_GROUP) ln_w[p] = u[p];
}

// over-relaxation (train.cc:157-171): w <- old * (em/old)^rate for unlocked arcs with old > 0
static __global__ void k_overrelax(uint32_t n_params, double rate, const uint32_t* __restrict__ param_tie,
                            const double* __restrict__ old, const double* __restrict__ ln_w, double* __restrict__ u) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_params) return;
  double w = ln_w[p];
  if (param_tie[p] != CML_LOCKED_GROUP && old[p] > -CUDART_INF) w = old[p] + (w - old[p]) * rate;
  u[p] = w;
}

// max |w_new - w_old| over unlocked parameters (train.cc:173-182), linear domain; result as the bit
// pattern of a non-negative double (monotone under unsigned compare).
static __global__ void k_max_change(uint32_t n_params, const uint32_t* __restrict__ param_tie,
                             const double* __restrict__ old, const double* __restrict__ ln_w,
                             unsigned long long* __restrict__ out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0;
  if (p < n_params && param_tie[p] != CML_LOCKED_GROUP) {
    const double a = ln_w[p], b = old[p];
    const double hi = fmax(a, b), lo = fmin(a, b);
    const double l = lsub(hi, lo);
    d = (l > -CUDART_INF) ? exp(l) : 0.;
  }
  d = warp_max(d);
  if ((threadIdx.x & 31) == 0 && d > 0) atomicMax(out, (unsigned long long)__double_as_longlong(d));
}


// The whole M-step of a SMALL model in one launch (one CTA): K4 + prep_new_weights + normalize + max_change, the
// phases of the kernels above separated by block barriers.  Used when there are no tie groups and the tables fit
// one CTA's patience (a few 10^4 parameters): the EM iteration of such models is launch bound (the cipher's
// dense-state E-step takes 57 us, six M-step launches cost 35 us more).
struct MstepArgs {
  uint32_t n_slots, n_params, n_groups;
  const uint32_t* slot_off;     // NULL: slot == parameter
  const uint32_t* slot_param;
  const double* counts;
  const double* prior;          // per slot or NULL
  const uint32_t* param_tie;
  const uint32_t* param_group;
  const uint32_t* group_off;
  const uint32_t* group_members;
  const double* group_add;      // or NULL
  double* acc;
  double* u;
  double* old;
  double* ln_w;
  double* gsum;
  double* glocked;
  unsigned long long* maxchg;
};
static __global__ void __launch_bounds__(1024) k_mstep_fused(MstepArgs A) {
  __shared__ unsigned long long smax;
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  const int lane = tid & 31;
  const uint32_t warp = tid >> 5, nw = nt >> 5;
  if (tid == 0) smax = 0ull;
  if (A.slot_off)
    for (uint32_t p = tid; p < A.n_params; p += nt) A.acc[p] = 0.;
  __syncthreads();
  for (uint32_t a = tid; a < A.n_slots; a += nt) {  // k_param_acc
    const double v = A.counts[a] + (A.prior ? A.prior[a] : 0.);
    if (A.slot_off) {
      for (uint32_t k = A.slot_off[a], e = A.slot_off[a + 1]; k < e; ++k) {
        const uint32_t p = A.slot_param[k];
        if (A.param_tie[p] != CML_LOCKED_GROUP && v != 0.) atomicAdd(&A.acc[p], v);
      }
    } else
      A.acc[a] = v;
  }
  __syncthreads();
  for (uint32_t p = tid; p < A.n_params; p += nt) {  // k_unnorm
    const double w = A.ln_w[p];
    A.old[p] = w;
    const bool keep = (A.param_tie[p] == CML_LOCKED_GROUP) || (A.param_group[p] == CML_NO_GROUP);
    A.u[p] = keep ? w : (A.acc[p] > 0 ? log(A.acc[p]) : -CUDART_INF);
  }
  __syncthreads();
  for (uint32_t g = warp; g < A.n_groups; g += nw) {  // k_norm_sums + k_norm_assign (no ties), a warp per group
    const double addc = A.group_add ? A.group_add[g] : -CUDART_INF;
    const uint32_t k0 = A.group_off[g], k1 = A.group_off[g + 1];
    double normal_sum = -CUDART_INF, reserved = -CUDART_INF;
    for (uint32_t k = k0 + lane; k < k1; k += 32) {
      const uint32_t p = A.group_members[k];
      const double v = lse2(A.u[p], addc);
      A.u[p] = v;
      if (A.param_tie[p] == CML_LOCKED_GROUP) {
        reserved = lse2(reserved, v);
        A.ln_w[p] = v;
      } else
        normal_sum = lse2(normal_sum, v);
    }
    normal_sum = warp_lse(normal_sum);
    reserved = warp_lse(reserved);
    if (lane == 0) {
      A.gsum[g] = normal_sum;
      A.glocked[g] = reserved;
    }
    const double fraction_remain = lsub(0., reserved);
    const bool give = (fraction_remain > -CUDART_INF) && (normal_sum > -CUDART_INF);
    __syncwarp();
    for (uint32_t k = k0 + lane; k < k1; k += 32) {
      const uint32_t p = A.group_members[k];
      if (A.param_tie[p] != CML_LOCKED_GROUP) A.ln_w[p] = give ? fraction_remain + A.u[p] - normal_sum : -CUDART_INF;
    }
  }
  __syncthreads();
  double d = 0;
  for (uint32_t p = tid; p < A.n_params; p += nt) {  // k_copy_ungrouped + k_max_change
    if (A.param_group[p] == CML_NO_GROUP) A.ln_w[p] = A.u[p];
    if (A.param_tie[p] != CML_LOCKED_GROUP) {
      const double a = A.ln_w[p], b = A.old[p];
      const double hi = fmax(a, b), lo = fmin(a, b);
      const double l = lsub(hi, lo);
      d = fmax(d, (l > -CUDART_INF) ? exp(l) : 0.);
    }
  }
  d = warp_max(d);
  if (lane == 0 && d > 0) atomicMax(&smax, (unsigned long long)__double_as_longlong(d));
  __syncthreads();
  if (tid == 0) *A.maxchg = smax;
}

}  // namespace cmlk
