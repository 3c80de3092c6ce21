// cml_gibbs.cu -- collapsed Gibbs sampling (carmel --crp) over the resident derivation lattices (K8).
//
// Reference semantics (restated, not copied):
//   gibbs_base::iteration            graehl/shared/gibbs.hpp:836-877   (remove block, resample, add block)
//   gibbs_param proposal_prob/addc   graehl/shared/gibbs.hpp:154-157,203-211
//   carmel_gibbs::operator()/choose  carmel/src/gibbs.cc:348-371       (arc prob = product of chain params)
//   derivations::random_path         carmel/src/derivations.h:345-375  (backward filter, forward sample)
//   pfor::global_normalize, choose_p derivations.h:318-337 ; graehl/shared/random.ipp:111-127
// One uniform per visited non-final lattice state, in path order.  The uniforms are counter based
// (seed, sweep, block, draw) so that any block can be sampled on any GPU thread and still reproduce the
// same derivation as the sequential CPU restatement.
//
// Modes
//   sequential (exact): ONE CTA walks the blocks in corpus order; every block sees the counts that include
//     the blocks sampled before it in this sweep, exactly like the reference (latency bound by design).
//   batched: one CTA per block, all blocks sampled against the counts of the previous sweep (the
//     reference's --include-self flavour of staleness), count deltas applied afterwards with fp64 REDs.
#include <algorithm>
#include <cmath>

#include "cml_ctx.cuh"
#include "cml_kernels_model.cuh"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double gibbs_uniform(uint64_t seed, uint32_t sweep, uint32_t block, uint32_t draw) {
  uint64_t h = mix64(seed ^ mix64(((uint64_t)sweep << 32) | block));
  h = mix64(h + draw);
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

struct GibbsArgs {
  const CmlExDesc* desc;
  uint32_t n_ex;
  const uint32_t* lvl_off;
  const uint32_t* out_off;
  const uint2* out_arc;        // {dst layered index, internal arc id}
  const uint32_t* arc_orig;    // internal arc id -> arc-table id
  const uint32_t* chain_off;   // arc-table id -> parameters (NULL: arc i is parameter i)
  const uint32_t* chain_param;
  const uint32_t* param_norm;  // CML_NO_GROUP: fixed probability = prior
  const double* prior;
  double* count;
  double* normsum;
  const double* arc_lnw;       // per internal arc id (batched mode / initial sample from EM weights), may be NULL
  double* beta;                // scratch: one double per lattice state
  const uint64_t* beta_base;
  const uint64_t* sample_base;
  const uint32_t* old_sample;  // arc-table ids of the previous sample
  const uint32_t* old_len;
  uint32_t* new_sample;
  uint32_t* new_len;
  double power;
  uint64_t seed;
  uint32_t sweep;
  int sequential;
};

__device__ __forceinline__ double param_prob(const GibbsArgs& A, uint32_t p) {
  const uint32_t g = A.param_norm[p];
  return g == CML_NO_GROUP ? A.prior[p] : A.count[p] / A.normsum[g];
}
__device__ double arc_lnprob(const GibbsArgs& A, uint32_t internal_id) {
  if (A.arc_lnw) return A.arc_lnw[internal_id];
  const uint32_t a = A.arc_orig[internal_id];
  if (!A.chain_off) return log(param_prob(A, a));
  double s = 0;
  for (uint32_t k = A.chain_off[a], e = A.chain_off[a + 1]; k < e; ++k) s += log(param_prob(A, A.chain_param[k]));
  return s;
}
__device__ void add_sample_counts(const GibbsArgs& A, const uint32_t* arcs, uint32_t n, double d, bool atomic) {
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t a = arcs[i];
    const uint32_t k0 = A.chain_off ? A.chain_off[a] : a, k1 = A.chain_off ? A.chain_off[a + 1] : a + 1;
    for (uint32_t k = k0; k < k1; ++k) {
      const uint32_t p = A.chain_off ? A.chain_param[k] : k;
      const uint32_t g = A.param_norm[p];
      if (g == CML_NO_GROUP) continue;
      if (atomic) {
        atomicAdd(&A.count[p], d);
        atomicAdd(&A.normsum[g], d);
      } else {
        A.count[p] += d;
        A.normsum[g] += d;
      }
    }
  }
}

// one block (example): backward filter over the layered CSR (log space), then forward sample by thread 0
__device__ void sample_block(const GibbsArgs& A, uint32_t e) {
  const CmlExDesc d = A.desc[e];
  const uint32_t* lvl = A.lvl_off + d.lvl_base;
  const uint32_t* ooff = A.out_off + d.row_base;
  const uint2* oarc = A.out_arc + d.arc_base;
  double* be = A.beta + A.beta_base[e];
  const double NI = -CUDART_INF;
  const int nl = (int)d.n_levels;
  for (int L = nl - 1; L >= 0; --L) {
    for (uint32_t s = lvl[L] + threadIdx.x; s < lvl[L + 1]; s += blockDim.x) {
      double m = NI, acc = 0;
      if (s == d.fin) {
        m = 0;
        acc = 1;
      }
      for (uint32_t k = ooff[s]; k < ooff[s + 1]; ++k) {
        const uint2 r = oarc[k];
        const double v = arc_lnprob(A, r.y) + be[r.x];
        if (v > m) {
          acc = acc * exp(m - v) + 1.;
          m = v;
        } else if (v > NI)
          acc += exp(v - m);
      }
      be[s] = (m > NI) ? m + log(acc) : NI;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    uint32_t* out = A.new_sample + A.sample_base[e];
    uint32_t n = 0, s = 0, draw = 0;  // the start state has layered index 0
    while (s != d.fin) {
      const uint32_t k0 = ooff[s], k1 = ooff[s + 1];
      if (k0 == k1) break;  // cannot happen on a pruned lattice
      // global_normalize: nw = (w*beta)^power ; p = nw / sum
      double m = NI;
      for (uint32_t k = k0; k < k1; ++k) {
        const uint2 r = oarc[k];
        m = fmax(m, A.power * (arc_lnprob(A, r.y) + be[r.x]));
      }
      double sum = 0;
      for (uint32_t k = k0; k < k1; ++k) {
        const uint2 r = oarc[k];
        const double v = A.power * (arc_lnprob(A, r.y) + be[r.x]);
        if (v > NI) sum += exp(v - m);
      }
      // choose_p: psum = sum of normalised p (~1); choice = psum * u; first arc where the running
      // remainder goes negative (or the last arc)
      double psum = 0;
      for (uint32_t k = k0; k < k1; ++k) {
        const uint2 r = oarc[k];
        const double v = A.power * (arc_lnprob(A, r.y) + be[r.x]);
        psum += (v > NI && sum > 0) ? exp(v - m) / sum : 0.;
      }
      double choice = psum * gibbs_uniform(A.seed, A.sweep, d.ex_index, draw++);
      uint32_t pick = k1 - 1;
      for (uint32_t k = k0; k < k1; ++k) {
        const uint2 r = oarc[k];
        const double v = A.power * (arc_lnprob(A, r.y) + be[r.x]);
        choice -= (v > NI && sum > 0) ? exp(v - m) / sum : 0.;
        if (choice < 0) {
          pick = k;
          break;
        }
      }
      out[n++] = A.arc_orig[oarc[pick].y];
      s = oarc[pick].x;
    }
    A.new_len[e] = n;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128) k_gibbs(GibbsArgs A) {
  if (A.sequential) {  // one CTA, blocks in corpus order, counts updated in place between blocks
    for (uint32_t e = 0; e < A.n_ex; ++e) {
      const double wt = A.desc[e].weight;
      if (threadIdx.x == 0) add_sample_counts(A, A.old_sample + A.sample_base[e], A.old_len[e], -wt, false);
      __syncthreads();
      sample_block(A, e);
      if (threadIdx.x == 0) add_sample_counts(A, A.new_sample + A.sample_base[e], A.new_len[e], wt, false);
      __syncthreads();
    }
  } else {
    for (uint32_t e = blockIdx.x; e < A.n_ex; e += gridDim.x) sample_block(A, e);
  }
}

// batched mode: apply (new - old) sample counts of every block
__global__ void k_gibbs_apply(GibbsArgs A) {
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.n_ex) return;
  const double wt = A.desc[e].weight;
  add_sample_counts(A, A.old_sample + A.sample_base[e], A.old_len[e], -wt, true);
  add_sample_counts(A, A.new_sample + A.sample_base[e], A.new_len[e], wt, true);
}

__global__ void k_gibbs_lnprob(uint32_t n_params, const uint32_t* __restrict__ param_norm, const double* __restrict__ prior,
                               const double* __restrict__ count, const double* __restrict__ normsum,
                               double* __restrict__ lnp) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_params) return;
  const uint32_t g = param_norm[p];
  const double v = g == CML_NO_GROUP ? prior[p] : count[p] / normsum[g];
  lnp[p] = v > 0 ? log(v) : -CUDART_INF;
}

__global__ void k_gibbs_accumulate(uint32_t n_params, const double* __restrict__ count, double* __restrict__ cum, double dt) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_params) cum[p] += dt * count[p];
}

}  // namespace

extern "C" int cml_gibbs_init(cml_ctx* ctx, const cml_gibbs_model* g) {
  if (!ctx || !g) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model, CML_ERR_STATE, "cml_set_model first");
  CML_REQUIRE(g->n_params == ctx->n_params && g->param_norm && g->param_prior, CML_ERR_ARG, "gibbs model does not match the model");
  CML_REQUIRE(ctx->batches.size() == 1 && ctx->batches[0]->ell_ex == 0, CML_ERR_STATE,
              "Gibbs sampling needs the lattices in ONE batch in the layered-CSR layout (log space or CML_OPT_NO_ELL)");
  cudaSetDevice(ctx->device);
  Batch& bt = *ctx->batches[0];
  cudaStream_t s = ctx->stream;
  std::vector<double> normsum(std::max<uint32_t>(1, g->n_norms), 0.), count(g->n_params);
  for (uint32_t p = 0; p < g->n_params; ++p) {
    const uint32_t n = g->param_norm[p];
    CML_REQUIRE(n == CML_NO_GROUP || n < g->n_norms, CML_ERR_ARG, "param_norm out of range");
    count[p] = g->param_prior[p];
    if (n != CML_NO_GROUP) normsum[n] += g->param_prior[p];  // restore_p0 (gibbs.hpp:618-623)
  }
  ctx->g_norms = g->n_norms;
  CML_CUDA(ctx->g_param_norm.upload(g->param_norm, g->n_params, s));
  CML_CUDA(ctx->g_prior.upload(g->param_prior, g->n_params, s));
  CML_CUDA(ctx->g_count.upload(count.data(), count.size(), s));
  CML_CUDA(ctx->g_normsum.upload(normsum.data(), normsum.size(), s));
  CML_CUDA(ctx->g_cum.alloc(g->n_params));
  CML_CUDA(cudaMemsetAsync(ctx->g_cum.p, 0, g->n_params * sizeof(double), s));
  CML_CUDA(ctx->g_lnp.alloc(g->n_params));
  std::vector<uint32_t> orig((size_t)ctx->n_arcs + 1);
  for (uint32_t a = 0; a <= ctx->n_arcs; ++a) orig[ctx->h_perm[a]] = a;
  CML_CUDA(ctx->g_arc_orig.upload(orig.data(), orig.size(), s));
  // per-example bases: sample slots (n_levels each) and beta scratch (n_states each)
  std::vector<uint64_t> sbase(bt.n_ex + 1, 0), bbase(bt.n_ex + 1, 0);
  for (uint64_t e = 0; e < bt.n_ex; ++e) {
    sbase[e + 1] = sbase[e] + bt.h_nlevels[e];
    bbase[e + 1] = bbase[e] + (bt.h_state_base[e + 1] - bt.h_state_base[e]);
  }
  ctx->h_sample_base = sbase;
  ctx->g_sample_cap = sbase[bt.n_ex];
  CML_CUDA(ctx->g_sample_base.upload(sbase.data(), sbase.size(), s));
  CML_CUDA(ctx->g_beta_base.upload(bbase.data(), bbase.size(), s));
  CML_CUDA(ctx->g_beta.alloc(std::max<uint64_t>(1, bbase[bt.n_ex])));
  for (int i = 0; i < 2; ++i) {
    CML_CUDA(ctx->g_sample[i].alloc(std::max<uint64_t>(1, ctx->g_sample_cap)));
    CML_CUDA(ctx->g_sample_len[i].alloc(bt.n_ex));
    CML_CUDA(cudaMemsetAsync(ctx->g_sample_len[i].p, 0, bt.n_ex * sizeof(uint32_t), s));
  }
  CML_CUDA(cudaStreamSynchronize(s));
  ctx->g_cur = 0;
  ctx->have_gibbs = true;
  return CML_OK;
}

extern "C" int cml_gibbs_sweep(cml_ctx* ctx, const cml_gibbs_sweep_opts* o) {
  if (!ctx || !o) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_gibbs, CML_ERR_STATE, "cml_gibbs_init first");
  cudaSetDevice(ctx->device);
  Batch& bt = *ctx->batches[0];
  cudaStream_t s = ctx->stream;
  GibbsArgs A;
  A.desc = bt.desc.p;
  A.n_ex = (uint32_t)bt.n_ex;
  A.lvl_off = bt.lvl_off.p;
  A.out_off = bt.out_off.p;
  A.out_arc = bt.out_arc.p;
  A.arc_orig = ctx->g_arc_orig.p;
  A.chain_off = ctx->trivial ? nullptr : ctx->chain_off.p;
  A.chain_param = ctx->chain_param.p;
  A.param_norm = ctx->g_param_norm.p;
  A.prior = ctx->g_prior.p;
  A.count = ctx->g_count.p;
  A.normsum = ctx->g_normsum.p;
  A.arc_lnw = nullptr;
  A.beta = ctx->g_beta.p;
  A.beta_base = ctx->g_beta_base.p;
  A.sample_base = ctx->g_sample_base.p;
  A.old_sample = ctx->g_sample[ctx->g_cur].p;
  A.old_len = ctx->g_sample_len[ctx->g_cur].p;
  A.new_sample = ctx->g_sample[ctx->g_cur ^ 1].p;
  A.new_len = ctx->g_sample_len[ctx->g_cur ^ 1].p;
  A.power = o->power;
  A.seed = o->seed;
  A.sweep = o->sweep;
  A.sequential = o->mode == CML_GIBBS_SEQUENTIAL;
  const bool table = o->init_from_params || !A.sequential;
  if (table) {  // per-arc ln probabilities as a table: from the EM weights, or from the frozen counts
    const double* lnp = ctx->ln_w.p;
    if (!o->init_from_params) {
      k_gibbs_lnprob<<<cdiv(ctx->n_params, 256), 256, 0, s>>>(ctx->n_params, ctx->g_param_norm.p, ctx->g_prior.p,
                                                            ctx->g_count.p, ctx->g_normsum.p, ctx->g_lnp.p);
      ++ctx->launches;
      lnp = ctx->g_lnp.p;
    } else
      CML_REQUIRE(ctx->have_params, CML_ERR_STATE, "init_from_params needs cml_set_params");
    CML_REQUIRE(ctx->precision == 64 && ctx->space == CML_SPACE_LOG, CML_ERR_STATE,
                "Gibbs sampling runs in fp64 log space (create the context with precision 64, CML_SPACE_LOG)");
    if (ctx->arc_slot_code.n < (size_t)ctx->n_arcs + 1) {
      CML_CUDA(ctx->arc_slot_code.alloc((size_t)ctx->n_arcs + 1));
      CML_CUDA(cudaMemsetAsync(ctx->arc_slot_code.p, 0xFF, ((size_t)ctx->n_arcs + 1) * sizeof(uint32_t), s));
    }
    cmlk::k_arc_weights<double, false, cmlk::WS<double>><<<cdiv(ctx->n_arcs + 1, 256), 256, 0, s>>>(
        ctx->n_arcs, ctx->trivial ? nullptr : ctx->chain_off.p, ctx->chain_param.p, lnp, ctx->arc_slot_code.p,
        ctx->arc_perm.p, ctx->arc_lnw.p, (double*)ctx->arc_w_real.p, (cmlk::WS<double>*)ctx->arc_ws.p);
    ++ctx->launches;
    A.arc_lnw = (const double*)ctx->arc_w_real.p;
  }
  if (A.sequential) {
    k_gibbs<<<1, 128, 0, s>>>(A);
    ++ctx->launches;
  } else {
    k_gibbs<<<std::min<unsigned>(A.n_ex, 148 * 16), 128, 0, s>>>(A);
    k_gibbs_apply<<<cdiv(A.n_ex, 128), 128, 0, s>>>(A);
    ctx->launches += 2;
  }
  if (o->accumulate_dt != 0.) {
    k_gibbs_accumulate<<<cdiv(ctx->n_params, 256), 256, 0, s>>>(ctx->n_params, ctx->g_count.p, ctx->g_cum.p, o->accumulate_dt);
    ++ctx->launches;
  }
  CML_CUDA(cudaGetLastError());
  CML_CUDA(cudaStreamSynchronize(s));
  ctx->g_cur ^= 1;
  return CML_OK;
}

extern "C" int cml_gibbs_get_samples(cml_ctx* ctx, uint32_t* path_len, uint32_t* path_arcs, uint64_t cap) {
  if (!ctx || !path_len) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_gibbs, CML_ERR_STATE, "cml_gibbs_init first");
  cudaSetDevice(ctx->device);
  Batch& bt = *ctx->batches[0];
  CML_CUDA(cudaMemcpyAsync(path_len, ctx->g_sample_len[ctx->g_cur].p, bt.n_ex * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                           ctx->stream));
  if (path_arcs) {
    CML_REQUIRE(cap >= ctx->g_sample_cap, CML_ERR_ARG, "path_arcs too small (need the total number of lattice levels)");
    CML_CUDA(cudaMemcpyAsync(path_arcs, ctx->g_sample[ctx->g_cur].p, ctx->g_sample_cap * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, ctx->stream));
  }
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  return CML_OK;
}

extern "C" uint64_t cml_gibbs_sample_capacity(cml_ctx* ctx) { return ctx ? ctx->g_sample_cap : 0; }

extern "C" int cml_gibbs_get_state(cml_ctx* ctx, double* count, double* cum, double* normsum) {
  if (!ctx) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_gibbs, CML_ERR_STATE, "cml_gibbs_init first");
  cudaSetDevice(ctx->device);
  if (count) CML_CUDA(cudaMemcpyAsync(count, ctx->g_count.p, ctx->n_params * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (cum) CML_CUDA(cudaMemcpyAsync(cum, ctx->g_cum.p, ctx->n_params * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (normsum)
    CML_CUDA(cudaMemcpyAsync(normsum, ctx->g_normsum.p, ctx->g_norms * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  return CML_OK;
}
