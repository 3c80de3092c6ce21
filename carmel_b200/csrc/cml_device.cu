// cml_device.cu -- C ABI (include/carmel_b200.h) of the B200-native carmel training hot path:
// context, model tables, trellis flattening into layered CSR, E-step and M-step launches.
// sm_100a only; there is no CPU fallback (every compute entry point needs the device).
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>

#include "cml_common.cuh"
#include "cml_kernels_fb.cuh"
#include "cml_kernels_model.cuh"

namespace {

enum ExClass { CLS_WARP0 = 0 /* .. CLS_WARP0+NWARPCLS-1 */, NWARPCLS = 6, CLS_CTA = 6, CLS_GLOBAL = 7, NCLS = 8 };
static const uint32_t kWarpCaps[NWARPCLS] = {64, 128, 256, 512, 1024, 2048};

struct Batch {
  uint64_t n_ex = 0, n_states = 0, n_arcs = 0, n_levels = 0;
  DevArray<CmlExDesc> desc;
  DevArray<uint32_t> lvl_off, in_off, out_off;
  DevArray<uint2> in_arc, out_arc;
  DevArray<double> ex_lnp;
  DevArray<uint32_t> ex_list;  // all classes concatenated
  uint32_t cls_begin[NCLS + 1] = {0};
  uint32_t cta_cap = 0;  // shared-memory capacity (states) needed by the CTA class
  DevArray<unsigned char> scratch;
  DevArray<int> scratch_lvl;
  cudaEvent_t ev_fb0 = nullptr, ev_fb1 = nullptr;  // bracket this batch's forward-backward kernels
  ~Batch() {
    if (ev_fb0) cudaEventDestroy(ev_fb0);
    if (ev_fb1) cudaEventDestroy(ev_fb1);
  }
  // host copies kept for introspection (cml_get_example_layout)
  std::vector<uint64_t> h_state_base;
  std::vector<uint32_t> h_level_of, h_local_of, h_nlevels;
};

}  // namespace

struct cml_ctx {
  int device = 0;
  int precision = 64;
  int space = CML_SPACE_LOG;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = CML_SM_COUNT_FALLBACK;
  size_t smem_optin = 0;
  std::string err;
  uint64_t launches = 0;

  // model
  bool have_model = false, trivial = true;
  uint32_t n_arcs = 0, n_params = 0, n_groups = 0, n_ties = 0;
  DevArray<uint32_t> chain_off, chain_param, param_group, param_tie, group_off, group_members, tie_off, tie_members;
  DevArray<double> arc_prior, group_add;
  bool have_prior = false, have_add = false;
  DevArray<double> ln_w, snap[4], arc_lnw, acc, u, old, gsum, glocked, tie_arc, tie_state, tie_maxl;
  DevArray<unsigned char> arc_w_real;
  DevArray<unsigned long long> maxchg;
  bool have_params = false;

  // reduce buffer: [n_arcs counts | sum_ln_p | sum_w_ln_p | n_zero]
  DevArray<double> reduce_own;
  double* reduce = nullptr;
  uint64_t reduce_n = 0;

  std::vector<std::unique_ptr<Batch>> batches;
  bool estimate_pending = false;
};

static thread_local std::string g_create_err;

#define CML_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) {                   \
      ctx->err = (msg);              \
      return (code);                 \
    }                                \
  } while (0)

static inline unsigned cdiv(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

// -------------------------------------------------------------------------------------------------
// context
// -------------------------------------------------------------------------------------------------
extern "C" const char* cml_version(void) { return "carmel_b200 0.1 (sm_100a)"; }

extern "C" int cml_create(cml_ctx** out, int device, int precision, int space) {
  if (!out) return CML_ERR_ARG;
  *out = nullptr;
  if (precision != 32 && precision != 64) {
    g_create_err = "precision must be 32 or 64";
    return CML_ERR_ARG;
  }
  if (space != CML_SPACE_LOG && space != CML_SPACE_SCALED) {
    g_create_err = "space must be CML_SPACE_LOG or CML_SPACE_SCALED";
    return CML_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_err = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (carmel_b200 has no CPU fallback)";
    return CML_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    g_create_err = "device index out of range";
    return CML_ERR_ARG;
  }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return CML_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_err = "carmel_b200 is built for sm_100a (Blackwell B200) only; found sm_" + std::to_string(prop.major) +
                   std::to_string(prop.minor);
    return CML_ERR_CUDA;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    return CML_ERR_CUDA;
  }
  cml_ctx* ctx = new cml_ctx();
  ctx->device = device;
  ctx->precision = precision;
  ctx->space = space;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_err = cudaGetErrorString(e);
    delete ctx;
    return CML_ERR_CUDA;
  }
  ctx->own_stream = true;
  *out = ctx;
  return CML_OK;
}

extern "C" void cml_destroy(cml_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->batches.clear();
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* cml_last_error(cml_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int cml_set_stream(cml_ctx* ctx, void* s) {
  if (!ctx) return CML_ERR_ARG;
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return CML_OK;
}

extern "C" int cml_synchronize(cml_ctx* ctx) {
  if (!ctx) return CML_ERR_ARG;
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  return CML_OK;
}

extern "C" uint64_t cml_launch_count(cml_ctx* ctx) { return ctx ? ctx->launches : 0; }

// -------------------------------------------------------------------------------------------------
// model
// -------------------------------------------------------------------------------------------------
static void build_csr(uint32_t n_keys, const std::vector<uint32_t>& key_of, std::vector<uint32_t>& off,
                      std::vector<uint32_t>& members) {
  off.assign(n_keys + 1, 0);
  for (uint32_t k : key_of)
    if (k != CML_NO_GROUP) ++off[k + 1];
  for (uint32_t i = 0; i < n_keys; ++i) off[i + 1] += off[i];
  members.resize(off[n_keys]);
  std::vector<uint32_t> cur(off.begin(), off.end() - 1);
  for (uint32_t p = 0; p < key_of.size(); ++p)
    if (key_of[p] != CML_NO_GROUP) members[cur[key_of[p]]++] = p;
}

extern "C" int cml_set_model(cml_ctx* ctx, const cml_model* m) {
  if (!ctx || !m) return CML_ERR_ARG;
  cudaSetDevice(ctx->device);
  CML_REQUIRE(m->n_arcs > 0 && m->n_params > 0, CML_ERR_ARG, "empty model");
  CML_REQUIRE(m->param_group && m->param_tie, CML_ERR_ARG, "param_group / param_tie are required");
  const bool trivial = (m->chain_off == nullptr);
  CML_REQUIRE(!trivial || m->n_arcs == m->n_params, CML_ERR_ARG, "trivial cascade needs n_arcs == n_params");
  std::vector<uint32_t> pg(m->param_group, m->param_group + m->n_params), tie_key(m->n_params);
  for (uint32_t p = 0; p < m->n_params; ++p) {
    CML_REQUIRE(pg[p] == CML_NO_GROUP || pg[p] < m->n_groups, CML_ERR_ARG, "param_group out of range");
    const uint32_t t = m->param_tie[p];
    CML_REQUIRE(t == CML_NO_GROUP || t <= m->n_ties, CML_ERR_ARG, "param_tie out of range");
    tie_key[p] = (t == CML_NO_GROUP || t == CML_LOCKED_GROUP) ? CML_NO_GROUP : t - 1;
  }
  if (!trivial) {
    CML_REQUIRE(m->chain_param != nullptr, CML_ERR_ARG, "chain_param missing");
    CML_REQUIRE(m->chain_off[0] == 0, CML_ERR_ARG, "chain_off[0] must be 0");
    for (uint32_t a = 0; a < m->n_arcs; ++a)
      CML_REQUIRE(m->chain_off[a] <= m->chain_off[a + 1], CML_ERR_ARG, "chain_off not monotone");
    for (uint32_t k = 0; k < m->chain_off[m->n_arcs]; ++k)
      CML_REQUIRE(m->chain_param[k] < m->n_params, CML_ERR_ARG, "chain_param out of range");
  }
  std::vector<uint32_t> goff, gmem, toff, tmem;
  build_csr(m->n_groups, pg, goff, gmem);
  build_csr(m->n_ties, tie_key, toff, tmem);

  cudaStream_t s = ctx->stream;
  ctx->trivial = trivial;
  ctx->n_arcs = m->n_arcs;
  ctx->n_params = m->n_params;
  ctx->n_groups = m->n_groups;
  ctx->n_ties = m->n_ties;
  if (!trivial) {
    CML_CUDA(ctx->chain_off.upload(m->chain_off, m->n_arcs + 1, s));
    CML_CUDA(ctx->chain_param.upload(m->chain_param, m->chain_off[m->n_arcs], s));
  } else {
    ctx->chain_off.release();
    ctx->chain_param.release();
  }
  ctx->have_prior = m->arc_prior != nullptr;
  if (ctx->have_prior) CML_CUDA(ctx->arc_prior.upload(m->arc_prior, m->n_arcs, s));
  ctx->have_add = m->group_add != nullptr && m->n_groups > 0;
  if (ctx->have_add) CML_CUDA(ctx->group_add.upload(m->group_add, m->n_groups, s));
  CML_CUDA(ctx->param_group.upload(pg.data(), m->n_params, s));
  CML_CUDA(ctx->param_tie.upload(m->param_tie, m->n_params, s));
  CML_CUDA(ctx->group_off.upload(goff.data(), goff.size(), s));
  CML_CUDA(ctx->group_members.upload(gmem.data(), gmem.size(), s));
  CML_CUDA(ctx->tie_off.upload(toff.data(), toff.size(), s));
  CML_CUDA(ctx->tie_members.upload(tmem.data(), tmem.size(), s));
  CML_CUDA(ctx->ln_w.alloc(m->n_params));
  for (auto& sn : ctx->snap) CML_CUDA(sn.alloc(m->n_params));
  CML_CUDA(ctx->acc.alloc(m->n_params));
  CML_CUDA(ctx->u.alloc(m->n_params));
  CML_CUDA(ctx->old.alloc(m->n_params));
  CML_CUDA(ctx->arc_lnw.alloc(m->n_arcs));
  CML_CUDA(ctx->arc_w_real.alloc((size_t)m->n_arcs * (ctx->precision / 8)));
  CML_CUDA(ctx->gsum.alloc(std::max<uint32_t>(1, m->n_groups)));
  CML_CUDA(ctx->glocked.alloc(std::max<uint32_t>(1, m->n_groups)));
  CML_CUDA(ctx->tie_arc.alloc(std::max<uint32_t>(1, m->n_ties)));
  CML_CUDA(ctx->tie_state.alloc(std::max<uint32_t>(1, m->n_ties)));
  CML_CUDA(ctx->tie_maxl.alloc(std::max<uint32_t>(1, m->n_ties)));
  CML_CUDA(ctx->maxchg.alloc(1));
  CML_CUDA(ctx->reduce_own.alloc((size_t)m->n_arcs + 3));
  ctx->reduce = ctx->reduce_own.p;
  ctx->reduce_n = (uint64_t)m->n_arcs + 3;
  CML_CUDA(cudaMemsetAsync(ctx->reduce, 0, ctx->reduce_n * sizeof(double), s));
  CML_CUDA(cudaStreamSynchronize(s));  // the host staging vectors go out of scope
  ctx->have_model = true;
  ctx->have_params = false;
  return CML_OK;
}

extern "C" int cml_set_params(cml_ctx* ctx, const double* ln_w) {
  if (!ctx || !ln_w) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model, CML_ERR_STATE, "cml_set_model first");
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaMemcpyAsync(ctx->ln_w.p, ln_w, ctx->n_params * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->have_params = true;
  return CML_OK;
}

extern "C" int cml_get_params(cml_ctx* ctx, double* ln_w) {
  if (!ctx || !ln_w) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_params, CML_ERR_STATE, "no parameters set");
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaMemcpyAsync(ln_w, ctx->ln_w.p, ctx->n_params * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  return CML_OK;
}

extern "C" int cml_snapshot_params(cml_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot > 3) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_params, CML_ERR_STATE, "no parameters set");
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaMemcpyAsync(ctx->snap[slot].p, ctx->ln_w.p, ctx->n_params * sizeof(double), cudaMemcpyDeviceToDevice,
                           ctx->stream));
  return CML_OK;
}

extern "C" int cml_restore_params(cml_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot > 3) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_params, CML_ERR_STATE, "no parameters set");
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaMemcpyAsync(ctx->ln_w.p, ctx->snap[slot].p, ctx->n_params * sizeof(double), cudaMemcpyDeviceToDevice,
                           ctx->stream));
  return CML_OK;
}

// -------------------------------------------------------------------------------------------------
// trellis flattening: reference-order adjacency lists -> topologically layered CSR
// -------------------------------------------------------------------------------------------------
namespace {

struct FlatEx {  // per-example sizes discovered in pass 1
  uint32_t n_levels = 0;
  bool cycle = false;
};

// pass 1: longest-path levels via Kahn's algorithm; fills level_of[] / local_of[] for the example
void levelize(uint32_t n, const uint32_t* off, const uint32_t* dst, uint32_t* level_of, uint32_t* local_of,
              FlatEx& fx, std::vector<uint32_t>& indeg, std::vector<uint32_t>& queue, std::vector<uint32_t>& cnt) {
  indeg.assign(n, 0);
  for (uint32_t k = 0, e = off[n]; k < e; ++k) ++indeg[dst[k]];
  queue.clear();
  for (uint32_t s = 0; s < n; ++s) {
    level_of[s] = 0;
    if (indeg[s] == 0) queue.push_back(s);
  }
  uint32_t maxl = 0;
  for (size_t h = 0; h < queue.size(); ++h) {
    const uint32_t s = queue[h];
    const uint32_t l1 = level_of[s] + 1;
    for (uint32_t k = off[s]; k < off[s + 1]; ++k) {
      const uint32_t d = dst[k];
      if (level_of[d] < l1) level_of[d] = l1;
      if (--indeg[d] == 0) {
        queue.push_back(d);
        if (level_of[d] > maxl) maxl = level_of[d];
      }
    }
  }
  if (queue.size() != n) {
    fx.cycle = true;
    return;
  }
  fx.n_levels = maxl + 1;
  cnt.assign(fx.n_levels + 1, 0);
  for (uint32_t s = 0; s < n; ++s) ++cnt[level_of[s] + 1];
  for (uint32_t l = 0; l < fx.n_levels; ++l) cnt[l + 1] += cnt[l];
  for (uint32_t s = 0; s < n; ++s) local_of[s] = cnt[level_of[s]]++;  // stable: (level, reference id)
}

}  // namespace

extern "C" int cml_add_trellises(cml_ctx* ctx, const cml_trellis_batch* b) {
  if (!ctx || !b) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model, CML_ERR_STATE, "cml_set_model first");
  CML_REQUIRE(b->n_ex > 0, CML_ERR_ARG, "empty batch");
  CML_REQUIRE(b->ex_states && b->ex_fin && b->arc_off && b->arc_dst && b->arc_id, CML_ERR_ARG, "null array in batch");
  CML_REQUIRE(b->n_ex < 0xFFFFFFFFull, CML_ERR_ARG, "too many examples in one batch");
  cudaSetDevice(ctx->device);
  const uint64_t n_ex = b->n_ex;
  const int rs = ctx->precision / 8;
  const bool scaled = ctx->space == CML_SPACE_SCALED;

  // prefix sums over the caller's arrays
  std::vector<uint64_t> state_base(n_ex + 1), arc_base(n_ex + 1);
  state_base[0] = arc_base[0] = 0;
  for (uint64_t e = 0; e < n_ex; ++e) {
    const uint32_t n = b->ex_states[e];
    CML_REQUIRE(n > 0, CML_ERR_ARG, "example with no states");
    CML_REQUIRE(b->ex_fin[e] < n, CML_ERR_ARG, "ex_fin out of range");
    state_base[e + 1] = state_base[e] + n;
    const uint32_t* off = b->arc_off + state_base[e] + e;
    CML_REQUIRE(off[0] == 0, CML_ERR_ARG, "arc_off must start at 0 for every example");
    arc_base[e + 1] = arc_base[e] + off[n];
  }
  const uint64_t tot_states = state_base[n_ex], tot_arcs = arc_base[n_ex];

  std::unique_ptr<Batch> bt(new Batch());
  bt->n_ex = n_ex;
  bt->n_states = tot_states;
  bt->n_arcs = tot_arcs;
  bt->h_state_base = state_base;
  bt->h_level_of.resize(tot_states);
  bt->h_local_of.resize(tot_states);
  bt->h_nlevels.resize(n_ex);

  // pass 1 (parallel over examples): levels
  const unsigned nthr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  std::atomic<int> bad_cycle{0}, bad_range{0};
  {
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
      std::vector<uint32_t> indeg, queue, cnt;
      for (;;) {
        const uint64_t e0 = next.fetch_add(256);
        if (e0 >= n_ex) break;
        for (uint64_t e = e0; e < std::min(n_ex, e0 + 256); ++e) {
          const uint32_t n = b->ex_states[e];
          const uint32_t* off = b->arc_off + state_base[e] + e;
          const uint32_t* dst = b->arc_dst + arc_base[e];
          const uint32_t* id = b->arc_id + arc_base[e];
          bool ok = true;
          for (uint32_t s = 0; s < n && ok; ++s) ok = off[s] <= off[s + 1];
          for (uint32_t k = 0; k < off[n] && ok; ++k) ok = dst[k] < n && id[k] < ctx->n_arcs;
          if (!ok) {
            bad_range = 1;
            continue;
          }
          FlatEx fx;
          levelize(n, off, dst, &bt->h_level_of[state_base[e]], &bt->h_local_of[state_base[e]], fx, indeg, queue, cnt);
          if (fx.cycle) bad_cycle = 1;
          bt->h_nlevels[e] = fx.n_levels;
        }
      }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthr; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  CML_REQUIRE(!bad_range, CML_ERR_ARG, "trellis arc destination or arc id out of range");
  CML_REQUIRE(!bad_cycle, CML_ERR_CYCLE,
              "derivation lattice has a cycle (the reference warns 'Forward/backward will miss some paths')");

  // layout offsets
  std::vector<uint64_t> lvl_base(n_ex + 1);
  lvl_base[0] = 0;
  for (uint64_t e = 0; e < n_ex; ++e) lvl_base[e + 1] = lvl_base[e] + 3ull * bt->h_nlevels[e] + 1;
  bt->n_levels = 0;
  for (uint64_t e = 0; e < n_ex; ++e) bt->n_levels += bt->h_nlevels[e];

  std::vector<CmlExDesc> desc(n_ex);
  std::vector<uint32_t> h_lvl(lvl_base[n_ex]), h_in_off(tot_states + n_ex), h_out_off(tot_states + n_ex);
  std::vector<uint2> h_in(tot_arcs), h_out(tot_arcs);

  // classes
  const size_t per_state = 2 * (size_t)rs + (scaled ? 8 : 0);
  const size_t smem_budget = std::min<size_t>(ctx->smem_optin ? ctx->smem_optin : 48 * 1024, 220 * 1024);
  const uint32_t cta_max_states = (uint32_t)((smem_budget - 64) / per_state);
  std::vector<std::vector<uint32_t>> cls(NCLS);
  uint64_t scratch_states = 0;
  std::vector<uint32_t> ex_width(n_ex);

  // pass 2 (parallel): fill CSR arrays
  {
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
      std::vector<uint32_t> ref_of, icur;
      for (;;) {
        const uint64_t e0 = next.fetch_add(256);
        if (e0 >= n_ex) break;
        for (uint64_t e = e0; e < std::min(n_ex, e0 + 256); ++e) {
          const uint32_t n = b->ex_states[e], nl = bt->h_nlevels[e];
          const uint32_t* off = b->arc_off + state_base[e] + e;
          const uint32_t* dst = b->arc_dst + arc_base[e];
          const uint32_t* id = b->arc_id + arc_base[e];
          const uint32_t* level_of = &bt->h_level_of[state_base[e]];
          const uint32_t* local_of = &bt->h_local_of[state_base[e]];
          uint32_t* lv = &h_lvl[lvl_base[e]];
          uint32_t* lmin = lv + nl + 1;
          uint32_t* lmax = lmin + nl;
          uint32_t* ioff = &h_in_off[state_base[e] + e];
          uint32_t* ooff = &h_out_off[state_base[e] + e];
          uint2* ia = &h_in[arc_base[e]];
          uint2* oa = &h_out[arc_base[e]];
          // level offsets
          for (uint32_t l = 0; l <= nl; ++l) lv[l] = 0;
          for (uint32_t s = 0; s < n; ++s) ++lv[level_of[s] + 1];
          uint32_t width = 0;
          for (uint32_t l = 0; l < nl; ++l) {
            width = std::max(width, lv[l + 1]);
            lv[l + 1] += lv[l];
          }
          ex_width[e] = width;
          for (uint32_t l = 0; l < nl; ++l) {
            lmin[l] = l ? l - 1 : 0;
            lmax[l] = std::min(nl - 1, l + 1);
          }
          // row sizes
          for (uint32_t s = 0; s <= n; ++s) ioff[s] = ooff[s] = 0;
          for (uint32_t s = 0; s < n; ++s) {
            ooff[local_of[s] + 1] = off[s + 1] - off[s];
            for (uint32_t k = off[s]; k < off[s + 1]; ++k) {
              ++ioff[local_of[dst[k]] + 1];
              const uint32_t ls = level_of[s], ld = level_of[dst[k]];
              if (ls < lmin[ld]) lmin[ld] = ls;
              if (ld > lmax[ls]) lmax[ls] = ld;
            }
          }
          for (uint32_t s = 0; s < n; ++s) {
            ioff[s + 1] += ioff[s];
            ooff[s + 1] += ooff[s];
          }
          // fill: outgoing in reference list order; incoming ordered by source layered index
          // visit sources in layered order so each in-list is sorted by source index (stable)
          ref_of.resize(n);
          for (uint32_t s = 0; s < n; ++s) ref_of[local_of[s]] = s;
          icur.assign(ioff, ioff + n);
          for (uint32_t j = 0; j < n; ++j) {
            const uint32_t s = ref_of[j];
            uint32_t o = ooff[j];
            for (uint32_t k = off[s]; k < off[s + 1]; ++k) {
              const uint32_t dj = local_of[dst[k]];
              oa[o++] = make_uint2(dj, id[k]);
              ia[icur[dj]++] = make_uint2(j, id[k]);
            }
          }
          CmlExDesc& d = desc[e];
          d.arc_base = arc_base[e];
          d.row_base = state_base[e] + e;
          d.lvl_base = lvl_base[e];
          d.scratch_base = 0;
          d.n_states = n;
          d.n_levels = nl;
          d.fin = local_of[b->ex_fin[e]];
          d.ex_index = (uint32_t)e;
          d.weight = b->ex_weight ? b->ex_weight[e] : 1.0;
          d.ln_weight = d.weight > 0 ? std::log(d.weight) : -INFINITY;
        }
      }
    };
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nthr; ++t) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
  }
  // classify (serial; cheap)
  for (uint64_t e = 0; e < n_ex; ++e) {
    const uint32_t n = desc[e].n_states;
    int c = -1;
    if (ex_width[e] <= 96) {
      for (int i = 0; i < NWARPCLS; ++i)
        if (n <= kWarpCaps[i]) {
          c = CLS_WARP0 + i;
          break;
        }
    }
    if (c < 0) {
      if (n <= cta_max_states) {
        c = CLS_CTA;
        bt->cta_cap = std::max(bt->cta_cap, n);
      } else {
        c = CLS_GLOBAL;
        desc[e].scratch_base = scratch_states;
        scratch_states += n;
      }
    }
    cls[c].push_back((uint32_t)e);
  }
  std::vector<uint32_t> ex_list;
  ex_list.reserve(n_ex);
  for (int c = 0; c < NCLS; ++c) {
    bt->cls_begin[c] = (uint32_t)ex_list.size();
    // longest examples first inside a class: better tail behaviour
    std::stable_sort(cls[c].begin(), cls[c].end(), [&](uint32_t a, uint32_t b2) {
      return desc[a].n_levels > desc[b2].n_levels;
    });
    ex_list.insert(ex_list.end(), cls[c].begin(), cls[c].end());
  }
  bt->cls_begin[NCLS] = (uint32_t)ex_list.size();

  cudaStream_t s = ctx->stream;
  CML_CUDA(bt->desc.upload(desc.data(), n_ex, s));
  CML_CUDA(bt->lvl_off.upload(h_lvl.data(), h_lvl.size(), s));
  CML_CUDA(bt->in_off.upload(h_in_off.data(), h_in_off.size(), s));
  CML_CUDA(bt->out_off.upload(h_out_off.data(), h_out_off.size(), s));
  CML_CUDA(bt->in_arc.upload(h_in.data(), h_in.size(), s));
  CML_CUDA(bt->out_arc.upload(h_out.data(), h_out.size(), s));
  CML_CUDA(bt->ex_list.upload(ex_list.data(), ex_list.size(), s));
  CML_CUDA(bt->ex_lnp.alloc(n_ex));
  if (scratch_states) {
    CML_CUDA(bt->scratch.alloc(scratch_states * 2 * rs));
    if (scaled) CML_CUDA(bt->scratch_lvl.alloc(scratch_states * 2));
  }
  CML_CUDA(cudaStreamSynchronize(s));
  ctx->batches.push_back(std::move(bt));
  return CML_OK;
}

extern "C" int cml_clear_trellises(cml_ctx* ctx) {
  if (!ctx) return CML_ERR_ARG;
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->batches.clear();
  return CML_OK;
}

extern "C" int cml_trellis_totals(cml_ctx* ctx, uint64_t* n_ex, uint64_t* n_states, uint64_t* n_arcs,
                                  uint64_t* n_levels) {
  if (!ctx) return CML_ERR_ARG;
  uint64_t a = 0, b = 0, c = 0, d = 0;
  for (auto& bt : ctx->batches) {
    a += bt->n_ex;
    b += bt->n_states;
    c += bt->n_arcs;
    d += bt->n_levels;
  }
  if (n_ex) *n_ex = a;
  if (n_states) *n_states = b;
  if (n_arcs) *n_arcs = c;
  if (n_levels) *n_levels = d;
  return CML_OK;
}

extern "C" int cml_get_example_layout(cml_ctx* ctx, uint64_t e, uint32_t* n_levels, uint32_t* level_of,
                                      uint32_t* local_of) {
  if (!ctx) return CML_ERR_ARG;
  for (auto& bt : ctx->batches) {
    if (e < bt->n_ex) {
      const uint64_t s0 = bt->h_state_base[e], s1 = bt->h_state_base[e + 1];
      if (n_levels) *n_levels = bt->h_nlevels[e];
      if (level_of) std::memcpy(level_of, &bt->h_level_of[s0], (s1 - s0) * 4);
      if (local_of) std::memcpy(local_of, &bt->h_local_of[s0], (s1 - s0) * 4);
      return CML_OK;
    }
    e -= bt->n_ex;
  }
  ctx->err = "example index out of range";
  return CML_ERR_ARG;
}

// -------------------------------------------------------------------------------------------------
// E-step
// -------------------------------------------------------------------------------------------------
template <typename Real, bool SCALED>
static int launch_fb(cml_ctx* ctx, Batch& bt) {
  using namespace cmlk;
  FbArgs A;
  A.desc = bt.desc.p;
  A.lvl_off = bt.lvl_off.p;
  A.in_off = bt.in_off.p;
  A.in_arc = bt.in_arc.p;
  A.out_off = bt.out_off.p;
  A.out_arc = bt.out_arc.p;
  A.arc_w = ctx->arc_w_real.p;
  A.counts = ctx->reduce;
  A.ex_lnp = bt.ex_lnp.p;
  A.scratch = bt.scratch.p;
  A.scratch_lvl = bt.scratch_lvl.p;
  const size_t per_state = 2 * sizeof(Real) + (SCALED ? 2 * sizeof(int) : 0);
  if (!bt.ev_fb0) {
    CML_CUDA(cudaEventCreate(&bt.ev_fb0));
    CML_CUDA(cudaEventCreate(&bt.ev_fb1));
  }
  CML_CUDA(cudaEventRecord(bt.ev_fb0, ctx->stream));
  for (int c = 0; c < NWARPCLS; ++c) {
    const uint32_t n = bt.cls_begin[c + 1] - bt.cls_begin[c];
    if (!n) continue;
    A.ex_list = bt.ex_list.p + bt.cls_begin[c];
    A.n_list = n;
    A.cap_states = kWarpCaps[c];
    const size_t smem = 4 * per_state * kWarpCaps[c];
    auto kern = k_fb_warp<Real, SCALED>;
    if (smem > 48 * 1024) CML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<cdiv(n, 4), 128, smem, ctx->stream>>>(A);
    ++ctx->launches;
  }
  {
    const uint32_t n = bt.cls_begin[CLS_CTA + 1] - bt.cls_begin[CLS_CTA];
    if (n) {
      A.ex_list = bt.ex_list.p + bt.cls_begin[CLS_CTA];
      A.n_list = n;
      A.cap_states = bt.cta_cap;
      const size_t smem = per_state * bt.cta_cap;
      auto kern = k_fb_cta<Real, SCALED, false>;
      if (smem > 48 * 1024)
        CML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<n, 256, smem, ctx->stream>>>(A);
      ++ctx->launches;
    }
  }
  {
    const uint32_t n = bt.cls_begin[CLS_GLOBAL + 1] - bt.cls_begin[CLS_GLOBAL];
    if (n) {
      A.ex_list = bt.ex_list.p + bt.cls_begin[CLS_GLOBAL];
      A.n_list = n;
      A.cap_states = 0;
      k_fb_cta<Real, SCALED, true><<<n, 256, 0, ctx->stream>>>(A);
      ++ctx->launches;
    }
  }
  CML_CUDA(cudaEventRecord(bt.ev_fb1, ctx->stream));
  k_reduce_lnp<<<std::min<unsigned>(cdiv(bt.n_ex, 256), 4 * ctx->sm_count), 256, 0, ctx->stream>>>(
      bt.ex_lnp.p, bt.desc.p, bt.n_ex, ctx->reduce + ctx->n_arcs);
  ++ctx->launches;
  CML_CUDA(cudaGetLastError());
  return CML_OK;
}

template <typename Real, bool SCALED>
static int launch_arc_weights(cml_ctx* ctx) {
  cmlk::k_arc_weights<Real, SCALED><<<cdiv(ctx->n_arcs, 256), 256, 0, ctx->stream>>>(
      ctx->n_arcs, ctx->trivial ? nullptr : ctx->chain_off.p, ctx->chain_param.p, ctx->ln_w.p, ctx->arc_lnw.p,
      (Real*)ctx->arc_w_real.p);
  ++ctx->launches;
  CML_CUDA(cudaGetLastError());
  return CML_OK;
}

extern "C" int cml_estimate_launch(cml_ctx* ctx) {
  if (!ctx) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model && ctx->have_params, CML_ERR_STATE, "cml_set_model and cml_set_params first");
  CML_REQUIRE(!ctx->batches.empty(), CML_ERR_NODERIV, "no trellises resident (no training example had a derivation)");
  cudaSetDevice(ctx->device);
  const bool sc = ctx->space == CML_SPACE_SCALED;
  int r;
  if (ctx->precision == 64)
    r = sc ? launch_arc_weights<double, true>(ctx) : launch_arc_weights<double, false>(ctx);
  else
    r = sc ? launch_arc_weights<float, true>(ctx) : launch_arc_weights<float, false>(ctx);
  if (r) return r;
  CML_CUDA(cudaMemsetAsync(ctx->reduce, 0, ctx->reduce_n * sizeof(double), ctx->stream));
  for (auto& bt : ctx->batches) {
    if (ctx->precision == 64)
      r = sc ? launch_fb<double, true>(ctx, *bt) : launch_fb<double, false>(ctx, *bt);
    else
      r = sc ? launch_fb<float, true>(ctx, *bt) : launch_fb<float, false>(ctx, *bt);
    if (r) return r;
  }
  ctx->estimate_pending = true;
  return CML_OK;
}

extern "C" int cml_estimate_finish(cml_ctx* ctx, cml_estimate_result* out) {
  if (!ctx) return CML_ERR_ARG;
  CML_REQUIRE(ctx->estimate_pending, CML_ERR_STATE, "cml_estimate_launch first");
  cudaSetDevice(ctx->device);
  double h[3];
  CML_CUDA(cudaMemcpyAsync(h, ctx->reduce + ctx->n_arcs, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->estimate_pending = false;
  if (out) {
    out->sum_ln_p = h[0];
    out->sum_w_ln_p = h[1];
    out->n_zero = (uint64_t)(h[2] + 0.5);
  }
  return CML_OK;
}

extern "C" int cml_estimate(cml_ctx* ctx, cml_estimate_result* out) {
  int r = cml_estimate_launch(ctx);
  if (r) return r;
  return cml_estimate_finish(ctx, out);
}

extern "C" int cml_last_fb_time_ms(cml_ctx* ctx, float* ms, uint32_t* n_kernels) {
  if (!ctx || !ms) return CML_ERR_ARG;
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  float tot = 0;
  uint32_t nk = 0;
  for (auto& bt : ctx->batches) {
    if (!bt->ev_fb0) continue;
    float t = 0;
    CML_CUDA(cudaEventElapsedTime(&t, bt->ev_fb0, bt->ev_fb1));
    tot += t;
    for (int c = 0; c < NCLS; ++c) nk += bt->cls_begin[c + 1] > bt->cls_begin[c];
  }
  *ms = tot;
  if (n_kernels) *n_kernels = nk;
  return CML_OK;
}

extern "C" int cml_get_example_logprob(cml_ctx* ctx, double* ln_p, uint64_t n) {
  if (!ctx || !ln_p) return CML_ERR_ARG;
  cudaSetDevice(ctx->device);
  uint64_t done = 0;
  for (auto& bt : ctx->batches) {
    if (done >= n) break;
    const uint64_t k = std::min<uint64_t>(bt->n_ex, n - done);
    CML_CUDA(cudaMemcpyAsync(ln_p + done, bt->ex_lnp.p, k * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    done += k;
  }
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  CML_REQUIRE(done == n, CML_ERR_ARG, "fewer examples resident than requested");
  return CML_OK;
}

extern "C" int cml_get_arc_counts(cml_ctx* ctx, double* counts) {
  if (!ctx || !counts) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model, CML_ERR_STATE, "cml_set_model first");
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaMemcpyAsync(counts, ctx->reduce, ctx->n_arcs * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  return CML_OK;
}

extern "C" int cml_reduce_buffer(cml_ctx* ctx, void** p, uint64_t* n) {
  if (!ctx) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model, CML_ERR_STATE, "cml_set_model first");
  if (p) *p = ctx->reduce;
  if (n) *n = ctx->reduce_n;
  return CML_OK;
}

extern "C" int cml_reduce_buffer_write(cml_ctx* ctx, const double* src, uint64_t n) {
  if (!ctx || !src) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model && n <= ctx->reduce_n, CML_ERR_ARG, "reduce buffer smaller than requested");
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaMemsetAsync(ctx->reduce, 0, ctx->reduce_n * sizeof(double), ctx->stream));
  CML_CUDA(cudaMemcpyAsync(ctx->reduce, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  return CML_OK;
}

extern "C" int cml_reduce_buffer_read(cml_ctx* ctx, double* dst, uint64_t n) {
  if (!ctx || !dst) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model && n <= ctx->reduce_n, CML_ERR_ARG, "reduce buffer smaller than requested");
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaMemcpyAsync(dst, ctx->reduce, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CML_CUDA(cudaStreamSynchronize(ctx->stream));
  return CML_OK;
}

extern "C" int cml_use_reduce_buffer(cml_ctx* ctx, void* p, uint64_t n) {
  if (!ctx) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model, CML_ERR_STATE, "cml_set_model first");
  if (!p) {
    ctx->reduce = ctx->reduce_own.p;
    return CML_OK;
  }
  CML_REQUIRE(n >= ctx->reduce_n, CML_ERR_ARG, "reduce buffer too small (need n_arcs + 3 doubles)");
  ctx->reduce = (double*)p;
  return CML_OK;
}

// -------------------------------------------------------------------------------------------------
// M-step
// -------------------------------------------------------------------------------------------------
static int run_normalize(cml_ctx* ctx) {  // u -> ln_w
  using namespace cmlk;
  cudaStream_t s = ctx->stream;
  if (ctx->n_groups) {
    k_norm_sums<<<cdiv((uint64_t)ctx->n_groups * 32, 256), 256, 0, s>>>(
        ctx->n_groups, ctx->group_off.p, ctx->group_members.p, ctx->have_add ? ctx->group_add.p : nullptr,
        ctx->param_tie.p, ctx->u.p, ctx->gsum.p, ctx->glocked.p);
    ++ctx->launches;
    if (ctx->n_ties) {
      k_tie_totals<<<cdiv((uint64_t)ctx->n_ties * 32, 256), 256, 0, s>>>(
          ctx->n_ties, ctx->tie_off.p, ctx->tie_members.p, ctx->param_group.p, ctx->u.p, ctx->gsum.p, ctx->glocked.p,
          ctx->tie_arc.p, ctx->tie_state.p, ctx->tie_maxl.p);
      ++ctx->launches;
    }
    k_norm_assign<<<cdiv((uint64_t)ctx->n_groups * 32, 256), 256, 0, s>>>(
        ctx->n_groups, ctx->group_off.p, ctx->group_members.p, ctx->param_tie.p, ctx->u.p, ctx->tie_arc.p,
        ctx->tie_state.p, ctx->tie_maxl.p, ctx->ln_w.p);
    ++ctx->launches;
  }
  k_copy_ungrouped<<<cdiv(ctx->n_params, 256), 256, 0, s>>>(ctx->n_params, ctx->param_group.p, ctx->u.p, ctx->ln_w.p);
  ++ctx->launches;
  CML_CUDA(cudaGetLastError());
  return CML_OK;
}

extern "C" int cml_normalize_params(cml_ctx* ctx) {
  if (!ctx) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model && ctx->have_params, CML_ERR_STATE, "cml_set_model and cml_set_params first");
  cudaSetDevice(ctx->device);
  CML_CUDA(cudaMemcpyAsync(ctx->u.p, ctx->ln_w.p, ctx->n_params * sizeof(double), cudaMemcpyDeviceToDevice,
                           ctx->stream));
  return run_normalize(ctx);
}

extern "C" int cml_maximize(cml_ctx* ctx, double rate, double* max_delta) {
  using namespace cmlk;
  if (!ctx) return CML_ERR_ARG;
  CML_REQUIRE(ctx->have_model && ctx->have_params, CML_ERR_STATE, "cml_set_model and cml_set_params first");
  cudaSetDevice(ctx->device);
  cudaStream_t s = ctx->stream;
  if (!ctx->trivial) CML_CUDA(cudaMemsetAsync(ctx->acc.p, 0, ctx->n_params * sizeof(double), s));
  k_param_acc<<<cdiv(ctx->n_arcs, 256), 256, 0, s>>>(ctx->n_arcs, ctx->trivial ? nullptr : ctx->chain_off.p,
                                                     ctx->chain_param.p, ctx->reduce,
                                                     ctx->have_prior ? ctx->arc_prior.p : nullptr, ctx->param_tie.p,
                                                     ctx->acc.p);
  k_unnorm<<<cdiv(ctx->n_params, 256), 256, 0, s>>>(ctx->n_params, ctx->acc.p, ctx->ln_w.p, ctx->param_tie.p,
                                                   ctx->param_group.p, ctx->u.p, ctx->old.p);
  ctx->launches += 2;
  int r = run_normalize(ctx);
  if (r) return r;
  if (rate > 1. && ctx->trivial) {  // over-relaxation is disabled for real cascades (train.cc:543-549)
    k_overrelax<<<cdiv(ctx->n_params, 256), 256, 0, s>>>(ctx->n_params, rate, ctx->param_tie.p, ctx->old.p,
                                                        ctx->ln_w.p, ctx->u.p);
    ++ctx->launches;
    if ((r = run_normalize(ctx))) return r;
  }
  CML_CUDA(cudaMemsetAsync(ctx->maxchg.p, 0, sizeof(unsigned long long), s));
  k_max_change<<<cdiv(ctx->n_params, 256), 256, 0, s>>>(ctx->n_params, ctx->param_tie.p, ctx->old.p, ctx->ln_w.p,
                                                       ctx->maxchg.p);
  ++ctx->launches;
  CML_CUDA(cudaGetLastError());
  unsigned long long bits = 0;
  CML_CUDA(cudaMemcpyAsync(&bits, ctx->maxchg.p, sizeof(bits), cudaMemcpyDeviceToHost, s));
  CML_CUDA(cudaStreamSynchronize(s));
  if (max_delta) std::memcpy(max_delta, &bits, sizeof(double));
  return CML_OK;
}

// -------------------------------------------------------------------------------------------------
extern "C" const char* const* cml_exported_symbols(size_t* n) {
  static const char* const syms[] = {
      "cml_version", "cml_create", "cml_destroy", "cml_last_error", "cml_set_stream", "cml_synchronize",
      "cml_launch_count", "cml_set_model", "cml_set_params", "cml_get_params", "cml_snapshot_params",
      "cml_restore_params", "cml_add_trellises", "cml_clear_trellises", "cml_trellis_totals",
      "cml_get_example_layout", "cml_estimate", "cml_estimate_launch", "cml_estimate_finish",
      "cml_get_example_logprob", "cml_get_arc_counts", "cml_reduce_buffer", "cml_use_reduce_buffer", "cml_maximize",
      "cml_normalize_params", "cml_exported_symbols", "cml_reduce_buffer_write", "cml_reduce_buffer_read",
      "cml_job_open", "cml_job_close", "cml_job_error", "cml_job_set_allreduce", "cml_job_prepare", "cml_job_context",
      "cml_job_train", "cml_job_write", "cml_job_stats", "cml_last_fb_time_ms"};
  if (n) *n = sizeof(syms) / sizeof(syms[0]);
  return syms;
}
