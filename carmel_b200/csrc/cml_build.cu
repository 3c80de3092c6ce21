// cml_build.cu -- device-side construction of the per-example derivation lattices.
//
// Reference: derivations::compute / derive / add_arcs / prune (carmel/src/derivations.h:479-513,572-629,640-704) and the
// uncached path's per-iteration rebuild (cached_derivs.h:77-95).  Output contract (bit for bit, checked by the byte-compare
// tests of the lattice dump): states (i, s, o) numbered in DFS pre-order from (0,0,0); at every state the label classes are
// tried in the order (eps:eps), (eps:out[o]), (in[i]:eps), (in[i]:out[o]) and within a class the WFST arcs in arc-table
// order; an arc is kept unless its destination is already known dead when it is examined; a state's stored arc list is the
// reverse of the order its arcs were kept; dead states are removed and the survivors renumbered in place.
//
// Design.  The DFS order is the contract, so an example is walked by ONE thread with an explicit stack -- the parallelism
// is over examples (a 1M-sentence corpus keeps every SM busy; the walk is pointer chasing in per-worker scratch, L2/HBM
// latency bound, hidden by tens of thousands of resident workers).  Workers are persistent threads that pull examples from
// an atomic counter; each owns a scratch region (open-addressing map (i,s,o) -> id with generation stamps, dead flags, DFS
// stack, kept-arc list, renumbering) sized for the round's capacity.  Pass 1 walks every example and records its sizes
// (examples whose walk overflows the round's capacity are retried in a later round with 4x the capacity and fewer workers);
// exact CSR offsets are prefix sums; pass 2 walks the kept examples again and writes rows and arcs at their final places.
// Nothing here touches the oracle or a CPU fallback: the host-side builder (host/trellis.cpp) stays as the product's
// default for small corpora and as the second implementation the tests compare with.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "cml_ctx.cuh"

namespace {

struct BIo {
  const uint32_t* state_off;    // [Q+1] per WFST state: range in keys / range_begin
  const uint64_t* keys;         // sorted (in << 32 | out) per state
  const uint32_t* range_begin;  // per key: begin in ids (end = next begin)
  const uint32_t* ids;          // arc-table ids, grouped by (state, label pair), each group in arc-table order
  const uint32_t* dest;         // [n_arcs] destination state per arc-table id
  uint32_t final_state;
};
struct BCorpus {
  const uint64_t* in_off;
  const uint32_t* in_sym;
  const uint64_t* out_off;
  const uint32_t* out_sym;
};
struct __align__(16) BSlot {
  uint32_t io, s, id, gen;
};
struct __align__(16) BFrame {
  uint32_t id, io, s, phase_dead;  // phase in bits 0..2, dead in bit 8
  uint32_t cur, end, pending, pad;
};
struct BKept {
  uint32_t src, dst, id;
};
enum { B_OK = 0, B_DROPPED = 1, B_OVERFLOW = 2, B_TOOLONG = 3 };
struct BArgs {
  BIo io;
  BCorpus c;
  const uint32_t* todo;  // example numbers of this launch
  uint32_t n_todo;
  uint32_t C, K, mask;   // capacities: states (= stack depth), kept arcs; map slots - 1
  BSlot* map;
  uint8_t* dead;
  BFrame* stack;
  BKept* kept;
  uint32_t* renum;
  uint32_t* cnt;
  unsigned long long* next;     // work counter
  // pass 1 outputs (indexed by example number)
  uint8_t* status;
  uint32_t* n_states;           // live states
  uint32_t* n_arcs;             // final arcs
  uint32_t* fin;
  uint32_t* peak;               // [3] max over examples: states before pruning, kept arcs before pruning, stack depth
  unsigned long long* pre_arcs; // arcs examined
  // pass 2
  int fill;
  const uint64_t* row_base;     // [example] first entry of its arc_off row block
  const uint64_t* arc_base;
  uint32_t* out_off;
  uint32_t* out_dst;
  uint32_t* out_id;
};

__device__ __forceinline__ uint32_t b_hash(uint32_t io, uint32_t s) {
  uint32_t h = io * 0x9E3779B1u ^ (s + 0x7F4A7C15u) * 0x85EBCA6Bu;
  h ^= h >> 15;
  h *= 0xC2B2AE35u;
  return h ^ (h >> 13);
}

__global__ void __launch_bounds__(128) k_build_trellis(BArgs A) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  BSlot* __restrict__ map = A.map + (size_t)w * (A.mask + 1);
  uint8_t* __restrict__ dead = A.dead + (size_t)w * A.C;
  BFrame* __restrict__ stack = A.stack + (size_t)w * A.C;
  BKept* __restrict__ kept = A.kept + (size_t)w * A.K;
  uint32_t* __restrict__ renum = A.renum + (size_t)w * A.C;
  uint32_t* __restrict__ cnt = A.cnt + (size_t)w * (A.C + 1);
  uint32_t gen = 0;  // the scratch is zeroed before the launch: stamp 0 = empty
  for (;;) {
    const unsigned long long t = atomicAdd(A.next, 1ull);
    if (t >= A.n_todo) break;
    const uint32_t e = A.todo[t];
    const uint64_t i0 = A.c.in_off[e], o0 = A.c.out_off[e];
    const uint32_t nin = (uint32_t)(A.c.in_off[e + 1] - i0), nout = (uint32_t)(A.c.out_off[e + 1] - o0);
    if (nin > 65534u || nout > 65534u) {
      A.status[e] = B_TOOLONG;
      continue;
    }
    const uint32_t* __restrict__ in = A.c.in_sym + i0;
    const uint32_t* __restrict__ out = A.c.out_sym + o0;
    const uint32_t goal_io = (nin << 16) | nout, goal_s = A.io.final_state;
    ++gen;
    uint32_t n_states = 0, n_kept = 0, sp = 0, sp_max = 0;
    unsigned long long pre = 0;
    bool overflow = false;
    // map: find (io, s) or insert it with id = n_states
    auto find_or_insert = [&](uint32_t io, uint32_t s, bool& inserted) -> uint32_t {
      uint32_t h = b_hash(io, s) & A.mask;
      for (;;) {
        const BSlot sl = map[h];
        if (sl.gen != gen) {
          map[h] = BSlot{io, s, n_states, gen};
          inserted = true;
          return n_states;
        }
        if (sl.io == io && sl.s == s) {
          inserted = false;
          return sl.id;
        }
        h = (h + 1) & A.mask;
      }
    };
    BFrame f;
    {
      bool ins;
      find_or_insert(0u, 0u, ins);
      f = BFrame{0u, 0u, 0u, (goal_io == 0u && goal_s == 0u) ? 0u : 0x100u, 0u, 0u, 0u, 0u};
      dead[0] = 0;
      n_states = 1;
      sp = sp_max = 1;
    }
    while (sp) {
      if (f.cur == f.end) {
        const uint32_t ph = f.phase_dead & 7u;
        if (ph == 4u) {  // every label class done: the state is finished
          const uint32_t done = f.id, done_dead = f.phase_dead >> 8;
          dead[done] = (uint8_t)done_dead;
          if (--sp == 0) break;
          f = stack[sp - 1];  // the parent was waiting for this destination (add_arcs, derivations.h:691-701)
          if (!done_dead) {
            if (n_kept < A.K) kept[n_kept] = BKept{f.id, done, f.pending};
            else overflow = true;
            ++n_kept;
            f.phase_dead &= 7u;
          }
          continue;
        }
        f.phase_dead += 1u;  // open label class ph (derivations.h:656-670)
        const uint32_t i = f.io >> 16, o = f.io & 0xFFFFu;
        const bool useI = i < nin, useO = o < nout;
        uint32_t lin = 0, lout = 0;  // 0 = *e*
        bool active = true;
        if (ph == 1u) {
          active = useO;
          if (active) lout = out[o];
        } else if (ph == 2u) {
          active = useI;
          if (active) lin = in[i];
        } else if (ph == 3u) {
          active = useI && useO;
          if (active) {
            lin = in[i];
            lout = out[o];
          }
        }
        f.cur = f.end = 0;
        if (active) {  // arcs of WFST state f.s with this label pair (wfst_io_index, derivations.h:142-155)
          const uint64_t key = ((uint64_t)lin << 32) | lout;
          uint32_t lo = A.io.state_off[f.s], hi = A.io.state_off[f.s + 1];
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (A.io.keys[mid] < key) lo = mid + 1;
            else hi = mid;
          }
          if (lo < A.io.state_off[f.s + 1] && A.io.keys[lo] == key) {
            f.cur = A.io.range_begin[lo];
            f.end = A.io.range_begin[lo + 1];
          }
        }
        continue;
      }
      // examine the next arc of the open class
      const uint32_t id = A.io.ids[f.cur++];
      ++pre;
      const uint32_t ph = (f.phase_dead & 7u) - 1u;
      const uint32_t nio = f.io + ((ph & 2u) ? 0x10000u : 0u) + (ph & 1u);
      const uint32_t ds = A.io.dest[id];
      if (n_states >= A.C) {  // (the map is kept at most half full: C <= (mask + 1) / 2)
        overflow = true;
        break;
      }
      bool ins;
      const uint32_t dst = find_or_insert(nio, ds, ins);
      if (ins) {
        f.pending = id;
        stack[sp - 1] = f;
        f = BFrame{n_states, nio, ds, (nio == goal_io && ds == goal_s) ? 0u : 0x100u, 0u, 0u, 0u, 0u};
        dead[n_states] = 0;
        ++n_states;
        ++sp;  // (sp <= n_states <= C)
        sp_max = max(sp_max, sp);
      } else if (!dead[dst]) {  // includes states still on the stack (their flag is 0 until they are finished)
        if (n_kept < A.K) kept[n_kept] = BKept{f.id, dst, id};
        else overflow = true;
        ++n_kept;
        f.phase_dead &= 7u;
      }
    }
    if (!A.fill) {
      atomicMax(A.peak + 0, n_states);
      atomicMax(A.peak + 1, n_kept);
      atomicMax(A.peak + 2, sp_max);
    }
    if (overflow) {
      A.status[e] = B_OVERFLOW;
      continue;
    }
    // the goal state
    uint32_t fin = 0xFFFFFFFFu;
    {
      uint32_t h = b_hash(goal_io, goal_s) & A.mask;
      for (;;) {
        const BSlot sl = map[h];
        if (sl.gen != gen) break;
        if (sl.io == goal_io && sl.s == goal_s) {
          fin = sl.id;
          break;
        }
        h = (h + 1) & A.mask;
      }
    }
    if (fin == 0xFFFFFFFFu || dead[fin] || dead[0]) {
      A.status[e] = B_DROPPED;
      if (!A.fill) atomicAdd(A.pre_arcs, pre);
      continue;
    }
    // prune: drop dead states keeping the order, drop arcs into them (derivations.h:612-628)
    uint32_t n_live = 0;
    for (uint32_t s = 0; s < n_states; ++s) renum[s] = dead[s] ? 0xFFFFFFFFu : n_live++;
    for (uint32_t s = 0; s <= n_live; ++s) cnt[s] = 0;
    for (uint32_t k = 0; k < n_kept; ++k) {
      const BKept a = kept[k];
      if (renum[a.src] != 0xFFFFFFFFu && renum[a.dst] != 0xFFFFFFFFu) ++cnt[renum[a.src] + 1];
    }
    for (uint32_t s = 0; s < n_live; ++s) cnt[s + 1] += cnt[s];
    if (!A.fill) {
      A.status[e] = B_OK;
      A.n_states[e] = n_live;
      A.n_arcs[e] = cnt[n_live];
      A.fin[e] = renum[fin];
      atomicAdd(A.pre_arcs, pre);
      continue;
    }
    uint32_t* __restrict__ off = A.out_off + A.row_base[e];
    uint32_t* __restrict__ adst = A.out_dst + A.arc_base[e];
    uint32_t* __restrict__ aid = A.out_id + A.arc_base[e];
    for (uint32_t s = 0; s <= n_live; ++s) off[s] = cnt[s];
    // stored list order = reverse keep order: every row is filled from its end
    for (uint32_t s = 0; s < n_live; ++s) cnt[s] = cnt[s + 1];
    for (uint32_t k = 0; k < n_kept; ++k) {
      const BKept a = kept[k];
      const uint32_t rs = renum[a.src], rd = renum[a.dst];
      if (rs == 0xFFFFFFFFu || rd == 0xFFFFFFFFu) continue;
      const uint32_t pos = --cnt[rs];
      adst[pos] = rd;
      aid[pos] = a.id;
    }
  }
}

struct BuiltOwner {
  std::vector<uint32_t> ex_states, ex_fin, arc_off, arc_dst, arc_id, kept_example, dropped;
  std::vector<double> ex_weight;
  cml_built_trellises pub;
};

template <typename T>
struct ScopedDev {  // cudaMalloc'ed scratch, released on scope exit
  T* p = nullptr;
  cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, std::max<size_t>(1, n) * sizeof(T)); }
  ~ScopedDev() {
    if (p) cudaFree(p);
  }
};

}  // namespace

extern "C" int cml_build_trellises(cml_ctx* ctx, const cml_wfst_view* x, const cml_corpus_view* c, cml_built_trellises** out) {
  if (!ctx || !x || !c || !out) return CML_ERR_ARG;
  *out = nullptr;
  CML_REQUIRE(x->n_states >= 1 && x->state_arc_off && (x->n_arcs == 0 || (x->arc_in && x->arc_out && x->arc_dest)), CML_ERR_ARG,
              "cml_build_trellises: incomplete transducer view");
  CML_REQUIRE(x->final_state < x->n_states, CML_ERR_ARG, "cml_build_trellises: final state out of range");
  CML_REQUIRE(c->n_ex == 0 || (c->in_off && c->out_off), CML_ERR_ARG, "cml_build_trellises: incomplete corpus view");
  CML_REQUIRE(c->n_ex < 0xFFFFFFFFull, CML_ERR_ARG, "cml_build_trellises: too many examples");
  const auto t_start = std::chrono::steady_clock::now();
  CML_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint32_t Q = x->n_states;
  const uint64_t n_arcs = x->n_arcs;
  const uint64_t n_ex = c->n_ex;
  CML_REQUIRE(x->state_arc_off[0] == 0 && x->state_arc_off[Q] == n_arcs, CML_ERR_ARG,
              "cml_build_trellises: state_arc_off must run from 0 to n_arcs");
  for (uint32_t s = 0; s < Q; ++s)
    CML_REQUIRE(x->state_arc_off[s] <= x->state_arc_off[s + 1], CML_ERR_ARG, "cml_build_trellises: state_arc_off not monotone");
  for (uint64_t a = 0; a < n_arcs; ++a)
    CML_REQUIRE(x->arc_dest[a] < Q, CML_ERR_ARG, "cml_build_trellises: arc destination out of range");
  for (uint64_t e = 0; e < n_ex; ++e)
    CML_REQUIRE(c->in_off[e] <= c->in_off[e + 1] && c->out_off[e] <= c->out_off[e + 1], CML_ERR_ARG,
                "cml_build_trellises: corpus offsets not monotone");
  // ---- (in,out)-label index of the transducer: per state the arc-table ids grouped by label pair, groups in key order,
  // ids of a group in arc-table order
  std::vector<uint32_t> state_off(Q + 1, 0), range_begin, ids;
  std::vector<uint64_t> keys;
  {
    std::vector<std::pair<uint64_t, uint32_t>> tmp;
    ids.reserve(n_arcs);
    for (uint32_t s = 0; s < Q; ++s) {
      tmp.clear();
      for (uint32_t a = x->state_arc_off[s]; a < x->state_arc_off[s + 1]; ++a)
        tmp.emplace_back(((uint64_t)x->arc_in[a] << 32) | x->arc_out[a], a);
      std::stable_sort(tmp.begin(), tmp.end(), [](auto const& p, auto const& q) { return p.first < q.first; });
      for (size_t i = 0; i < tmp.size(); ++i) {
        if (i == 0 || tmp[i].first != tmp[i - 1].first) {
          keys.push_back(tmp[i].first);
          range_begin.push_back((uint32_t)ids.size());
        }
        ids.push_back(tmp[i].second);
      }
      state_off[s + 1] = (uint32_t)keys.size();
    }
    range_begin.push_back((uint32_t)ids.size());
  }
  DevArray<uint32_t> d_state_off, d_range_begin, d_ids, d_dest, d_in_sym, d_out_sym, d_todo;
  DevArray<uint64_t> d_keys, d_in_off, d_out_off;
  CML_CUDA(d_state_off.upload(state_off.data(), state_off.size(), st));
  CML_CUDA(d_keys.upload(keys.data(), keys.size(), st));
  CML_CUDA(d_range_begin.upload(range_begin.data(), range_begin.size(), st));
  CML_CUDA(d_ids.upload(ids.data(), ids.size(), st));
  CML_CUDA(d_dest.upload(x->arc_dest, n_arcs, st));
  const uint64_t n_in = n_ex ? c->in_off[n_ex] : 0, n_out = n_ex ? c->out_off[n_ex] : 0;
  static const uint64_t zero1[1] = {0};
  CML_CUDA(d_in_off.upload(n_ex ? c->in_off : zero1, n_ex + 1, st));
  CML_CUDA(d_out_off.upload(n_ex ? c->out_off : zero1, n_ex + 1, st));
  CML_CUDA(d_in_sym.upload(c->in_sym, n_in, st));
  CML_CUDA(d_out_sym.upload(c->out_sym, n_out, st));
  DevArray<uint8_t> d_status;
  DevArray<uint32_t> d_nst, d_narc, d_fin, d_peak;
  DevArray<unsigned long long> d_next, d_pre;
  CML_CUDA(d_status.alloc(n_ex + 1));
  CML_CUDA(d_nst.alloc(n_ex + 1));
  CML_CUDA(d_narc.alloc(n_ex + 1));
  CML_CUDA(d_fin.alloc(n_ex + 1));
  CML_CUDA(d_peak.alloc(3));
  CML_CUDA(d_next.alloc(1));
  CML_CUDA(d_pre.alloc(1));
  CML_CUDA(cudaMemsetAsync(d_peak.p, 0, 3 * sizeof(uint32_t), st));
  CML_CUDA(cudaMemsetAsync(d_pre.p, 0, sizeof(unsigned long long), st));
  CML_CUDA(cudaMemsetAsync(d_status.p, B_OVERFLOW, n_ex + 1, st));

  size_t budget = (size_t)8 << 30;  // scratch bytes of one launch
  if (const char* e = getenv("CML_BUILD_SCRATCH_MB")) budget = (size_t)std::strtoull(e, nullptr, 10) << 20;
  const uint32_t max_workers = (uint32_t)ctx->sm_count * 2048u;
  auto per_worker = [](uint64_t C, uint64_t K, uint64_t M) {
    return M * sizeof(BSlot) + C + C * sizeof(BFrame) + K * sizeof(BKept) + C * 4 + (C + 1) * 4;
  };
  // one launch over `todo` with capacities (C, K); fill = pass 2
  auto launch = [&](std::vector<uint32_t> const& todo, uint64_t C, uint64_t K, int fill, const uint64_t* row_base, const uint64_t* arc_base,
                    uint32_t* o_off, uint32_t* o_dst, uint32_t* o_id) -> int {
    uint64_t M = 64;
    while (M < 2 * C) M <<= 1;
    uint64_t W = std::min<uint64_t>(std::min<uint64_t>(todo.size(), max_workers), std::max<uint64_t>(1, budget / per_worker(C, K, M)));
    W = (W + 127) / 128 * 128;
    ScopedDev<BSlot> map;
    ScopedDev<uint8_t> dead;
    ScopedDev<BFrame> stack;
    ScopedDev<BKept> kept;
    ScopedDev<uint32_t> renum, cnt;
    CML_CUDA(map.alloc(W * M));
    CML_CUDA(dead.alloc(W * C));
    CML_CUDA(stack.alloc(W * C));
    CML_CUDA(kept.alloc(W * K));
    CML_CUDA(renum.alloc(W * C));
    CML_CUDA(cnt.alloc(W * (C + 1)));
    CML_CUDA(cudaMemsetAsync(map.p, 0, W * M * sizeof(BSlot), st));
    CML_CUDA(d_todo.upload(todo.data(), todo.size(), st));
    CML_CUDA(cudaMemsetAsync(d_next.p, 0, sizeof(unsigned long long), st));
    BArgs A{};
    A.io = BIo{d_state_off.p, d_keys.p, d_range_begin.p, d_ids.p, d_dest.p, x->final_state};
    A.c = BCorpus{d_in_off.p, d_in_sym.p, d_out_off.p, d_out_sym.p};
    A.todo = d_todo.p;
    A.n_todo = (uint32_t)todo.size();
    A.C = (uint32_t)C;
    A.K = (uint32_t)K;
    A.mask = (uint32_t)(M - 1);
    A.map = map.p;
    A.dead = dead.p;
    A.stack = stack.p;
    A.kept = kept.p;
    A.renum = renum.p;
    A.cnt = cnt.p;
    A.next = d_next.p;
    A.status = d_status.p;
    A.n_states = d_nst.p;
    A.n_arcs = d_narc.p;
    A.fin = d_fin.p;
    A.peak = d_peak.p;
    A.pre_arcs = d_pre.p;
    A.fill = fill;
    A.row_base = row_base;
    A.arc_base = arc_base;
    A.out_off = o_off;
    A.out_dst = o_dst;
    A.out_id = o_id;
    k_build_trellis<<<(unsigned)(W / 128), 128, 0, st>>>(A);
    ++ctx->launches;
    CML_CUDA(cudaGetLastError());
    CML_CUDA(cudaStreamSynchronize(st));  // (the scratch is released on return)
    return CML_OK;
  };

  // ---- pass 1: sizes.  First capacity guess from the longest example; overflowing examples go to the next round.
  std::vector<uint8_t> status(n_ex + 1, B_OVERFLOW);
  std::vector<uint32_t> todo(n_ex);
  std::iota(todo.begin(), todo.end(), 0u);
  uint64_t max_len = 0;
  for (uint64_t e = 0; e < n_ex; ++e)
    max_len = std::max<uint64_t>(max_len, (c->in_off[e + 1] - c->in_off[e]) + (c->out_off[e + 1] - c->out_off[e]));
  uint64_t C = std::max<uint64_t>(64, (max_len + 2) * std::min<uint32_t>(Q, 8u));
  if (const char* e = getenv("CML_BUILD_FIRST_STATES")) C = std::max<uint64_t>(2, std::strtoull(e, nullptr, 10));
  uint64_t K = 8 * C;
  uint32_t rounds = 0;
  while (!todo.empty()) {
    CML_REQUIRE(per_worker(C, K, 4 * C) <= budget && C < (1ull << 30), CML_ERR_ARG,
                "cml_build_trellises: an example's lattice does not fit the scratch budget (CML_BUILD_SCRATCH_MB)");
    const int rc = launch(todo, C, K, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    ++rounds;
    CML_CUDA(cudaMemcpy(status.data(), d_status.p, n_ex, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> again;
    for (uint32_t e : todo) {
      CML_REQUIRE(status[e] != B_TOOLONG, CML_ERR_ARG, "cml_build_trellises: a string is longer than 65,534 symbols");
      if (status[e] == B_OVERFLOW) again.push_back(e);
    }
    todo.swap(again);
    // the overflowing walks stopped early: their peaks are lower bounds, so grow geometrically
    C *= 4;
    K *= 4;
  }
  std::unique_ptr<BuiltOwner> own(new BuiltOwner());
  std::vector<uint32_t> nst(n_ex), narc(n_ex), fin(n_ex);
  uint32_t peak[3] = {0, 0, 0};
  unsigned long long pre = 0;
  if (n_ex) {
    CML_CUDA(cudaMemcpy(nst.data(), d_nst.p, n_ex * 4, cudaMemcpyDeviceToHost));
    CML_CUDA(cudaMemcpy(narc.data(), d_narc.p, n_ex * 4, cudaMemcpyDeviceToHost));
    CML_CUDA(cudaMemcpy(fin.data(), d_fin.p, n_ex * 4, cudaMemcpyDeviceToHost));
  }
  CML_CUDA(cudaMemcpy(peak, d_peak.p, sizeof(peak), cudaMemcpyDeviceToHost));
  CML_CUDA(cudaMemcpy(&pre, d_pre.p, sizeof(pre), cudaMemcpyDeviceToHost));
  std::vector<uint64_t> row_base(n_ex + 1, 0), arc_base(n_ex + 1, 0);
  uint64_t rows = 0, arcs = 0;
  for (uint64_t e = 0; e < n_ex; ++e) {
    row_base[e] = rows;
    arc_base[e] = arcs;
    if (status[e] == B_OK) {
      own->kept_example.push_back((uint32_t)e);
      own->ex_states.push_back(nst[e]);
      own->ex_fin.push_back(fin[e]);
      own->ex_weight.push_back(c->weight ? c->weight[e] : 1.0);
      rows += (uint64_t)nst[e] + 1;
      arcs += narc[e];
    } else
      own->dropped.push_back((uint32_t)e);
  }
  // ---- pass 2: the kept examples again, rows and arcs written at their final places
  own->arc_off.resize(rows);
  own->arc_dst.resize(arcs);
  own->arc_id.resize(arcs);
  if (!own->kept_example.empty()) {
    DevArray<uint64_t> d_row_base, d_arc_base;
    DevArray<uint32_t> d_off, d_dst, d_id;
    CML_CUDA(d_row_base.upload(row_base.data(), n_ex + 1, st));
    CML_CUDA(d_arc_base.upload(arc_base.data(), n_ex + 1, st));
    CML_CUDA(d_off.alloc(rows));
    CML_CUDA(d_dst.alloc(std::max<uint64_t>(1, arcs)));
    CML_CUDA(d_id.alloc(std::max<uint64_t>(1, arcs)));
    const int rc = launch(own->kept_example, (uint64_t)peak[0] + 1, std::max<uint64_t>(1, peak[1]), 1, d_row_base.p, d_arc_base.p, d_off.p,
                          d_dst.p, d_id.p);
    if (rc) return rc;
    ++rounds;
    CML_CUDA(cudaMemcpy(own->arc_off.data(), d_off.p, rows * 4, cudaMemcpyDeviceToHost));
    if (arcs) {
      CML_CUDA(cudaMemcpy(own->arc_dst.data(), d_dst.p, arcs * 4, cudaMemcpyDeviceToHost));
      CML_CUDA(cudaMemcpy(own->arc_id.data(), d_id.p, arcs * 4, cudaMemcpyDeviceToHost));
    }
  }
  cml_built_trellises& P = own->pub;
  std::memset(&P, 0, sizeof(P));
  P.batch.n_ex = own->kept_example.size();
  P.batch.ex_states = own->ex_states.data();
  P.batch.ex_fin = own->ex_fin.data();
  P.batch.ex_weight = own->ex_weight.data();
  P.batch.arc_off = own->arc_off.data();
  P.batch.arc_dst = own->arc_dst.data();
  P.batch.arc_id = own->arc_id.data();
  P.kept_example = own->kept_example.data();
  P.n_dropped = own->dropped.size();
  P.dropped = own->dropped.data();
  P.pre_arcs = pre;
  P.launches = rounds;
  P.peak_states = peak[0];
  P.peak_kept_arcs = peak[1];
  P.peak_depth = peak[2];
  P.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  P.owner = own.get();
  *out = &own.release()->pub;
  return CML_OK;
}

extern "C" void cml_free_built_trellises(cml_built_trellises* b) {
  if (b && b->owner) delete static_cast<BuiltOwner*>(b->owner);
}
