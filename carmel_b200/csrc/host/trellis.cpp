// trellis.cpp -- per-example derivation lattice construction: input string x WFST x output string.
//
// Output contract = the reference's derivations::compute (carmel/src/derivations.h:479-513,640-704
// derive/add_arcs, :572-629 prune): states (i, s, o) numbered in DFS pre-order from (0,0,0); at every
// state labels are tried in the order (eps:eps), (eps:out[o]), (in[i]:eps), (in[i]:out[o]) and within
// a label the WFST arcs in arc-table order; an arc is kept unless its destination is already known
// dead when the arc is examined; a state's stored arc list is the reverse of the order its arcs were
// kept; dead states are then removed and the survivors renumbered in place.  This file reproduces
// that contract bit for bit (state ids, arc order, arc-table ids) with an explicit-stack DFS over
// flat arrays (no recursion depth limit, no per-arc allocation), parallel over examples.
#include <algorithm>
#include <atomic>
#include <ostream>
#include <thread>

#include "carmel_host.hpp"

namespace cb {

void TrellisBatch::clear() {
  ex_states.clear();
  ex_fin.clear();
  ex_weight.clear();
  arc_off.clear();
  arc_dst.clear();
  arc_id.clear();
  kept_example.clear();
  pre_arcs = 0;
}

// binary dump shared with the test-suite (same record layout as the CPU oracle's dump)
void TrellisBatch::dump(std::ostream& o, uint32_t n_arcs_table) const {
  auto u32 = [&](uint32_t v) { o.write((const char*)&v, 4); };
  u32((uint32_t)ex_states.size());
  u32(n_arcs_table);
  size_t row = 0, arc = 0;
  for (size_t e = 0; e < ex_states.size(); ++e) {
    const uint32_t n = ex_states[e];
    const uint32_t* off = &arc_off[row];
    u32(n);
    u32(off[n]);
    u32(ex_fin[e]);
    o.write((const char*)&ex_weight[e], 8);
    for (uint32_t s = 0; s < n; ++s) {
      u32(off[s + 1] - off[s]);
      for (uint32_t k = off[s]; k < off[s + 1]; ++k) {
        u32(arc_dst[arc + k]);
        u32(arc_id[arc + k]);
      }
    }
    row += n + 1;
    arc += off[n];
  }
}

namespace {

// (in,out)-label index of the transducer: per WFST state the arc-table ids grouped by label pair,
// each group in arc-table order (derivations.h:142-155 wfst_io_index)
struct IoIndex {
  std::vector<uint32_t> state_off;            // per state: range in keys/range arrays
  std::vector<uint64_t> keys;                 // sorted (in<<32|out) per state
  std::vector<uint32_t> range_begin;          // per key: begin in ids (end = next begin)
  std::vector<uint32_t> ids;                  // arc-table ids
  std::vector<uint32_t> dest;                 // destination WFST state per arc-table id
  explicit IoIndex(Wfst const& x) {
    const uint32_t n = x.num_states();
    state_off.assign(n + 1, 0);
    uint32_t id = 0;
    std::vector<std::pair<uint64_t, uint32_t>> tmp;
    for (uint32_t s = 0; s < n; ++s) {
      tmp.clear();
      for (Arc const& a : x.states[s]) {
        tmp.emplace_back(((uint64_t)a.in << 32) | a.out, id++);
        dest.push_back(a.dest);
      }
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
  // ids with label (in,out) leaving WFST state s
  inline void match(uint32_t s, uint32_t in, uint32_t out, const uint32_t*& b, const uint32_t*& e) const {
    const uint64_t k = ((uint64_t)in << 32) | out;
    const uint64_t* lo = keys.data() + state_off[s];
    const uint64_t* hi = keys.data() + state_off[s + 1];
    const uint64_t* it = std::lower_bound(lo, hi, k);
    if (it == hi || *it != k) {
      b = e = nullptr;
      return;
    }
    const size_t j = it - keys.data();
    b = ids.data() + range_begin[j];
    e = ids.data() + range_begin[j + 1];
  }
};

// open-addressing map (i,s,o) -> state id, reset per example by bumping a generation stamp
struct StateMap {
  struct Slot {
    uint32_t i, s, o, id, gen;
  };
  std::vector<Slot> slots;
  uint32_t mask = 0, gen = 0, used = 0;
  StateMap() { resize(1024); }
  void resize(uint32_t n) {
    slots.assign(n, Slot{0, 0, 0, 0, 0});
    mask = n - 1;
    gen = 1;
    used = 0;
  }
  void reset() {
    ++gen;
    used = 0;
    if (gen == 0xFFFFFFFFu) resize((uint32_t)slots.size());
  }
  static inline uint32_t hash(uint32_t i, uint32_t s, uint32_t o) {
    uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull ^ ((uint64_t)s + 0x632BE59BD9B4E019ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 31;
    h += (uint64_t)o * 0x165667B19E3779F9ull;
    h ^= h >> 29;
    return (uint32_t)(h ^ (h >> 32));
  }
  void grow() {
    std::vector<Slot> old;
    old.swap(slots);
    const uint32_t g = gen;
    resize((uint32_t)old.size() * 2);
    for (Slot const& sl : old)
      if (sl.gen == g) {
        bool ins;
        *find_or_insert(sl.i, sl.s, sl.o, sl.id, ins) = sl.id;
      }
  }
  // returns pointer to the id; inserted=true when the key was new (id set to new_id)
  uint32_t* find_or_insert(uint32_t i, uint32_t s, uint32_t o, uint32_t new_id, bool& inserted) {
    if (used * 2 >= slots.size()) grow();
    uint32_t h = hash(i, s, o) & mask;
    for (;;) {
      Slot& sl = slots[h];
      if (sl.gen != gen) {
        sl = Slot{i, s, o, new_id, gen};
        ++used;
        inserted = true;
        return &sl.id;
      }
      if (sl.i == i && sl.s == s && sl.o == o) {
        inserted = false;
        return &sl.id;
      }
      h = (h + 1) & mask;
    }
  }
  bool find(uint32_t i, uint32_t s, uint32_t o, uint32_t& id) const {
    uint32_t h = hash(i, s, o) & mask;
    for (;;) {
      Slot const& sl = slots[h];
      if (sl.gen != gen) return false;
      if (sl.i == i && sl.s == s && sl.o == o) {
        id = sl.id;
        return true;
      }
      h = (h + 1) & mask;
    }
  }
};

struct Builder {
  IoIndex const& io;
  Wfst const& x;
  StateMap map;
  struct Frame {
    uint32_t id, i, s, o;
    int phase;             // next label class to open (0..4)
    const uint32_t *cur, *end;
    uint32_t ni, no;       // lattice coordinates reached by the arcs of the open label class
    uint32_t pending_id;   // arc-table id of the arc whose destination is being derived
    bool dead;
  };
  std::vector<Frame> stack;
  std::vector<char> dead_state;
  struct Kept {
    uint32_t src, dst, id;
  };
  std::vector<Kept> kept;  // in keep order
  std::vector<uint32_t> renum, cnt;
  uint64_t pre_arcs = 0;

  Builder(IoIndex const& io, Wfst const& x) : io(io), x(x) {}

  // returns false when the goal is unreachable.  Appends the example to `out`.
  bool build(Example const& ex, TrellisBatch& out) {
    const uint32_t nin = (uint32_t)ex.in.size(), nout = (uint32_t)ex.out.size();
    const uint32_t gi = nin, gs = x.final_state, go = nout;
    map.reset();
    stack.clear();
    dead_state.clear();
    kept.clear();
    uint32_t n_states = 0;

    auto open_state = [&](uint32_t i, uint32_t s, uint32_t o) {
      stack.push_back(Frame{n_states++, i, s, o, 0, nullptr, nullptr, 0, 0, 0, !(i == gi && s == gs && o == go)});
      dead_state.push_back(0);
    };
    bool ins;
    map.find_or_insert(0, 0, 0, 0, ins);
    open_state(0, 0, 0);

    while (!stack.empty()) {
      Frame& f = stack.back();
      if (f.cur == f.end) {  // open the next label class (derivations.h:656-670)
        if (f.phase == 4) {  // all classes done: state is finished
          dead_state[f.id] = f.dead;
          const uint32_t done = f.id;
          stack.pop_back();
          if (!stack.empty()) {  // the parent was waiting for this destination (add_arcs :691-701)
            Frame& p = stack.back();
            if (!dead_state[done]) {
              kept.push_back(Kept{p.id, done, p.pending_id});
              p.dead = false;
            }
          }
          continue;
        }
        const int ph = f.phase++;
        const bool useO = f.o < nout, useI = f.i < nin;
        uint32_t lin = kEps, lout = kEps;
        f.ni = f.i;
        f.no = f.o;
        bool active = true;
        switch (ph) {
          case 0: break;
          case 1:
            active = useO;
            if (active) {
              lout = ex.out[f.o];
              f.no = f.o + 1;
            }
            break;
          case 2:
            active = useI;
            if (active) {
              lin = ex.in[f.i];
              f.ni = f.i + 1;
            }
            break;
          default:
            active = useI && useO;
            if (active) {
              lin = ex.in[f.i];
              lout = ex.out[f.o];
              f.ni = f.i + 1;
              f.no = f.o + 1;
            }
        }
        if (active)
          io.match(f.s, lin, lout, f.cur, f.end);
        else
          f.cur = f.end = nullptr;
        continue;
      }
      // examine the next arc of the open class
      const uint32_t id = *f.cur++;
      ++pre_arcs;
      const uint32_t ds = io.dest[id];
      const uint32_t ni = f.ni, no = f.no, me = f.id;
      uint32_t* pid = map.find_or_insert(ni, ds, no, n_states, ins);
      if (ins) {
        f.pending_id = id;
        open_state(ni, ds, no);  // invalidates f
      } else {
        const uint32_t dst = *pid;
        if (!dead_state[dst]) {  // includes states still on the stack (dead_state is 0 until finished)
          kept.push_back(Kept{me, dst, id});
          stack.back().dead = false;
        }
      }
    }

    uint32_t fin;
    if (!map.find(gi, gs, go, fin)) return false;
    // prune: drop dead states keeping order, drop arcs into them (derivations.h:612-628)
    renum.assign(n_states, 0);
    uint32_t n_kept = 0;
    for (uint32_t s = 0; s < n_states; ++s) renum[s] = dead_state[s] ? ~0u : n_kept++;
    if (!~renum[fin] || !~renum[0]) return false;
    cnt.assign(n_kept + 1, 0);
    for (Kept const& k : kept)
      if (~renum[k.src] && ~renum[k.dst]) ++cnt[renum[k.src] + 1];
    for (uint32_t s = 0; s < n_kept; ++s) cnt[s + 1] += cnt[s];
    const size_t arc_base = out.arc_dst.size();
    const uint32_t n_arcs = cnt[n_kept];
    out.arc_dst.resize(arc_base + n_arcs);
    out.arc_id.resize(arc_base + n_arcs);
    out.arc_off.insert(out.arc_off.end(), cnt.begin(), cnt.end());
    // stored list order = reverse keep order: fill each row from its end
    for (uint32_t s = 0; s < n_kept; ++s) cnt[s] = cnt[s + 1];
    for (Kept const& k : kept) {
      if (!~renum[k.src] || !~renum[k.dst]) continue;
      const uint32_t pos = --cnt[renum[k.src]];
      out.arc_dst[arc_base + pos] = renum[k.dst];
      out.arc_id[arc_base + pos] = k.id;
    }
    out.ex_states.push_back(n_kept);
    out.ex_fin.push_back(renum[fin]);
    out.ex_weight.push_back(ex.weight);
    return true;
  }
};

}  // namespace

void build_trellises(Wfst const& x, Corpus const& corpus, TrellisBatch& out, std::vector<uint32_t>& dropped,
                     unsigned n_threads) {
  out.clear();
  dropped.clear();
  IoIndex io(x);
  const size_t n = corpus.examples.size();
  if (!n_threads) n_threads = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  n_threads = (unsigned)std::min<size_t>(n_threads, std::max<size_t>(1, n / 64));
  const size_t chunk = std::max<size_t>(64, (n + n_threads * 8 - 1) / (n_threads * 8));
  const size_t n_chunks = (n + chunk - 1) / chunk;
  std::vector<TrellisBatch> parts(n_chunks);
  std::vector<std::vector<uint32_t>> part_dropped(n_chunks);
  std::atomic<size_t> next{0};
  auto work = [&]() {
    Builder b(io, x);
    for (;;) {
      const size_t c = next.fetch_add(1);
      if (c >= n_chunks) break;
      for (size_t e = c * chunk; e < std::min(n, (c + 1) * chunk); ++e) {
        if (b.build(corpus.examples[e], parts[c]))
          parts[c].kept_example.push_back((uint32_t)e);
        else
          part_dropped[c].push_back((uint32_t)e);
      }
      parts[c].pre_arcs = b.pre_arcs;
      b.pre_arcs = 0;
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < n_threads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  size_t ns = 0, na = 0, no = 0;
  for (auto const& p : parts) {
    ns += p.ex_states.size();
    na += p.arc_dst.size();
    no += p.arc_off.size();
  }
  out.ex_states.reserve(ns);
  out.ex_fin.reserve(ns);
  out.ex_weight.reserve(ns);
  out.kept_example.reserve(ns);
  out.arc_off.reserve(no);
  out.arc_dst.reserve(na);
  out.arc_id.reserve(na);
  for (size_t c = 0; c < n_chunks; ++c) {
    auto& p = parts[c];
    out.ex_states.insert(out.ex_states.end(), p.ex_states.begin(), p.ex_states.end());
    out.ex_fin.insert(out.ex_fin.end(), p.ex_fin.begin(), p.ex_fin.end());
    out.ex_weight.insert(out.ex_weight.end(), p.ex_weight.begin(), p.ex_weight.end());
    out.kept_example.insert(out.kept_example.end(), p.kept_example.begin(), p.kept_example.end());
    out.arc_off.insert(out.arc_off.end(), p.arc_off.begin(), p.arc_off.end());
    out.arc_dst.insert(out.arc_dst.end(), p.arc_dst.begin(), p.arc_dst.end());
    out.arc_id.insert(out.arc_id.end(), p.arc_id.begin(), p.arc_id.end());
    out.pre_arcs += p.pre_arcs;
    dropped.insert(dropped.end(), part_dropped[c].begin(), part_dropped[c].end());
    p = TrellisBatch();
  }
}

// The same batch built on the GPU (cml_build_trellises, csrc/cml_build.cu): the transducer's arc table and the corpus are
// handed over as flat arrays, the result comes back in the form build_trellises produces (byte-identical dumps).
void build_trellises_device(cml_ctx* ctx, Wfst const& x, Corpus const& corpus, TrellisBatch& out, std::vector<uint32_t>& dropped,
                            double* seconds) {
  out.clear();
  dropped.clear();
  std::vector<uint32_t> soff, ain, aout, adest;
  x.arc_offsets(soff);
  const size_t na = x.num_arcs();
  ain.reserve(na);
  aout.reserve(na);
  adest.reserve(na);
  for (auto const& st : x.states)
    for (Arc const& a : st) {
      ain.push_back(a.in);
      aout.push_back(a.out);
      adest.push_back(a.dest);
    }
  const size_t n = corpus.examples.size();
  std::vector<uint64_t> in_off(n + 1, 0), out_off(n + 1, 0);
  std::vector<double> weight(n);
  for (size_t e = 0; e < n; ++e) {
    in_off[e + 1] = in_off[e] + corpus.examples[e].in.size();
    out_off[e + 1] = out_off[e] + corpus.examples[e].out.size();
    weight[e] = corpus.examples[e].weight;
  }
  std::vector<uint32_t> in_sym(in_off[n]), out_sym(out_off[n]);
  for (size_t e = 0; e < n; ++e) {
    std::copy(corpus.examples[e].in.begin(), corpus.examples[e].in.end(), in_sym.begin() + in_off[e]);
    std::copy(corpus.examples[e].out.begin(), corpus.examples[e].out.end(), out_sym.begin() + out_off[e]);
  }
  cml_wfst_view xv{};
  xv.n_states = x.num_states();
  xv.final_state = x.final_state;
  xv.n_arcs = na;
  xv.state_arc_off = soff.data();
  xv.arc_in = ain.data();
  xv.arc_out = aout.data();
  xv.arc_dest = adest.data();
  cml_corpus_view cv{};
  cv.n_ex = n;
  cv.in_off = in_off.data();
  cv.in_sym = in_sym.data();
  cv.out_off = out_off.data();
  cv.out_sym = out_sym.data();
  cv.weight = weight.data();
  cml_built_trellises* b = nullptr;
  if (cml_build_trellises(ctx, &xv, &cv, &b) != CML_OK || !b)
    throw std::runtime_error(std::string("cml_build_trellises: ") + cml_last_error(ctx));
  const cml_trellis_batch& t = b->batch;
  uint64_t rows = 0;
  for (uint64_t e = 0; e < t.n_ex; ++e) rows += (uint64_t)t.ex_states[e] + 1;
  const uint64_t arcs = rows ? 0 : 0;
  (void)arcs;
  out.ex_states.assign(t.ex_states, t.ex_states + t.n_ex);
  out.ex_fin.assign(t.ex_fin, t.ex_fin + t.n_ex);
  out.ex_weight.assign(t.ex_weight, t.ex_weight + t.n_ex);
  out.arc_off.assign(t.arc_off, t.arc_off + rows);
  uint64_t n_arcs = 0, row = 0;
  for (uint64_t e = 0; e < t.n_ex; ++e) {
    n_arcs += t.arc_off[row + t.ex_states[e]];
    row += (uint64_t)t.ex_states[e] + 1;
  }
  out.arc_dst.assign(t.arc_dst, t.arc_dst + n_arcs);
  out.arc_id.assign(t.arc_id, t.arc_id + n_arcs);
  out.kept_example.assign(b->kept_example, b->kept_example + t.n_ex);
  out.pre_arcs = b->pre_arcs;
  dropped.assign(b->dropped, b->dropped + b->n_dropped);
  if (seconds) *seconds = b->seconds;
  cml_free_built_trellises(b);
}

}  // namespace cb
