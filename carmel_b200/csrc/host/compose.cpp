// compose.cpp -- WFST composition with the 3-state epsilon filter, recording for every composed arc
// the chain of original parameters whose product is its weight (host side).
//
// Behavioural contract taken from the reference (carmel/src/compose.cc:316-531,
// carmel/src/cascade.h:489-599), restated on flat arrays:
//   * composed states are numbered in discovery order; the work list is LIFO;
//   * a composed state's arcs end up in reverse creation order (the reference pushes to the front);
//   * per source pair, arcs are created by walking the SMALLER state's arc list and looking matches
//     up in a label index of the larger one (index lists hold arcs in reverse arc order) when the
//     larger state has more than `index_threshold` arcs, else by a double loop;
//   * filter state 0/1/2 semantics of the epsilon filter; several reachable final triples get a
//     fresh final state reached by locked *e*:*e* arcs of weight one.
// Parameters are identified by integer ids (member base + arc-table index) instead of pointers.
#include <algorithm>

#include "carmel_host.hpp"

namespace cb {

void Cascade::number_members() {
  member_param_base.clear();
  member_state_base.clear();
  uint32_t base = 0;
  for (Wfst* w : members) {
    member_param_base.push_back(base);
    std::vector<uint32_t> off;
    w->arc_offsets(off);
    member_state_base.push_back(off);
    base += off.back();
  }
  n_params = base;
}

namespace {

struct ArcRef {  // one arc of an operand plus what is needed to build chains
  const Arc* arc;
  uint32_t param;  // parameter id (original member) -- unused when the operand is itself composed
};

struct Composer {
  Cascade& c;
  Wfst &a, &b;
  bool a_is_composed;
  uint32_t a_member, b_member;
  std::unordered_map<uint32_t, uint32_t> eps_chain;  // parameter id -> chain id of its singleton chain

  Composer(Cascade& c, Wfst& a, Wfst& b, bool ac, uint32_t am, uint32_t bm)
      : c(c), a(a), b(b), a_is_composed(ac), a_member(am), b_member(bm) {}

  uint32_t param_a(uint32_t s, uint32_t k) const { return c.member_param_base[a_member] + c.member_state_base[a_member][s] + k; }
  uint32_t param_b(uint32_t s, uint32_t k) const { return c.member_param_base[b_member] + c.member_state_base[b_member][s] + k; }
  static bool locked_one(Arc const& e) { return e.group == kLocked && e.ln_w == 0.; }  // cascade.h:569

  // chain for an arc taken alone (epsilon moves): cascade.h:573-586
  uint32_t record_single(Arc const& e, uint32_t param, bool operand_is_composed) {
    if (c.trivial) return e.group;
    if (operand_is_composed) return locked_one(e) ? 0u : e.group;
    auto it = eps_chain.find(param);
    if (it != eps_chain.end()) return it->second;
    uint32_t id = 0;
    if (!locked_one(e)) {
      id = (uint32_t)c.chains.size();
      c.chains.push_back({param});
    }
    eps_chain.emplace(param, id);
    return id;
  }
  // chain for a matched pair: cascade.h:544-559,588-599
  uint32_t record_pair(Arc const& l, uint32_t pl, Arc const& r, uint32_t pr) {
    if (c.trivial) return kNoGroup;
    std::vector<uint32_t> v;
    if (a_is_composed) {
      if (!locked_one(r)) v.push_back(pr);
      auto const& ca = c.chains[l.group];
      v.insert(v.end(), ca.begin(), ca.end());
    } else {
      if (!locked_one(l)) v.push_back(pl);
      if (!locked_one(r)) v.push_back(pr);
    }
    if (v.empty()) return 0;
    c.chains.push_back(std::move(v));
    return (uint32_t)c.chains.size() - 1;
  }
};

}  // namespace

std::unique_ptr<Wfst> compose(Cascade& c, Wfst& a, Wfst& b, bool a_is_composed, uint32_t a_member, uint32_t b_member,
                              uint32_t index_threshold) {
  std::unique_ptr<Wfst> rp(new Wfst());
  Wfst& r = *rp;
  r.alph[0] = a.alph[0];
  r.alph[1] = b.alph[1];
  r.named = false;
  if (!(a.valid && b.valid)) return rp;
  Composer rec(c, a, b, a_is_composed, a_member, b_member);

  // symbol maps across the interface alphabet (a's output <-> b's input), matched by spelling
  Alphabet &aout = *a.alph[1], &bin = *b.alph[0];
  const uint32_t NOMATCH = 0xFFFFFFFEu;
  std::vector<uint32_t> a2b(aout.names.size()), b2a(bin.names.size());
  for (size_t i = 0; i < a2b.size(); ++i) {
    const int j = bin.find(aout.names[i]);
    a2b[i] = j < 0 ? NOMATCH : (uint32_t)j;
  }
  for (size_t i = 0; i < b2a.size(); ++i) {
    const int j = aout.find(bin.names[i]);
    b2a[i] = j < 0 ? NOMATCH : (uint32_t)j;
  }

  const uint64_t nA = a.num_states(), nB = b.num_states();
  auto key = [&](uint32_t qa, uint32_t qb, uint32_t f) { return ((uint64_t)f * nA + qa) * nB + qb; };
  std::unordered_map<uint64_t, uint32_t> state_of;
  struct Work {
    uint32_t id, qa, qb, f;
  };
  std::vector<Work> stack;
  state_of.emplace(key(0, 0, 0), 0);
  r.states.emplace_back();
  stack.push_back({0, 0, 0, 0});

  // label index of one state: arc positions grouped by label, each group in REVERSE arc order
  struct LabelIndex {
    std::vector<uint32_t> order;  // arc positions sorted by (label, descending position)
    std::vector<std::pair<uint32_t, std::pair<uint32_t, uint32_t>>> groups;  // label -> [begin,end) in order
    bool built = false;
    std::pair<uint32_t, uint32_t> find(uint32_t label) const {
      auto it = std::lower_bound(groups.begin(), groups.end(), label,
                                 [](auto const& g, uint32_t l) { return g.first < l; });
      if (it == groups.end() || it->first != label) return {0, 0};
      return it->second;
    }
  };
  std::vector<LabelIndex> idxA(nA), idxB(nB);
  auto build_index = [](std::vector<Arc> const& arcs, bool by_out, LabelIndex& ix) {
    if (ix.built) return;
    ix.built = true;
    ix.order.resize(arcs.size());
    for (uint32_t i = 0; i < arcs.size(); ++i) ix.order[i] = i;
    std::sort(ix.order.begin(), ix.order.end(), [&](uint32_t x, uint32_t y) {
      const uint32_t lx = by_out ? arcs[x].out : arcs[x].in, ly = by_out ? arcs[y].out : arcs[y].in;
      return lx != ly ? lx < ly : x > y;
    });
    for (uint32_t i = 0; i < ix.order.size();) {
      const uint32_t l = by_out ? arcs[ix.order[i]].out : arcs[ix.order[i]].in;
      uint32_t j = i;
      while (j < ix.order.size() && (by_out ? arcs[ix.order[j]].out : arcs[ix.order[j]].in) == l) ++j;
      ix.groups.push_back({l, {i, j}});
      i = j;
    }
  };

  std::vector<Arc> created;  // arcs of the current source state in creation order
  auto emit = [&](uint32_t in, uint32_t out, uint32_t qa, uint32_t qb, uint32_t f, double ln_w, uint32_t group) {
    auto ins = state_of.emplace(key(qa, qb, f), r.num_states());
    if (ins.second) {
      stack.push_back({ins.first->second, qa, qb, f});
      r.states.emplace_back();
    }
    created.push_back(Arc{in, out, ins.first->second, ln_w, group});
  };

  while (!stack.empty()) {
    const Work w = stack.back();
    stack.pop_back();
    auto const &qa = a.states[w.qa], &qb = b.states[w.qb];
    created.clear();
    const bool a_larger = qa.size() > qb.size();
    const size_t larger = a_larger ? qa.size() : qb.size();
    if (larger > index_threshold && !a_larger) {  // walk a's arcs, look up in b's input index
      LabelIndex& ib = idxB[w.qb];
      build_index(qb, false, ib);
      const auto beps = ib.find(kEps);
      for (uint32_t i = 0; i < qa.size(); ++i) {
        Arc const& l = qa[i];
        const uint32_t pl = a_is_composed ? 0 : rec.param_a(w.qa, i);
        if (l.out == kEps) {
          if (w.f != 2) emit(l.in, kEps, l.dest, w.qb, 1, l.ln_w, rec.record_single(l, pl, a_is_composed));
          if (w.f == 0)
            for (uint32_t k = beps.first; k < beps.second; ++k) {
              const uint32_t j = ib.order[k];
              emit(l.in, qb[j].out, l.dest, qb[j].dest, 0, l.ln_w + qb[j].ln_w, rec.record_pair(l, pl, qb[j], rec.param_b(w.qb, j)));
            }
        } else {
          const auto m = ib.find(a2b[l.out]);
          for (uint32_t k = m.first; k < m.second; ++k) {
            const uint32_t j = ib.order[k];
            emit(l.in, qb[j].out, l.dest, qb[j].dest, 0, l.ln_w + qb[j].ln_w, rec.record_pair(l, pl, qb[j], rec.param_b(w.qb, j)));
          }
        }
      }
      if (w.f != 1)
        for (uint32_t k = beps.first; k < beps.second; ++k) {
          const uint32_t j = ib.order[k];
          emit(kEps, qb[j].out, w.qa, qb[j].dest, 2, qb[j].ln_w, rec.record_single(qb[j], rec.param_b(w.qb, j), false));
        }
    } else if (larger > index_threshold) {  // walk b's arcs, look up in a's output index
      LabelIndex& ia = idxA[w.qa];
      build_index(qa, true, ia);
      const auto aeps = ia.find(kEps);
      for (uint32_t j = 0; j < qb.size(); ++j) {
        Arc const& rr = qb[j];
        const uint32_t pr = rec.param_b(w.qb, j);
        if (rr.in == kEps) {
          if (w.f != 1) emit(kEps, rr.out, w.qa, rr.dest, 2, rr.ln_w, rec.record_single(rr, pr, false));
          if (w.f == 0)
            for (uint32_t k = aeps.first; k < aeps.second; ++k) {
              const uint32_t i = ia.order[k];
              emit(qa[i].in, rr.out, qa[i].dest, rr.dest, 0, qa[i].ln_w + rr.ln_w,
                   rec.record_pair(qa[i], a_is_composed ? 0 : rec.param_a(w.qa, i), rr, pr));
            }
        } else {
          const uint32_t want = rr.in < b2a.size() ? b2a[rr.in] : NOMATCH;
          const auto m = ia.find(want);
          for (uint32_t k = m.first; k < m.second; ++k) {
            const uint32_t i = ia.order[k];
            emit(qa[i].in, rr.out, qa[i].dest, rr.dest, 0, qa[i].ln_w + rr.ln_w,
                 rec.record_pair(qa[i], a_is_composed ? 0 : rec.param_a(w.qa, i), rr, pr));
          }
        }
      }
      if (w.f != 2)
        for (uint32_t k = aeps.first; k < aeps.second; ++k) {
          const uint32_t i = ia.order[k];
          emit(qa[i].in, kEps, qa[i].dest, w.qb, 1, qa[i].ln_w,
               rec.record_single(qa[i], a_is_composed ? 0 : rec.param_a(w.qa, i), a_is_composed));
        }
    } else {  // both small: plain double loop
      for (uint32_t i = 0; i < qa.size(); ++i) {
        Arc const& l = qa[i];
        const uint32_t pl = a_is_composed ? 0 : rec.param_a(w.qa, i);
        if (l.out == kEps) {
          if (w.f != 2) emit(l.in, kEps, l.dest, w.qb, 1, l.ln_w, rec.record_single(l, pl, a_is_composed));
          if (w.f == 0)
            for (uint32_t j = 0; j < qb.size(); ++j)
              if (qb[j].in == kEps)
                emit(l.in, qb[j].out, l.dest, qb[j].dest, 0, l.ln_w + qb[j].ln_w, rec.record_pair(l, pl, qb[j], rec.param_b(w.qb, j)));
        } else {
          const uint32_t want = a2b[l.out];
          for (uint32_t j = 0; j < qb.size(); ++j)
            if (qb[j].in == want)
              emit(l.in, qb[j].out, l.dest, qb[j].dest, 0, l.ln_w + qb[j].ln_w, rec.record_pair(l, pl, qb[j], rec.param_b(w.qb, j)));
        }
      }
      if (w.f != 1)
        for (uint32_t j = 0; j < qb.size(); ++j)
          if (qb[j].in == kEps)
            emit(kEps, qb[j].out, w.qa, qb[j].dest, 2, qb[j].ln_w, rec.record_single(qb[j], rec.param_b(w.qb, j), false));
    }
    r.states[w.id].assign(created.rbegin(), created.rend());
  }

  // final states (compose.cc:503-528)
  uint32_t found[3], n_found = 0;
  for (uint32_t f = 0; f < 3; ++f) {
    auto it = state_of.find(key(a.final_state, b.final_state, f));
    if (it != state_of.end()) {
      found[n_found++] = it->second;
      r.final_state = it->second;
    }
  }
  if (!n_found) return rp;
  if (n_found > 1) {
    r.final_state = r.num_states();
    r.states.emplace_back();
    for (uint32_t i = 0; i < n_found; ++i) {
      auto& st = r.states[found[i]];
      st.insert(st.begin(), Arc{kEps, kEps, r.final_state, 0., c.trivial ? kLocked : 0u});
    }
  }
  r.valid = true;
  return rp;
}

}  // namespace cb
