// fem.cpp -- carmel -> forest-em bridge on the host: --fem-forest / --fem-norm / --fem-param export and
// --load-fem-param import, so that carmel's derivation lattices can be trained by the forest kernels
// (forest-em-b200) and the two programs can be checked against each other on the same model.
//
// Reference: cascade_parameters::fem_deriv / fem_norms / print_params / read_params (carmel/src/cascade.h:85-202),
// backrefs (graehl/shared/graph.h:165-194), option handling carmel/src/carmel.cc:756-830.  A lattice state with two or
// more arcs is an OR node, an arc is an AND node over the parameters of its chain (1-based parameter ids = visit order
// over the cascade members) followed by the forest of its destination; a state reached more than once is written once
// as #k(...) and referenced as #k afterwards.
#include <fstream>
#include <iostream>
#include <sstream>

#include "carmel_host.hpp"

namespace cb {

// model arrays (parameters, chains, normalisation groups) -- host only, no GPU
void TrainJob::build_model() {
  if (!M.ln_w.empty()) return;
  using_cascade = !cascade.trivial;
  members = using_cascade ? cascade.members : std::vector<Wfst*>{x};
  for (size_t i = 0; i < members.size(); ++i) add_model_member(*members[i], i < methods.size() ? methods[i] : NormalizeMethod(), M);
  M.n_params = (uint32_t)M.ln_w.size();
  M.n_arcs = (uint32_t)x->num_arcs();
  if (using_cascade) {
    M.chain_off.push_back(0);
    for (auto const& st : x->states)
      for (Arc const& a : st) {
        auto const& ch = cascade.chains.at(a.group);
        M.chain_param.insert(M.chain_param.end(), ch.begin(), ch.end());
        M.chain_off.push_back((uint32_t)M.chain_param.size());
      }
  }
}

// --load-fem-param (cascade.h:183-202 read_params): one weight per cascade-member arc, visit order
void TrainJob::load_fem_param(std::string const& file) {
  std::ifstream in(file);
  if (!in) throw std::runtime_error("Missing --load-fem-param file.\n");
  std::vector<Wfst*> ms = cascade.trivial ? std::vector<Wfst*>{x} : cascade.members;
  std::string tok;
  for (Wfst* m : ms)
    for (auto& st : m->states)
      for (Arc& a : st) {
        double w;
        if (!(in >> tok) || !parse_weight(tok.c_str(), w))
          throw std::runtime_error("--load-fem-param file doesn't have enough params; make sure it was --fem-param saved for the same cascade");
        a.ln_w = w;
      }
}

namespace {

// one lattice in forest-em's text syntax (cascade.h:119-166), without recursion (a 100k-letter line nests 100k deep)
void fem_deriv(std::ostream& o, uint32_t n, const uint32_t* off, const uint32_t* dst, const uint32_t* id, uint32_t fin,
               ModelArrays const& M, bool using_cascade) {
  struct BR {
    uint32_t uses = 0, id = 0;
  };
  std::vector<BR> br(n);
  uint32_t nextid = 1;
  {  // backrefs(g, start): depth-first use counts (graph.h:165-194)
    std::vector<std::pair<uint32_t, uint32_t>> st;  // (state, next arc)
    auto use = [&](uint32_t s) -> bool {
      if (br[s].uses++ > 0) {
        br[s].id = nextid++;
        return false;
      }
      return true;
    };
    if (use(0)) st.emplace_back(0u, off[0]);
    while (!st.empty()) {
      auto& top = st.back();
      if (top.second == off[top.first + 1]) {
        st.pop_back();
        continue;
      }
      const uint32_t d = dst[top.second++];
      if (use(d)) st.emplace_back(d, off[d]);
    }
  }
  struct Frame {
    uint32_t s, k;       // state, next arc
    bool ornode, open;   // state is an OR node; the current arc's "(" is open
  };
  std::vector<Frame> st;
  auto enter = [&](uint32_t s) -> bool {  // prints the state's prefix; false = nothing more to print for it
    BR& b = br[s];
    if (b.uses > 1) {
      o << "#" << b.id;
      b.uses = 0;  // BACKREF_DEFINED
    } else if (b.uses == 0) {
      o << "#" << b.id;
      return false;
    }
    const bool ornode = off[s + 1] - off[s] >= 2;
    if (ornode) o << "(OR";
    st.push_back({s, off[s], ornode, false});
    return true;
  };
  std::vector<char> backdef(n, 0);
  for (uint32_t s = 0; s < n; ++s) backdef[s] = br[s].uses > 1;
  enter(0);
  while (!st.empty()) {
    Frame& f = st.back();
    if (f.open) {  // back from the destination of arc k-1
      o << ")";
      f.open = false;
    }
    if (f.k == off[f.s + 1]) {
      if (f.ornode) o << ")";
      st.pop_back();
      continue;
    }
    const uint32_t k = f.k++;
    if (f.ornode) o << " ";
    const uint32_t a = id[k], d = dst[k];
    const uint32_t p0 = using_cascade ? M.chain_off[a] : a, p1 = using_cascade ? M.chain_off[a + 1] : a + 1;
    const bool mid = d != fin;
    const bool nonleaf1 = backdef[f.s] || (p1 > p0 && (p1 - p0 > 1 || mid));
    if (nonleaf1) o << "(";
    bool sp = false;
    for (uint32_t c = p0; c < p1; ++c) {
      if (sp) o << ' ';
      sp = true;
      o << (using_cascade ? M.chain_param[c] : c) + 1;
    }
    bool descended = false;
    if (mid) {
      if (sp) o << ' ';
      const size_t depth = st.size();
      st[depth - 1].open = nonleaf1;  // (enter may reallocate st: do not touch f afterwards)
      descended = enter(d);
      if (!descended && nonleaf1) {
        o << ")";
        st[depth - 1].open = false;
      }
    } else if (nonleaf1)
      o << ")";
    (void)descended;
  }
  o << "\n";
}

}  // namespace

// --fem-param / --fem-norm (carmel.cc:810-830 fem_out, after training: with -M -1 the weights are the normalised input
// weights, which is how carmel/sample/decipher/to-fem.sh uses it; cascade.h:85-117,168-181)
void TrainJob::export_fem_tables(std::ostream& log) {
  build_model();
  if (lopt.count("fem-param")) {
    log << "Writing cascade weights to --fem-param=" << lopt["fem-param"] << std::endl;
    std::ofstream o(lopt["fem-param"]);
    for (Wfst* m : members)
      for (auto const& st : m->states)
        for (Arc const& a : st) o << format_weight(a.ln_w) << "\n";
  }
  if (lopt.count("fem-norm")) {
    log << "Writing forest-em normgroups to --fem-norm=" << lopt["fem-norm"] << std::endl;
    std::ofstream o(lopt["fem-norm"]);
    // groups in first-appearance order per member (fem_norms walks NormGroupIter per member; the member order of a
    // conditional group is that of a hash table in the reference: the groups are sets)
    std::vector<std::vector<uint32_t>> groups(M.n_groups);
    std::vector<uint32_t> order;
    o << "(";
    uint32_t p = 0;
    for (Wfst* m : members) {
      o << "\n";
      order.clear();
      for (auto const& st : m->states)
        for (size_t k = 0; k < st.size(); ++k, ++p) {
          const uint32_t g = M.param_group[p];
          if (g == kNoGroup) continue;
          if (groups[g].empty()) order.push_back(g);
          groups[g].push_back(p + 1);
        }
      for (uint32_t g : order) {
        o << '(';
        for (uint32_t q : groups[g]) o << ' ' << q;
        o << " )\n";
      }
    }
    o << ")\n";
  }
}

// --fem-forest (cascade.h:119-166; written while the derivations are computed, cached_derivs.h out_derivfile): one
// forest per example with a derivation, from the lattices prepare() has just built
void TrainJob::export_fem_forest(TrellisBatch const& tb, std::ostream& log) {
  build_model();
  log << "Writing forest-em derivation forests to --fem-forest=" << lopt["fem-forest"] << std::endl;
  std::ofstream o(lopt["fem-forest"]);
  uint64_t sb = 0, ab = 0;
  for (size_t e = 0; e < tb.ex_states.size(); ++e) {
    const uint32_t n = tb.ex_states[e];
    const uint32_t* off = tb.arc_off.data() + sb + e;
    fem_deriv(o, n, off, tb.arc_dst.data() + ab, tb.arc_id.data() + ab, tb.ex_fin[e], M, using_cascade);
    sb += n;
    ab += off[n];
  }
}

}  // namespace cb
