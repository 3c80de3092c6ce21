// oracle_cli.cpp -- command-line driver for the CPU ORACLE (test infrastructure, not the product).
// Accepts the subset of carmel's argv grammar (carmel/src/carmel.cc:929-1066) that the training
// hot path needs, runs the restated reference algorithm from carmel_oracle.hpp, and can dump
// per-example trellises / per-iteration history for the parity tests and time E-steps for the
// CPU baseline in bench.py.
//
//   carmel_oracle [-t] [-HJ] [-M n] [-e w] [-X w] [-f w] [-U] [-j|-u] [-:] [-F out] [--train-cascade]
//                 [--normby=JCN..] [--priors=a,b,..] [--dump-trellis=file] [--history=file]
//                 [--time-estimate=K] corpus wfst [wfst ...]
#include <functional>

#include "carmel_oracle.hpp"
#include "gibbs_oracle.hpp"
#include <chrono>
#include <sys/resource.h>
#include <unistd.h>

using namespace orc;

static void write_u32(std::ostream& o, uint32_t v) { o.write((char const*)&v, 4); }
static void write_f64(std::ostream& o, double v) { o.write((char const*)&v, 8); }

// Binary trellis dump (little endian), one record per kept example:
//   u32 n_states, u32 n_arcs, u32 fin, f64 weight, then per state (id order): u32 n_out,
//   then n_out x (u32 dest, u32 arc_id) in the stored list order (reverse of discovery order).
static void dump_trellis(std::ostream& o, Derivations const& d) {
  write_u32(o, (uint32_t)d.n_states());
  write_u32(o, (uint32_t)d.n_arcs());
  write_u32(o, d.fin);
  write_f64(o, d.weight);
  for (auto const& st : d.g) {
    write_u32(o, (uint32_t)st.size());
    for (auto const& a : st) {
      write_u32(o, a.dest);
      write_u32(o, a.id);
    }
  }
}

static std::vector<std::string> split(std::string const& s, char c) {
  std::vector<std::string> r;
  std::string cur;
  for (char ch : s) {
    if (ch == c) {
      r.push_back(cur);
      cur.clear();
    } else
      cur.push_back(ch);
  }
  r.push_back(cur);
  return r;
}

int real_main(int argc, char** argv) {
  bool flags[256] = {0};
  std::map<std::string, std::string> lopt;
  std::vector<std::string> files;
  std::vector<char> pending;
  TrainOpts topt;
  std::string outfile;
  NormGroupBy default_group = CONDITIONAL;
  for (int i = 1; i < argc; ++i) {
    std::string arg = argv[i];
    if (!pending.empty()) {
      char p = pending.front();
      pending.erase(pending.begin());
      W w;
      switch (p) {
        case 'M': topt.max_iter = (unsigned)atol(arg.c_str()); break;
        case 'e':
          parse_weight(arg.c_str(), w);
          topt.converge_arc_delta = w;
          break;
        case 'X':
          parse_weight(arg.c_str(), w);
          topt.converge_perplexity_ratio = w;
          break;
        case 'f':
          parse_weight(arg.c_str(), w);
          topt.smoothFloor = w;
          break;
        case 'o':
          topt.learning_rate_growth_factor = std::max(1., atof(arg.c_str()));
          break;
        case 'F': outfile = arg; break;
        case 'R': break;  // seed: unused (no RNG in the EM path)
        default: break;
      }
      continue;
    }
    if (arg.size() > 1 && arg[0] == '-') {
      if (arg[1] == '-') {
        auto eq = arg.find('=');
        std::string k = arg.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
        lopt[k] = eq == std::string::npos ? "" : arg.substr(eq + 1);
      } else {
        for (size_t k = 1; k < arg.size(); ++k) {
          char c = arg[k];
          flags[(unsigned char)c] = true;
          if (strchr("MeXfoFR", c)) pending.push_back(c);
          if (c == 'j') default_group = JOINT;
          if (c == 'u') default_group = NONE;
        }
      }
    } else
      files.push_back(arg);
  }
  bool trainc = lopt.count("train-cascade") || lopt.count("crp");
  if (trainc) flags[(unsigned)'t'] = true;
  if (flags[(unsigned)':']) topt.cache_derivations = true;
  topt.weight_is_prior_count = flags[(unsigned)'U'];
  if (!flags[(unsigned)'t'] || files.size() < 2) {
    std::cerr << "usage: carmel_oracle -t [options] corpus wfst [wfst...]\n";
    return 1;
  }
  std::string corpus_file = files[0];
  std::vector<std::string> fst_files(files.begin() + 1, files.end());
  unsigned nChain = (unsigned)fst_files.size();
  std::vector<std::unique_ptr<WFST>> chain;
  for (auto const& f : fst_files) {
    std::unique_ptr<WFST> w(new WFST());
    if (!w->read_file(f, !flags[(unsigned)'K'])) {
      std::cerr << "Bad format of transducer file: " << f << "\n";
      return 2;
    }
    if (nChain > 1) w->named = false;  // carmel.cc:1197 unNameStates unless -m
    chain.push_back(std::move(w));
  }
  // normalization methods per transducer: carmel.cc:453-477 set_vector
  std::vector<NormalizeMethod> methods(nChain);
  for (auto& m : methods) m.group = default_group;
  if (lopt.count("normby")) {
    std::string const& s = lopt["normby"];
    for (unsigned i = 0; i < nChain; ++i) {
      char c = i < s.size() ? s[i] : (s.empty() ? 'C' : s.back());
      methods[i].group = (c == 'j' || c == 'J') ? JOINT : (c == 'c' || c == 'C') ? CONDITIONAL : NONE;
    }
  }
  if (lopt.count("priors")) {
    auto v = split(lopt["priors"], ',');
    for (unsigned i = 0; i < nChain; ++i) {
      std::string const& t = i < v.size() ? v[i] : v.back();
      W w;
      parse_weight(t.c_str(), w);
      methods[i].add_count = w;
    }
  }
  Cascade cascade(trainc && nChain >= 2);
  if (nChain < 2 && !cascade.trivial) cascade.set_trivial();
  // carmel.cc:1286-1355: result = chain[0]; minimize (reduce) ; compose left to right
  WFST* result = chain[0].get();
  if (!flags[(unsigned)'d']) result->reduce();
  std::vector<std::unique_ptr<WFST>> composed_keep;
  cascade.add(result);
  bool first = true;
  for (unsigned i = 1; i < nChain && result->valid(); ++i, first = false) {
    cascade.add(chain[i].get());
    if (first)
      cascade.prepare_compose(false, false);
    else
      cascade.prepare_compose(true, false);
    std::unique_ptr<WFST> next = compose(cascade, *result, *chain[i]);
    std::cerr << "\n\t(" << next->numStates() << " states / " << next->numArcs() << " arcs";
    if (!next->valid()) {
      std::cerr << ")\nEmpty or invalid result of composition with transducer \"" << fst_files[i] << "\".\n";
      return 3;
    }
    unsigned st = next->numStates();
    size_t na = next->numArcs();
    if (!flags[(unsigned)'d']) next->reduce();
    if (next->numStates() != st || next->numArcs() != na)
      std::cerr << " reduce-> " << next->numStates() << "/" << next->numArcs();
    std::cerr << ")";
    composed_keep.push_back(std::move(next));
    result = composed_keep.back().get();
    cascade.done_composing(result);
  }
  if (nChain == 1) cascade.set_composed(result);
  std::cerr << std::endl;
  if (!result->valid()) {
    std::cerr << "invalid transducer\n";
    return 3;
  }
  if (lopt.count("write-composed")) {
    std::ofstream o(lopt["write-composed"]);
    result->write(o, true, true, true);
  }
  Corpus corpus;
  {
    std::ifstream cf(corpus_file);
    if (!cf) {
      std::cerr << "File " << corpus_file << " could not be opened for input.\n";
      return 9;
    }
    corpus.read(cf, *result);
  }
  if (cascade.trivial) methods.resize(1);

  if (lopt.count("crp")) return gibbs_main(*result, cascade, corpus, methods, topt, lopt, flags, fst_files, chain);

  Trainer tr(*result, cascade, corpus, methods, topt, std::cerr);

  if (lopt.count("dump-trellis")) {  // dump the pruned trellises (before training)
    std::ofstream o(lopt["dump-trellis"], std::ios::binary);
    IOIndex io(*result);
    uint32_t n = 0;
    std::ostringstream body;
    for (auto const& e : corpus.examples) {
      Derivations d;
      d.in = e.in;
      d.out = e.out;
      d.weight = e.weight;
      if (d.compute(*result, io, tr.arcs)) {
        dump_trellis(body, d);
        ++n;
      }
    }
    write_u32(o, n);
    write_u32(o, (uint32_t)tr.arcs.size());
    o << body.str();
  }
  if (lopt.count("fem-forest") || lopt.count("fem-norm") || lopt.count("fem-param")) {
    // forest-em export of the cascade (cascade.h:34-166; carmel.cc:756-830): parameter ids are 1-based
    // visit order over the cascade members; one forest per example with a derivation.
    std::vector<WFST*> members;
    if (cascade.trivial)
      members.push_back(result);
    else
      for (auto& w : chain) members.push_back(w.get());
    std::unordered_map<Arc const*, unsigned> aid;
    unsigned next_id = 1;
    for (WFST* w : members) w->visit_arcs([&](unsigned, Arc& a) { aid.emplace(&a, next_id++); });
    if (lopt.count("fem-param")) {  // cascade.h:168-181 print_params
      std::ofstream o(lopt["fem-param"]);
      for (WFST* w : members) w->visit_arcs([&](unsigned, Arc& a) { o << fmt_weight(a.weight) << "\n"; });
    }
    if (lopt.count("fem-norm")) {  // cascade.h:85-117 fem_norms
      std::ofstream o(lopt["fem-norm"]);
      o << "(";
      for (size_t i = 0; i < members.size(); ++i) {
        o << "\n";
        NormalizeMethod const& nm = methods[i < methods.size() ? i : methods.size() - 1];
        if (nm.group == NONE) continue;
        std::vector<std::vector<Arc*>> groups;
        members[i]->norm_groups(nm.group, groups);
        for (auto const& g : groups) {
          o << '(';
          for (Arc* a : g) o << ' ' << aid[a];
          o << " )\n";
        }
      }
      o << ")\n";
    }
    if (lopt.count("fem-forest")) {  // cascade.h:119-166 fem_deriv + graph.h:165-194 backrefs
      std::ofstream o(lopt["fem-forest"]);
      IOIndex io(*result);
      for (auto const& e : corpus.examples) {
        Derivations d;
        d.in = e.in;
        d.out = e.out;
        d.weight = e.weight;
        if (!d.compute(*result, io, tr.arcs)) continue;
        struct BR {
          unsigned uses = 0, id = 0;
        };
        std::vector<BR> br(d.g.size());
        unsigned nextid = 1;
        std::function<void(unsigned)> use = [&](unsigned s) {
          if (br[s].uses++ > 0) {
            br[s].id = nextid++;
            return;
          }
          for (auto const& a : d.g[s]) use(a.dest);
        };
        use(0);
        std::function<void(unsigned)> emit = [&](unsigned s) {
          BR& b = br[s];
          const bool backdef = b.uses > 1;
          if (backdef) {
            o << "#" << b.id;
            b.uses = 0;
          } else if (b.uses == 0) {
            o << "#" << b.id;
            return;
          }
          auto const& st = d.g[s];
          const bool ornode = st.size() >= 2;
          if (ornode) o << "(OR";
          for (auto const& a : st) {
            if (ornode) o << " ";
            Arc* arc = tr.arcs.t[a.id].arc;
            std::vector<Arc*> p;
            if (cascade.trivial)
              p.push_back(arc);
            else
              p = cascade.chains[arc->group];
            const bool mid = a.dest != d.fin;
            const bool nonleaf1 = backdef || (!p.empty() && (p.size() > 1 || mid));
            if (nonleaf1) o << "(";
            bool sp = false;
            for (Arc* x : p) {
              if (sp) o << ' ';
              sp = true;
              o << aid[x];
            }
            if (mid) {
              if (sp) o << ' ';
              emit(a.dest);
            }
            if (nonleaf1) o << ")";
          }
          if (ornode) o << ")";
        };
        emit(0);
        o << "\n";
      }
    }
    return 0;
  }
  if (lopt.count("dump-estimate")) {
    // one E-step at the initial (normalised) weights: arc-table ln weights, ln counts, per-example ln P
    std::ofstream o(lopt["dump-estimate"], std::ios::binary);
    cascade.update();
    IOIndex io(*result);
    for (auto& a : tr.arcs.t) a.counts.setZero();
    std::vector<double> lnp;
    for (auto const& e : corpus.examples) {
      Derivations d;
      d.in = e.in;
      d.out = e.out;
      d.weight = e.weight;
      if (d.compute(*result, io, tr.arcs)) lnp.push_back(d.collect_counts(tr.arcs).w);
    }
    write_u32(o, (uint32_t)tr.arcs.size());
    write_u32(o, (uint32_t)lnp.size());
    for (auto const& a : tr.arcs.t) write_f64(o, a.arc->weight.w);
    for (auto const& a : tr.arcs.t) write_f64(o, a.counts.w);
    for (double v : lnp) write_f64(o, v);
    return 0;
  }
  if (lopt.count("dump-viterbi")) {
    // best derivation of every training pair at the initial (normalised) weights: what `carmel -k 1` finds on the
    // composed string x transducer x string machine (fst.h:769-800 bestPaths over makeGraph; graehl/shared/kbest.h).
    // Restated as a max-plus pass over the derivation lattice in its topological order (reverse of the DFS post-order,
    // graph.h:241-288), out-arcs in stored order, strict improvement; then the walk back from the goal.
    std::ofstream o(lopt["dump-viterbi"]);
    o.precision(17);
    cascade.update();
    IOIndex io(*result);
    for (auto const& e : corpus.examples) {
      Derivations d;
      d.in = e.in;
      d.out = e.out;
      d.weight = e.weight;
      if (!d.compute(*result, io, tr.arcs)) continue;
      const unsigned n = (unsigned)d.g.size();
      std::vector<unsigned> order;
      d.make_order(order);
      const double NI = -std::numeric_limits<double>::infinity();
      std::vector<double> best(n, NI);
      std::vector<unsigned> from(n, 0), via(n, 0);
      best[0] = 0;
      for (auto t = order.rbegin(); t != order.rend(); ++t) {
        const unsigned s = *t;
        if (!(best[s] > NI)) continue;
        for (auto const& a : d.g[s]) {
          const double c = best[s] + tr.arcs.t[a.id].arc->weight.w;
          if (c > best[a.dest]) {
            best[a.dest] = c;
            from[a.dest] = s;
            via[a.dest] = a.id;
          }
        }
      }
      std::vector<unsigned> path;
      if (best[d.fin] > NI)
        for (unsigned s = d.fin; s != 0; s = from[s]) path.push_back(via[s]);
      o << best[d.fin] << " " << path.size();
      for (size_t k = path.size(); k-- > 0;) o << " " << path[k];
      o << "\n";
    }
    return 0;
  }
  if (lopt.count("time-estimate")) {  // CPU baseline: time K E-steps (+M-steps) on this corpus
    unsigned K = (unsigned)atoi(lopt["time-estimate"].c_str());
    cascade.update();
    W up;
    tr.estimate(up);  // warm-up (also drops examples without derivations, builds cache if -:)
    tr.maximize(1);
    auto t0 = std::chrono::steady_clock::now();
    for (unsigned k = 0; k < K; ++k) {
      cascade.update();
      tr.estimate(up);
      tr.maximize(1);
    }
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("{\"iters\": %u, \"seconds\": %.6f, \"trellis_arcs\": %zu, \"trellis_states\": %zu, \"examples\": %u, "
                "\"ln_prob\": %.17g, \"cached\": %s}\n",
                K, dt, tr.total_trellis_arcs, tr.total_trellis_states, corpus.n_pairs, up.w,
                topt.cache_derivations ? "true" : "false");
    return 0;
  }

  tr.train();

  if (lopt.count("history")) {
    std::ofstream o(lopt["history"]);
    o.precision(17);
    for (auto const& h : tr.history) o << h.iter << " " << h.ln_prob << " " << h.ln_weighted_prob << " " << h.max_change << "\n";
  }
  bool full = flags[(unsigned)'J'], onearc = flags[(unsigned)'H'];
  WeightFormat wf;
  if (flags[(unsigned)'B']) wf.base = WeightFormat::LOG10;
  else if (flags[(unsigned)'2']) wf.base = WeightFormat::LN;
  if (flags[(unsigned)'Z']) wf.thresh = WeightFormat::ALWAYS;
  if (flags[(unsigned)'D']) wf.thresh = WeightFormat::NEVER;
  if (trainc && !cascade.trivial) {  // cascade.h:23-32 write_trained
    for (unsigned i = 0; i < nChain; ++i) {
      std::string ft = fst_files[i] + ".trained";
      std::cerr << "Writing trained " << fst_files[i] << " to " << ft << std::endl;
      std::ofstream of(ft);
      chain[i]->write(of, full, onearc, false, wf);
    }
  } else if (trainc) {
    std::string ft = fst_files[0] + ".trained";
    std::ofstream of(ft);
    result->write(of, full, onearc, false, wf);
  } else {
    if (!outfile.empty()) {
      std::ofstream of(outfile);
      result->write(of, full, onearc, false, wf);
    } else
      result->write(std::cout, full, onearc, false, wf);
  }
  return 0;
}

int main(int argc, char** argv) {
  // the reference's trellis construction is recursive (derivations.h:640-704): give it stack
  struct rlimit rl;
  if (getrlimit(RLIMIT_STACK, &rl) == 0 && rl.rlim_cur != RLIM_INFINITY && rl.rlim_cur < (1ull << 30)) {
    rl.rlim_cur = std::min<rlim_t>(rl.rlim_max, 1ull << 30);
    if (setrlimit(RLIMIT_STACK, &rl) == 0 && !getenv("ORC_REEXEC")) {
      setenv("ORC_REEXEC", "1", 1);
      execv("/proc/self/exe", argv);
    }
  }
  try {
    return real_main(argc, argv);
  } catch (std::exception& e) {
    std::cerr << "ERROR: " << e.what() << "\n";
    return 11;
  }
}
