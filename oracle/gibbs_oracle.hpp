// gibbs_oracle.hpp -- CPU ORACLE for carmel --crp Gibbs sampling (test infrastructure, not the product).
//
// Restates graehl/shared/gibbs.hpp (gibbs_param :106-227, gibbs_base :229-1079: define_param :582-597,
// restore_p0 :618-623, cache prob :712-742, addc :769-792, run :803-828, iteration :836-877,
// finalize_cumulative_counts :626-638), graehl/shared/delta_sum.hpp:49-106, carmel/src/gibbs.cc
// (add_gibbs_params :114-186, resample_block :306-326, proposal weight :348-359, choose_arc :362-371,
// train_gibbs :386-430), carmel/src/derivations.h:306-375 (pfor::global_normalize, random_path) and
// graehl/shared/random.ipp:111-127 (choose_p).
//
// PARITY UNPINNED by the reference: its golden log is RNG dependent (boost lagged_fibonacci607, seed not
// recorded) and no sampled derivations are stored.  This restatement is pinned only to itself; the
// uniform draws are INJECTED (counter-based generator below, shared with the product) so that the
// product's sequential mode can be compared derivation by derivation.
#pragma once
#include "carmel_oracle.hpp"

namespace orc {

// counter-based uniforms shared by oracle and product: u(seed, sweep, block, draw) in [0,1)
inline uint64_t gibbs_mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
inline double gibbs_uniform(uint64_t seed, uint32_t sweep, uint32_t block, uint32_t draw) {
  uint64_t h = gibbs_mix64(seed ^ gibbs_mix64(((uint64_t)sweep << 32) | block));
  h = gibbs_mix64(h + draw);
  return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

struct DeltaSum {  // delta_sum.hpp:49-106
  double x = 0, tmax = 0, s = 0;
  void clear(double x0) {
    x = x0;
    s = tmax = 0;
  }
  void add_delta(double d, double t) {
    double moret = t - tmax;
    if (moret > 0) {
      tmax = t;
      s += moret * x;
    } else if (moret < 0)
      s += d * (-moret);
    x += d;
  }
  void extend(double t) {
    double moret = t - tmax;
    tmax = t;
    s += x * moret;
  }
};
static const unsigned NONORM = 0xFFFFFFFFu;
struct GibbsParam {  // gibbs.hpp:106-227
  double prior = 0;
  unsigned norm = NONORM;
  DeltaSum sumcount;
  bool has_norm() const { return norm != NONORM; }
};

struct GibbsOpts {
  unsigned iter = 0, burnin = 0;
  bool uniformp0 = false, dirichlet_p0 = false, final_counts = false, exclude_prior = false;
  double high_temp = 1, low_temp = 1;
  uint64_t seed = 1;
  unsigned init_em = 0;
  bool em_p0 = false;
  // --sample-prob (carmel.cc:1869 "show the sample prob given model, previous sample"): the proposal probability of each
  // new sample under the current counts, which is what the older binary behind the golden log
  // carmel-tutorial/commands.trace:6976-12996 printed by default ("sample prob=").
  bool sample_prob = false;
  // --expectation (gibbs_opts.hpp; gibbs.cc:311-316, derivations.h:381-398 collect_counts_gibbs): a block's "sample" is
  // every lattice arc with its posterior under the current proposal probabilities -- incremental EM over the CRP counts
  bool expectation = false;
};

struct Gibbs {
  WFST& composed;
  Cascade& cascade;
  std::vector<NormalizeMethod> methods;
  GibbsOpts gopt;
  std::vector<GibbsParam> gps;
  std::vector<double> normsum;
  unsigned nnorm = 0;
  std::vector<Derivations> derivs;
  std::vector<std::vector<unsigned>> chain_of_arc;  // arc-table id -> param ids (chain order)
  std::vector<std::vector<unsigned>> sample;        // per block: param ids in path order
  std::vector<std::vector<unsigned>> sample_arcs;   // per block: arc-table ids of the sampled path
  std::vector<std::vector<double>> sample_wt;       // --expectation: weight of every entry of sample[b] (block_delta::wt)
  std::vector<Arc*> arc_of_param;
  std::vector<double> init_arc_weight;  // composed arc weights for the iteration-0 sample (--init-em)
  bool init_prob = false;
  double time = 0;
  unsigned iter = 0;
  double n_sym = 0;
  std::vector<double> iter_ln_prob;

  Gibbs(WFST& x, Cascade& c, std::vector<NormalizeMethod> const& m, GibbsOpts const& o)
      : composed(x), cascade(c), methods(m), gopt(o) {}

  unsigned define_param(unsigned norm, double prior, Arc* a) {
    if (norm != NONORM) nnorm = std::max(nnorm, norm + 1);
    GibbsParam p;
    p.norm = norm;
    p.prior = prior;
    gps.push_back(p);
    arc_of_param.push_back(a);
    return (unsigned)gps.size() - 1;
  }
  // gibbs.cc:114-186
  unsigned add_gibbs_params(unsigned id, WFST& w, NormalizeMethod const& nm) {
    if (nm.group == NONE) {
      w.visit_arcs([&](unsigned, Arc& a) { a.group = define_param(NONORM, a.weight.getReal(), &a); });
      return id;
    }
    std::vector<std::vector<Arc*>> groups;
    w.norm_groups(nm.group, groups);
    double alpha = nm.add_count.getReal();
    bool cond = nm.group == CONDITIONAL;
    for (auto& g : groups) {
      W sum;
      std::vector<Arc*> unlocked;
      for (Arc* a : g) {
        if (a->isLocked())
          a->group = define_param(NONORM, a->weight.getReal(), a);
        else {
          unlocked.push_back(a);
          sum += a->weight;
        }
      }
      unsigned N = (unsigned)unlocked.size();
      if (gopt.dirichlet_p0) sum = W::one();
      if (cond) std::reverse(unlocked.begin(), unlocked.end());
      for (Arc* a : unlocked) {
        double prob = (a->weight / sum).getReal();
        double prior = gopt.uniformp0 ? alpha : alpha * prob * N;
        a->group = define_param(id, prior, a);
      }
      ++id;
    }
    return id;
  }
  void setup(Corpus& corpus, std::ostream& log) {
    // chains must be captured BEFORE groupIds are overwritten by parameter ids
    ArcsTable atab(composed, false, W());
    chain_of_arc.resize(atab.size());
    std::vector<std::vector<Arc*>> chain_arcs(atab.size());
    for (size_t i = 0; i < atab.size(); ++i) {
      Arc* a = atab.t[i].arc;
      if (cascade.trivial)
        chain_arcs[i].push_back(a);
      else
        chain_arcs[i] = cascade.chains[a->group];
    }
    // derivations (cached, pruned): gibbs.cc:23
    IOIndex io(composed);
    for (auto i = corpus.examples.begin(); i != corpus.examples.end();) {
      Derivations d;
      d.in = i->in;
      d.out = i->out;
      d.weight = i->weight;
      if (d.compute(composed, io, atab)) {
        derivs.push_back(std::move(d));
        ++i;
      } else {
        log << "No derivations in transducer for input/output\n";
        i = corpus.examples.erase(i);
      }
    }
    corpus.count();
    n_sym = corpus.n_output;
    unsigned norm = 0;
    for (unsigned i = 0; i < cascade.cascade.size(); ++i) norm = add_gibbs_params(norm, *cascade.cascade[i], methods[i]);
    for (size_t i = 0; i < atab.size(); ++i)
      for (Arc* a : chain_arcs[i]) chain_of_arc[i].push_back(a->group);
    sample.assign(derivs.size(), {});
    sample_arcs.assign(derivs.size(), {});
    sample_wt.assign(derivs.size(), {});
  }
  double proposal_prob(unsigned id) const {
    GibbsParam const& p = gps[id];
    return p.has_norm() ? p.sumcount.x / normsum[p.norm] : p.prior;
  }
  void addc_weighted(std::vector<unsigned> const& b, std::vector<double> const& w, double d) {  // gibbs.hpp:779-786
    for (size_t i = 0; i < b.size(); ++i) {
      GibbsParam& p = gps[b[i]];
      if (p.has_norm()) {
        normsum[p.norm] += w[i] * d;
        p.sumcount.add_delta(w[i] * d, time);
      }
    }
  }
  // derivations.h:381-398 collect_counts_gibbs: forward-backward with the proposal weights; every arc's posterior goes
  // to every parameter of its chain.  Returns the block probability (sum over all derivations).
  W expectation_block(unsigned b) {
    Derivations& d = derivs[b];
    auto wt = [&](GraphArc const& a) {
      W prob = W::one();
      for (unsigned id : chain_of_arc[a.id]) prob *= W(proposal_prob(id));
      return prob;
    };
    std::vector<W> f, bw;
    W prob = d.compute_fb(f, bw, wt);
    for (unsigned s = 0; s < d.g.size(); ++s)
      for (GraphArc const& a : d.g[s]) {
        W contrib = wt(a) * f[a.src] * bw[a.dest];
        const double w = (contrib / prob).getReal();
        for (unsigned id : chain_of_arc[a.id]) {
          sample[b].push_back(id);
          sample_wt[b].push_back(w);
        }
      }
    return prob;
  }
  void addc(std::vector<unsigned> const& b, double d) {
    for (unsigned id : b) {
      GibbsParam& p = gps[id];
      if (p.has_norm()) {
        normsum[p.norm] += d;
        p.sumcount.add_delta(d, time);
      }
    }
  }
  // derivations.h:345-375 random_path with injected uniforms
  void resample_block(unsigned b, double power, uint32_t sweep) {
    Derivations& d = derivs[b];
    unsigned nst = (unsigned)d.g.size();
    auto wt = [&](GraphArc const& a) {
      if (init_prob) return W(init_arc_weight[a.id]);
      W prob = W::one();
      for (unsigned id : chain_of_arc[a.id]) prob *= W(proposal_prob(id));
      return prob;
    };
    std::vector<unsigned> reverse_order;
    d.make_order(reverse_order);
    std::vector<GArcs> r;
    d.make_reverse(r);
    std::vector<W> bw(nst, W());
    bw[d.fin] = W::one();
    for (auto t = reverse_order.begin(); t != reverse_order.end(); ++t)
      for (GraphArc const& a : r[*t]) bw[a.dest] += bw[*t] * wt(a);
    unsigned s = 0;
    uint32_t draw = 0;
    while (s != d.fin) {
      GArcs& arcs = d.g[s];
      // pfor::global_normalize (derivations.h:318-337)
      W sum;
      std::vector<W> nw;
      for (GraphArc& a : arcs) {
        W v = (wt(a) * bw[a.dest]).pow(power);
        sum += v;
        nw.push_back(v);
      }
      if (sum.isZero()) sum.setOne();
      std::vector<double> p;
      for (W v : nw) p.push_back((v / sum).getReal());
      // choose_p (random.ipp:111-127)
      double psum = 0;
      for (double x : p) psum += x;
      double choice = psum * gibbs_uniform(gopt.seed, sweep, b, draw++);
      auto it = arcs.begin();
      size_t k = 0;
      for (;;) {
        choice -= p[k];
        auto cur = it;
        ++it;
        ++k;
        if (choice < 0 || it == arcs.end()) {
          it = cur;
          break;
        }
      }
      for (unsigned id : chain_of_arc[it->id]) sample[b].push_back(id);
      sample_arcs[b].push_back(it->id);
      s = it->dest;
    }
  }
  void iteration(std::ostream& log) {
    double temperature = gopt.high_temp;
    if (gopt.iter > 0 && gopt.high_temp != gopt.low_temp)
      temperature = gopt.high_temp + (gopt.low_temp - gopt.high_temp) * std::min(1.0, (double)iter / gopt.iter);
    double power = temperature > 0 ? 1. / temperature : 1;
    std::vector<double> ccount(gps.size()), csum(nnorm, 0.);  // cache reset (gibbs.hpp:700-705,656-667)
    for (size_t i = 0; i < gps.size(); ++i)
      if (gps[i].has_norm()) csum[gps[i].norm] += (ccount[i] = gps[i].prior);
    W p = W::one();
    if (iter > 0) init_prob = false;
    for (unsigned b = 0; b < derivs.size(); ++b) {
      double wt = derivs[b].weight;
      if (gopt.expectation) {  // gibbs.hpp:851-871 with block_delta weights
        addc_weighted(sample[b], sample_wt[b], -wt);
        sample[b].clear();
        sample_wt[b].clear();
        p *= expectation_block(b);
        addc_weighted(sample[b], sample_wt[b], wt);
        continue;
      }
      addc(sample[b], -wt);
      sample[b].clear();
      sample_arcs[b].clear();
      resample_block(b, power, iter);
      W bp = W::one();
      if (gopt.sample_prob) {
        // scored AFTER the new sample's counts are back in (gibbs.hpp:866 comment: "do it after to get overestimate"):
        // the only reading under which the golden log is possible -- its i=0 sample has 2^-207028 > the EM optimum 2^-212071
        addc(sample[b], wt);
        for (unsigned id : sample[b]) bp *= W(proposal_prob(id));
        p *= bp;
        continue;
      }
      for (unsigned id : sample[b]) {
        GibbsParam const& gp = gps[id];
        bp *= W(gp.has_norm() ? ccount[id]++ / csum[gp.norm]++ : gp.prior);
      }
      p *= bp;
      addc(sample[b], wt);
    }
    iter_ln_prob.push_back(p.w);
    log << "Gibbs i=" << iter << (gopt.expectation ? " sum-all-derivations prob=" : gopt.sample_prob ? " sample prob=" : " cache-model prob=")
        << fmt_base2(p);
    if (n_sym) log << " per-point-ppx(N=" << n_sym << ")=" << fmt_base2(p.ppxper(n_sym));
    log << " per-block-ppx(N=" << derivs.size() << ")=" << fmt_base2(p.ppxper((double)derivs.size())) << "\n";
  }
  void run(std::ostream& log) {
    normsum.assign(nnorm, 0.);
    for (auto& p : gps)
      if (p.has_norm()) {
        normsum[p.norm] += p.prior;
        p.sumcount.clear(p.prior);
      }
    iter = 0;
    time = 0;
    iteration(log);
    for (iter = 1; iter <= gopt.iter; ++iter) {
      time = (double)iter - (double)gopt.burnin;
      if (time < 0) time = 0;
      iteration(log);
    }
    // finalize_cumulative_counts (gibbs.hpp:626-644)
    if (!(gopt.final_counts && !gopt.exclude_prior)) {
      double tmax1 = ((double)gopt.iter - (double)gopt.burnin) + 1;
      for (auto& p : gps) {
        if (!p.has_norm()) continue;
        if (gopt.exclude_prior) {
          p.sumcount.s += -p.prior * p.sumcount.tmax;
          p.sumcount.x += -p.prior;
        }
        if (!gopt.final_counts) {
          p.sumcount.extend(tmax1);
          p.sumcount.x = p.sumcount.s;
        }
      }
      normsum.assign(nnorm, 0.);
      for (auto& p : gps)
        if (p.has_norm()) normsum[p.norm] += p.sumcount.x;
    }
    // probs_to_cascade (gibbs.cc:66-76): arc.weight = final_prob(param)
    for (size_t i = 0; i < gps.size(); ++i) {
      GibbsParam const& p = gps[i];
      double fp = p.has_norm() ? (p.sumcount.x > 0 ? p.sumcount.x / normsum[p.norm] : 0.) : p.prior;
      arc_of_param[i]->weight = W(fp);
    }
  }
};

inline int gibbs_main(WFST& result, Cascade& cascade, Corpus& corpus, std::vector<NormalizeMethod>& methods,
                      TrainOpts const& topt, std::map<std::string, std::string>& lopt, bool* flags,
                      std::vector<std::string> const& fst_files, std::vector<std::unique_ptr<WFST>>& chain) {
  cascade.set_composed(&result);
  for (auto& m : methods)  // gibbs.cc:390-397
    if (!(m.add_count.w > -ORC_INF)) {
      std::cerr << "Gibbs sampling requires positive --priors for base model / initial sample.  Setting to 0.01\n";
      m.add_count = W(1e-2);
    }
  GibbsOpts g;
  g.iter = topt.max_iter;
  if (lopt.count("crp") && !lopt["crp"].empty()) g.iter = (unsigned)atoi(lopt["crp"].c_str());
  if (lopt.count("burnin")) g.burnin = (unsigned)atoi(lopt["burnin"].c_str());
  if (lopt.count("uniform-p0")) g.uniformp0 = true;
  if (lopt.count("dirichlet-p0")) g.dirichlet_p0 = true;
  if (lopt.count("final-counts")) g.final_counts = true;
  if (lopt.count("crp-exclude-prior")) g.exclude_prior = true;
  if (lopt.count("high-temp")) g.high_temp = atof(lopt["high-temp"].c_str());
  g.sample_prob = lopt.count("sample-prob") > 0;
  g.expectation = lopt.count("expectation") > 0;
  if (lopt.count("low-temp")) g.low_temp = atof(lopt["low-temp"].c_str());
  if (lopt.count("seed")) g.seed = strtoull(lopt["seed"].c_str(), nullptr, 10);
  if (g.final_counts) g.burnin = g.iter;
  if (g.burnin > g.iter) g.burnin = g.iter;
  Gibbs gb(result, cascade, methods, g);
  gb.setup(corpus, std::cerr);
  gb.run(std::cerr);
  if (lopt.count("dump-samples")) {  // final sample of every block: arc-table ids in path order
    std::ofstream o(lopt["dump-samples"]);
    for (auto const& s : gb.sample_arcs) {
      for (size_t k = 0; k < s.size(); ++k) o << (k ? " " : "") << s[k];
      o << "\n";
    }
  }
  if (lopt.count("history")) {
    std::ofstream o(lopt["history"]);
    o.precision(17);
    for (size_t i = 0; i < gb.iter_ln_prob.size(); ++i) o << i << " " << gb.iter_ln_prob[i] << "\n";
  }
  bool full = flags[(unsigned)'J'], onearc = flags[(unsigned)'H'];
  for (auto& w : chain)  // cascade.clear_groups()
    w->visit_arcs([](unsigned, Arc& a) { a.group = NO_GROUP; });
  if (cascade.trivial) result.visit_arcs([](unsigned, Arc& a) { a.group = NO_GROUP; });
  for (size_t i = 0; i < fst_files.size(); ++i) {
    std::string ft = fst_files[i] + ".trained";
    std::cerr << "Writing trained " << fst_files[i] << " to " << ft << std::endl;
    std::ofstream of(ft);
    (cascade.trivial ? &result : chain[i].get())->write(of, full, onearc, false);
  }
  return 0;
}

}  // namespace orc
