// forest_cli.cpp -- command-line driver of the forest-em CPU ORACLE (test infrastructure, not the product).
//
// Accepts the subset of forest-em's options (forest-em/forest-em-params.hpp:69-176) that concern the
// training path: -f/--forests-file  -n/--normgroups-file  -I/--initparam-file  -o/--outparam-file
// -O/--outcounts-file  -S/--out-per-forest-inside-sum  -i/--max-iter  -e/--converge  -d/--deltaparam-epsilon
// -p/--prior-counts-per  -k/--add-k-smoothing  -z  -u  -N  -U  -H  -L/--log-level, plus the test hooks
//   --history=file        iteration, average log prob, max delta, index, N  (17 digits)
//   --print-forests=file  parse -> print round trip of every forest (forest.hpp:245-320)
//   --time-estimate=K     time K E-steps (inside + outside + counts) and print a JSON line
#include <chrono>

#include "forest_oracle.hpp"

using namespace forc;

struct Args {
  std::string forests, norm, initparam, outparam, outcounts, outinside, history, print_forests, outviterbi;
  ForestOpts opt;
  bool dbl = false;
  int time_estimate = 0;
  // --crp=n Gibbs sampling (gibbs_opts.hpp:34-130 as used by forest-em-params.hpp:172-175)
  unsigned crp = 0, burnin = 0;
  double alpha = .1, high_temp = 1, low_temp = 1, n_sym = 0;
  bool uniformp0 = false, final_counts = false, exclude_prior = false, sample_prob = false;
  unsigned long long seed = 1;
  std::string outsample;
};

template <class Real>
static int run(Args const& a) {
  Forests<Real> F;
  F.opt = a.opt;
  std::ostream& log = std::cerr;
  if (!a.initparam.empty()) {
    std::ifstream in(a.initparam);
    if (!in) throw std::runtime_error("can't open " + a.initparam);
    F.read_params(in);
  }
  std::ifstream nin;
  if (!a.norm.empty()) {
    nin.open(a.norm);
    if (!nin) throw std::runtime_error("can't open " + a.norm);
    F.read_norm_groups(nin);
    if (!a.initparam.empty() && a.opt.normalize_initial) F.normalize_params();
  }
  if (!a.forests.empty()) {
    if (a.forests == a.norm) {
      F.read_forests(nin);  // "norm_and_forests": groups then forests in one stream
    } else {
      std::ifstream fin(a.forests);
      if (!fin) throw std::runtime_error("can't open " + a.forests);
      F.read_forests(fin);
    }
    log << F.n_nodes << " forest nodes total, max #nodes " << F.max_nodes << ", " << F.forests.size() << " forests\n "
        << F.norm_groups.groups.size() << " normalization groups, " << F.norm_groups.num_params() << " parameters\n"
        << " largest rule index was " << F.max_forest_ruleid << ".\n";
    if (!a.print_forests.empty()) {
      std::ofstream o(a.print_forests);
      for (auto const& f : F.forests) {
        f.print(o);
        o << "\n";
      }
    }
    F.prepare();
    if (a.time_estimate > 0) {
      std::ostringstream sink;
      F.estimate(true, sink);
      auto t0 = std::chrono::steady_clock::now();
      double alp = 0;
      for (int k = 0; k < a.time_estimate; ++k) alp = F.estimate(false, sink);
      double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      size_t he = 0;
      for (auto const& f : F.forests)
        for (auto const& n : f.nodes) he += (!n.backref && n.label != 0);
      std::printf("{\"iters\": %d, \"seconds\": %.6f, \"hyperedges\": %zu, \"nodes\": %zu, \"forests\": %zu, \"avg_logprob\": %.17g}\n",
                  a.time_estimate, sec, he, F.n_nodes, F.forests.size(), alp);
      return 0;
    }
    if (a.crp) {  // forest-em-params.cpp:112-113: gibbs replaces EM
      typename Forests<Real>::GibbsOpts g;
      g.iter = a.crp;
      g.burnin = std::min(a.burnin, a.crp);
      if (a.final_counts) g.burnin = a.crp;  // gibbs_opts.hpp:259
      g.alpha = a.alpha;
      g.uniformp0 = a.uniformp0;
      g.final_counts = a.final_counts;
      g.exclude_prior = a.exclude_prior;
      g.sample_prob = a.sample_prob;
      g.high_temp = a.high_temp;
      g.low_temp = a.low_temp;
      g.seed = a.seed;
      g.n_sym = a.n_sym;
      F.run_gibbs(g, log);
      if (!a.outsample.empty()) {  // print_sample (forest-em.hpp:775-784): rule ids in record order, one forest per line
        std::ofstream o(a.outsample);
        for (auto const& smp : F.g_sample) {
          for (size_t k = 0; k < smp.size(); ++k) o << (k ? " " : "") << smp[k];
          o << "\n";
        }
      }
      if (!a.history.empty()) {
        std::ofstream o(a.history);
        o.precision(17);
        for (size_t i = 0; i < F.g_iter_ln_prob.size(); ++i) o << i << ' ' << F.g_iter_ln_prob[i] << "\n";
      }
    } else if (a.opt.max_iter)
      F.train(log);
  }
  if (!a.outparam.empty()) {
    log << "Writing trained parameters to " << a.outparam << "\n";
    std::ofstream o(a.outparam);
    F.write_range(o, F.rule_weights);
  }
  if (!a.outcounts.empty()) {
    log << "Writing trained counts to " << a.outcounts << "\n";
    std::ofstream o(a.outcounts);
    F.write_range(o, F.counts);
  }
  if (!a.outinside.empty()) {  // final_iteration (forest-em.hpp:500-509): inside only, no counts
    std::ofstream o(a.outinside);
    std::ostringstream sink;
    F.last_inside.clear();
    F.forest_no = 0;
    F.n_zeroprob = 0;
    for (auto const& f : F.forests) F.visit_forest(f, false, false, sink);
    for (double l : F.last_inside) o << fmt_weight(LW<Real>::ln((Real)l), a.opt.human_probs) << "\n";
  }
  if (!a.outviterbi.empty()) {  // forest-em-params.cpp:125-131 final viterbi forests decoding (-v)
    log << "Running final viterbi forests decoding.\n";
    std::ofstream o(a.outviterbi);
    for (auto const& f : F.forests) F.write_viterbi(o, f, a.opt.human_probs);
  }
  if (!a.history.empty() && !a.crp) {
    std::ofstream o(a.history);
    o.precision(17);
    for (auto const& h : F.history) o << h.i << ' ' << h.alp << ' ' << h.max_delta << ' ' << h.max_index << ' ' << h.n << "\n";
  }
  return 0;
}

int main(int argc, char** argv) {
  Args a;
  static const std::map<std::string, char> longs = {
      {"forests-file", 'f'},  {"normgroups-file", 'n'},   {"initparam-file", 'I'},     {"outparam-file", 'o'},
      {"outcounts-file", 'O'}, {"max-iter", 'i'},          {"converge", 'e'},           {"deltaparam-epsilon", 'd'},
      {"prior-counts-per", 'p'}, {"add-k-smoothing", 'k'}, {"zero-zerocounts", 'z'},    {"initial-1-params", 'u'},
      {"normalize-initial", 'N'}, {"use-double-precision", 'U'}, {"human-probs", 'H'},  {"log-level", 'L'},
      {"out-per-forest-inside-sum", 'S'}, {"outviterbi-file", 'v'}};
  try {
    for (int i = 1; i < argc; ++i) {
      std::string s = argv[i];
      char c = 0;
      std::string val;
      bool have_val = false;
      if (s.rfind("--", 0) == 0) {
        size_t eq = s.find('=');
        std::string key = s.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
        if (eq != std::string::npos) {
          val = s.substr(eq + 1);
          have_val = true;
        }
        if (key == "history") { a.history = val; continue; }
        if (key == "print-forests") { a.print_forests = val; continue; }
        if (key == "time-estimate") { a.time_estimate = std::atoi(val.c_str()); continue; }
        if (key == "crp") { a.crp = (unsigned)std::atol(val.c_str()); continue; }
        if (key == "burnin") { a.burnin = (unsigned)std::atol(val.c_str()); continue; }
        if (key == "const-alpha") { a.alpha = std::atof(val.c_str()); continue; }
        if (key == "high-temp") { a.high_temp = std::atof(val.c_str()); continue; }
        if (key == "low-temp") { a.low_temp = std::atof(val.c_str()); continue; }
        if (key == "n-symbols") { a.n_sym = std::atof(val.c_str()); continue; }
        if (key == "seed") { a.seed = std::strtoull(val.c_str(), nullptr, 10); continue; }
        if (key == "outsample-file") { a.outsample = val; continue; }
        if (key == "uniform-p0") { a.uniformp0 = true; continue; }
        if (key == "final-counts") { a.final_counts = true; continue; }
        if (key == "crp-exclude-prior") { a.exclude_prior = true; continue; }
        if (key == "sample-prob") { a.sample_prob = true; continue; }
        auto it = longs.find(key);
        if (it == longs.end()) throw std::runtime_error("unknown option --" + key);
        c = it->second;
      } else if (s.size() >= 2 && s[0] == '-') {
        c = s[1];
        if (s.size() > 2) {
          val = s.substr(2);
          have_val = true;
        }
      } else
        throw std::runtime_error("unexpected argument " + s);
      auto need = [&]() -> std::string {
        if (have_val) return val;
        if (i + 1 >= argc) throw std::runtime_error(std::string("option -") + c + " needs a value");
        return argv[++i];
      };
      switch (c) {
        case 'f': a.forests = need(); break;
        case 'n': a.norm = need(); break;
        case 'I': a.initparam = need(); break;
        case 'o': a.outparam = need(); break;
        case 'O': a.outcounts = need(); break;
        case 'S': a.outinside = need(); break;
        case 'v': a.outviterbi = need(); break;
        case 'i': a.opt.max_iter = (unsigned)std::atol(need().c_str()); break;
        case 'e': a.opt.converge_ratio = std::atof(need().c_str()); break;
        case 'd': a.opt.converge_delta = std::atof(need().c_str()); break;
        case 'p': a.opt.prior_counts = std::atof(need().c_str()); break;
        case 'k': a.opt.add_k_smoothing = std::atof(need().c_str()); break;
        case 'L': a.opt.log_level = (unsigned)std::atol(need().c_str()); break;
        case 'z': a.opt.zero_zerocounts = true; break;
        case 'u': a.opt.initial_1_params = true; break;
        case 'N': a.opt.normalize_initial = true; break;
        case 'U': a.dbl = true; break;
        case 'H': a.opt.human_probs = true; break;
        default: throw std::runtime_error(std::string("unknown option -") + c);
      }
    }
    if (a.opt.max_iter && a.forests.empty()) throw std::runtime_error("Missing forests-file.");
    if (a.norm.empty() && (a.opt.max_iter || a.opt.normalize_initial)) throw std::runtime_error("Missing normgroups-file.\n");
    return a.dbl ? run<double>(a) : run<float>(a);
  } catch (std::exception& e) {
    std::cerr << "ERROR: " << e.what() << "\n\nTry 'forest-em -h' for documentation\n";
    return 1;
  }
}
