// main.cpp -- carmel-b200: carmel's training command line on the B200.
//
// Keeps the reference's argv grammar for the training path (carmel/src/carmel.cc:929-1066): bundled
// single-character flags, value flags that consume the NEXT argument, --key[=value] long options
// (unknown keys are accepted silently, as in the reference), first file = training corpus, the
// rest = transducers composed left to right.  Output: trained transducer on stdout or -F file, or
// <input>.trained per cascade member with --train-cascade (cascade.h:23-32); log lines on stderr.
//
//   carmel-b200 -t [-HJ] [-M n] [-e w] [-X w] [-f w] [-U] [-o g] [-j|-u] [-d] [-K] [-F out]
//               [--train-cascade] [--normby=JCN] [--priors=w,w] [--float] [--scaled] [--gpu=n]
//               [--dump-trellis=file] [--history=file] corpus wfst [wfst ...]
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "carmel_host.hpp"

using namespace cb;

static std::vector<std::string> split(std::string const& s, char c) {
  std::vector<std::string> r;
  std::stringstream ss(s);
  std::string t;
  while (std::getline(ss, t, c)) r.push_back(t);
  if (r.empty()) r.push_back("");
  return r;
}

static void usage() {
  std::cerr << "carmel-b200 (" << cml_version() << "): B200-native carmel training path\n"
               "usage: carmel-b200 -t [options] corpus wfst [wfst ...]\n"
               "  -t            train (EM forward-backward) on the pair corpus\n"
               "  -M n          maximum iterations (default 500)     -e w  converge on max weight change (1e-4)\n"
               "  -X w          converge on perplexity ratio (.999)  -f w  add floor count w to every arc\n"
               "  -U            initial weights are prior counts     -o g  over-relaxation growth factor\n"
               "  -j / -u       joint / no normalisation (default conditional on input)\n"
               "  -d            do not reduce    -K  state names are indices    -F file  output file\n"
               "  -H -J -B -2 -Z -D   output formatting as in carmel\n"
               "  --train-cascade     train the cascade members, write <file>.trained\n"
               "  --normby=JCN --priors=w,w   per-transducer normalisation / additive priors\n"
               "  --float  fp32 state scores (default fp64)   --scaled  scaled linear space (default log)\n"
               "  --gpu=n  CUDA device   --history=file  --dump-trellis=file  --write-composed=file\n";
}

int main(int argc, char** argv) {
  try {
    if (argc == 1) {
      usage();
      return 0;
    }
    bool flags[256] = {false};
    std::map<std::string, std::string> lopt;
    std::vector<std::string> files;
    std::vector<char> pending;  // value flags waiting for their argument
    TrainOpts topt;
    std::string outfile;
    NormGroupBy default_group = CONDITIONAL;
    for (int i = 1; i < argc; ++i) {
      const std::string arg = argv[i];
      if (!pending.empty()) {
        const char p = pending.front();
        pending.erase(pending.begin());
        double w = 0;
        switch (p) {
          case 'M': topt.max_iter = (uint32_t)std::atol(arg.c_str()); break;
          case 'e':
            if (parse_weight(arg.c_str(), w)) topt.ln_converge_delta = w;
            break;
          case 'X':
            if (parse_weight(arg.c_str(), w)) topt.ln_converge_ratio = w;
            break;
          case 'f':
            if (parse_weight(arg.c_str(), w)) topt.ln_smooth_floor = w;
            break;
          case 'o': topt.rate_growth = std::max(1., std::atof(arg.c_str())); break;
          case 'F': outfile = arg; break;
          default: break;  // -R seed etc.: accepted, unused on this path
        }
        continue;
      }
      if (arg.size() > 1 && arg[0] == '-') {
        if (arg[1] == '-') {
          const size_t eq = arg.find('=');
          const std::string key = arg.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
          const std::string val = eq == std::string::npos ? "" : arg.substr(eq + 1);
          lopt[key] = val;
          std::cerr << "option " << key << " = " << val << std::endl;
        } else {
          for (size_t k = 1; k < arg.size(); ++k) {
            const char c = arg[k];
            flags[(unsigned char)c] = true;
            if (std::strchr("MeXfoFR!kTpwzgLN=+", c)) pending.push_back(c);
            if (c == 'j') default_group = JOINT;
            if (c == 'u') default_group = NONE;
          }
        }
      } else
        files.push_back(arg);
    }
    if (lopt.count("help")) {
      usage();
      return 0;
    }
    const bool trainc = lopt.count("train-cascade") > 0;
    if (trainc) flags[(unsigned)'t'] = true;
    if (lopt.count("crp")) {
      std::cerr << "carmel-b200: --crp Gibbs sampling is not available in this build\n";
      return -11;
    }
    if (!flags[(unsigned)'t']) {
      std::cerr << "carmel-b200 implements carmel's training path only: use -t (or --train-cascade)\n";
      return -11;
    }
    if (files.size() < 2) {
      usage();
      return -9;
    }
    topt.weight_is_prior = flags[(unsigned)'U'];
    if (lopt.count("float")) topt.precision = 32;
    if (lopt.count("scaled")) topt.space = CML_SPACE_SCALED;
    if (lopt.count("gpu")) topt.device = std::atoi(lopt["gpu"].c_str());
    if (lopt.count("history")) topt.history_file = lopt["history"];
    if (lopt.count("dump-trellis")) topt.dump_trellis_file = lopt["dump-trellis"];

    const std::string corpus_file = files[0];
    std::vector<std::string> fst_files(files.begin() + 1, files.end());
    const uint32_t n_chain = (uint32_t)fst_files.size();
    std::vector<std::unique_ptr<Wfst>> chain;
    for (auto const& f : fst_files) {
      std::ifstream in(f);
      if (!in) {
        std::cerr << "File " << f << " could not be opened for input.\n";
        return -9;
      }
      std::unique_ptr<Wfst> w(new Wfst());
      if (!w->read(in, !flags[(unsigned)'K'])) {
        std::cerr << "Bad format of transducer file: " << f << "\n";
        return -2;
      }
      if (n_chain > 1 && !flags[(unsigned)'m']) w->named = false;  // carmel.cc:1197
      chain.push_back(std::move(w));
    }
    std::vector<NormalizeMethod> methods(n_chain);
    for (auto& m : methods) m.group = default_group;
    if (lopt.count("normby")) {
      const std::string s = lopt["normby"];
      for (uint32_t i = 0; i < n_chain && !s.empty(); ++i) {
        const char c = i < s.size() ? s[i] : s.back();
        methods[i].group = (c == 'j' || c == 'J') ? JOINT : (c == 'c' || c == 'C') ? CONDITIONAL : NONE;
      }
    }
    if (lopt.count("priors")) {
      const auto v = split(lopt["priors"], ',');
      for (uint32_t i = 0; i < n_chain; ++i) {
        double w;
        if (parse_weight((i < v.size() ? v[i] : v.back()).c_str(), w)) methods[i].ln_add_count = w;
      }
    }

    Cascade cascade;
    cascade.trivial = !(trainc && n_chain >= 2);
    if (!cascade.trivial) cascade.chains.emplace_back();  // chain 0 = nil (cascade.h:366-383)
    Wfst* result = chain[0].get();
    if (!flags[(unsigned)'d']) result->reduce();  // carmel.cc:1286 cm.minimize(result)
    for (auto& w : chain) cascade.members.push_back(w.get());
    cascade.number_members();  // after reducing the first member: parameter ids = final arc order
    std::vector<std::unique_ptr<Wfst>> keep;
    for (uint32_t i = 1; i < n_chain && result->valid; ++i) {
      std::unique_ptr<Wfst> next = compose(cascade, *result, *chain[i], i > 1, 0, i);
      if (!flags[(unsigned)'q']) std::cerr << "\n\t(" << next->num_states() << " states / " << next->num_arcs() << " arcs";
      if (!next->valid) {
        std::cerr << ")\nEmpty or invalid result of composition with transducer \"" << fst_files[i] << "\".\n";
        return -3;
      }
      const uint32_t st = next->num_states();
      const size_t na = next->num_arcs();
      if (!flags[(unsigned)'d']) next->reduce();
      if (!flags[(unsigned)'q']) {
        if (next->num_states() != st || next->num_arcs() != na)
          std::cerr << " reduce-> " << next->num_states() << "/" << next->num_arcs();
        std::cerr << ")";
      }
      keep.push_back(std::move(next));
      result = keep.back().get();
    }
    if (!flags[(unsigned)'q']) std::cerr << std::endl;
    if (!result->valid) {
      std::cerr << "Empty or invalid transducer.\n";
      return -3;
    }
    cascade.composed = result;
    if (cascade.trivial) methods.resize(1);
    if (lopt.count("write-composed")) {
      std::ofstream o(lopt["write-composed"]);
      result->write(o, true, true, true, WeightFormat());
    }

    Corpus corpus;
    {
      std::ifstream cf(corpus_file);
      if (!cf) {
        std::cerr << "File " << corpus_file << " could not be opened for input.\n";
        return -9;
      }
      corpus.read(cf, *result);
    }

    if (lopt.count("trellis-only")) {  // host-side lattice construction only (no GPU work): for parity tests
      TrellisBatch tb;
      std::vector<uint32_t> dropped;
      build_trellises(*result, corpus, tb, dropped);
      uint64_t ns = 0;
      for (uint32_t n : tb.ex_states) ns += n;
      std::cerr << "Built " << tb.ex_states.size() << " derivation lattices (" << ns << " states, " << tb.arc_dst.size()
                << " arcs; " << dropped.size() << " examples without derivations; " << tb.pre_arcs
                << " arcs examined)\n";
      if (!topt.dump_trellis_file.empty()) {
        std::ofstream o(topt.dump_trellis_file, std::ios::binary);
        tb.dump(o, (uint32_t)result->num_arcs());
      }
      return 0;
    }

    train(*result, cascade, corpus, methods, topt, std::cerr);

    WeightFormat wf;
    if (flags[(unsigned)'B'])
      wf.base = WeightFormat::LOG10;
    else if (flags[(unsigned)'2'])
      wf.base = WeightFormat::LN;
    if (flags[(unsigned)'Z']) wf.thresh = WeightFormat::ALWAYS;
    if (flags[(unsigned)'D']) wf.thresh = WeightFormat::NEVER;
    const bool full = flags[(unsigned)'J'], onearc = flags[(unsigned)'H'];
    if (trainc) {
      for (uint32_t i = 0; i < n_chain; ++i) {
        const std::string ft = fst_files[i] + ".trained";
        std::cerr << "Writing trained " << fst_files[i] << " to " << ft << std::endl;
        std::ofstream of(ft);
        (cascade.trivial ? result : chain[i].get())->write(of, full, onearc, false, wf);
      }
    } else if (!outfile.empty()) {
      std::ofstream of(outfile);
      if (!of) {
        std::cerr << "Could not create file " << outfile << ".\n";
        return -8;
      }
      result->write(of, full, onearc, false, wf);
    } else
      result->write(std::cout, full, onearc, false, wf);
    return 0;
  } catch (std::exception& e) {
    std::cerr << "ERROR: " << e.what() << std::endl;
    return -11;  // carmel.cc:1558-1561
  }
}
