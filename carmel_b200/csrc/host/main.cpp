// main.cpp -- carmel-b200: carmel's training command line on the B200.
//
// A thin shell over the job object (job.cpp): same argv grammar, log lines and outputs as
// `carmel -t ...` / `carmel --train-cascade ...` (carmel/src/carmel.cc:868-1561).  Exit codes
// follow the reference: 0 ok, -11 on an exception (carmel.cc:1558-1561), -2/-3/-8/-9 input errors.
//
//   carmel-b200 -t [-HJ] [-M n] [-e w] [-X w] [-f w] [-U] [-o g] [-j|-u] [-d] [-K] [-F out]
//               [--train-cascade] [--normby=JCN] [--priors=w,w] [--float] [--scaled] [--gpu=n]
//               [--dump-trellis=file] [--history=file] [--trellis-only] corpus wfst [wfst ...]
#include <fstream>
#include <iostream>

#include "carmel_host.hpp"

using namespace cb;

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
               "  --gpu=n  CUDA device   --history=file  --dump-trellis=file  --write-composed=file\n"
               "  --trellis-only  build (and dump) the derivation lattices on the host, then stop\n";
}

int main(int argc, char** argv) {
  try {
    if (argc == 1) {
      usage();
      return 0;
    }
    TrainJob job;
    const int rc = open_job(argc, argv, job, std::cerr);
    if (job.lopt.count("help")) {
      usage();
      return 0;
    }
    if (rc == -9 && job.fst_files.empty()) usage();
    if (rc != 0) return rc;
    if (job.lopt.count("trellis-only")) {  // host-side lattice construction only (no GPU work): parity tests
      TrellisBatch tb;
      std::vector<uint32_t> dropped;
      size_t e0, e1;  // with --shard=r/N: this rank's block only (what TrainJob::prepare keeps on its GPU)
      shard_range(job.corpus, job.opt.shard_rank, job.opt.shard_count, e0, e1);
      if (job.opt.shard_count > 1) {
        Corpus local;
        local.examples.assign(job.corpus.examples.begin() + e0, job.corpus.examples.begin() + e1);
        job.corpus.examples.swap(local.examples);
        std::cerr << "Shard " << job.opt.shard_rank << "/" << job.opt.shard_count << ": examples [" << e0 << ", " << e1 << ")\n";
      }
      build_trellises(*job.x, job.corpus, tb, dropped);
      uint64_t ns = 0;
      for (uint32_t n : tb.ex_states) ns += n;
      std::cerr << "Built " << tb.ex_states.size() << " derivation lattices (" << ns << " states, " << tb.arc_dst.size()
                << " arcs; " << dropped.size() << " examples without derivations; " << tb.pre_arcs
                << " arcs examined)\n";
      if (!job.opt.dump_trellis_file.empty()) {
        std::ofstream o(job.opt.dump_trellis_file, std::ios::binary);
        tb.dump(o, (uint32_t)job.x->num_arcs());
      }
      return 0;
    }
    if (job.gopt.enabled)
      job.run_gibbs(std::cerr);
    else
      job.run(std::cerr);
    job.write_outputs(std::cout);
    return 0;
  } catch (std::exception& e) {
    std::cerr << "ERROR: " << e.what() << std::endl;
    return -11;
  }
}
