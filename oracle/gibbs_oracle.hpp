// gibbs_oracle.hpp -- CPU ORACLE for carmel --crp Gibbs sampling (test infrastructure, not the product).
#pragma once
#include "carmel_oracle.hpp"
namespace orc {
inline int gibbs_main(WFST&, Cascade&, Corpus&, std::vector<NormalizeMethod>&, TrainOpts const&,
                      std::map<std::string, std::string>&, bool*, std::vector<std::string> const&,
                      std::vector<std::unique_ptr<WFST>>&) {
  std::cerr << "gibbs oracle not built yet\n";
  return 12;
}
}  // namespace orc
