// forest_cli.cpp -- driver for the forest-em CPU ORACLE (test infrastructure, not the product).
#include "forest_oracle.hpp"
int main() {
  std::cerr << "forest oracle not built yet\n";
  return 12;
}
