// forest_main.cpp -- `forest-em-b200`: forest-em's command line (training subset) on the GPU library.
// Mirrors ForestEmParams::main / perform_forest_em (forest-em/forest-em-params.hpp:262-293,
// forest-em-params.cpp:62-148): exit code 0 on success, 1 on any error with "ERROR: ..." on stderr.
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "carmel_b200.h"

int main(int argc, char** argv) {
  cml_forest_job* job = nullptr;
  int rc = cml_forest_job_open(&job, argc, (const char* const*)argv);
  if (rc != CML_OK) {
    cml_forest_job_close(job);
    return 1;
  }
  rc = cml_forest_job_train(job);
  if (rc == CML_OK) rc = cml_forest_job_write(job);
  if (rc != CML_OK) std::cerr << "ERROR: " << cml_forest_job_error(job) << "\n\nTry 'forest-em -h' for documentation\n";
  cml_forest_job_close(job);
  return rc == CML_OK ? 0 : 1;
}
