"""carmel_b200 -- B200-native (sm_100a) implementation of carmel's training hot path.

The product is the CUDA library carmel_b200/_build/libcarmel_b200.so (C ABI in
include/carmel_b200.h) and the host C++ command line carmel_b200/_build/carmel-b200; this Python
package is only the ctypes binding used by tests, bench.py and the multi-GPU driver.
"""
from .api import (Context, Job, comm_unique_id, CarmelB200Error, load_library, exported_symbols, SPACE_LOG, SPACE_SCALED, NO_GROUP,
                  LOCKED_GROUP, LIB_PATH, CLI_PATH, FOREST_CLI_PATH, OPT_ARC_COUNTS, OPT_NO_ELL, OPT_LANE_MIN, OPT_NO_COUNTS, OPT_NO_FACTOR, OPT_NO_WIDE,
                  ERR_NOT_DENSE)

__all__ = ["Context", "Job", "CarmelB200Error", "load_library", "exported_symbols", "SPACE_LOG", "SPACE_SCALED",
           "NO_GROUP", "LOCKED_GROUP", "LIB_PATH", "CLI_PATH", "FOREST_CLI_PATH"]
