// forest_oracle.hpp -- CPU ORACLE for forest-em inside-outside (test infrastructure, not the product).
#pragma once
#include "carmel_oracle.hpp"
