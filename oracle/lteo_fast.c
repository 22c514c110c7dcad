/* placeholder for the multi-threaded / SIMD CPU-baseline variants (filled in later) */
#include "lte_oracle.h"
