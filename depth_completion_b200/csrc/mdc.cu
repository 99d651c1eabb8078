// Single translation unit of libmdc_b200.so (C ABI declared in include/mdc.h, include/mdc_debug.h).
#include "api.cuh"
