#include "common.cuh"
extern "C" const char* falwa_b200_version(void) { return "falwa_b200 0.1.0 sm_100a fp64"; }
