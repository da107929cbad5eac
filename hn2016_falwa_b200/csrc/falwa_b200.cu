// Single translation unit of libfalwa_b200.so: the stage files share kernels and host helpers.
#include "version.cu"
#include "qgpv.cu"
#include "refstate.cu"
#include "flux_scan.cu"
#include "flux.cu"
#include "nhn22_solve.cu"
#include "pipeline.cu"
#include "baro2d.cu"
