#include "common.cuh"
#include <string>
#include <map>
#include <cstring>
extern "C" const char* falwa_b200_version(void) { return "falwa_b200 0.4.0 sm_100a fp64"; }

// ---- launch accounting / per-kernel timing (diagnostics for bench.py; not on the data path) ----
extern "C" long long falwa_b200_launch_count(void) { return falwa::profile_state().launches; }

extern "C" int falwa_b200_profile_enable(int on) {
    falwa::ProfileState& P = falwa::profile_state();
    for (auto& r : P.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    P.recs.clear();
    P.enabled = on != 0;
    return 0;
}

// Synchronises the device, sums the recorded event pairs per kernel name and writes lines
// "name count total_ms\n" into buf (truncated to buflen).  Clears the records.
extern "C" int falwa_b200_profile_read(char* buf, int buflen) {
    falwa::ProfileState& P = falwa::profile_state();
    if (!buf || buflen < 1) return FALWA_ERR_NULL;
    FALWA_CUDA_TRY(cudaDeviceSynchronize());
    std::map<std::string, std::pair<int, double>> acc;
    std::vector<std::string> order;
    for (auto& r : P.recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) ms = 0.f;
        auto it = acc.find(r.name);
        if (it == acc.end()) { acc[r.name] = {1, (double)ms}; order.push_back(r.name); }
        else { it->second.first += 1; it->second.second += ms; }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    P.recs.clear();
    std::string out;
    for (auto& n : order) {
        char line[256];
        snprintf(line, sizeof line, "%s %d %.6f\n", n.c_str(), acc[n].first, acc[n].second);
        out += line;
    }
    const size_t n = std::min(out.size(), (size_t)buflen - 1);
    memcpy(buf, out.data(), n);
    buf[n] = 0;
    return 0;
}

// ---- FP64 pipe micro-benchmark (the DFMA rate every ALU-bound fraction is quoted against) --------
namespace falwa {
__global__ void __launch_bounds__(256)
dfma_rate_kernel(int iters, double seed, double* __restrict__ sink) {
    // eight independent dependency chains per thread: the pipe, not the latency, is what is measured
    double a0 = seed + threadIdx.x, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6., a7 = a0 + 7.;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;   // keeps the chains alive without a store per thread
}
}  // namespace falwa

// Runs `blocks` CTAs of 256 threads, each thread 8 * iters DFMAs; returns the elapsed milliseconds in *ms_out.
extern "C" int falwa_b200_debug_dfma_rate(int blocks, int iters, double* sink_dev, float* ms_out, void* stream) {
    if (!sink_dev || !ms_out || blocks < 1 || iters < 1) return FALWA_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    FALWA_CUDA_TRY(cudaEventCreate(&e0));
    FALWA_CUDA_TRY(cudaEventCreate(&e1));
    falwa::dfma_rate_kernel<<<blocks, 256, 0, st>>>(iters, 1.0, sink_dev);   // warm-up
    FALWA_CUDA_TRY(cudaEventRecord(e0, st));
    falwa::dfma_rate_kernel<<<blocks, 256, 0, st>>>(iters, 1.0, sink_dev);
    FALWA_CUDA_TRY(cudaEventRecord(e1, st));
    FALWA_CUDA_TRY(cudaEventSynchronize(e1));
    FALWA_CUDA_TRY(cudaEventElapsedTime(ms_out, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    FALWA_LAUNCH_CHECK();
    return 0;
}
