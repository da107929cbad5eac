#include "common.cuh"
#include <string>
#include <map>
#include <cstring>
extern "C" const char* falwa_b200_version(void) { return "falwa_b200 0.3.0 sm_100a fp64"; }

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
