// Shared helpers for the falwa_b200 kernels (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <vector>
#include <type_traits>
#include <thread>
#include <atomic>
#include <algorithm>
#include "../../include/falwa_b200.h"

#define FALWA_CUDA_TRY(expr)                         \
    do {                                             \
        cudaError_t _e = (expr);                     \
        if (_e != cudaSuccess) return (int)_e;       \
    } while (0)

#define FALWA_LAUNCH_CHECK() FALWA_CUDA_TRY(cudaGetLastError())

// Launch accounting + optional per-kernel CUDA-event timing (falwa_b200_profile_* in the C ABI).
// `FALWA_KERNEL("name", stream);` directly before a launch counts it and, when profiling is on,
// brackets everything up to the end of the enclosing scope with two events on that stream.
#define FALWA_CAT2(a, b) a##b
#define FALWA_CAT(a, b) FALWA_CAT2(a, b)
#define FALWA_KERNEL(name, st) falwa::KernelScope FALWA_CAT(_falwa_ks_, __LINE__)(name, st)

namespace falwa {

constexpr int kWarp = 32;

struct ProfileRecord {
    const char* name;
    cudaEvent_t e0, e1;
};
struct ProfileState {
    bool enabled = false;
    long long launches = 0;
    std::vector<ProfileRecord> recs;
};
inline ProfileState& profile_state() {
    static ProfileState s;
    return s;
}
struct KernelScope {
    cudaStream_t st;
    cudaEvent_t e1 = nullptr;
    KernelScope(const char* name, cudaStream_t stream) : st(stream) {
        ProfileState& P = profile_state();
        ++P.launches;
        if (P.enabled) {
            ProfileRecord r;
            r.name = name;
            cudaEventCreate(&r.e0);
            cudaEventCreate(&r.e1);
            cudaEventRecord(r.e0, st);
            e1 = r.e1;
            P.recs.push_back(r);
        }
    }
    ~KernelScope() {
        if (e1) cudaEventRecord(e1, st);
    }
};

// Streaming (read-once) global load: do not pollute L1.
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double* p, double v) { __stcs(p, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; result valid in every thread.  `red` = shared scratch of >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// Order-preserving map double <-> uint64 so that atomicMax/Min on integers gives the exact fp max/min.
__device__ __forceinline__ unsigned long long f64_to_ordered(double d) {
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_to_f64(unsigned long long u) {
    u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)u);
}

// Area-analysis bin of a value: 1 + int((qmax - q) / dq), clamped to [1, nb]  (compute_reference_states.f90:82-86).
// The quotient is first formed with the reciprocal (one DMUL instead of a ~40-instruction DDIV sequence); both
// roundings are within 2 ulp of the true quotient (|t| <= nb), so they can truncate differently only when the
// product lies within ~1e-12 of an integer — then, and only then, the exact division decides.  Bin assignment is
// therefore identical to the reference's arithmetic for every input.
__device__ __noinline__ double area_bin_exact(double x, double dq) { return x / dq; }  // kept out of line: rare

__device__ __forceinline__ int area_bin(double qmax, double q, double dq, double rdq, int nb, bool* near_edge = nullptr) {
    const double x = qmax - q;
    double t = x * rdq;
    const double fr = t - floor(t);
    const bool edge = !(fr > 1e-9 && fr < 1. - 1e-9);
    if (__builtin_expect(edge, 0)) t = area_bin_exact(x, dq);
    if (near_edge) *near_edge = edge;
    return max(1, min(nb, 1 + (int)t));
}

// Small stream-ordered device buffer filled from a host vector (tables of a few KB).
struct DeviceTable {
    double* d = nullptr;
    cudaStream_t s = nullptr;
    int upload(const std::vector<double>& h, cudaStream_t stream) {
        s = stream;
        FALWA_CUDA_TRY(cudaMallocAsync((void**)&d, h.size() * sizeof(double), stream));
        // pageable source: the copy is staged before the call returns, so `h` may die afterwards
        FALWA_CUDA_TRY(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
        return 0;
    }
    ~DeviceTable() {
        if (d) cudaFreeAsync(d, s);
    }
};

// Raw stream-ordered scratch.
template <typename T>
struct DeviceScratch {
    T* d = nullptr;
    cudaStream_t s = nullptr;
    int alloc(size_t n, cudaStream_t stream, bool zero_fill = false) {
        s = stream;
        FALWA_CUDA_TRY(cudaMallocAsync((void**)&d, n * sizeof(T), stream));
        if (zero_fill) FALWA_CUDA_TRY(cudaMemsetAsync(d, 0, n * sizeof(T), stream));
        return 0;
    }
    ~DeviceScratch() {
        if (d) cudaFreeAsync(d, s);
    }
};

inline double host_pi() { return std::acos(-1.0); }

// Integer tuning knob from the environment (benchmark experiments only; defaults are the shipped values).
inline int tune_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace falwa
