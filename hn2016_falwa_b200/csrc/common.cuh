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
#include <cstring>
#include <deque>
#include <mutex>
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
    std::atomic<long long> launches{0};   // kernels launched by this library (any host thread)
    std::mutex mu;                        // guards recs (the C ABI may be driven from one host thread per stream)
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
            std::lock_guard<std::mutex> lock(P.mu);
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

// ---- small host <-> device transfers that bypass the copy engines ------------------------------------------------
// The pipeline's per-batch tables (static-stability profiles, per-problem argument blocks, a few KB) used to travel
// with cudaMemcpyAsync.  On a stream that overlaps with bulk PCIe traffic of other streams (HostStream / QGDataset:
// hundreds of MB per chunk in each direction) those tiny copies queue in the SAME DMA engine behind the bulk
// transfer, and every stage stalled ~10 ms for it.  Instead the host writes into a pinned, device-mapped ring and a
// small kernel on the compute stream pulls the bytes over PCIe (and, in the other direction, kernels store results
// straight into mapped host memory), so no copy engine is involved.
struct PinnedRing {
    static constexpr size_t kCap = 32u << 20;
    struct Fence { size_t b, e; cudaEvent_t ev; };
    char* base = nullptr;
    size_t head = 0;
    std::mutex mu;
    std::deque<Fence> fences;
    std::vector<cudaEvent_t> pool;

    // -> host pointer (== device pointer under unified addressing) of `bytes` ring bytes, or nullptr when the request
    // is too large for the ring (callers fall back to cudaMemcpyAsync).  Blocks only if the region is still in use.
    void* take(size_t bytes, size_t* off) {
        if (bytes > kCap / 4) return nullptr;
        if (!base && cudaHostAlloc((void**)&base, kCap, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
            base = nullptr;
            (void)cudaGetLastError();
            return nullptr;
        }
        const size_t sz = (bytes + 255) & ~(size_t)255;
        if (head + sz > kCap) head = 0;
        const size_t b = head, e = head + sz;
        for (auto it = fences.begin(); it != fences.end();) {
            if (it->b < e && b < it->e) {
                cudaEventSynchronize(it->ev);
                pool.push_back(it->ev);
                it = fences.erase(it);
            } else ++it;
        }
        head = e;
        *off = b;
        return base + b;
    }
    // the ring bytes [off, off + bytes) are in use until everything queued on `st` so far has run
    void fence(size_t off, size_t bytes, cudaStream_t st) {
        cudaEvent_t ev;
        bool ok = true;
        if (!pool.empty()) { ev = pool.back(); pool.pop_back(); }
        else ok = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
        if (ok && cudaEventRecord(ev, st) != cudaSuccess) { pool.push_back(ev); ok = false; }
        if (!ok) {   // no usable event: the bytes are free once the stream has drained
            (void)cudaGetLastError();
            cudaStreamSynchronize(st);
            return;
        }
        fences.push_back({off, off + ((bytes + 255) & ~(size_t)255), ev});
    }
};
// One ring (and one event pool) per device: an event belongs to the device it was created on, and a process may
// drive several GPUs through the C ABI.
inline PinnedRing& pinned_ring() {
    static PinnedRing rings[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); dev = 0; }
    return rings[dev & 15];
}

__global__ void __launch_bounds__(256)
pull_words_kernel(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src, size_t n) {
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
        dst[t] = src[t];
}

// dst (device) <- src (any host memory), stream-ordered; `src` may die when the call returns.
inline int upload_bytes(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    PinnedRing& R = pinned_ring();
    if (bytes % 8 == 0) {
        std::lock_guard<std::mutex> lock(R.mu);
        size_t off = 0;
        if (void* stage = R.take(bytes, &off)) {
            std::memcpy(stage, src, bytes);
            const size_t n = bytes / 8;
            const int blocks = (int)std::min<size_t>((n + 255) / 256, 296);
            { FALWA_KERNEL("pull_words_kernel", st); pull_words_kernel<<<blocks, 256, 0, st>>>((unsigned long long*)dst, (const unsigned long long*)stage, n); }
            FALWA_LAUNCH_CHECK();
            R.fence(off, bytes, st);
            return 0;
        }
    }
    // pageable source: the copy is staged before the call returns
    FALWA_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
}

// Small stream-ordered device buffer filled from a host vector (tables of a few KB).
struct DeviceTable {
    double* d = nullptr;
    cudaStream_t s = nullptr;
    int upload(const std::vector<double>& h, cudaStream_t stream) {
        s = stream;
        FALWA_CUDA_TRY(cudaMallocAsync((void**)&d, h.size() * sizeof(double), stream));
        return upload_bytes(d, h.data(), h.size() * sizeof(double), stream);
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
