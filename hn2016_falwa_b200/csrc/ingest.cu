// Ingest of archived fields: packed int16 (CF scale_factor / add_offset), float32 or float64 values as they lie in
// the file  ->  the float64 [batch][level][lat][lon] layout of the pipeline, latitude ascending, pressure descending.
//
// Reference behaviour restated (not ported): the reference receives already decoded arrays from xarray (CF decoding,
// value = raw * scale_factor + add_offset in float64, _FillValue -> NaN) and reorders them on the host with np.flip
// (src/falwa/xarrayinterface.py:384-405; the recipe for the packed test snapshot is tests/test_output_results.py:26-45).
// Here the raw values cross PCIe (2 bytes per value instead of 8) and are decoded and reordered by one HBM pass.
#include "common.cuh"

namespace falwa {

struct IngestArgs {
    int nlev, nlat, nlon;
    int flip_lat, flip_lev;
    int has_fill;
    double scale, offset, fill;
};

template <typename T> struct IngestPair;
template <> struct IngestPair<short>  { using type = short2; };
template <> struct IngestPair<float>  { using type = float2; };
template <> struct IngestPair<double> { using type = double2; };

// storage byte order -> native (NetCDF-3 files are big-endian)
__device__ __forceinline__ short ingest_bswap(short v) {
    const unsigned u = (unsigned short)v;
    return (short)(unsigned short)(((u & 0xffu) << 8) | (u >> 8));
}
__device__ __forceinline__ float ingest_bswap(float v) {
    return __uint_as_float(__byte_perm(__float_as_uint(v), 0u, 0x0123));
}
__device__ __forceinline__ double ingest_bswap(double v) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    return __hiloint2double((int)__byte_perm(lo, 0u, 0x0123), (int)__byte_perm(hi, 0u, 0x0123));
}

template <typename T, bool SWAP>
__device__ __forceinline__ double ingest_decode(T raw, const IngestArgs& a) {
    if (SWAP) raw = ingest_bswap(raw);
    double x = (double)raw;
    if (!std::is_same<T, double>::value || a.scale != 1.0 || a.offset != 0.0) {
        x = x * a.scale;      // two roundings like the CF decoder (no FMA: the library is built with -fmad=false)
        x = x + a.offset;
    }
    if (a.has_fill && (double)raw == a.fill) x = __longlong_as_double(0x7ff8000000000000LL);
    return x;
}

// grid: (pairs of one level / (256*UNROLL), nlev, batch).  V = 2: one lane converts two neighbouring longitudes per
// access, so a warp reads 128 B (int16) / 256 B (float) / 512 B and writes 512 contiguous bytes.
template <typename T, int V, bool SWAP>
__global__ void __launch_bounds__(256)
ingest_kernel(IngestArgs a, const T* __restrict__ src, double* __restrict__ dst) {
    constexpr int UNROLL = 4;
    const int k = blockIdx.y, b = blockIdx.z;
    const int per_row = a.nlon / V;
    const unsigned n = (unsigned)a.nlat * (unsigned)per_row;
    const int ks = a.flip_lev ? a.nlev - 1 - k : k;
    const size_t level = (size_t)a.nlat * a.nlon;
    const T* s = src + ((size_t)b * a.nlev + ks) * level;
    double* d = dst + ((size_t)b * a.nlev + k) * level;
    const unsigned p0 = blockIdx.x * (256u * UNROLL) + threadIdx.x;
    if (!a.flip_lat) {
        // rows keep their order: the level is one contiguous run
        if (V == 2) {
            using P = typename IngestPair<T>::type;
            P raw[UNROLL];
#pragma unroll
            for (int r = 0; r < UNROLL; ++r) {
                const unsigned p = p0 + r * 256u;
                if (p < n) raw[r] = __ldcs(reinterpret_cast<const P*>(s) + p);
            }
#pragma unroll
            for (int r = 0; r < UNROLL; ++r) {
                const unsigned p = p0 + r * 256u;
                if (p < n) __stcs(reinterpret_cast<double2*>(d) + p,
                                  make_double2(ingest_decode<T, SWAP>(raw[r].x, a), ingest_decode<T, SWAP>(raw[r].y, a)));
            }
        } else {
#pragma unroll
            for (int r = 0; r < UNROLL; ++r) {
                const unsigned p = p0 + r * 256u;
                if (p < n) d[p] = ingest_decode<T, SWAP>(s[p], a);
            }
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < UNROLL; ++r) {
        const unsigned p = p0 + r * 256u;
        if (p >= n) continue;
        const unsigned j = p / (unsigned)per_row, c = p - j * (unsigned)per_row;
        const size_t so = (size_t)(a.nlat - 1 - (int)j) * a.nlon, dofs = (size_t)j * a.nlon;
        if (V == 2) {
            using P = typename IngestPair<T>::type;
            const P raw = __ldcs(reinterpret_cast<const P*>(s + so) + c);
            __stcs(reinterpret_cast<double2*>(d + dofs) + c,
                   make_double2(ingest_decode<T, SWAP>(raw.x, a), ingest_decode<T, SWAP>(raw.y, a)));
        } else {
            d[dofs + c] = ingest_decode<T, SWAP>(s[so + c], a);
        }
    }
}

template <typename T>
static int ingest_launch(const IngestArgs& a, bool swap, int batch, const void* src, double* dst, cudaStream_t st) {
    const bool pairs = (a.nlon % 2 == 0) && ((uintptr_t)src % (2 * sizeof(T)) == 0) && ((uintptr_t)dst % 16 == 0);
    const long long n = (long long)a.nlat * (a.nlon / (pairs ? 2 : 1));
    dim3 grid(ceil_div(n, 256 * 4), a.nlev, batch);
    FALWA_KERNEL("ingest_kernel", st);
    if (pairs && !swap)      ingest_kernel<T, 2, false><<<grid, 256, 0, st>>>(a, (const T*)src, dst);
    else if (pairs && swap)  ingest_kernel<T, 2, true><<<grid, 256, 0, st>>>(a, (const T*)src, dst);
    else if (!swap)          ingest_kernel<T, 1, false><<<grid, 256, 0, st>>>(a, (const T*)src, dst);
    else                     ingest_kernel<T, 1, true><<<grid, 256, 0, st>>>(a, (const T*)src, dst);
    FALWA_LAUNCH_CHECK();
    return 0;
}

}  // namespace falwa

using namespace falwa;

// src_type: 0 float64, 1 float32, 2 int16; big_endian: the stored values are byte-swapped (NetCDF-3 classic files).
// value = raw * scale + offset (float64; pass 1, 0 for plain values);
// raw == fill_value -> NaN when has_fill.  flip_lat / flip_lev reverse the latitude / level axis while copying.
// src and dst are device pointers, [batch][nlev][nlat][nlon] each, and must not overlap.
extern "C" int falwa_b200_ingest_field(int src_type, int batch, int nlev, int nlat, int nlon, const void* src,
                                       int big_endian, double scale, double offset, int has_fill,
                                       double fill_value, int flip_lat, int flip_lev, double* dst, void* stream) {
    if (!src || !dst) return FALWA_ERR_NULL;
    if (batch < 1 || nlev < 1 || nlat < 1 || nlon < 1 || src_type < 0 || src_type > 2) return FALWA_ERR_ARG;
    if (batch > 65535 || nlev > 65535) return FALWA_ERR_UNSUPPORTED;
    IngestArgs a{nlev, nlat, nlon, flip_lat != 0, flip_lev != 0, has_fill != 0, scale, offset, fill_value};
    cudaStream_t st = (cudaStream_t)stream;
    switch (src_type) {
        case 0: return ingest_launch<double>(a, big_endian != 0, batch, src, dst, st);
        case 1: return ingest_launch<float>(a, big_endian != 0, batch, src, dst, st);
        default: return ingest_launch<short>(a, big_endian != 0, batch, src, dst, st);
    }
}
