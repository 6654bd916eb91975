// types.cuh — value types of the kernels (float, double, complex64, complex128)
// behind one tiny arithmetic vocabulary, so every kernel is written once and
// instantiated for the s/d/c/z letters of the MKL routine it replaces.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace sdb {

struct alignas(8) cf32 {
    float re, im;
};
struct alignas(16) cf64 {
    double re, im;
};

template <typename T> struct Num;

template <> struct Num<float> {
    static constexpr int dtype = 0;
    __host__ __device__ static float zero() { return 0.f; }
    __host__ __device__ static float make(double re, double) { return float(re); }
    __host__ __device__ static bool is_zero(float v) { return v == 0.f; }
    __host__ __device__ static bool is_one(float v) { return v == 1.f; }
};
template <> struct Num<double> {
    static constexpr int dtype = 1;
    __host__ __device__ static double zero() { return 0.0; }
    __host__ __device__ static double make(double re, double) { return re; }
    __host__ __device__ static bool is_zero(double v) { return v == 0.0; }
    __host__ __device__ static bool is_one(double v) { return v == 1.0; }
};
template <> struct Num<cf32> {
    static constexpr int dtype = 2;
    __host__ __device__ static cf32 zero() { return cf32{0.f, 0.f}; }
    __host__ __device__ static cf32 make(double re, double im) { return cf32{float(re), float(im)}; }
    __host__ __device__ static bool is_zero(cf32 v) { return v.re == 0.f && v.im == 0.f; }
    __host__ __device__ static bool is_one(cf32 v) { return v.re == 1.f && v.im == 0.f; }
};
template <> struct Num<cf64> {
    static constexpr int dtype = 3;
    __host__ __device__ static cf64 zero() { return cf64{0.0, 0.0}; }
    __host__ __device__ static cf64 make(double re, double im) { return cf64{re, im}; }
    __host__ __device__ static bool is_zero(cf64 v) { return v.re == 0.0 && v.im == 0.0; }
    __host__ __device__ static bool is_one(cf64 v) { return v.re == 1.0 && v.im == 0.0; }
};

// c + a*b
__device__ __forceinline__ float madd(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double madd(double a, double b, double c) { return fma(a, b, c); }
__device__ __forceinline__ cf32 madd(cf32 a, cf32 b, cf32 c) {
    c.re = fmaf(a.re, b.re, c.re);
    c.re = fmaf(-a.im, b.im, c.re);
    c.im = fmaf(a.re, b.im, c.im);
    c.im = fmaf(a.im, b.re, c.im);
    return c;
}
__device__ __forceinline__ cf64 madd(cf64 a, cf64 b, cf64 c) {
    c.re = fma(a.re, b.re, c.re);
    c.re = fma(-a.im, b.im, c.re);
    c.im = fma(a.re, b.im, c.im);
    c.im = fma(a.im, b.re, c.im);
    return c;
}

__device__ __forceinline__ float mul(float a, float b) { return a * b; }
__device__ __forceinline__ double mul(double a, double b) { return a * b; }
__device__ __forceinline__ cf32 mul(cf32 a, cf32 b) { return madd(a, b, cf32{0.f, 0.f}); }
__device__ __forceinline__ cf64 mul(cf64 a, cf64 b) { return madd(a, b, cf64{0.0, 0.0}); }

__device__ __forceinline__ float add(float a, float b) { return a + b; }
__device__ __forceinline__ double add(double a, double b) { return a + b; }
__device__ __forceinline__ cf32 add(cf32 a, cf32 b) { return cf32{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cf64 add(cf64 a, cf64 b) { return cf64{a.re + b.re, a.im + b.im}; }

__device__ __forceinline__ float conj_(float a) { return a; }
__device__ __forceinline__ double conj_(double a) { return a; }
__device__ __forceinline__ cf32 conj_(cf32 a) { return cf32{a.re, -a.im}; }
__device__ __forceinline__ cf64 conj_(cf64 a) { return cf64{a.re, -a.im}; }

// read-only (non-coherent) scalar loads
__device__ __forceinline__ float ldg(const float* p) { return __ldg(p); }
__device__ __forceinline__ double ldg(const double* p) { return __ldg(p); }
__device__ __forceinline__ cf32 ldg(const cf32* p) {
    const float2 q = __ldg(reinterpret_cast<const float2*>(p));
    return cf32{q.x, q.y};
}
__device__ __forceinline__ cf64 ldg(const cf64* p) {
    const double2 q = __ldg(reinterpret_cast<const double2*>(p));
    return cf64{q.x, q.y};
}

// streaming loads (evict-first): data read exactly once, keep it from displacing reusable lines
__device__ __forceinline__ float ldcs(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ldcs(const double* p) { return __ldcs(p); }
__device__ __forceinline__ cf32 ldcs(const cf32* p) {
    const float2 q = __ldcs(reinterpret_cast<const float2*>(p));
    return cf32{q.x, q.y};
}
__device__ __forceinline__ cf64 ldcs(const cf64* p) {
    const double2 q = __ldcs(reinterpret_cast<const double2*>(p));
    return cf64{q.x, q.y};
}

// L2-coherent loads (bypass L1): for data other threads update with atomics
__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ cf32 ldcg(const cf32* p) {
    const float2 q = __ldcg(reinterpret_cast<const float2*>(p));
    return cf32{q.x, q.y};
}
__device__ __forceinline__ cf64 ldcg(const cf64* p) {
    const double2 q = __ldcg(reinterpret_cast<const double2*>(p));
    return cf64{q.x, q.y};
}

// atomic accumulate (no return value needed -> RED in SASS)
__device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(cf32* p, cf32 v) {
    atomicAdd(&p->re, v.re);
    atomicAdd(&p->im, v.im);
}
__device__ __forceinline__ void atomic_add(cf64* p, cf64 v) {
    atomicAdd(&p->re, v.re);
    atomicAdd(&p->im, v.im);
}

// warp shuffles for every value type
__device__ __forceinline__ float shfl(unsigned m, float v, int src, int w) { return __shfl_sync(m, v, src, w); }
__device__ __forceinline__ double shfl(unsigned m, double v, int src, int w) { return __shfl_sync(m, v, src, w); }
__device__ __forceinline__ cf32 shfl(unsigned m, cf32 v, int src, int w) {
    return cf32{__shfl_sync(m, v.re, src, w), __shfl_sync(m, v.im, src, w)};
}
__device__ __forceinline__ cf64 shfl(unsigned m, cf64 v, int src, int w) {
    return cf64{__shfl_sync(m, v.re, src, w), __shfl_sync(m, v.im, src, w)};
}

// Host-side dispatch on the runtime dtype code: calls f(T{}) with the static type.
#define SDB_DISPATCH_DTYPE(dtype, T, ...)                                       \
    [&]() -> sdb_status {                                                       \
        switch (dtype) {                                                        \
            case SDB_F32: { using T = float; return __VA_ARGS__(); }            \
            case SDB_F64: { using T = double; return __VA_ARGS__(); }           \
            case SDB_C64: { using T = ::sdb::cf32; return __VA_ARGS__(); }      \
            case SDB_C128: { using T = ::sdb::cf64; return __VA_ARGS__(); }     \
            default:                                                            \
                ::sdb::set_error("unsupported dtype code %d", int(dtype));      \
                return SDB_STATUS_NOT_SUPPORTED;                                \
        }                                                                       \
    }()

}  // namespace sdb
