// Shared device helpers for the gnntf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gnntf_b200.h"

#define GNNTF_CUDA_TRY(expr)                      \
    do {                                          \
        cudaError_t _e = (expr);                  \
        if (_e != cudaSuccess) return (int)_e;    \
    } while (0)

// Kernel launches report configuration errors through cudaGetLastError: it does not synchronise,
// and it CLEARS a non-sticky error (bad grid, shared-memory attribute), so one recoverable
// configuration error is reported once instead of poisoning every later call of this thread (the
// library links cudart statically: nobody else can clear this runtime's error state).
#define GNNTF_LAUNCH_CHECK()                      \
    do {                                          \
        cudaError_t _e = cudaGetLastError();      \
        if (_e != cudaSuccess) return (int)_e;    \
    } while (0)

namespace gnntf {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

// Streaming loads: CSR arrays and the teleport term are touched once per step, so they must
// not displace gathered feature rows from L1.
__device__ __forceinline__ int ld_stream(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
// Gathered feature rows: read-only path, allocate in L1 (neighbouring rows share neighbours).
__device__ __forceinline__ float ld_gather(const float* p) { return __ldg(p); }
__device__ __forceinline__ float4 ld_gather4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
// Output rows are written once and next read by another kernel: do not allocate in L1.
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ static Vec gather(const float* p) { return Vec{{ld_gather(p)}}; }
    __device__ __forceinline__ static Vec stream(const float* p) { return Vec{{ld_stream(p)}}; }
    __device__ __forceinline__ void store(float* p) const { st_stream(p, v[0]); }
};
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ static Vec gather(const float* p) {
        float4 t = ld_gather4(p);
        return Vec{{t.x, t.y, t.z, t.w}};
    }
    __device__ __forceinline__ static Vec stream(const float* p) {
        float4 t = ld_stream4(p);
        return Vec{{t.x, t.y, t.z, t.w}};
    }
    __device__ __forceinline__ void store(float* p) const {
        st_stream4(p, make_float4(v[0], v[1], v[2], v[3]));
    }
    __device__ __forceinline__ float4 as_float4() const { return make_float4(v[0], v[1], v[2], v[3]); }
};

}  // namespace gnntf
