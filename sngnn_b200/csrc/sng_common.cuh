// Shared helpers for libsng.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sng.h"

namespace sng {

void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaPeekAtLastError -> SNG_ERR_CUDA + message
int sm_count();                       // cached; <=0 if no device
int debug_env_int(const char* name, int lo, int hi);   // tuning override; always 0 unless sng_set_debug_env(1) was called

#define SNG_REQUIRE(cond, ...)                                  \
    do {                                                        \
        if (!(cond)) {                                          \
            ::sng::set_error(__VA_ARGS__);                      \
            return SNG_ERR_ARG;                                 \
        }                                                       \
    } while (0)

constexpr float kNormEps = 1e-12f;   // F.normalize eps, R: models/models.py:122

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
    acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
// butterfly sum over the G lanes of a group (G power of two, groups are lane-aligned)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the 32/G groups of a warp (lanes with equal lane % G)
template <int G>
__device__ __forceinline__ float cross_group_sum(float v) {
#pragma unroll
    for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// 1 / max(||x||, 1e-12) from the squared norm (one MUFU.RSQ; 2 ulp, far inside the 1e-5 contract)
__device__ __forceinline__ float inv_norm_of(float ss) { return rsqrtf(fmaxf(ss, kNormEps * kNormEps)); }

// lanes-per-row group size for a padded channel count c (multiple of 4): smallest power of two >= c/4
inline int group_lanes(int64_t c) {
    int g = 1;
    while (g * 4 < c) g <<= 1;
    return g;
}

}  // namespace sng
