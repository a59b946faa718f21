// Helpers shared by the decode kernels (decode.cu, decode_pipe.cu): 128-bit packs, streaming accesses,
// the DFL side reduction.  sm_100a only.
#pragma once
#include "cerb_kernels.h"

#ifndef DEC_THREADS
#define DEC_THREADS 128
#endif
#ifndef DEC_MINB
#define DEC_MINB 5  // 96 registers -> 20 warps per SM: best of the measured variants (profiles/r01_decode.md)
#endif
#define CLS_CHUNK 32
#define LOG2E_F 1.4426950408889634f

template <int BYTES> struct RawOf;
template <> struct RawOf<16> { typedef uint4 type; };
template <> struct RawOf<8> { typedef uint2 type; };
template <> struct RawOf<4> { typedef uint32_t type; };
template <> struct RawOf<2> { typedef uint16_t type; };

template <typename T, int VEC> union Pack {
    typename RawOf<sizeof(T) * VEC>::type raw;
    T e[VEC];
};

template <typename T, int VEC> __device__ __forceinline__ Pack<T, VEC> load_pack(const T* p) {
    Pack<T, VEC> r;
    if constexpr (sizeof(T) * VEC == 16) {
        r.raw = ldg_stream16(p);
    } else {
        r.raw = __ldg(reinterpret_cast<const typename RawOf<sizeof(T) * VEC>::type*>(p));
    }
    return r;
}
template <typename T, int VEC> __device__ __forceinline__ void store_pack(T* p, const Pack<T, VEC>& v) {
    if constexpr (sizeof(T) * VEC == 16) {
        stg_stream16(p, v.raw);
    } else {
        *reinterpret_cast<typename RawOf<sizeof(T) * VEC>::type*>(p) = v.raw;
    }
}

// maximum of the VEC stored (already rounded) scores of one vector; scores are sigmoids, so no NaN unless the
// logit was NaN, which fmaxf / __hmax2 drop -- a NaN score is never a candidate either
template <typename T, int VEC> __device__ __forceinline__ T pack_max(const Pack<T, VEC>& v);
template <> __device__ __forceinline__ __half pack_max<__half, 8>(const Pack<__half, 8>& v) {
    const __half2* h = reinterpret_cast<const __half2*>(&v.raw);
    const __half2 m = __hmax2(__hmax2(h[0], h[1]), __hmax2(h[2], h[3]));
    return __hmax(__low2half(m), __high2half(m));
}
template <> __device__ __forceinline__ float pack_max<float, 4>(const Pack<float, 4>& v) {
    return fmaxf(fmaxf(v.e[0], v.e[1]), fmaxf(v.e[2], v.e[3]));
}
template <typename T, int VEC> __device__ __forceinline__ T pack_max(const Pack<T, VEC>& v) { return v.e[0]; }

// Expected DFL distance of one box side for VEC anchors: sum_k k * softmax(logits)_k.
// Rounding points follow the reference for half tensors: probabilities are rounded to
// half (softmax output), the 1x1 conv accumulates in fp32 and rounds once.
template <typename T, int VEC>
__device__ __forceinline__ void dfl_load(const T* __restrict__ side_base, int hw, Pack<T, VEC> (&v)[CERB_REG_MAX]) {
#pragma unroll
    for (int k = 0; k < CERB_REG_MAX; ++k) v[k] = load_pack<T, VEC>(side_base + (size_t)k * hw);
}
template <typename T, int VEC>
__device__ __forceinline__ void dfl_reduce(const Pack<T, VEC> (&v)[CERB_REG_MAX], float (&d)[VEC]) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        float x[CERB_REG_MAX];
#pragma unroll
        for (int k = 0; k < CERB_REG_MAX; ++k) x[k] = to_f32<T>(v[k].e[i]);
        d[i] = dfl_expectation<T>(x);
    }
}
template <typename T, int VEC>
__device__ __forceinline__ void dfl_side(const T* __restrict__ side_base, int hw, float (&d)[VEC]) {
    Pack<T, VEC> v[CERB_REG_MAX];
    dfl_load<T, VEC>(side_base, hw, v);
    dfl_reduce<T, VEC>(v, d);
}
