// Device-side scalar/vector helpers for the sm_100a radiance kernels.
//
// Arithmetic policy (DESIGN.md "Numerics"): the translation unit is compiled with -fmad=false, IEEE
// division and square root, so that every expression evaluates exactly as written (the CPU oracle is
// compiled with -ffp-contract=off and evaluates the same expressions).  Fused multiply-adds appear only
// where fmaf() is spelled out — in the BVH traversal, which has no reference arithmetic to follow.
// Evaluation order of the helpers below is LuisaCompute's device math
// (src/compute/src/backends/cuda/cuda_builtin/cuda_device_math.h in the reference tree).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lrk {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 v3(float x, float y, float z) { return {x, y, z}; }
__device__ __forceinline__ V3 v3(float s) { return {s, s, s}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator+(V3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
__device__ __forceinline__ V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }// true divisions, as the oracle's
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// A product that the compiler must not fuse into a neighbouring add (shade.cu is compiled with FMA contraction): where a
// difference of two products is EXACTLY zero in IEEE arithmetic - every cross product of axis-aligned edges is - a fused
// multiply-add leaves the rounding error of one product instead, and the sign of that residue then picks a branch (the
// orthonormal-basis construction of a shading frame switches on the sign of n.z: a wall whose normal has n.z = 0 exactly would
// get another - equally valid, but different - frame than the oracle's, and with it other sampled directions).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float mul_exact(float a, float b) { return __fmul_rn(a, b); }
#else
#if defined(__GNUC__) && !defined(__clang__)
__attribute__((optimize("fp-contract=off")))
#endif
inline float mul_exact(float a, float b) { return a * b; }
#endif
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return {mul_exact(a.y, b.z) - mul_exact(b.y, a.z), mul_exact(a.z, b.x) - mul_exact(b.z, a.x), mul_exact(a.x, b.y) - mul_exact(b.x, a.y)};
}
__device__ __forceinline__ float length(V3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
__device__ __forceinline__ float sqr(float x) { return x * x; }
__device__ __forceinline__ float saturate(float x) { return fminf(fmaxf(x, 0.f), 1.f); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float lerp(float a, float b, float t) { return t * (b - a) + a; }
__device__ __forceinline__ V3 lerp(V3 a, V3 b, float t) { return t * (b - a) + a; }
// The reference backends' builtin pow (cuda_device_math.h:21-36): a whole-number exponent (decided at run time) means
// square-and-multiply with exactly rounded products, anything else powf.  A gamma-2.0 texture decode is x * x, not powf(x, 2).
__device__ inline float builtin_pow(float x, float y) {
    const int n = static_cast<int>(y);
    if (static_cast<float>(n) != y) return powf(x, y);
    float acc = 1.0f, base = x;
    for (unsigned bits = n < 0 ? 0u - static_cast<unsigned>(n) : static_cast<unsigned>(n); bits != 0u; bits >>= 1) {
        if (bits & 1u) acc *= base;
        base *= base;
    }
    return n < 0 ? 1.0f / acc : acc;
}
__device__ __forceinline__ float sign(float x) { return copysignf(1.0f, x); }// never 0
__device__ __forceinline__ V3 reflect(V3 v, V3 n) { return v - 2.0f * dot(v, n) * n; }
__device__ __forceinline__ V3 face_forward(V3 v, V3 n) { return dot(v, n) < 0.f ? -v : v; }
__device__ __forceinline__ float max3(V3 a) { return fmaxf(fmaxf(fmaxf(0.f, a.x), a.y), a.z); }
__device__ __forceinline__ bool any_nonzero(V3 w) { return w.x != 0.f || w.y != 0.f || w.z != 0.f; }

constexpr float kPi = 3.14159265358979323846264338327950288f;
constexpr float kPiOverTwo = 1.57079632679489661923132169163975144f;
constexpr float kPiOverFour = 0.785398163397448309615660845819875721f;
constexpr float kInvPi = 0.318309886183790671537767526745028724f;
constexpr float kOneMinusEpsilon = 0x1.fffffep-1f;
constexpr float kFltMax = 3.402823466e+38f;

// explicit-FMA forms (traversal only)
__device__ __forceinline__ float fdot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 fcross(V3 a, V3 b) {
    return {fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x))};
}

// ---- RNG: reference src/util/rng.cpp:53-68,128-140 ; src/samplers/independent.cpp:57-82 ----------
__device__ __forceinline__ uint32_t rotl32(uint32_t x, uint32_t r) { return (x << r) | (x >> (32u - r)); }

__device__ __forceinline__ uint32_t xxhash32_uint4(uint32_t px, uint32_t py, uint32_t pz, uint32_t pw) {
    constexpr uint32_t PRIME32_2 = 2246822519u, PRIME32_3 = 3266489917u;
    constexpr uint32_t PRIME32_4 = 668265263u, PRIME32_5 = 374761393u;
    uint32_t h32 = pw + PRIME32_5 + px * PRIME32_3;
    h32 = PRIME32_4 * rotl32(h32, 17u);
    h32 += py * PRIME32_3;
    h32 = PRIME32_4 * rotl32(h32, 17u);
    h32 += pz * PRIME32_3;
    h32 = PRIME32_4 * rotl32(h32, 17u);
    h32 = PRIME32_2 * (h32 ^ (h32 >> 15u));
    h32 = PRIME32_3 * (h32 ^ (h32 >> 13u));
    return h32 ^ (h32 >> 16u);
}

__device__ __forceinline__ uint32_t xxhash32_uint3(uint32_t px, uint32_t py, uint32_t pz) {// src/util/rng.cpp:38-51
    constexpr uint32_t PRIME32_2 = 2246822519u, PRIME32_3 = 3266489917u, PRIME32_4 = 668265263u, PRIME32_5 = 374761393u;
    uint32_t h32 = pz + PRIME32_5 + px * PRIME32_3;
    h32 = PRIME32_4 * ((h32 << 17u) | (h32 >> 15u));
    h32 += py * PRIME32_3;
    h32 = PRIME32_4 * ((h32 << 17u) | (h32 >> 15u));
    h32 = PRIME32_2 * (h32 ^ (h32 >> 15u));
    h32 = PRIME32_3 * (h32 ^ (h32 >> 13u));
    return h32 ^ (h32 >> 16u);
}
__device__ __forceinline__ float lcg(uint32_t &state) {
    state = 1664525u * state + 1013904223u;
    return fminf(kOneMinusEpsilon, static_cast<float>(state) * 0x1p-32f);
}

}// namespace lrk
