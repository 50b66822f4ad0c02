// The samplers of the path integrator: src/base/sampler.h:42-48 as implemented by
// src/samplers/independent.cpp:57-82, pmj02bn.cpp:46-213, sobol.cpp:38-170, padded_sobol.cpp:34-146, zsobol.cpp:40-177.
//
// A path carries ONE 32-bit word of sampler state through the wavefront queues (PathBuffers::id_rng.y): the LCG state of the
// Independent sampler, or the dimension counter of a table-driven one - everything else those samplers keep per path (pixel,
// sample index, Sobol' index) is a function of the path's generation slot and is recomputed where numbers are drawn.  The sampler
// type is uniform over a launch, so the switch below never diverges; the table-driven bodies are out of line so that the
// Independent path of the headline configurations keeps its register allocation.  All of it is integer arithmetic plus one
// int-to-float conversion: bit-identical to the oracle (oracle/oracle.cpp `struct Sampler`), which is bit-identical to the
// reference's renders with each sampler (tests/test_ref_render.py).
#pragma once
#include "scene.cuh"

namespace lrk {

__device__ __forceinline__ uint32_t xxhash32_uint2(uint32_t px, uint32_t py) {
    constexpr uint32_t PRIME32_2 = 2246822519u, PRIME32_3 = 3266489917u, PRIME32_4 = 668265263u, PRIME32_5 = 374761393u;
    uint32_t h32 = py + PRIME32_5 + px * PRIME32_3;
    h32 = PRIME32_4 * rotl32(h32, 17u);
    h32 = PRIME32_2 * (h32 ^ (h32 >> 15u));
    h32 = PRIME32_3 * (h32 ^ (h32 >> 13u));
    return h32 ^ (h32 >> 16u);
}

__device__ __forceinline__ uint32_t permutation_element(uint32_t i, uint32_t l, uint32_t w, uint32_t p) {// pmj02bn.cpp:60-86
    do {
        i ^= p;
        i *= 0xe170893du;
        i ^= p >> 16u;
        i ^= (i & w) >> 4u;
        i ^= p >> 8u;
        i *= 0x0929eb3fu;
        i ^= p >> 23u;
        i ^= (i & w) >> 1u;
        i *= 1u | p >> 27u;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11u;
        i *= 0x74dcb303u;
        i ^= (i & w) >> 2u;
        i *= 0x9e501cc3u;
        i ^= (i & w) >> 2u;
        i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5u;
    } while (i >= l);
    return (i + p) % l;
}

__device__ __forceinline__ uint32_t fast_owen_scramble(uint32_t seed, uint32_t v) {// sobol.cpp:40-48
    v = __brev(v);
    v ^= v * 0x3d20adeau;
    v += seed;
    v *= (seed >> 16u) | 1u;
    v ^= v * 0x05526c56u;
    v ^= v * 0x53a22864u;
    return __brev(v);
}

__device__ __forceinline__ uint32_t mask_covering(uint32_t x) {// the bit mask covering x (x = l - 1)
    x |= x >> 1u; x |= x >> 2u; x |= x >> 4u; x |= x >> 8u; x |= x >> 16u;
    return x;
}

// What a kernel hands to the sampler: three scalars out of its DeviceScene parameter (never a reference to the parameter itself -
// a kernel parameter whose address is taken is copied to every thread's local memory).
struct SamplerRef {
    uint32_t type, seed;
    const lrk_sampler *rec;
};
__device__ __forceinline__ SamplerRef sampler_ref(const DeviceScene &sc) { return SamplerRef{sc.sampler_type, sc.sampler_seed, sc.sampler}; }

struct PathSampler {
    uint32_t state;// Independent: LCG state; table-driven samplers: the dimension counter
    uint32_t px, py, sample_index;
    uint64_t index;// Sobol: index of the sample in the global sequence; ZSobol: Morton index

    // ---- table-driven bodies (out of line) ------------------------------------------------------------------------------------
    __device__ __noinline__ static uint32_t sobol_bits(const lrk_sampler &c, uint64_t a, uint32_t dim) {// sobol.cpp:52-62
        uint32_t v = 0u;
        for (uint32_t i = dim * 52u; a != 0u; a >>= 1u, i++)
            if (a & 1u) v ^= __ldg(c.sobol_matrices + i);
        return v;
    }
    __device__ __forceinline__ static float blue_noise(const lrk_sampler &c, uint32_t tex_index, uint32_t x, uint32_t y) {// pmj02bn.cpp:46-51
        const uint32_t u = y % 128u, v = x % 128u, wv = tex_index % 48u;
        return static_cast<float>(__ldg(c.blue_noise + (static_cast<size_t>(wv) * 128u + v) * 128u + u)) / 65535.f;
    }
    __device__ __forceinline__ static uint64_t mix_bits(uint64_t v) {// zsobol.cpp:112-119
        v ^= v >> 31u;
        v *= 0x7fb5d329728ea185ull;
        v ^= v >> 27u;
        v *= 0x81dadef4bc2dd44dull;
        v ^= (v >> 32u) >> 1u;
        return v;
    }
    __device__ __forceinline__ static uint64_t left_shift2(uint64_t x) {// zsobol.cpp:144-152
        x = (x ^ (x << 16u)) & 0x0000ffff0000ffffull;
        x = (x ^ (x << 8u)) & 0x00ff00ff00ff00ffull;
        x = (x ^ (x << 4u)) & 0x0f0f0f0f0f0f0f0full;
        x = (x ^ (x << 2u)) & 0x3333333333333333ull;
        x = (x ^ (x << 1u)) & 0x5555555555555555ull;
        return x;
    }
    __device__ __noinline__ static uint64_t zsobol_sample_index(const lrk_sampler &c, uint64_t index, uint32_t dimension) {// zsobol.cpp:104-140
        // the 24 permutations of four digits, one byte each (two bits per digit, digit 0 in the low bits)
        constexpr uint8_t perms[24] = {
            0xe4, 0xb4, 0xd8, 0x78, 0x6c, 0x9c, 0xe1, 0xb1, 0xc9, 0x39, 0x2d, 0x8d,
            0xc6, 0x36, 0xd2, 0x72, 0x4e, 0x1e, 0x27, 0x87, 0x1b, 0x4b, 0x63, 0x93};
        uint64_t sample = 0u;
        const bool pow2_samples = (c.log2_spp & 1u) != 0u;
        const int last_digit = pow2_samples ? 1 : 0;
        const uint64_t morton = index;
        for (int i = static_cast<int>(c.num_base4_digits) - 1; i >= last_digit; i--) {
            const uint32_t digit_shift = 2u * static_cast<uint32_t>(i) - (pow2_samples ? 1u : 0u);
            const uint32_t digit = static_cast<uint32_t>(morton >> digit_shift) & 3u;
            const uint64_t higher = morton >> (digit_shift + 2u);
            const uint32_t p = static_cast<uint32_t>((mix_bits(higher ^ static_cast<uint64_t>(dimension * 0x55555555u)) >> 24u) % 24u);
            sample |= static_cast<uint64_t>((perms[p] >> (2u * digit)) & 3u) << digit_shift;
        }
        if (pow2_samples) sample |= (morton & 1u) ^ (mix_bits((morton >> 1u) ^ static_cast<uint64_t>(dimension * 0x55555555u)) & 1u);
        return sample;
    }
    // the per-path values a table-driven sampler derives from (pixel, sample index): Sampler::Instance::start
    __device__ __noinline__ static uint64_t derive(const lrk_sampler &c, uint32_t px, uint32_t py, uint32_t sample_index) {
        uint64_t index = 0u;
        if (c.type == LRK_SAMPLER_SOBOL) {// sobol.cpp:64-96,132-137
            uint32_t m = 0u;
            while ((1u << m) < c.scale) m++;
            if (m == 0u) return sample_index;
            uint64_t idx = static_cast<uint64_t>(sample_index) << (2u * m);
            uint64_t delta = 0u;
            uint32_t frame = sample_index;
            for (uint32_t k = 0u; frame != 0u; frame >>= 1u, k++)
                if (frame & 1u) delta ^= __ldg(c.vdc + k);
            uint64_t b = delta ^ ((static_cast<uint64_t>(px) << m) | py);
            for (uint32_t d = 0u; b != 0u; b >>= 1u, d++)
                if (b & 1u) idx ^= __ldg(c.vdc_inv + d);
            index = idx;
        } else if (c.type == LRK_SAMPLER_ZSOBOL) {// zsobol.cpp:142-159
            index = (((left_shift2(py) << 1u) | left_shift2(px)) << c.log2_spp) | sample_index;
        }
        return index;
    }
    // by value in, by value out (number(s) + the advanced dimension counter): nothing of the caller's state has its address taken,
    // so the Independent path keeps its LCG word in a register
    struct Draw {
        float x, y;
        uint32_t dimension;
    };
    __device__ __noinline__ static Draw table_1d(const lrk_sampler &c, uint32_t seed, uint32_t dimension, uint32_t px, uint32_t py,
                                                 uint32_t sample_index, uint64_t index) {
        float result;
        switch (c.type) {
            case LRK_SAMPLER_PMJ02BN: {// pmj02bn.cpp:179-191
                const uint32_t hash = xxhash32_uint4(px, py, dimension, seed);
                const uint32_t idx = permutation_element(sample_index, c.spp, c.w, hash);
                const float delta = blue_noise(c, dimension, px, py);
                const float u = (static_cast<float>(idx) + delta) * (1.f / static_cast<float>(c.spp));
                dimension += 1u;
                result = fminf(fmaxf(u, 0.f), kOneMinusEpsilon);
                break;
            }
            case LRK_SAMPLER_SOBOL: {// sobol.cpp:148-154
                if (dimension >= 1024u) dimension = 2u;
                const uint32_t hash = xxhash32_uint2(dimension, seed);
                const float u = static_cast<float>(fast_owen_scramble(hash, sobol_bits(c, index, dimension))) * 0x1p-32f;
                dimension += 1u;
                result = fminf(fmaxf(u, 0.f), kOneMinusEpsilon);
                break;
            }
            case LRK_SAMPLER_PADDED_SOBOL: {// padded_sobol.cpp:124-133
                const uint32_t hash = xxhash32_uint4(px, py, sample_index ^ seed, dimension);
                const uint32_t idx = permutation_element(sample_index, c.spp, mask_covering(c.spp - 1u), hash);
                dimension += 1u;
                result = fminf(static_cast<float>(fast_owen_scramble(hash, sobol_bits(c, idx, 0u))) * 0x1p-32f, kOneMinusEpsilon);
                break;
            }
            default: {// ZSOBOL, zsobol.cpp:164-169
                const uint64_t si = zsobol_sample_index(c, index, dimension);
                const uint32_t hash = __ldg(c.zsobol_hash + dimension * 2u);
                dimension = (dimension + 1u) % 1024u;
                result = fminf(static_cast<float>(fast_owen_scramble(hash, sobol_bits(c, si, 0u))) * 0x1p-32f, kOneMinusEpsilon);
                break;
            }
        }
        return Draw{result, 0.f, dimension};
    }
    __device__ __noinline__ static Draw table_2d(const lrk_sampler &c, uint32_t seed, uint32_t dimension, uint32_t px, uint32_t py,
                                                 uint32_t sample_index, uint64_t index) {
        float2 result;
        switch (c.type) {
            case LRK_SAMPLER_PMJ02BN: {// pmj02bn.cpp:192-207
                uint32_t idx = sample_index;
                const uint32_t pmj_instance = dimension / 2u;
                if (pmj_instance >= 5u) idx = permutation_element(sample_index, c.spp, c.w, xxhash32_uint4(px, py, dimension, seed));
                const uint32_t *e = c.pmj_samples + (static_cast<size_t>(pmj_instance % 5u) * 65536u + idx) * 2u;
                const float u0 = static_cast<float>(__ldg(e)) * 0x1p-32f + blue_noise(c, dimension, px, py);
                const float u1 = static_cast<float>(__ldg(e + 1)) * 0x1p-32f + blue_noise(c, dimension + 1u, px, py);
                dimension += 2u;
                result = make_float2(u0 - floorf(u0), u1 - floorf(u1));
                break;
            }
            case LRK_SAMPLER_SOBOL: {// sobol.cpp:155-163
                if (dimension + 1u >= 1024u) dimension = 2u;
                const uint32_t hx = xxhash32_uint2(dimension, seed), hy = xxhash32_uint2(dimension + 1u, seed);
                const float x = static_cast<float>(fast_owen_scramble(hx, sobol_bits(c, index, dimension))) * 0x1p-32f;
                const float y = static_cast<float>(fast_owen_scramble(hy, sobol_bits(c, index, dimension + 1u))) * 0x1p-32f;
                dimension += 2u;
                result = make_float2(fminf(fmaxf(x, 0.f), kOneMinusEpsilon), fminf(fmaxf(y, 0.f), kOneMinusEpsilon));
                break;
            }
            case LRK_SAMPLER_PADDED_SOBOL: {// padded_sobol.cpp:134-146
                const uint32_t hx = xxhash32_uint4(px, py, sample_index ^ seed, dimension);
                const uint32_t hy = xxhash32_uint4(px, py, sample_index ^ seed, dimension + 1u);
                const uint32_t idx = permutation_element(sample_index, c.spp, mask_covering(c.spp - 1u), hx);
                dimension += 2u;
                result = make_float2(fminf(static_cast<float>(fast_owen_scramble(hx, sobol_bits(c, idx, 0u))) * 0x1p-32f, kOneMinusEpsilon),
                                     fminf(static_cast<float>(fast_owen_scramble(hy, sobol_bits(c, idx, 1u))) * 0x1p-32f, kOneMinusEpsilon));
                break;
            }
            default: {// ZSOBOL, zsobol.cpp:170-177
                const uint64_t si = zsobol_sample_index(c, index, dimension);
                const uint32_t hx = __ldg(c.zsobol_hash + dimension * 2u), hy = __ldg(c.zsobol_hash + dimension * 2u + 1u);
                dimension = (dimension + 2u) % 1024u;
                result = make_float2(fminf(static_cast<float>(fast_owen_scramble(hx, sobol_bits(c, si, 0u))) * 0x1p-32f, kOneMinusEpsilon),
                                     fminf(static_cast<float>(fast_owen_scramble(hy, sobol_bits(c, si, 1u))) * 0x1p-32f, kOneMinusEpsilon));
                break;
            }
        }
        return Draw{result.x, result.y, dimension};
    }
    __device__ __noinline__ static Draw table_pixel_2d(const lrk_sampler &c, uint32_t seed, uint32_t dimension, uint32_t px, uint32_t py,
                                                       uint32_t sample_index, uint64_t index) {
        if (c.type == LRK_SAMPLER_PMJ02BN) {// pmj02bn.cpp:209-213
            const uint32_t tx = px % c.tile, ty = py % c.tile;
            const size_t offset = static_cast<size_t>(tx + ty * c.tile) * c.spp + sample_index;
            return Draw{__ldg(c.pmj_pixel_samples + offset * 2u), __ldg(c.pmj_pixel_samples + offset * 2u + 1u), dimension};
        }
        if (c.type == LRK_SAMPLER_SOBOL) {// sobol.cpp:164-170
            const float x = static_cast<float>(sobol_bits(c, index, 0u)) * 0x1p-32f, y = static_cast<float>(sobol_bits(c, index, 1u)) * 0x1p-32f;
            const float s = static_cast<float>(c.scale);
            return Draw{fminf(fmaxf(x * s - static_cast<float>(px), 0.f), kOneMinusEpsilon),
                        fminf(fmaxf(y * s - static_cast<float>(py), 0.f), kOneMinusEpsilon), dimension};
        }
        return table_2d(c, seed, dimension, px, py, sample_index, index);
    }

    // ---- the interface the kernels use -------------------------------------------------------------------------------------------
    // Sampler::Instance::start(pixel, sample_index)
    __device__ __forceinline__ void start(SamplerRef sc, uint32_t x, uint32_t y, uint32_t s) {
        px = x;
        py = y;
        sample_index = s;
        if (sc.type == LRK_SAMPLER_INDEPENDENT) {
            state = xxhash32_uint4(x, y, sc.seed, s);
        } else {
            state = (sc.type == LRK_SAMPLER_PMJ02BN || sc.type == LRK_SAMPLER_SOBOL) ? 2u : 0u;
            index = derive(*sc.rec, x, y, s);
        }
    }
    // Sampler::Instance::load_state: the word the path carried + what follows from its generation slot
    __device__ __forceinline__ void resume(SamplerRef sc, uint32_t word, uint32_t x, uint32_t y, uint32_t s) {
        state = word;
        if (sc.type != LRK_SAMPLER_INDEPENDENT) {
            px = x;
            py = y;
            sample_index = s;
            index = derive(*sc.rec, x, y, s);
        }
    }
    __device__ __forceinline__ float next1d(SamplerRef sc) {
        if (sc.type == LRK_SAMPLER_INDEPENDENT) return lcg(state);
        const Draw d = table_1d(*sc.rec, sc.seed, state, px, py, sample_index, index);
        state = d.dimension;
        return d.x;
    }
    __device__ __forceinline__ float2 next2d(SamplerRef sc) {
        if (sc.type == LRK_SAMPLER_INDEPENDENT) {
            const float a = lcg(state);
            const float b = lcg(state);
            return make_float2(a, b);
        }
        const Draw d = table_2d(*sc.rec, sc.seed, state, px, py, sample_index, index);
        state = d.dimension;
        return make_float2(d.x, d.y);
    }
    __device__ __forceinline__ float2 pixel2d(SamplerRef sc) {// generate_pixel_2d, sampler.h:48
        if (sc.type == LRK_SAMPLER_INDEPENDENT) return next2d(sc);
        const Draw d = table_pixel_2d(*sc.rec, sc.seed, state, px, py, sample_index, index);
        state = d.dimension;
        return make_float2(d.x, d.y);
    }
};

}// namespace lrk
