// shim: <curanddx.hpp>. The reference's OptiX sampler draws from cuRANDDx's PCG; the CPU renderer (the parity target) draws
// from rand_pcg::Pcg32 with another seeding and another hash (SURVEY §8c lists the divergence), so sampler STREAMS are not
// part of the _ref comparisons. This stand-in only has to make `sample.hpp` compile and hand out uniform floats: a PCG32
// XSH-RR with the 24-byte footprint that makes the reference's static_assert(sizeof(PerRayData) == 64) hold.
#pragma once
#include <stdint.h>
namespace curanddx {
struct pcg {};
template <int N> struct SM {};
struct Thread {};
template <typename G> struct Generator {
    uint64_t state = 0, inc = 1, spare = 0;
    Generator() = default;
    Generator(unsigned long long seed, unsigned long long subsequence, unsigned long long /*offset*/) {
        inc = (subsequence << 1) | 1u;
        state = seed + inc;
        generate();
    }
    unsigned int generate() {
        uint64_t old = state;
        state = old * 6364136223846793005ull + inc;
        uint32_t xsh = (uint32_t)(((old >> 18) ^ old) >> 27), rot = (uint32_t)(old >> 59);
        return (xsh >> rot) | (xsh << ((32 - rot) & 31));
    }
};
template <typename G, typename X> Generator<G> operator+(Generator<G> g, X) { return g; }
struct uniform {
    float lo, hi;
    template <typename R> float generate(R& rng) const { return lo + (hi - lo) * ((float)(rng.generate() >> 8) * (1.0f / 16777216.0f)); }
};
}  // namespace curanddx
