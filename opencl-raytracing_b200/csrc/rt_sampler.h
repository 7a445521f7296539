// rt_sampler.h — the reference's sampler sequences on the device (crates/raytracing-cpu/src/sample.rs).
//
// The per-sample stream is a pure function of (seed, x, y, sample_index): stream = FxHash(x, y, s),
// rng = Pcg32::new(FxHash(seed), stream) (sample.rs:29-87). Carrying only the 64-bit PCG state (and
// the stratified `dimension`) in the path state is therefore enough to resume a path in any kernel
// of the wavefront, and any image / sample partition across GPUs reproduces the same draws.
//
// Third-party arithmetic restated from the published algorithms: rustc-hash 2.1.1 FxHasher,
// rand_pcg 0.9.0 Lcg64Xsh32, rand 0.9.2 StandardUniform<f32> and UniformInt<u32>::sample_single.
#pragma once
#include "rt_common.h"

namespace rt {

struct FxHasher {
    uint64_t hash;
    RT_HD FxHasher() : hash(0) {}
    RT_HD void add(uint64_t i) { hash = (hash + i) * 0xf1357aea2e62a9c5ull; }
    RT_HD void write_u32(uint32_t v) { add(v); }
    RT_HD void write_u64(uint64_t v) { add(v); }
    RT_HD uint64_t finish() const { return (hash << 26) | (hash >> 38); }
};

struct Pcg32 {
    uint64_t state, inc;
    RT_HD void seed(uint64_t st, uint64_t stream) {
        inc = (stream << 1) | 1;
        state = st + inc;
        step();
    }
    RT_HD void step() { state = state * 6364136223846793005ull + inc; }
    RT_HD uint32_t next_u32() {
        uint64_t old = state;
        step();
        uint32_t rot = (uint32_t)(old >> 59);
        uint32_t xsh = (uint32_t)(((old >> 18) ^ old) >> 27);
        return (xsh >> rot) | (xsh << ((32 - rot) & 31));
    }
    RT_HD float next_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
    RT_HD uint32_t range_u32(uint32_t lo, uint32_t hi) {  // Canon's method, one bias-reduction draw
        uint32_t range = hi - lo;
        if (range == 0) return next_u32();
        uint64_t m = (uint64_t)next_u32() * range;
        uint32_t result = (uint32_t)(m >> 32), lo_order = (uint32_t)m;
        if (lo_order > (uint32_t)(0u - range)) {
            uint64_t m2 = (uint64_t)next_u32() * range;
            uint32_t new_hi = (uint32_t)(m2 >> 32);
            if ((uint64_t)lo_order + new_hi > 0xffffffffull) result += 1;
        }
        return lo + result;
    }
};

// sample.rs:228-254
RT_HD uint32_t permute(uint32_t index, uint32_t length, uint32_t seed) {
    uint32_t npot = 1;
    while (npot < length) npot <<= 1;
    uint32_t mask = npot - 1;
    for (;;) {
        index ^= seed;
        index *= 0xe170893du;
        index ^= seed >> 16;
        index ^= (index & mask) >> 4;
        index ^= seed >> 8;
        index *= 0x0929eb3fu;
        index ^= seed >> 23;
        index ^= (index & mask) >> 1;
        index *= (1u | seed >> 27);
        index *= 0x6935fa69u;
        index ^= (index & mask) >> 11;
        index *= 0x74dcb303u;
        index ^= (index & mask) >> 2;
        index *= 0x9e501cc3u;
        index ^= (index & mask) >> 2;
        index *= 0xc860a3dfu;
        index &= mask;
        index ^= index >> 5;
        if (index < length) return (index + seed) % length;
    }
}

// Launch-constant sampler parameters (RaytracerSettings::sampler + hashed seed).
struct SamplerParams {
    uint64_t seed_hashed;  // FxHash(seed.unwrap_or(42)), sample.rs:30-35
    uint32_t stratified;
    uint32_t jitter;
    uint32_t x_strata, y_strata;
};

RT_HD uint64_t hash_seed(uint64_t seed) {
    FxHasher h;
    h.write_u64(seed);
    return h.finish();
}

// CpuSampler (sample.rs:8-181)
struct Sampler {
    Pcg32 rng;
    uint64_t seed;  // hashed
    uint32_t dimension, sample_index;
    uint32_t stratified, jitter, x_strata, y_strata;

    RT_HD void init(const SamplerParams& p) {
        seed = p.seed_hashed;
        stratified = p.stratified;
        jitter = p.jitter;
        x_strata = p.x_strata;
        y_strata = p.y_strata;
        dimension = 0;
        sample_index = 0;
        rng.seed(seed, 0);
    }
    RT_HD void init_one_off(uint64_t s) {  // sample.rs:59-64 (seed used as is)
        seed = s;
        stratified = 0;
        jitter = 1;
        x_strata = y_strata = 1;
        dimension = 0;
        sample_index = 0;
        rng.seed(s, 0);
    }
    RT_HD static uint64_t stream_of(uint32_t px, uint32_t py, uint32_t sidx) {
        FxHasher h;
        h.write_u32(px);
        h.write_u32(py);
        h.write_u32(sidx);
        return h.finish();
    }
    RT_HD void start_sample(uint32_t px, uint32_t py, uint32_t sidx) {  // sample.rs:69-87
        rng.seed(seed, stream_of(px, py, sidx));
        dimension = 0;
        sample_index = sidx;
    }
    // resume a path: state carried by the wavefront, inc recomputed from (x, y, s)
    RT_HD void resume(uint32_t px, uint32_t py, uint32_t sidx, uint64_t state, uint32_t dim) {
        rng.inc = (stream_of(px, py, sidx) << 1) | 1;
        rng.state = state;
        dimension = dim;
        sample_index = sidx;
    }
    RT_HD uint32_t dim_hash() const {
        FxHasher h;
        h.write_u32(dimension);
        h.write_u64(seed);
        return (uint32_t)h.finish();
    }
    RT_HD float uniform() {  // sample.rs:89-121
        if (!stratified) return rng.next_f32();
        uint32_t total = x_strata * y_strata;
        uint32_t strata = permute(sample_index, total, dim_hash());
        float delta = jitter ? rng.next_f32() : 0.5f;
        dimension += 1;
        return ((float)strata + delta) / (float)total;
    }
    RT_HD uint32_t u32_range(uint32_t lo, uint32_t hi) {  // sample.rs:123-138
        if (!stratified) return rng.range_u32(lo, hi);
        float u = uniform();
        float offset = u * (float)(hi - lo);
        return lo + rs_as_u32(offset);
    }
    RT_HD V2 uniform2() {  // sample.rs:140-180
        if (!stratified) {
            float a = rng.next_f32();
            float b = rng.next_f32();
            return mk2(a, b);
        }
        uint32_t hash = dim_hash();
        uint32_t total = x_strata * y_strata;
        uint32_t strata = permute(sample_index, total, hash);
        dimension += 2;
        uint32_t y = strata / x_strata, x = strata % x_strata;
        float dx = 0.5f, dy = 0.5f;
        if (jitter) { dx = rng.next_f32(); dy = rng.next_f32(); }
        return mk2(((float)x + dx) / (float)x_strata, ((float)y + dy) / (float)y_strata);
    }
};

// sample.rs:184-224
RT_HD V2 sample_unit_disk(V2 u) {
    float r = sqrtf(u.x);
    float theta = 2.0f * PI * u.y;
    return mk2(r * cosf(theta), r * sinf(theta));
}
RT_HD V2 sample_unit_disk_concentric(V2 u) {
    V2 o = 2.0f * u - mk2(1.0f, 1.0f);
    if (o.x == 0.0f && o.y == 0.0f) return mk2(0, 0);
    float theta, r;
    if (fabsf(o.x) > fabsf(o.y)) { theta = FRAC_PI_4 * (o.y / o.x); r = o.x; }
    else { theta = FRAC_PI_2 - FRAC_PI_4 * (o.x / o.y); r = o.y; }
    return r * mk2(cosf(theta), sinf(theta));
}
RT_HD V3 sample_cosine_hemisphere(V2 u) {
    V2 d = sample_unit_disk(u);
    float z = sqrtf(fmaxf(1.0f - d.x * d.x - d.y * d.y, 0.0f));
    return mk3(d.x, d.y, z);
}
RT_HD float sample_exponential(float u, float a) { return -logf(1.0f - u) / a; }
RT_HD float power_heuristic(uint32_t na, float pa, uint32_t nb, float pb) {
    float wa = ((float)na * pa) * ((float)na * pa);
    float wb = ((float)nb * pb) * ((float)nb * pb);
    return wa / (wa + wb);
}

}  // namespace rt
