// ParticleRng = rand_xoshiro::Xoshiro256PlusPlus (examples/stress_tests/particles.rs:125-128), kept host-side by the
// engine and rolled back with every snapshot like the reference's rollback_resource_with_clone::<ParticleRng>() (:200).
//
// Third-party arithmetic (not under /root/reference), restated from the published algorithms:
//   SplitMix64 (Vigna, splitmix64.c)            rand_xoshiro 0.7 `SplitMix64` — seeds the generator below
//   xoshiro256++ (Blackman & Vigna)             rand_xoshiro 0.7 `Xoshiro256PlusPlus`
//       seed_from_u64(s): state = the first four outputs of SplitMix64(s)   (rand_xoshiro overrides rand_core's default)
//       next_u64 = rotl(s0 + s3, 23) + s0 followed by the xoshiro256 state update; next_u32 = upper half of next_u64
//   rand 0.9 `random_range(low..high)` for f32 (UniformFloat::sample_single):
//       value1_2 = f32::from_bits((next_u32 >> 9) | 0x3f80_0000); value0_1 = value1_2 - 1.0
//       res = value0_1 * (high - low) + low   (Rust never contracts: mul and add are rounded separately)
//       returned when res < high.  rand's edge-case branch (shrink `scale` by one ulp and draw again) is unreachable
//       for the example's (-200.0, 200.0): the largest value0_1 = 1 - 2^-23 gives 199.99994 < 200
//       (tests/test_third_party_kats.py proves it over all 2^23 mantissas), so it is not restated.
// Pins: tests/golden/third_party_kats.json holds the published known-answer vectors of splitmix64.c and
// xoshiro256plusplus.c as rand_xoshiro's own unit tests quote them; bgr_particle_rng_stream exposes this code to them.
#pragma once
#include <cstdint>
#include <cstring>

namespace bgr {

struct SplitMix64 {
    uint64_t x = 0;
    uint64_t next_u64() {
        x += 0x9E3779B97F4A7C15ULL;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
};

struct ParticleRng {
    uint64_t s[4] = {0, 0, 0, 0};
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    void seed_from_u64(uint64_t seed) {
        SplitMix64 sm{seed};
        for (int i = 0; i < 4; ++i) s[i] = sm.next_u64();
    }
    uint64_t next_u64() {
        const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return result;
    }
    uint32_t next_u32() { return uint32_t(next_u64() >> 32); }
    float random_range(float low, float high) {  // rng.random_range(low..high)
        const float scale = high - low;
        for (;;) {
            const uint32_t bits = (next_u32() >> 9) | 0x3f800000u;
            float v12; std::memcpy(&v12, &bits, 4);
            volatile float v01 = v12 - 1.0f;
            volatile float prod = v01 * scale;   // mul and add rounded separately (no contraction)
            const float res = prod + low;
            if (res < high) return res;
        }
    }
};

}  // namespace bgr
