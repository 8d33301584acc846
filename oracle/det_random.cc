// TEST INFRASTRUCTURE — deterministic stand-in for the reference's src/utils/random.cc.
// The reference seeds its thread_local generators from the hash of the thread id
// (/root/reference/src/utils/random.cc:21-26, random.h:63-66) and has no seed flag, so two runs (or two
// backends) never see the same random stream.  For the "identical MCTS visit counts under a fixed seed" check
// the harness binaries (oracle/_ref/sayuri_*_det) link THIS translation unit instead: every thread's generator
// is seeded from the environment variable SAYURI_SEED (default 1), everything else — xoroshiro128+ / splitmix64
// as published by Blackman & Vigna / Steele et al., the same generators the reference names in random.h:13 — is
// unchanged in behaviour.  Implements the member functions declared in the reference's utils/random.h.
#include <cstdlib>

#include "utils/random.h"

namespace {

inline std::uint64_t Mix(std::uint64_t z) {  // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

inline std::uint64_t RotL(std::uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

std::uint64_t FixedSeed(std::uint64_t requested) {
    if (requested != kThreadSeed && requested != kTimeSeed) return requested;
    const char* env = std::getenv("SAYURI_SEED");
    return env ? std::strtoull(env, nullptr, 10) : 1ull;
}

}  // namespace

template <RandomMethod R> thread_local std::uint64_t Random<R>::seeds_[Random<R>::kMaxSeedSize];

template <> void Random<RandomMethod::kSplitMix64>::InitSeed(std::uint64_t seed) {
    seed = FixedSeed(seed);
    for (auto i = size_t{0}; i < kMaxSeedSize; ++i) {
        seed = Mix(seed);
        seeds_[i] = seed;
    }
}

template <> void Random<RandomMethod::kXoroShiro128Plus>::InitSeed(std::uint64_t seed) {
    seed = FixedSeed(seed);
    for (auto i = size_t{0}; i < kMaxSeedSize; ++i) {
        seed = Mix(seed);
        seeds_[i] = seed;
    }
}

template <> std::uint64_t Random<RandomMethod::kSplitMix64>::Generate() {
    seeds_[kMaxSeedSize - 1] += 0x9e3779b97f4a7c15ull;
    std::uint64_t z = seeds_[kMaxSeedSize - 1];
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

template <> std::uint64_t Random<RandomMethod::kXoroShiro128Plus>::Generate() {
    const std::uint64_t a = seeds_[0];
    std::uint64_t b = seeds_[1];
    const std::uint64_t out = a + b;
    b ^= a;
    seeds_[0] = RotL(a, 55) ^ b ^ (b << 14);
    seeds_[1] = RotL(b, 36);
    return out;
}
